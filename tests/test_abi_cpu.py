"""CPU: the C-ABI library builds, loads and exports every function include/coskad_b200.h declares;
without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'coskad_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(coskad_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/coskad_b200.h but not exported'


def test_binding_table_covers_header(built_lib):
    from coskad_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.coskad_abi_version() == 2


def test_sm100_code_only(built_lib):
    import subprocess
    out = subprocess.run(['cuobjdump', '-lelf', built_lib], capture_output=True, text=True).stdout
    assert 'sm_100a' in out and not re.search(r'sm_(?!100a)\d+', out), out


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-GPU failure mode')
def test_no_cpu_fallback(built_lib):
    from coskad_b200 import _lib, sts
    with pytest.raises(_lib.CoskadError):
        _lib.Context(0)
    m = sts.STSE(2, [32, 16, 32], 64, 16, 12, 17).eval()
    with pytest.raises(_lib.CoskadError):
        m(torch.zeros(4, 2, 12, 17))


def test_module_tree_matches_reference_names():
    from coskad_b200 import sts
    from oracle import stsgcn as onet
    for kind, cls, d in (('stse', sts.STSE, 16), ('stsae', sts.STSAE, 8)):
        m = cls(2, [32, 16, 32], 64, d, 12, 17)
        sd = onet.init_state_dict(kind, latent_dim=d)
        assert sorted(m.state_dict().keys()) == sorted(sd.keys())
        for k, v in m.state_dict().items():
            assert tuple(v.shape) == tuple(sd[k].shape), k
        m.load_state_dict(sd, strict=True)
    assert sum(p.numel() for p in sts.STSE(2, [32, 16, 32], 64, 16, 12, 17).parameters()) == 239716


def test_unsupported_configs_fail_loudly():
    from coskad_b200 import sts
    with pytest.raises(ValueError):
        sts.STSE(2, [32, 16, 32], 64, 16, 12, 17, projector='mlp')
    with pytest.raises(ValueError):
        sts.STSE(2, [32, 16, 32], 64, 16, 12, 17, encoder_type='st_gcn')
    with pytest.raises(NotImplementedError):
        sts.STSE(2, [32, 16, 32], 64, 16, 12, 17, dropout=0.1)

"""The drop-in boundary against the REAL thing: the reference's own five Lightning task modules (unmodified files in
baseline/_ref, copied there by oracle/install_ref.py; /root/reference when it is visible) import under
coskad_b200.compat.install() and construct from the reference's own YAML configs (projector: linear -- 'mlp' is broken
upstream, models/common/components.py:218).  CPU only: construction builds parameter containers, no kernel runs."""
import argparse
import importlib
import os
import subprocess
import sys

import pytest
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ref_root():
    for cand in (os.path.join(ROOT, 'baseline', '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(cand, 'models')):
            return cand
    return None


REF = _ref_root()
needs_ref = pytest.mark.skipif(REF is None, reason='no reference tree (baseline/_ref absent: run oracle/install_ref.py)')

MODULES = [('models.hyperbolic_encoder', 'LitEncoder', 'config/UBnormal/hyperbolic_encoder.yaml'),
           ('models.euclidean_encoder_dynamicCenter', 'LitEncoder', 'config/UBnormal/euclidean_encoder.yaml'),
           ('models.euclidean_encoder_staticCenter', 'LitEncoder', 'config/STC/euclidean_encoder.yaml'),
           ('models.euclidean_autoencoder', 'LitAutoEncoder', 'config/UBnormal/euclidean_autoencoder.yaml'),
           ('models.spherical_vae', 'LitEncoder', 'config/UBnormal/spherical_vae.yaml')]


def _args(cfg: str) -> argparse.Namespace:
    with open(os.path.join(REF, cfg)) as f:
        text = f.read()
    # reference defect: config/UBnormal/euclidean_autoencoder.yaml:14 has an unescaped quote ('/path_to_model's_checkpoint')
    text = '\n'.join("load_ckpt: ''" if ln.startswith('load_ckpt:') else ln for ln in text.splitlines())
    d = yaml.load(text, Loader=yaml.FullLoader)
    d['projector'] = 'linear'
    d['encoder_type'] = 'STS_GCN'      # config/UBnormal/euclidean_encoder.yaml:35 asks for an ablation encoder (out of scope)
    d.setdefault('num_centers', 1)
    return argparse.Namespace(**d)


@needs_ref
def test_reference_task_modules_import_and_construct_in_a_fresh_interpreter():
    """a fresh interpreter with NOTHING of the reference imported before install(): the case that failed in round 1
    ('models' is not a package / no LightningDataModule)"""
    code = f'''
import sys
sys.path.insert(0, {ROOT!r})
import coskad_b200.compat as compat
compat.install(reference={REF!r})
import importlib
for name in {[m for m, _, _ in MODULES]!r}:
    mod = importlib.import_module(name)
    assert hasattr(mod, "LitDataModule"), name
    assert mod.__file__.startswith({REF!r}), mod.__file__
import models.sts.ae, models.sts.vae, utils.eval_utils
print("ok")
'''
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd='/tmp')
    assert res.returncode == 0 and res.stdout.strip().endswith('ok'), res.stdout + res.stderr


@needs_ref
@pytest.mark.parametrize('modname,clsname,cfg', MODULES)
def test_reference_task_module_constructs_from_its_yaml(modname, clsname, cfg):
    import coskad_b200.compat as compat
    from coskad_b200 import sts
    compat.install(reference=REF)
    mod = importlib.import_module(modname)
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REF))          # the reference's file, not a re-host
    lit = getattr(mod, clsname)(_args(cfg))
    assert isinstance(lit.model, sts.STSE)                                            # ... running on the CUDA-backed classes
    keys = lit.state_dict().keys()
    assert 'model.encoder.model.0.gcn.A' in keys and 'model.encoder.model.3.tcn.1.running_var' in keys
    if clsname == 'LitAutoEncoder':
        assert 'model.rev_btlnk.weight' in keys and 'model.decoder.model.3.prelu.weight' in keys
    if modname.endswith('spherical_vae'):
        assert 'model.fc_mean.weight' in keys and 'model.fc_var.bias' in keys
    # the hooks the trainer drives exist on the reference's class
    for hook in ('training_step', 'validation_step', 'configure_optimizers', 'forward'):
        assert callable(getattr(lit, hook))
    dm = mod.LitDataModule(batch_size=4, train_dataset=[1, 2, 3])
    assert dm.hparams.batch_size == 4


@needs_ref
def test_install_extends_the_reference_models_package_instead_of_replacing_it():
    import coskad_b200.compat as compat
    compat.install(reference=REF)
    import models
    assert hasattr(models, '__path__')
    from models.sts.ae import STSE as RefSTSE                 # the reference's in-tree torch module still imports
    from models.stse.stse_hidden_hypersphere import STSE     # ... beside the shim at the missing path
    assert RefSTSE is not STSE and RefSTSE.__module__ == 'models.sts.ae'
    compat.install(reference=REF)                             # idempotent
    from models.stse.stse_hidden_hypersphere import STSE as again
    assert again is STSE


def test_install_without_any_reference_creates_an_empty_models_package():
    code = f'''
import sys
sys.path.insert(0, {ROOT!r})
import coskad_b200.compat as compat
compat.install(reference="")
compat.reference_root = lambda: None
from models.stse.stse_hidden_hypersphere import STSE
from models.stsae.stsae_hidden_hypersphere import STSAE
from models.stsve.stsve_hidden_hypersphere import STSVE
import pytorch_lightning as pl
from pytorch_lightning.callbacks import ModelCheckpoint
from pytorch_lightning.strategies import DDPStrategy
assert hasattr(pl, "LightningDataModule") and hasattr(pl.Trainer, "from_argparse_args")
print("ok")
'''
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd='/tmp')
    assert res.returncode == 0 and res.stdout.strip().endswith('ok'), res.stdout + res.stderr

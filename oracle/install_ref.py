"""Populate baseline/_ref with the UNMODIFIED reference tree (test / benchmark infrastructure, not product code).

The reference is pure Python without setup.py / pyproject.toml, so ``pip install --target baseline/_ref /root/reference``
has nothing to build; this script copies the importable packages (models/, utils/) and the YAML configs as they are.
baseline/_ref is git-ignored (no reference source enters the history) but not gpurun-ignored, so the copy travels to the
GPU box, where /root/reference does not exist.  Used by: bench.py --impl reference (the reference's own
models.sts.ae.STSE on the host cores), bench.py's ref_gpu_eager leg, tests/test_compat_*.py (the reference's own
Lightning modules driven through coskad_b200.compat).

    python oracle/install_ref.py [--src /root/reference]
"""
from __future__ import annotations

import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, 'baseline', '_ref')
KEEP = ('models', 'utils', 'config', 'train_COSKAD.py', 'eval_COSKAD.py', 'environment.yml')


def install(src: str = '/root/reference') -> bool:
    if not os.path.isdir(os.path.join(src, 'models')):
        return os.path.isdir(os.path.join(DST, 'models'))
    os.makedirs(DST, exist_ok=True)
    for name in KEEP:
        s, d = os.path.join(src, name), os.path.join(DST, name)
        if os.path.isdir(s):
            shutil.rmtree(d, ignore_errors=True)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
        elif os.path.isfile(s):
            shutil.copy2(s, d)
    return True


if __name__ == '__main__':
    src = sys.argv[sys.argv.index('--src') + 1] if '--src' in sys.argv else '/root/reference'
    print('baseline/_ref', 'installed' if install(src) else 'NOT available', 'from', src)

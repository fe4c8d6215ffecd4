"""Build recipe for libcoskad_b200.so (in-tree, sm_100a only).

``python -m coskad_b200._build`` or ``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB_PATH = os.path.join(HERE, 'libcoskad_b200.so')
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _sources():
    out = []
    for root in (CSRC, os.path.join(os.path.dirname(HERE), 'include')):
        for fn in sorted(os.listdir(root)):
            if fn.endswith(('.cu', '.cuh', '.h')):
                out.append(os.path.join(root, fn))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in _sources())


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source into coskad_b200/libcoskad_b200.so; returns the ptxas log."""
    if not force and not needs_build():
        return ''
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libcoskad_b200.so')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', LIB_PATH + '.tmp', os.path.join(CSRC, 'abi.cu')]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    os.replace(LIB_PATH + '.tmp', LIB_PATH)
    log = res.stdout + res.stderr
    if verbose:
        print(log)
    return log


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))

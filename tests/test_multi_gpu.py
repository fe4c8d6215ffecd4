"""Multi-GPU correctness under pytest -m gpu: spawns tests/multi_gpu_check.py with 2 ranks (torchrun, NCCL) when the box
has at least 2 GPUs -- sharded scores bit-identical to the single-GPU scores, all-reduced center, averaged gradients, and
the graph-replayed data-parallel training step.  On a 1-GPU box the test is skipped; bench.py's `parity_multi` key covers
the sharded-score identity in every N > 1 benchmark run."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_nccl_check():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (gpurun --gpus 2)')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(_free_port()), os.path.join(ROOT, 'tests', 'multi_gpu_check.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    out = res.stdout + res.stderr
    assert res.returncode == 0 and 'multi_gpu_check world=2' in out and '-> OK' in out, out[-4000:]

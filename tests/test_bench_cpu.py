"""CPU: the reference arm of bench.py runs without a GPU and prints exactly one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--cpu-sample', '512'], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'pose windows/sec scored' and d['unit'] == 'windows/s'
    assert d['higher_is_better'] is True and d['value'] > 0 and d['steps'] == 1 and d['warmup'] == 1
    # the unmodified reference network from baseline/_ref when that copy exists (oracle/install_ref.py), else the oracle port
    have_ref = os.path.isdir(os.path.join(ROOT, 'baseline', '_ref', 'models', 'sts'))
    assert d['cpu_baseline']['kind'] == ('reference' if have_ref else 'port'), d['cpu_baseline']
    assert d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': 'windows/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert d['gpu_launches'] == 0 and d['vs_baseline'] is None


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1',
                          '--warmup', '1'], capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ''

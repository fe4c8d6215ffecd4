#!/bin/bash
# quick GPU check: fused-eval parity + short bench (run under gpurun)
timeout 300 python -m pytest tests/test_fused_eval_gpu.py -x -q 2>&1 | tail -4
timeout 300 python bench.py --steps 4 --warmup 3 --windows-per-step 262144 --resident-windows 1048576 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('windows/s', d['value'], 'ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"

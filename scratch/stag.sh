#!/bin/bash
for rep in 1 2; do
for v in "$@"; do
  COSKAD_STAGGER=$v timeout 300 python bench.py --steps 6 --warmup 3 --windows-per-step 262144 --resident-windows 1048576 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stagger $v', 'windows/s %.0f' % d['value'], 'ms %.3f' % d['ms_per_step'])"
done
done

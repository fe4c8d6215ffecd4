#!/bin/bash
for c in "$@"; do
  python bench.py --steps 4 --warmup 3 --windows-per-step 1048576 --resident-windows 2097152 --no-cpu-baseline --e2e-chunk $c 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('chunk $c', 'device %.0f' % d['value'], 'e2e %.0f' % d['e2e']['value'])"
done

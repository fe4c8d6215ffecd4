#!/bin/bash
# A/B bench of prebuilt library variants on ONE box: tools/ab.sh name1 name2 ... (tools/variants/<name>.so)
cp coskad_b200/libcoskad_b200.so /tmp/orig.so
for rep in 1 2; do
for v in "$@"; do
  cp tools/variants/$v.so coskad_b200/libcoskad_b200.so
  timeout 300 python bench.py --steps 6 --warmup 3 --windows-per-step 262144 --resident-windows 1048576 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', 'windows/s %.0f' % d['value'], 'ms %.3f' % d['ms_per_step'])"
done
done
cp /tmp/orig.so coskad_b200/libcoskad_b200.so

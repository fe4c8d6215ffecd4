#!/bin/bash
# dram bytes of the fused kernel per variant (one small launch under ncu), then the usual A/B
cp coskad_b200/libcoskad_b200.so /tmp/orig.so
for v in "$@"; do
  cp tools/variants/$v.so coskad_b200/libcoskad_b200.so
  timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:fused_eval_tc -s 3 -c 1 python bench.py --steps 2 --warmup 3 --windows-per-step 131072 --resident-windows 131072 --no-cpu-baseline --no-e2e 2>&1 | grep -E "dram__bytes" | awk -v v=$v '{print v, $1, $2, $3}'
done
cp /tmp/orig.so coskad_b200/libcoskad_b200.so

#!/bin/bash
cp coskad_b200/libcoskad_b200.so /tmp/orig.so
for rep in 1 2; do for v in "$@"; do
  cp tools/variants/$v.so coskad_b200/libcoskad_b200.so
  echo "== $v"; timeout 200 python tools/ae_bench.py 2>&1 | tail -6
done; done
cp /tmp/orig.so coskad_b200/libcoskad_b200.so

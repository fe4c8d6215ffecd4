#!/bin/bash
# usage: tools/ncu_csv.sh <name> <kernel regex> <skip> <count> <command...>  -- ncu --set full capture, raw page exported as CSV on the box
# (the .ncu-rep stays in /tmp on the box: gpurun_out/ is capped at 64 MiB)
name=$1; regex=$2; skip=$3; count=$4; shift 4
mkdir -p gpurun_out
ncu --set full --clock-control none -k "regex:$regex" -s "$skip" -c "$count" -o /tmp/$name "$@" > gpurun_out/$name.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
ls -la /tmp/$name.ncu-rep gpurun_out/$name.raw.csv

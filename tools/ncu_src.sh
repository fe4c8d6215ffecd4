#!/bin/bash
# usage: tools/ncu_src.sh <name> <kernel regex> <skip> <command...> -- one launch, --set full with source, raw + source pages as CSV
name=$1; regex=$2; skip=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$regex" -s "$skip" -c 1 -o /tmp/$name "$@" > gpurun_out/$name.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
ls -la gpurun_out/$name.*

#!/bin/bash
# ncu --set full capture of ONE timed launch of the shipped fused eval kernel inside bench.py (1 Mi windows per launch),
# exported on the box as raw + source CSV (the .ncu-rep stays in /tmp: gpurun_out/ is capped at 64 MiB)
name=${1:-r02_eval}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/$name.bench.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:fused_eval_tc -s 4 -c 1 \
    --metrics lts__t_bytes.sum,lts__t_sectors_srcunit_tex.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sector_hit_rate.pct,sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active \
    -o /tmp/$name python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary > gpurun_out/$name.ncu.log 2>&1
ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
ncu -i /tmp/$name.ncu-rep --page source --csv > gpurun_out/$name.source.csv 2>/dev/null
ls -la /tmp/$name.ncu-rep gpurun_out/$name.*

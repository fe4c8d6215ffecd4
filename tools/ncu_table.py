"""usage: python tools/ncu_table.py <raw.csv> -- one line per captured launch from `ncu --page raw --csv`"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]; col = {k: i for i, k in enumerate(h)}
keys = [('us', 'gpu__time_duration.sum', 1e-3), ('grid', 'launch__grid_size', 1), ('regs', 'launch__registers_per_thread', 1),
        ('rdMB', 'dram__bytes_read.sum', None), ('wrMB', 'dram__bytes_write.sum', None),
        ('dram%', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 1),
        ('warps%', 'sm__warps_active.avg.pct_of_peak_sustained_active', 1),
        ('issue%', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 1),
        ('fma%', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 1),
        ('lsu%', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 1),
        ('l1%', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 1),
        ('st_long', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 1),
        ('st_bar', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 1),
        ('st_short', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 1),
        ('st_mio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 1),
        ('st_lg', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 1),
        ('st_wait', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 1)]
units = rows[1]
print('kernel'.ljust(44), ' '.join(k[0].rjust(8) for k in keys))
for r in rows[2:]:
    out = []
    for name, k, sc in keys:
        if k not in col or not r[col[k]]:
            out.append('-'.rjust(8)); continue
        v = float(r[col[k]].replace(',', ''))
        u = units[col[k]]
        if sc is None:
            v = v / 1e6 if u == 'byte' else (v / 1e3 if u == 'Kbyte' else (v if u == 'Mbyte' else v * 1e3 if u == 'Gbyte' else v))
        elif name == 'us':
            v = {'ns': v * 1e-3, 'us': v, 'ms': v * 1e3, 'usecond': v, 'msecond': v*1e3, 'nsecond': v*1e-3}.get(u, v)
        out.append(f'{v:8.1f}')
    print(r[col['Kernel Name']][:44].ljust(44), ' '.join(out))

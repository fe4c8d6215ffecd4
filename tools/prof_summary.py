"""usage: python tools/prof_summary.py gpurun_out/prof.ncu-rep profiles/name.md "title"
       python tools/prof_summary.py gpurun_out/name.raw.csv profiles/name.md "title"    (CSV pages exported on the GPU box by
       tools/ncu_eval.sh / tools/ncu_csv.sh: name.raw.csv + name.source.csv; the .ncu-rep itself exceeds the gpurun_out/ cap)"""
import csv, os, subprocess, sys, io, collections
rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
from_csv = rep.endswith('.raw.csv')
raw = open(rep).read() if from_csv else subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keys = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.max', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum']
lines = [f'# {title}', '', f'source: `{rep}` (ncu --set full --clock-control none --import-source on), one launch', '',
         '| metric | value | unit |', '|---|---|---|']
for k in keys:
    if k in m:
        lines.append(f'| {k} | {m[k][0]} | {m[k][1]} |')
lines += ['', '## warp stall reasons (warps stalled per issue-active cycle)', '', '| reason | ratio |', '|---|---|']
st = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), float(m[h][0] or 0))
      for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for n, v in sorted(st, key=lambda kv: -kv[1])[:10]:
    lines.append(f'| {n} | {v:.3f} |')
# per-barrier-segment breakdown from the SASS page
src_csv = rep[:-len('.raw.csv')] + '.source.csv'
src = (open(src_csv).read() if os.path.exists(src_csv) else '') if from_csv else \
    subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hrow = next((i for i, r in enumerate(rows[:5]) if 'Source' in r), 1)
h2 = rows[hrow] if rows else []; col = {h: i for i, h in enumerate(h2)}
segs = collections.defaultdict(collections.Counter); seg = 0
for r in rows[hrow + 1:]:
    if len(r) < len(h2): continue
    s_ = r[col['Source']].strip()
    if not s_: continue
    op = (s_.split()[1] if s_.startswith('@') else s_.split()[0]).split('.')[0]
    segs[seg]['samples'] += int(r[col['# Samples']] or 0)
    e = int(r[col['Instructions Executed']] or 0)
    segs[seg]['inst'] += e; segs[seg]['op_' + op] += e
    if op == 'BAR': seg += 1
tot = sum(s['samples'] for s in segs.values()) or 1; toti = sum(s['inst'] for s in segs.values()) or 1
lines += ['', '## time by barrier-delimited code segment (SASS sampling; segments follow the stage order of the kernel)', '',
          '| segment | samples % | warp-inst % | FFMA % of inst | LDS % | tensor/UTC % |', '|---|---|---|---|---|---|']
for k in sorted(segs):
    s = segs[k]
    if s['inst'] == 0: continue
    utc = sum(v for o, v in s.items() if o.startswith('op_UTC'))
    lines.append(f"| {k} | {100*s['samples']/tot:.1f} | {100*s['inst']/toti:.1f} | {100*s['op_FFMA']/s['inst']:.0f} | {100*s['op_LDS']/s['inst']:.0f} | {100*utc/s['inst']:.2f} |")
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:40]))

"""per barrier-delimited segment: samples, instructions, stall-reason shares, smem wavefronts. usage: seg_stalls.py src.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; col = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
segs = collections.defaultdict(collections.Counter); seg = 0
for r in rows[2:]:
    if len(r) < len(h): continue
    s = r[col['Source']].strip()
    op = (s.split()[1] if s.startswith('@') else s.split()[0]).split('.')[0]
    c = segs[seg]
    c['samples'] += int(r[col['# Samples']] or 0)
    c['inst'] += int(r[col['Instructions Executed']] or 0)
    c['wf'] += int(r[col['L1 Wavefronts Shared']] or 0)
    c['n_sass'] += 1
    for st in stalls: c[st] += int(r[col[st]] or 0)
    c['op_' + op] += int(r[col['Instructions Executed']] or 0)
    if op == 'BAR': seg += 1
tot = sum(c['samples'] for c in segs.values())
cyc = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print('seg  samp%  kcyc/tile  sass  kinst/tile(warp)  wf/tile  top stalls')
ntile = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
for k in sorted(segs):
    c = segs[k]
    if not c['samples']: continue
    ss = sorted(((c[st], st[6:]) for st in stalls), reverse=True)[:5]
    tops = ' '.join(f'{n}:{100*v/max(1,c["samples"]):.0f}' for v, n in ss)
    ops = sorted(((v, o[3:]) for o, v in c.items() if o.startswith('op_')), reverse=True)[:4]
    opss = ' '.join(f'{o}:{100*v/max(1,c["inst"]):.0f}' for v, o in ops)
    print(f'{k:3d} {100*c["samples"]/tot:6.1f} {cyc*c["samples"]/tot/1e3:8.1f} {c["n_sass"]:6d} {c["inst"]/ntile/1e3:8.2f} {c["wf"]/ntile/1e3:8.2f}  {tops} | {opss}')

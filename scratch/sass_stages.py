import csv, sys, collections, re
fn = sys.argv[1]
rows = list(csv.reader(open(fn)))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
seg = 0
segs = collections.defaultdict(lambda: collections.Counter())
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[col['Source']].strip()
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op0 = op.split('.')[0]
    samples = int(r[col['# Samples']] or 0)
    execd = int(r[col['Instructions Executed']] or 0)
    s = segs[seg]
    s['samples'] += samples
    s['inst'] += execd
    s['op_' + op0] += execd
    for h in stall_cols:
        s[h] += int(r[col[h]] or 0)
    if op0 == 'BAR':
        seg += 1
tot = sum(s['samples'] for s in segs.values())
toti = sum(s['inst'] for s in segs.values())
print(f'total samples {tot}, total warp-inst {toti}')
for k in sorted(segs):
    s = segs[k]
    if s['inst'] == 0: continue
    ops = {o[3:]: v for o, v in s.items() if o.startswith('op_')}
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
    st = sorted(((h[6:], s[h]) for h in stall_cols), key=lambda kv: -kv[1])[:5]
    print(f'seg {k:2d}: samples {100*s["samples"]/tot:5.1f}%  inst {100*s["inst"]/toti:5.1f}%  ffma {100*ops.get("FFMA",0)/max(1,s["inst"]):4.0f}%  '
          + ' '.join(f'{o}:{100*v/s["inst"]:.0f}%' for o, v in top) + ' | ' + ' '.join(f'{h}:{100*v/max(1,s["samples"]):.0f}%' for h, v in st))

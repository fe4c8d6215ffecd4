"""usage: python tools/src_hot.py <name.source.csv> [n] -- barrier-delimited segments and the hottest SASS lines of an ncu source page"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows[:5]) if 'Source' in r][0]
h = rows[hi]; col = {k: i for i, k in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) >= len(h) and r[col['Source']].strip()]
tot = sum(int(r[col['# Samples']] or 0) for r in data) or 1
stall = [k for k in h if k.startswith('stall_') and 'Not' not in k]
seg = 0; segs = collections.defaultdict(lambda: [0, collections.Counter(), collections.Counter(), 0])
for i, r in enumerate(data):
    s = int(r[col['# Samples']] or 0); src = r[col['Source']].strip()
    op = (src.split()[1] if src.startswith('@') else src.split()[0]).split('.')[0]
    d = segs[seg]; d[0] += s; d[3] += int(r[col['Instructions Executed']] or 0)
    for k in stall:
        if r[col[k]]: d[1][k[6:]] += int(float(r[col[k]]))
    d[2][op] += int(r[col['Instructions Executed']] or 0)
    if op == 'BAR': seg += 1
for k, d in segs.items():
    if d[0] / tot < 0.01: continue
    ss = sum(d[1].values()) or 1
    print(f'seg {k:2d} {100*d[0]/tot:5.1f}% inst {d[3]:>10d} | ' + ' '.join(f'{n}:{100*v/ss:.0f}' for n, v in d[1].most_common(4)) + ' | ' + ' '.join(f'{o}:{c}' for o, c in d[2].most_common(5)))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
print('--- hottest lines')
for s, i, src in sorted(((int(r[col['# Samples']] or 0), i, r[col['Source']].strip()[:80]) for i, r in enumerate(data)), reverse=True)[:n]:
    print(f'{100*s/tot:5.2f}% @{i} {src}')

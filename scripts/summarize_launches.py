"""Summarise an ncu --csv launch list (gpu__time_duration.sum) by kernel name (and grid for conv_tc)."""
import csv, collections, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
def sh(n):
    n = n.replace('void ', '').replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
    return re.sub(r'\(.*', '', re.sub(r'<.*', '', n))
agg = collections.defaultdict(lambda: [0, 0.0]); agg2 = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    n, v, g = sh(r[idx['Kernel Name']]), float(r[idx['Metric Value']]), r[idx['Grid Size']]
    agg[n][0] += 1; agg[n][1] += v
    if 'conv_tc' in n: agg2[g][0] += 1; agg2[g][1] += v
tot = sum(v for _, v in agg.values())
print(f'{len(rows)-1} launches, {tot/1e3:.1f} us total (ncu gpu__time_duration: cold-cache, serialised)')
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]): print(f'{v/1e3:10.1f} us {100*v/tot:5.1f}%  x{c:4d}  {n}')
print('conv_tc_kernel by grid:')
for g, (c, v) in sorted(agg2.items(), key=lambda x: -x[1][1]): print(f'{v/1e3:10.1f} us x{c:3d} grid {g}')

"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
agg = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if len(r) >= 8 and r[0].isdigit():
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:100], int(r[6]), int(r[7])))
        except ValueError:
            pass
ts = sum(a[3] for a in agg) or 1
ti = sum(a[4] for a in agg) or 1
print('total samples', ts, 'total warp-instructions', ti)
byfile = collections.Counter()
for a in agg:
    byfile[a[0]] += a[4]
print(dict(byfile))
agg.sort(key=lambda a: -a[3])
for a in agg[:top]:
    print('%-24s %4d inst %5.1f%% samp %5.1f%%  %s' % (a[0], a[1], 100 * a[4] / ti, 100 * a[3] / ts, a[2]))

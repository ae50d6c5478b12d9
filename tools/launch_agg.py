"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel and grid, for the LAST pass that starts with `marker`.
python tools/launch_agg.py file.csv [marker]"""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
marker = sys.argv[2] if len(sys.argv) > 2 else None
if marker:
    idx = [i for i, r in enumerate(rows) if marker in r[4]]
    rows = rows[idx[-1]:] if idx else rows
agg = collections.OrderedDict()
for r in rows:
    k = r[4][:64] + " grid=" + r[8]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[-1]) / 1e3
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
    print(f"{v[1]:9.1f} us {v[0]:3d} x {v[1] / v[0]:8.1f}  {100 * v[1] / tot:5.1f}%  {k}")
print(f"total {tot:.1f} us over {len(rows)} launches")

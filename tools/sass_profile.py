"""Condense `ncu -i rep --page source --csv` into runs of SASS lines with a similar execution count:
python tools/sass_profile.py src.csv [kernel index in the file]   (instruction share and stall-sample share per run)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s0 = starts[which]; s1 = starts[which + 1] if which + 1 < len(starts) else len(rows)
print(rows[s0][1][:100])
hdr = rows[s0 + 1]; data = [r for r in rows[s0 + 2:s1] if len(r) > 10]
iA = hdr.index("Source"); iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
tot = sum(int(r[iE]) for r in data); ts = sum(int(r[iS]) for r in data)
print("warp instructions", tot, "samples", ts, "SASS lines", len(data))
seg = []; cur = None
for n, r in enumerate(data):
    e = int(r[iE]); s = int(r[iS])
    if cur and abs(cur["e"] - e) <= 0.05 * max(cur["e"], e, 1):
        cur["n"] += 1; cur["sum"] += e; cur["s"] += s; cur["end"] = n
    else:
        cur = {"e": e, "n": 1, "sum": e, "s": s, "start": n, "end": n, "first": r[iA].strip()}
        seg.append(cur)
for g in seg:
    if g["sum"] > 0.004 * tot or g["s"] > 0.004 * ts:
        print(f"{g['start']:5d}-{g['end']:5d} n={g['n']:4d} exec/inst={g['e']:9d} inst%={100 * g['sum'] / tot:5.1f} samp%={100 * g['s'] / ts:5.1f}  {g['first'][:70]}")

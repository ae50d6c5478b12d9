"""Summarise a layer-kernel timeline written by `tools/gpu_probe.py trace` (gpurun_out/probe.json or a copy of it).
Per traced tile of cluster 0 (cycles, relative to kernel entry): x-tile load issue, MMA issue stamps, epilogue stamps."""
import json
import sys

d = json.load(open(sys.argv[1]))["trace"]["result"]
rows = d["layer"]["tiles"]
names = ["xload", "?1", "c01_start", "c01_issued", "c23_issued", "=", "mma2_issued", "g0_acc", "g1_acc", "g0_end", "e2_acc", "e2_end", "wfull", "g1_end"]
print("layer kernel: exit", d["layer"]["exit"], "cycles; wall", d.get("layer_wall_ns"), "ns; effective SM MHz", d.get("layer_sm_mhz_effective"))
print(" ".join(f"{n:>11}" for n in names))
prev = None
for r in rows[:32]:
    print(" ".join(f"{v:>11}" for v in r))
print("per-tile deltas (sorted by c01_start): tile period, c01 issue span, c23 issue span, c23_issued->mma2_issued, g1_acc->g1_end, e2_acc->e2_end")
for a, b in zip(rows[:-1], rows[1:]):
    print(f"period {b[2] - a[2]:>7}  c01 {a[3] - a[2]:>6}  c23 {a[4] - a[3]:>6}  mma2wait {a[6] - a[4]:>6}  g0 {a[9] - a[7]:>6}  g1 {a[13] - a[8]:>6}  e2 {a[11] - a[10]:>6}  g1acc-c23iss {a[8] - a[4]:>6}")

"""Summaries of ncu output for profiles/: `launches <csv>` groups a gpu__time_duration launch list by kernel;
`report <file.ncu-rep>` extracts the metrics DESIGN.md quotes from one `--set full` capture.  Runs on the CPU box."""
import csv
import re
import subprocess
import sys
from collections import defaultdict

KEEP = (r"dram__bytes_(read|write)\.sum$|dram__bytes_(read|write)\.sum\.per_second|gpu__time_duration\.sum|gpu__dram_throughput\.avg\.pct"
        r"|sm__pipe_tensor_cycles_active\.avg\.pct|sm__pipe_tensor_cycles_active_realtime\.avg\.pct|launch__registers_per_thread$"
        r"|launch__(grid_size|block_size|cluster|shared_mem_per_block_dynamic|occupancy_limit)|sm__inst_executed_pipe_(xu|alu|fma|lsu|tma|tmem)\.avg\.pct_of_peak_sustained_elapsed"
        r"|lts__t_sectors\.avg\.pct|lts__t_sector_hit_rate|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__data_pipe_tc_wavefronts_mem_shared\.sum\.pct"
        r"|l1tex__m_xbar2l1tex_read_bytes\.sum$|l1tex__m_l1tex2xbar_write_bytes\.sum$|gpc__cycles_elapsed\.avg\.per_second|sm__throughput\.avg\.pct"
        r"|smsp__warp_issue_stalled.*_per_warp_active\.pct|sm__warps_active\.avg\.pct")


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(list)
    for r in rows:
        if r is hdr or len(r) <= vi:
            continue
        try:
            agg[r[ki]].append(float(r[vi].replace(",", "")))
        except ValueError:
            pass
    unit = 1e-3  # ns -> us
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v) * unit:10.1f} us {len(v):5d} x {sum(v) / len(v) * unit:9.2f} us {100 * sum(v) / tot:6.1f}%  {k[:110]}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("kernel:", vals[hdr.index("Kernel Name")][:140])
        for h, u, v in zip(hdr, units, vals):
            if re.search(KEEP, h):
                print(f"  {h} [{u}] {v}")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])

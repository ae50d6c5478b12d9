"""Host<->device copy bandwidth of pinned buffers vs CPU affinity / NUMA placement (why e2e varies between boxes)."""
import os, sys, time, subprocess
import torch
print("affinity", sorted(os.sched_getaffinity(0)))
print(subprocess.run("lscpu | grep -i -E 'numa|socket|^CPU\\(s\\)'; nvidia-smi topo -m | head -8; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null", shell=True, capture_output=True, text=True).stdout)
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, 8)
    cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
    print("nvml cpu affinity of GPU0:", cpus[:8], "...", len(cpus))
except Exception as e:
    cpus = []
    print("nvml affinity failed", e)

def bw(tag):
    x = torch.empty(64, 1000, 256).pin_memory()
    y = torch.empty(64, 128, 1000).pin_memory()
    d = torch.empty_like(x, device="cuda"); dy = torch.empty(64, 128, 1000, device="cuda")
    for _ in range(3):
        d.copy_(x, non_blocking=True); y.copy_(dy, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        d.copy_(x, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    h2d = x.numel() * 4 * 10 / (a.elapsed_time(b) * 1e-3) / 1e9
    a.record()
    for _ in range(10):
        y.copy_(dy, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    d2h = y.numel() * 4 * 10 / (a.elapsed_time(b) * 1e-3) / 1e9
    print(f"{tag}: H2D {h2d:.1f} GB/s, D2H {d2h:.1f} GB/s", flush=True)

bw("default affinity")
allc = sorted(os.sched_getaffinity(0))
for sub in (allc[:len(allc) // 2], allc[len(allc) // 2:]):
    os.sched_setaffinity(0, sub)
    bw(f"affinity {sub[0]}-{sub[-1]}")
os.sched_setaffinity(0, allc)
loc = [c for c in cpus if c in allc]
if loc:
    os.sched_setaffinity(0, loc)
    bw(f"GPU-local cpus ({len(loc)})")

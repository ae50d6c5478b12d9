"""Timeline of CTA 0 of the fused ResBlock-round kernel: python tools/voc_pair_trace.py C k dil B T -> gpurun_out/voc_pair_trace_*.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import torch
from prodiff_b200 import _lib

lib = _lib.load()
Cc, k, dil, B, T = [int(v) for v in sys.argv[1:6]]
g = torch.Generator().manual_seed(1)
x = torch.randn(B, Cc, T, generator=g).cuda()
w1 = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).cuda()
w2 = (torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5).cuda()
b1 = torch.zeros(Cc).cuda(); b2 = torch.zeros(Cc).cuda()
y = torch.zeros(B, Cc, T, device="cuda")
n = lib.pd_voc_debug_pair_scratch(B, T, Cc, k)
scratch = torch.empty(n + 1024, dtype=torch.uint8, device="cuda")
off = (-scratch.data_ptr()) % 1024
buf = torch.zeros(_lib.N_KINDS * 1024, dtype=torch.int64, device="cuda")
def run():
    _lib.check(lib.pd_voc_debug_pair(x.data_ptr(), B, T, Cc, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), k, dil, y.data_ptr(),
                                     scratch.data_ptr() + off, n, torch.cuda.current_stream().cuda_stream), "pair")
run(); torch.cuda.synchronize()
lib.pd_debug_set_trace(buf.data_ptr())
run(); torch.cuda.synchronize()
lib.pd_debug_set_trace(None)
t = buf[13 * 1024:14 * 1024].cpu().view(64, 16)
t0 = int(t[0, 0])
names = ["x_ready", "next_gather_issued", "acc1_seen", "epi1_done", "y_published", "acc2_seen", "epi2_done", "bar_done", "mma_x_seen", "mma_conv1_issued", "mma_y_seen", "mma_conv2_issued"]
rows = []
for i in range(64):
    if int(t[i, 0]) == 0: break
    rows.append({n_: int(t[i, j]) - t0 for j, n_ in enumerate(names)})
for r in rows[:12]:
    print(" ".join(f"{k_}={v}" for k_, v in r.items()))
if len(rows) > 3:
    per = (rows[-1]["x_ready"] - rows[1]["x_ready"]) / (len(rows) - 2)
    print("tiles", len(rows), "cycles per tile (steady)", per)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"voc_pair_trace_c{Cc}_k{k}_d{dil}.json"), "w"))

"""Timeline of CTA 0 of the vocoder's convolution kernel (pd_voc_debug_conv): python tools/voc_conv_trace.py Cin Cout k dil B T"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import torch
from prodiff_b200 import _lib
lib = _lib.load()
Cin, Cout, k, dil, B, T = [int(v) for v in sys.argv[1:7]]
g = torch.Generator().manual_seed(1)
x = torch.randn(B, Cin, T, generator=g).cuda()
w = (torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5).cuda()
b = torch.zeros(Cout).cuda()
y = torch.empty(B, Cout, T, device="cuda")
n = lib.pd_voc_debug_conv_scratch(B, T, Cin, Cout, k)
scratch = torch.empty(n + 1024, dtype=torch.uint8, device="cuda")
off = (-scratch.data_ptr()) % 1024
buf = torch.zeros(_lib.N_KINDS * 1024, dtype=torch.int64, device="cuda")
def run():
    _lib.check(lib.pd_voc_debug_conv(x.data_ptr(), B, T, Cin, w.data_ptr(), b.data_ptr(), Cout, k, dil, y.data_ptr(), scratch.data_ptr() + off, n,
                                     torch.cuda.current_stream().cuda_stream), "conv")
run(); torch.cuda.synchronize()
lib.pd_debug_set_trace(buf.data_ptr())
run(); torch.cuda.synchronize()
lib.pd_debug_set_trace(None)
t = buf[13 * 1024:14 * 1024].cpu().view(64, 16)
names = ["tma_first", "tma_last", "acc_free", "first_block", "last_block", "waited_for_blocks", "epi_sees_acc", "epi_done", "a_tile_seen"]
t0 = int(t[0, 2])
rows = []
for i in range(64):
    if int(t[i, 2]) == 0: break
    rows.append({nm: (int(t[i, j]) - t0 if j != 5 else int(t[i, j])) if int(t[i, j]) != 0 else 0 for j, nm in enumerate(names)})
for r in rows[8:14]:
    print(" ".join(f"{a}={v}" for a, v in r.items()))
if len(rows) > 12:
    per = (rows[-1]["acc_free"] - rows[8]["acc_free"]) / (len(rows) - 9)
    mma = sum(r["last_block"] - r["acc_free"] for r in rows[8:]) / len(rows[8:])
    wt = sum(r["waited_for_blocks"] for r in rows[8:]) / len(rows[8:])
    epi = sum(r["epi_done"] - r["epi_sees_acc"] for r in rows[8:]) / len(rows[8:])
    print(f"tiles {len(rows)}: {per:.0f} cycles per tile; K loop {mma:.0f} of which waiting for operand blocks {wt:.0f}; epilogue {epi:.0f}")
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"voc_conv_trace_c{Cin}_{Cout}_k{k}_d{dil}.json"), "w"))

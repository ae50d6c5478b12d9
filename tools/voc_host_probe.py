"""Host enqueue time vs GPU time of one generator call: python tools/voc_host_probe.py [B] [T]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.dont_write_bytecode = True
import torch
import prodiff_b200 as pb
from prodiff_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
h = synth.VOCODER_44K
gen = pb.Generator(h); gen.load_state_dict(synth.make_vocoder_state(h, 7)); gen = gen.cuda().eval()
c = torch.randn(B, 128, T, device="cuda") * 2 - 5
f0 = torch.full((B, T), 220.0, device="cuda")
with torch.no_grad():
    for _ in range(3): gen(c, f0, seed=1)
    torch.cuda.synchronize()
    n = 30
    t0 = time.perf_counter()
    for _ in range(n): gen(c, f0, seed=1)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print(json.dumps(dict(B=B, T=T, host_enqueue_ms=(t1 - t0) / n * 1e3, total_ms=(t2 - t0) / n * 1e3)))

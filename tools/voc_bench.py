"""Time the NSF-HiFiGAN generator (44.1 kHz config) on the GPU: python tools/voc_bench.py [B] [T] [reps]."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import torch
import prodiff_b200 as pb
from prodiff_b200 import synth, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
h = synth.VOCODER_44K
gen = pb.Generator(h)
gen.load_state_dict(synth.make_vocoder_state(h, 7))
gen = gen.cuda().eval()
c = torch.randn(B, 128, T, device="cuda") * 2 - 5
f0 = torch.full((B, T), 220.0, device="cuda")
with torch.no_grad():
    gen(c, f0, seed=1)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        w = gen(c, f0, seed=1)
    b.record()
    torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(json.dumps(dict(B=B, T=T, ms=ms, frames_per_s=B * T / ms * 1e3, x_realtime=B * T * 512 / 44100 / (ms / 1e3), tflops=B * T * 648e6 / (ms * 1e-3) / 1e12)))

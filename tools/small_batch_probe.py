"""Per-kernel time of one 100-step teacher sampling call at small per-GPU batches (the 8-GPU share of BASELINE configs[2])."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import torch
import prodiff_b200 as pb
from prodiff_b200 import synth, _lib

cfg = synth.DiffNetConfig(128, 256, 20, 256, 1)
wn = pb.WaveNet(128, 256, 20, 256, 1)
wn.load_state_dict(synth.make_state(cfg, 1234, "trained"), strict=True)
gd = pb.GaussianDiffusion(128, wn, timesteps=100, time_scale=1000, schedule_type="vpsde", max_beta=0.06, spec_min=[-12.0], spec_max=[0.0]).cuda().eval()
for B in (8, 16, 32):
    cond = torch.randn(B, 1000, 256, device="cuda")
    with torch.no_grad():
        for _ in range(2):
            gd(cond, infer_step=100, infer=True, seed=1)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gd(cond, infer_step=100, infer=True, seed=1)
        b.record()
        torch.cuda.synchronize()
        total = a.elapsed_time(b)
        _lib.profile_begin()
        gd(cond, infer_step=100, infer=True, seed=1)
        kinds = _lib.profile_end()
    print(f"B={B}: {total:.2f} ms per 100-step call; per kernel (ms, launches):", {k: (round(v[0], 2), v[1]) for k, v in kinds.items()}, flush=True)

"""Soak test of the persistent layer-stack kernel and the sampler: many calls over random shapes, each run twice with the
same seed (the results must be bit-identical: any race on the inter-cluster flags would show up as a mismatch, a lost wake-up
as a watchdog trap).  Usage on the GPU box:  python tools/soak.py [iterations]"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import numpy as np
import torch
import prodiff_b200 as pb
from prodiff_b200 import synth

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 150
rng = np.random.default_rng(0)
models = {}
for cycle in (1, 3, 5):
    cfg = synth.DiffNetConfig(128, 256, 20, 256, cycle)
    wn = pb.WaveNet(128, 256, 20, 256, cycle)
    wn.load_state_dict(synth.make_state(cfg, 7 + cycle, "trained"), strict=True)
    models[cycle] = pb.GaussianDiffusion(128, wn, timesteps=4, time_scale=1000, schedule_type="vpsde", max_beta=40,
                                         spec_min=[-12.0], spec_max=[0.0]).cuda().eval()
t0 = time.time()
frames = 0
for it in range(n_iter):
    cycle = int(rng.choice([1, 3, 5]))
    B = int(rng.integers(1, 71))
    T = int(rng.integers(1, 1501))
    K = int(rng.integers(1, 5))
    ragged = bool(rng.integers(0, 2))
    lens = [int(v) for v in rng.integers(1, T + 1, size=B)] if ragged else None
    cond = torch.randn(B, T, 256, device="cuda")
    xT = torch.rand(B, 1, 128, T, device="cuda")
    with torch.no_grad():
        a = models[cycle](cond, infer_step=K, infer=True, x_T=xT, seed=it, lengths=lens).clone()
        b = models[cycle](cond, infer_step=K, infer=True, x_T=xT, seed=it, lengths=lens).clone()
    torch.cuda.synchronize()
    assert torch.isfinite(a).all(), (it, B, T, K, cycle, ragged)
    assert torch.equal(a, b), (it, B, T, K, cycle, ragged, float((a - b).abs().max()))
    frames += 2 * B * T
print(f"soak ok: {n_iter} shapes x 2 runs, {frames} frames, {time.time() - t0:.1f} s", flush=True)
# back-to-back full-size calls (the bench shape), determinism per seed
cond = torch.randn(64, 1000, 256, device="cuda")
xT = torch.rand(64, 1, 128, 1000, device="cuda")
with torch.no_grad():
    ref = models[1](cond, infer_step=4, infer=True, x_T=xT, seed=1).clone()
    for i in range(200):
        out = models[1](cond, infer_step=4, infer=True, x_T=xT, seed=1)
        if i % 50 == 49:
            assert torch.equal(out, ref), i
torch.cuda.synchronize()
print("soak ok: 200 full-size calls identical", flush=True)

"""configs[3] per-GPU share (1 x 6000 frames, C = 512, cycle 4, K = 100) under kernel-selection toggles: python tools/c512_probe.py"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import torch
import prodiff_b200 as pb
from prodiff_b200 import synth
dev = torch.device("cuda:0")
cfg = synth.DiffNetConfig(128, 256, 20, 512, 4)
wn = pb.WaveNet(cfg.in_dims, cfg.hidden_size, cfg.residual_layers, cfg.residual_channels, cfg.dilation_cycle_length)
wn.load_state_dict(synth.make_state(cfg, 1234, "trained"), strict=True)
K = int(os.environ.get("PROBE_K", "20"))
gd = pb.GaussianDiffusion(128, wn, timesteps=100, time_scale=1000, schedule_type="vpsde", max_beta=0.06, spec_min=[-12.0], spec_max=[0.0]).to(dev).eval()
B, T = int(os.environ.get("PROBE_B", "1")), int(os.environ.get("PROBE_T", "6000"))
g = torch.Generator(device=dev).manual_seed(1)
cond = torch.randn(B, T, 256, device=dev, generator=g)
xt = torch.rand(B, 1, 128, T, device=dev, generator=g)
with torch.no_grad():
    for _ in range(2): gd(cond, infer_step=K, infer=True, seed=3, x_T=xt)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): gd(cond, infer_step=K, infer=True, seed=3, x_T=xt)
    b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
F = K * (4 * 128 * 512 + 16 * 20 * 512 * 512 + 2 * 512 * 512) + 4 * 20 * 256 * 512
print(json.dumps(dict(env={k: v for k, v in os.environ.items() if k.startswith("PD_")}, B=B, T=T, K=K, ms=ms, ms_per_step=ms / K, frames_per_s=B * T / ms * 1e3, frac=B * T / ms * 1e3 * F / 1e12 / 1415.4)))

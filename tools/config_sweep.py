"""Throughput of the other BASELINE.json configs on ONE B200 (device-resident inputs, CUDA events, in-kernel Philox noise):
configs[2] (teacher 100-step DDPM, 64 x 1000), configs[3] per-GPU share (1 x 6000 frames, C = 512, K = 100, x_T supplied),
configs[4] sweep (B in {1 ... 1024} x 1000 frames, K in {2, 4}; K = 100 at B in {1, 8, 64}) and the rectified-flow samplers (N2).
Writes gpurun_out/config_sweep.json.  Usage on the GPU box:  python tools/config_sweep.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import torch  # noqa: E402


def flops_per_frame(M, H, L, C, K, evals=None):
    return (evals if evals is not None else K) * (4 * M * C + 16 * L * C * C + 2 * C * C) + 4 * L * H * C


def timed(fn, warmup=2, iters=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    import prodiff_b200 as pb
    from prodiff_b200 import synth
    dev = torch.device("cuda:0")
    rows = []

    def model(C, cycle, timesteps, max_beta, M=128):
        cfg = synth.DiffNetConfig(M, 256, 20, C, cycle)
        wn = pb.WaveNet(cfg.in_dims, cfg.hidden_size, cfg.residual_layers, cfg.residual_channels, cfg.dilation_cycle_length)
        wn.load_state_dict(synth.make_state(cfg, 1234, "trained"), strict=True)
        gd = pb.GaussianDiffusion(M, wn, timesteps=timesteps, time_scale=1000, schedule_type="vpsde", max_beta=max_beta,
                                  spec_min=[-12.0], spec_max=[0.0]).to(dev).eval()
        return cfg, wn, gd

    def run(label, gd, C, B, T, K, x_T=False):
        g = torch.Generator(device=dev).manual_seed(1)
        cond = torch.randn(B, T, 256, device=dev, generator=g)
        xt = torch.rand(B, 1, 128, T, device=dev, generator=g) if x_T else None
        with torch.no_grad():
            ms = timed(lambda: gd(cond, infer_step=K, infer=True, seed=3, x_T=xt))
        fps = B * T / (ms * 1e-3)
        tf = fps * flops_per_frame(128, 256, 20, C, K) / 1e12
        rows.append(dict(config=label, B=B, T=T, C=C, K=K, ms_per_sample=round(ms, 3), frames_per_s=round(fps),
                         algorithmic_tflops=round(tf, 1), frac_of_sustained_bf16_peak=round(tf / 1415.4, 3)))
        print(rows[-1], flush=True)

    _, _, gd4 = model(256, 1, 4, 40.0)
    _, _, gd100 = model(256, 1, 100, 0.06)
    run("configs[1] 4-step ProDiff", gd4, 256, 64, 1000, 4)
    run("configs[2] teacher 100-step DDPM (per GPU: full batch)", gd100, 256, 64, 1000, 100)
    run("configs[2] 8-GPU share (8 x 1000)", gd100, 256, 8, 1000, 100)
    for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 1024):
        for K, gd in ((2, gd4), (4, gd4), (100, gd100)):
            if K == 100 and B not in (1, 8, 64):
                continue
            run("configs[4] sweep", gd, 256, B, 1000, K)
    del gd4, gd100
    torch.cuda.empty_cache()
    _, _, gd512 = model(512, 4, 100, 0.06)
    run("configs[3] long-form, per-GPU share (1 x 6000, C=512, K_step=100, x_T supplied)", gd512, 512, 1, 6000, 100, x_T=True)
    del gd512
    torch.cuda.empty_cache()
    # rectified flow (N2): 20 euler steps / 5 rk4 steps = 20 evaluations each
    cfg, wn, _ = model(256, 1, 4, 40.0)
    for algo, n, evals in (("euler", 20, 20), ("rk4", 5, 20), ("rk5", 4, 24)):
        rf = pb.RectifiedFlow(128, wn, time_scale=1000, sampling_algorithm=algo, spec_min=[-12.0], spec_max=[0.0]).to(dev).eval()
        g = torch.Generator(device=dev).manual_seed(1)
        cond = torch.randn(64, 1000, 256, device=dev, generator=g)
        with torch.no_grad():
            ms = timed(lambda: rf(cond, infer_step=n, infer=True))
        fps = 64 * 1000 / (ms * 1e-3)
        tf = fps * flops_per_frame(128, 256, 20, 256, None, evals) / 1e12
        rows.append(dict(config=f"reflow {algo}, {n} steps ({evals} evaluations)", B=64, T=1000, C=256, K=evals, ms_per_sample=round(ms, 3),
                         frames_per_s=round(fps), algorithmic_tflops=round(tf, 1), frac_of_sustained_bf16_peak=round(tf / 1415.4, 3)))
        print(rows[-1], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "config_sweep.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()

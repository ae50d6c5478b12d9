"""GPU bring-up probe: runs each check in its own subprocess (a device trap kills only that check) with a timeout,
and writes a JSON report to gpurun_out/probe.json.  Usage on the GPU box:  python tools/gpu_probe.py [check ...]"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def check_gemm(dtype="bf16", B=2, T=300, Cin=256, taps=3, dil=2, N=512):
    import ctypes as C
    import torch
    from prodiff_b200 import _lib
    lib = _lib.load()
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float16
    g = torch.Generator().manual_seed(0)
    a = torch.randn(B, T, Cin, generator=g).to(tdt).cuda()
    w = (torch.randn(N, taps, Cin, generator=g) / (taps * Cin) ** 0.5).to(tdt).cuda()
    bias = torch.randn(N, generator=g).cuda()
    out = torch.full((B * T, N), float("nan"), device="cuda")
    rc = lib.pd_debug_gemm(0 if dtype == "bf16" else 1, a.data_ptr(), B, T, Cin, taps, dil, w.data_ptr(), N,
                           bias.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "pd_debug_gemm")
    torch.cuda.synchronize()
    # fp32 reference: conv over frames with zero padding per utterance
    af = a.float().transpose(1, 2)                      # [B, Cin, T]
    wf = w.float().permute(0, 2, 1).contiguous()        # [N, Cin, taps]
    ref = torch.nn.functional.conv1d(af, wf, bias, padding=dil * (taps // 2), dilation=dil)
    ref = ref.transpose(1, 2).reshape(B * T, N)
    err = (out - ref).abs()
    return dict(max_err=float(err.max()), mean_err=float(err.mean()), nan=int(torch.isnan(out).sum()),
                ref_absmax=float(ref.abs().max()),
                worst_row=int(err.max(dim=1).values.argmax()), worst_col=int(err.max(dim=0).values.argmax()))


def check_layers(dtype="bf16", B=2, T=300, L=4, C=256, cycle=2, M=128):
    """Full denoiser, compared buffer by buffer with the oracle's trace to localise a wrong stage."""
    import torch
    import prodiff_b200 as pb
    from prodiff_b200 import synth
    from oracle import prodiff_oracle as oracle
    cfg = synth.DiffNetConfig(M, 256, L, C, cycle)
    state = synth.make_state(cfg, 7, "kaiming")
    cond, x_T, _ = synth.make_inputs(cfg, B, T, 1, 8)
    steps = torch.arange(B, dtype=torch.float32) * 1.5 + 1.0
    tr = {}
    with torch.no_grad():
        ref = oracle.wavenet_forward(state, x_T, steps, cond.transpose(1, 2), cycle, trace=tr)
    wn = pb.WaveNet(M, 256, L, C, cycle, operand_dtype=dtype)
    wn.load_state_dict(state)
    wn = wn.cuda().eval()
    with torch.no_grad():
        out = wn(x_T.cuda(), steps.cuda(), cond.cuda().transpose(1, 2))
    torch.cuda.synchronize()
    plan = next(iter(wn.engine().plans.values()))
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float16
    N = B * T
    res = {}
    def unperm(packed):  # packed gate-paired column order -> original channel order
        idx = torch.tensor([((p & 255) >> 7) * C + (p >> 8) * 128 + (p & 127) for p in range(2 * C)])
        o = torch.empty_like(packed)
        o[..., idx] = packed
        return o
    mtr = (T + 127) // 128
    cpb = plan.buffer("cp").view(torch.float16).view(L, B, mtr, 2 * C // 8, 128, 8)   # [layer][utt][tile][chunk][row][8]
    cp = cpb.permute(0, 1, 2, 4, 3, 5).reshape(L, B, mtr * 128, 2 * C)[:, :, :T].reshape(L, N, 2 * C).float().cpu()
    sb = plan.buffer("sb").view(torch.float32).view(-1, L, 4, 2 * C).cpu()
    y = plan.buffer("ybuf").view(tdt).view(L, N, C).float().cpu()
    for l in range(L):
        p = f"residual_layers.{l}."
        c_ref = (tr["cond_proj"][l] + state[p + "dilated_conv.bias"][None, :, None]).transpose(1, 2).reshape(N, 2 * C)
        res[f"cp{l}"] = float((unperm(cp[l]) - c_ref).abs().max())
        y_ref = tr["gated"][l].transpose(1, 2).reshape(N, C)
        res[f"y{l}"] = float((y[l] - y_ref).abs().max())
    # step-bias table: slot 0 of entry b, layer 0 == W_dil . d summed over taps
    e = oracle.sinusoidal_pos_emb(steps, C); e = torch.nn.functional.linear(e, state["mlp.0.weight"], state["mlp.0.bias"])
    e = oracle.mish(e); e = torch.nn.functional.linear(e, state["mlp.2.weight"], state["mlp.2.bias"])
    d0 = torch.nn.functional.linear(e, state["residual_layers.0.diffusion_projection.weight"], state["residual_layers.0.diffusion_projection.bias"])
    w = state["residual_layers.0.dilated_conv.weight"]
    s_all = torch.einsum("ock,bc->bo", w, d0)
    res["sb_all"] = float((unperm(sb[:B, 0, 0]) - s_all).abs().max())
    res["sb_tap0"] = float((unperm(sb[:B, 0, 1]) - torch.einsum("ock,bc->bo", w[:, :, 1:], d0)).abs().max())   # slot 1 = taps 1 + 2
    xl = plan.buffer("xa" if L % 2 == 0 else "xb").view(tdt).view(N, C).float().cpu()
    res["x_last"] = float((xl - tr["x"][-1].transpose(1, 2).reshape(N, C)).abs().max())
    res["out"] = float((out.cpu() - ref).abs().max())
    res["out_mean"] = float((out.cpu() - ref).abs().mean())
    res["ref_std"] = float(ref.std())
    res["nan"] = int(torch.isnan(out).sum())
    return res


def check_shift():
    """Row-shifted views of one SW128 tile as UMMA A operand: which base_offset convention is right?"""
    import torch
    from prodiff_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    R = 160
    a = torch.randn(R, 64, generator=g).to(torch.bfloat16).cuda()
    w = torch.randn(64, 64, generator=g).to(torch.bfloat16).cuda()
    res = {}
    for mode in (0, 1):
        errs = []
        for shift in range(0, 33):
            out = torch.full((128, 64), float("nan"), device="cuda")
            _lib.check(lib.pd_debug_shift(0, a.data_ptr(), R, w.data_ptr(), shift, mode, out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "pd_debug_shift")
            torch.cuda.synchronize()
            ref = a[shift:shift + 128].float() @ w.float().t()
            errs.append(round(float((out - ref).abs().max()), 4))
        res[f"mode{mode}"] = errs
    return res


def check_trace():
    """Pipeline timeline of CTA 0 for the conv+gate (kind 2) and res (kind 3) kernels at the bench shape."""
    import torch
    sys.argv = ["bench"]
    import bench
    from prodiff_b200 import _lib
    lib = _lib.load()
    gd, cfg, _ = bench.build_model(torch.device("cuda"))
    from prodiff_b200 import synth
    cond, _, _ = synth.make_inputs(cfg, int(os.environ.get("PROBE_B", "64")), 1000, 1, 5)
    cond = cond.cuda()
    with torch.no_grad():
        gd(cond, infer_step=4, infer=True, seed=1)
        buf = torch.zeros(_lib.N_KINDS * 1024, dtype=torch.int64, device="cuda")
        lib.pd_debug_set_trace(buf.data_ptr())
        gd(cond, infer_step=4, infer=True, seed=1)
        torch.cuda.synchronize()
        lib.pd_debug_set_trace(None)
    out = {}
    for kind, name in ((8, "layer"), (2, "gate"), (3, "res"), (0, "cond"), (6, "out")):
        t = buf[kind * 1024:(kind + 1) * 1024].cpu().view(64, 16)
        t0 = int(t[0, 15])
        rows = []
        for i in range(32 if kind == 8 else 8):
            if int(t[i, 6]) == 0:
                break
            rows.append([int(v) - t0 if int(v) else -1 for v in t[i, :12]] + [int(t[i, 12]), int(t[i, 13]) - t0 if int(t[i, 13]) else 0])
        if kind == 8:
            rows.sort(key=lambda r: r[2])
        out[name] = dict(exit=int(t[0, 14]) - t0, tiles=rows)
        if kind == 8:
            out["layer_mma2"] = [[int(v) - t0 if int(v) else -1 for v in t[32 + i, :13]] for i in range(4)]
            w = buf[9 * 1024:10 * 1024].cpu().view(8, 128)
            out["weight_loads"] = [dict(issue=[int(v) - t0 for v in w[i, 0:32]], landed=[int(v) - t0 for v in w[i, 32:64]],
                                        waited=[int(v) for v in w[i, 64:96]]) for i in range(8)]
            ns = int(t[63, 1]) - int(t[63, 0])
            out["layer_wall_ns"] = ns
            out["layer_sm_mhz_effective"] = round((int(t[0, 14]) - t0) / max(ns, 1) * 1e3, 1)

    t = buf[10 * 1024:11 * 1024].cpu().view(64, 16)
    t0 = int(t[0, 15])
    out["tail"] = dict(exit=int(t[0, 14]) - t0, tiles=[[int(v) - t0 if int(v) else -1 for v in t[i, :11]] for i in range(8)])
    return out


def check_pytest(expr):
    r = subprocess.run([sys.executable, "-m", "pytest", "tests", "-x", "-q", "-m", "gpu", "-k", expr], capture_output=True,
                       text=True, cwd=ROOT, timeout=900)
    return dict(rc=r.returncode, tail=(r.stdout + r.stderr)[-1200:])


CHECKS = {
    "shift": check_shift,
    "trace": check_trace,
    "layers_bf16": lambda: check_layers("bf16"),
    "layers_fp16": lambda: check_layers("fp16"),
    "layers_c512": lambda: check_layers("bf16", B=1, T=200, L=3, C=512, cycle=3),
    "layers_m64": lambda: check_layers("bf16", B=1, T=100, L=2, C=256, cycle=1, M=64),
    "gemm_k1": lambda: check_gemm(B=1, T=128, Cin=64, taps=1, dil=1, N=256),
    "gemm_k4": lambda: check_gemm(B=1, T=128, Cin=256, taps=1, dil=1, N=256),
    "gemm_multi_tile": lambda: check_gemm(B=3, T=300, Cin=256, taps=1, dil=1, N=512),
    "gemm_taps_d1": lambda: check_gemm(B=2, T=300, Cin=256, taps=3, dil=1, N=512),
    "gemm_taps_d8": lambda: check_gemm(B=2, T=300, Cin=256, taps=3, dil=8, N=512),
    "gemm_fp16": lambda: check_gemm(dtype="fp16", B=2, T=300, Cin=256, taps=3, dil=2, N=512),
    "gemm_short_T": lambda: check_gemm(B=3, T=7, Cin=256, taps=3, dil=4, N=256),
    "gemm_big": lambda: check_gemm(B=16, T=1000, Cin=512, taps=3, dil=1, N=1024),
}


def main():
    names = sys.argv[1:] or list(CHECKS)
    if len(names) == 2 and names[0] == "--one":
        print("RESULT " + json.dumps(CHECKS[names[1]]()))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    report = {}
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], capture_output=True, text=True,
                               timeout=180, cwd=ROOT)
            res = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            report[n] = dict(rc=r.returncode, result=json.loads(res[-1][7:]) if res else None,
                             tail=(r.stdout + r.stderr)[-1500:] if (r.returncode or not res) else "")
        except subprocess.TimeoutExpired as e:
            report[n] = dict(rc="timeout", result=None, tail=str(e)[-500:])
        report[n]["secs"] = round(time.time() - t0, 1)
        print(n, json.dumps(report[n])[:600], flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
            json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()

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


CHECKS = {
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

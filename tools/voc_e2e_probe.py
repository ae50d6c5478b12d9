"""GPU probe: the whole generator against the CPU oracle on the golden cases (errors per case) + timing at the bench shape."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import numpy as np
import torch
import prodiff_b200 as pb
from prodiff_b200 import synth
from oracle import vocoder_oracle as vo
from oracle.golden_cases import VOCODER_CASES

out = {}
for name, case in VOCODER_CASES.items():
    h = case["h"]
    state = synth.make_vocoder_state(h, case["wseed"])
    c, f0, ri, nz = synth.make_vocoder_inputs(h, case["B"], case["T"], case["iseed"])
    gen = pb.Generator(h)
    gen.load_state_dict(state)
    gen = gen.cuda().eval()
    with torch.no_grad():
        wav = gen(c.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda()).cpu()
        tr = {}
        ref = vo.generator_forward(state, h, c, f0, ri, nz, trace=tr)
    g = np.load(os.path.join(ROOT, "tests", "golden", f"vocoder_{name}.npz"))
    out[name] = dict(max_abs=float((wav - ref).abs().max()), mean_abs=float((wav - ref).abs().mean()), vs_golden=float(np.abs(wav.numpy() - g["wav"]).max()),
                     ref_std=float(ref.std()), nan=int(torch.isnan(wav).sum()))
    print(name, out[name], flush=True)
if "--time" in sys.argv:
    h = synth.VOCODER_44K
    gen = pb.Generator(h)
    gen.load_state_dict(synth.make_vocoder_state(h, 7))
    gen = gen.cuda().eval()
    for (B, T) in ((1, 1000), (8, 1000), (64, 1000)):
        c = torch.randn(B, 128, T, device="cuda") * 2 - 5
        f0 = torch.full((B, T), 220.0, device="cuda")
        with torch.no_grad():
            for _ in range(2):
                gen(c, f0, seed=1)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(3):
                w = gen(c, f0, seed=1)
            b.record()
            torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        out[f"time_B{B}_T{T}"] = dict(ms=ms, frames_per_s=B * T / ms * 1e3, x_realtime=B * T * 512 / 44100 / (ms / 1e3), finite=bool(torch.isfinite(w).all()))
        print(f"B={B} T={T}: {ms:.2f} ms", out[f"time_B{B}_T{T}"], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "voc_e2e_probe.json"), "w"), indent=1)

"""Soak test of the NSF-HiFiGAN generator kernels: many calls over random batch sizes / lengths / ragged patterns on the 44.1 kHz
config, each run twice with the same inputs (bit-identical or a race shows), every result finite and |wav| <= 1, ragged rows checked
against the single-utterance call now and then.  Usage on the GPU box:  python tools/voc_soak.py [iterations]"""
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

n_iter = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(0)
h = synth.VOCODER_44K
gen = pb.Generator(h)
gen.load_state_dict(synth.make_vocoder_state(h, 7))
gen = gen.cuda().eval()
upp = gen.upp
t0 = time.time()
frames = 0
for it in range(n_iter):
    B = int(rng.integers(1, 13))
    T = int(rng.integers(1, 400))
    c, f0, ri, nz = synth.make_vocoder_inputs(h, B, T, 1000 + it)
    ragged = bool(rng.integers(0, 2))
    lens = [int(rng.integers(1, T + 1)) for _ in range(B)] if ragged else None
    args = dict(rand_ini=ri.cuda(), noise=nz.cuda(), lengths=lens)
    with torch.no_grad():
        a = gen(c.cuda(), f0.cuda(), **args)
        b = gen(c.cuda(), f0.cuda(), **args)
        torch.cuda.synchronize()
        assert torch.equal(a, b), f"iteration {it}: two runs differ (B={B}, T={T}, ragged={ragged})"
        assert torch.isfinite(a).all() and float(a.abs().max()) <= 1.0, f"iteration {it}: bad values"
        if ragged and it % 5 == 0:
            r = int(rng.integers(0, B))
            n = lens[r]
            one = gen(c[r:r + 1, :, :n].contiguous().cuda(), f0[r:r + 1, :n].contiguous().cuda(), rand_ini=ri.cuda(),
                      noise=nz[r:r + 1, :n * upp].contiguous().cuda())
            assert torch.equal(a[r, 0, :n * upp], one[0, 0]), f"iteration {it}: ragged row {r} differs from the single run"
    frames += B * T
print(f"voc soak ok: {n_iter} iterations, {frames} frames, {time.time() - t0:.1f} s")

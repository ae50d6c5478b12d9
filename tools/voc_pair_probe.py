"""GPU probe of the fused ResBlock-round kernel (pd_voc_debug_pair) against torch.conv1d on fp16-rounded operands.
Usage on the GPU box: python tools/voc_pair_probe.py  -> gpurun_out/voc_pair_probe.json"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import torch
import torch.nn.functional as F
from prodiff_b200 import _lib

lib = _lib.load()
CASES = [  # C, k, dil, B, T
    (16, 3, 1, 2, 900), (16, 7, 5, 1, 2048), (16, 11, 3, 2, 1111), (16, 11, 5, 1, 3), (32, 3, 1, 2, 777), (32, 7, 3, 1, 1500),
    (32, 11, 5, 2, 640), (64, 3, 1, 2, 400), (64, 7, 3, 1, 1000), (64, 11, 5, 1, 1000), (64, 3, 1, 1, 5), (64, 11, 1, 3, 246),
    (16, 7, 1, 3, 250), (32, 11, 1, 40, 4000), (64, 7, 1, 30, 3000), (16, 3, 1, 64, 8000),
]


def reference(x, w1, b1, w2, b2, k, dil):
    h = (k - 1) // 2
    xa = F.leaky_relu(x, 0.1).half().double()
    y1 = F.conv1d(xa, w1.half().double(), b1.double(), dilation=dil, padding=h * dil)
    y1 = F.leaky_relu(y1, 0.1).half().double()
    y2 = F.conv1d(y1, w2.half().double(), b2.double(), padding=h)
    raw = torch.where(xa >= 0, xa, xa * 10.0)
    return (y2 + raw).float()


out = {}
for _ in [0]:
    for (Cc, k, dil, B, T) in CASES:
        g = torch.Generator().manual_seed(Cc * 1000 + k + dil + T)
        x = torch.randn(B, Cc, T, generator=g)
        w1 = torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5
        w2 = torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5
        b1 = torch.randn(Cc, generator=g) * 0.1
        b2 = torch.randn(Cc, generator=g) * 0.1
        ref = reference(x, w1, b1, w2, b2, k, dil)
        d = [t.cuda().contiguous() for t in (x, w1, b1, w2, b2)]
        y = torch.zeros(B, Cc, T, device="cuda")
        n = lib.pd_voc_debug_pair_scratch(B, T, Cc, k)
        scratch = torch.empty(n + 1024, dtype=torch.uint8, device="cuda")
        off = (-scratch.data_ptr()) % 1024
        name = f"c{Cc}_k{k}_d{dil}_B{B}_T{T}"
        rc = lib.pd_voc_debug_pair(d[0].data_ptr(), B, T, Cc, d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(), k, dil,
                                   y.data_ptr(), scratch.data_ptr() + off, n, torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            out[name] = dict(rc=rc, err=lib.pd_last_error().decode())
            print(name, out[name], flush=True)
            continue
        torch.cuda.synchronize()
        yc = y.cpu()
        err = (yc - ref).abs()
        out[name] = dict(rc=0, max_abs=float(err.max()), mean_abs=float(err.mean()), ref_std=float(ref.std()), nan=int(torch.isnan(yc).sum()),
                         worst_t=int(err.amax(dim=(0, 1)).argmax()))
        print(name, out[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "voc_pair_probe.json"), "w"), indent=1)

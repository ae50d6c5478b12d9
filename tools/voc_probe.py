"""GPU probe of the vocoder's convolution kernel (pd_voc_debug_conv) against torch.conv1d on fp16-rounded operands.
Usage on the GPU box: python tools/voc_probe.py   -> gpurun_out/voc_probe.json"""
import ctypes as C
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
import torch
from prodiff_b200 import _lib

lib = _lib.load()
CASES = [  # Cin, Cout, k, dil, B, T
    (128, 512, 7, 1, 2, 300), (256, 256, 3, 1, 2, 333), (256, 256, 11, 5, 1, 700), (128, 128, 7, 3, 2, 500),
    (64, 64, 3, 1, 2, 400), (64, 64, 11, 5, 1, 1000), (32, 32, 3, 1, 2, 777), (32, 32, 7, 3, 1, 1500), (32, 32, 11, 5, 2, 640),
    (16, 16, 3, 1, 2, 900), (16, 16, 7, 5, 1, 2048), (16, 16, 11, 3, 2, 1111), (16, 1, 7, 1, 2, 1000), (512, 2048, 3, 1, 1, 130),
    (64, 64, 3, 1, 1, 5), (16, 16, 11, 5, 1, 3),
]
out = {}
for (Cin, Cout, k, dil, B, T) in CASES:
    g = torch.Generator().manual_seed(Cin * 1000 + Cout + k + dil)
    x = torch.randn(B, Cin, T, generator=g).half().float().cuda()
    w = (torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5).half().float().cuda()
    b = torch.randn(Cout, generator=g).cuda()
    y = torch.empty(B, Cout, T, device="cuda")
    n = lib.pd_voc_debug_conv_scratch(B, T, Cin, Cout, k)
    scratch = torch.empty(n + 1024, dtype=torch.uint8, device="cuda")
    off = (-scratch.data_ptr()) % 1024
    rc = lib.pd_voc_debug_conv(x.data_ptr(), B, T, Cin, w.data_ptr(), b.data_ptr(), Cout, k, dil, y.data_ptr(),
                               scratch.data_ptr() + off, n, torch.cuda.current_stream().cuda_stream)
    name = f"cin{Cin}_cout{Cout}_k{k}_d{dil}_B{B}_T{T}"
    if rc != 0:
        out[name] = dict(rc=rc, err=lib.pd_last_error().decode())
        print(name, out[name], flush=True)
        continue
    torch.cuda.synchronize()
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        ref = torch.nn.functional.conv1d(x.double(), w.double(), b.double(), dilation=dil, padding=(k - 1) // 2 * dil).float()
    err = float((y - ref).abs().max())
    out[name] = dict(rc=0, max_abs=err, ref_std=float(ref.std()), nan=int(torch.isnan(y).sum()))
    print(name, out[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "voc_probe.json"), "w"), indent=1)

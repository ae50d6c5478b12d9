"""In-kernel timeline (clock64 at pipeline events, CTA 0) of the text encoder's FFN-conv GEMM at 64 x 200 tokens.
Usage (GPU box): python tools/encoder_trace.py -> gpurun_out/encoder_trace.json"""
import json
import sys

import torch

sys.path.insert(0, ".")
from prodiff_b200 import _lib, synth               # noqa: E402
import prodiff_b200 as pb                          # noqa: E402

cfg = synth.EncoderConfig(64, 256, 4, 9, 2)
enc = pb.FastspeechEncoder(64, 256, 4, 9, num_heads=2)
enc.load_state_dict(synth.make_encoder_state(cfg, 1234), strict=False)
enc = enc.cuda().eval()
tok, extra = synth.make_encoder_inputs(cfg, [200] * 64, 200, 4321)
lib = _lib.load()
with torch.no_grad():
    enc(tok.cuda(), extra.cuda())
    buf = torch.zeros(16 * 1024, dtype=torch.int64, device="cuda")
    lib.pd_debug_set_trace(buf.data_ptr())
    enc(tok.cuda(), extra.cuda())
    torch.cuda.synchronize()
    lib.pd_debug_set_trace(None)
out = {}
for kind, name in ((12, "ffn_1"), (11, "ffn_2")):
    t = buf[kind * 1024:(kind + 1) * 1024].cpu().view(64, 16)
    t0 = int(t[0, 15])
    out[name] = [[int(v) - t0 if int(v) else -1 for v in t[i, :8]] for i in range(6) if int(t[i, 6])]
print(json.dumps(out))
json.dump(out, open("gpurun_out/encoder_trace.json", "w"), indent=1)

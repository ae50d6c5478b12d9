"""GPU probe for the text encoder: final error against the oracle per fixture case and, for one-layer cases, the error of
the intermediates the C-ABI call leaves in its workspace (x before the last LayerNorm, q|k|v, GELU output).
Usage (on the GPU box): python tools/encoder_probe.py [bf16|fp16]"""
import json
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import encoder_oracle as eo            # noqa: E402  (tools/ is test infrastructure)
from oracle.golden_cases import ENCODER_CASES      # noqa: E402
from prodiff_b200 import _lib, synth               # noqa: E402
import prodiff_b200 as pb                          # noqa: E402


def main():
    dt = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    res = {}
    for name, case in ENCODER_CASES.items():
        cfg = case["cfg"]
        state = synth.make_encoder_state(cfg, case["wseed"])
        tok, extra = synth.make_encoder_inputs(cfg, case["lengths"], case["Tph"], case["iseed"])
        ex = extra if case["extra"] else None
        tr = {}
        with torch.no_grad():
            ref = eo.encoder_forward(state, tok, ex, cfg.num_heads, tr)
        enc = pb.FastspeechEncoder(cfg.vocab_size, cfg.hidden_size, cfg.num_layers, cfg.kernel_size, num_heads=cfg.num_heads,
                                   operand_dtype=dt)
        enc.load_state_dict(state, strict=False)
        enc = enc.cuda().eval()
        eng = enc.engine()
        B, T = tok.shape
        H = cfg.hidden_size
        N = B * T
        nbytes = eng.lib.pd_encoder_workspace_bytes(eng.handle, B, T)
        ws = torch.zeros(nbytes + 1024, dtype=torch.uint8, device="cuda")
        base = (ws.data_ptr() + 1023) // 1024 * 1024
        off0 = base - ws.data_ptr()
        out = torch.empty(B, T, H, device="cuda")
        tk = tok.cuda()
        exd = ex.cuda() if ex is not None else None
        _lib.check(eng.lib.pd_encoder_forward(eng.handle, tk.data_ptr(), exd.data_ptr() if exd is not None else None, out.data_ptr(),
                                              B, T, base, nbytes, torch.cuda.current_stream().cuda_stream), "pd_encoder_forward")
        torch.cuda.synchronize()
        r = dict(final_max=float((out.cpu() - ref).abs().max()), final_mean=float((out.cpu() - ref).abs().mean()),
                 finite=bool(torch.isfinite(out).all()), ref_absmax=float(ref.abs().max()))
        al = lambda v: (v + 1023) // 1024 * 1024
        o_x = 0
        o_xn = al(o_x + N * H * 4)
        o_qkv = al(o_xn + N * H * 2)
        o_att = al(o_qkv + N * 3 * H * 2)
        o_h1 = al(o_att + N * H * 2)
        view = lambda o, n, d: ws[off0 + o: off0 + o + n * d.itemsize].view(d)
        x_gpu = view(o_x, N * H, torch.float32).view(B, T, H).cpu()
        r["x_pre_last_ln_max"] = float((x_gpu - tr[f"layer{cfg.num_layers - 1}"]).abs().max())
        if cfg.num_layers == 1:
            p = "layers.0.op."
            x0 = tr["embed"]
            ln1 = F.layer_norm(x0, (H,), state[p + "layer_norm1.weight"], state[p + "layer_norm1.bias"], 1e-5)
            qkv_ref = F.linear(ln1, state[p + "self_attn.in_proj_weight"])
            qkv_gpu = view(o_qkv, N * 3 * H, torch.float16).view(B, T, 3 * H).float().cpu()
            r["qkv_max"] = float((qkv_gpu - qkv_ref).abs().max())
            x1 = tr["attn0"]
            d16 = torch.float16 if dt == "fp16" else torch.bfloat16
            h1_gpu = view(o_h1, N * 4 * H, d16).view(B, T, 4 * H).float().cpu()
            ln2 = F.layer_norm(x1, (H,), state[p + "layer_norm2.weight"], state[p + "layer_norm2.bias"], 1e-5)
            h1_ref = F.gelu(F.conv1d(ln2.transpose(1, 2), state[p + "ffn.ffn_1.weight"], state[p + "ffn.ffn_1.bias"],
                                     padding=cfg.kernel_size // 2).transpose(1, 2) * cfg.kernel_size ** -0.5)
            r["h1_max"] = float((h1_gpu - h1_ref).abs().max())
        res[name] = r
        print(name, json.dumps(r), flush=True)
    json.dump(res, open(f"gpurun_out/encoder_probe_{dt}.json", "w"), indent=1)


if __name__ == "__main__":
    main()

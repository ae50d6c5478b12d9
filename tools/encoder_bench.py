"""Timing of the text encoder: pd_encoder_forward through prodiff_b200.FastspeechEncoder against the same
algorithm in eager PyTorch on the same GPU (the oracle restatement moved to CUDA: cuBLAS / cuDNN kernels, what the
reference module executes) and on the host cores.  Workload: the text side of BASELINE configs[1] -- 64 utterances x 200
tokens (1000 frames at ~5 frames per phoneme), the configured 4-layer / 2-head / kernel-9 encoder (base_config.yaml:106-116).
Usage (GPU box): python tools/encoder_bench.py [B] [Tph] -> gpurun_out/encoder_bench.json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from oracle import encoder_oracle as eo            # noqa: E402  (checker / eager baseline only)
from prodiff_b200 import synth                     # noqa: E402
import prodiff_b200 as pb                          # noqa: E402


def cuda_ms(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    Tph = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    cfg = synth.EncoderConfig(64, 256, 4, 9, 2)
    state = synth.make_encoder_state(cfg, 1234)
    rng_len = [Tph - (7 * b) % (Tph // 4) for b in range(B)]
    tok, extra = synth.make_encoder_inputs(cfg, rng_len, Tph, 4321)
    n_tok = B * Tph
    H, L, k = cfg.hidden_size, cfg.num_layers, cfg.kernel_size
    flops = n_tok * L * (2 * H * 3 * H + 2 * H * H + 2 * k * H * 4 * H + 2 * 4 * H * H + 4 * Tph * H)
    res = dict(workload=f"{B} utt x {Tph} tokens, {L} layers, {cfg.num_heads} heads, kernel {k}, H {H}", tokens=n_tok,
               algorithmic_gflop=flops / 1e9)
    tk, ex = tok.cuda(), extra.cuda()
    for dt in ("bf16", "fp16"):
        enc = pb.FastspeechEncoder(cfg.vocab_size, H, L, k, num_heads=cfg.num_heads, operand_dtype=dt)
        enc.load_state_dict(state, strict=False)
        enc = enc.cuda().eval()
        with torch.no_grad():
            ms = cuda_ms(lambda: enc(tk, ex))
            out = enc(tk, ex)
        res[f"ours_{dt}_ms"] = ms
        res[f"ours_{dt}_tokens_per_s"] = n_tok / ms * 1e3
        res[f"ours_{dt}_tflops"] = flops / ms / 1e9
    if os.environ.get("ENC_BENCH_OURS_ONLY") == "1":
        print(json.dumps(res))
        return
    st_gpu = {k_: v.cuda() for k_, v in state.items()}
    with torch.no_grad():
        ref_gpu = _eager(st_gpu, tk, ex, cfg.num_heads)
        res["eager_torch_gpu_fp32_ms"] = cuda_ms(lambda: _eager(st_gpu, tk, ex, cfg.num_heads))
        res["max_abs_vs_eager_gpu"] = float((out - ref_gpu).abs().max())
        torch.set_num_threads(os.cpu_count() or 1)
        nb = min(B, 8)
        t0 = time.perf_counter()
        eo.encoder_forward(state, tok[:nb], extra[:nb], cfg.num_heads)
        res["cpu_oracle_ms_scaled"] = (time.perf_counter() - t0) * 1e3 * B / nb
        res["cpu_cores"] = os.cpu_count()
    res["speedup_vs_eager_gpu"] = res["eager_torch_gpu_fp32_ms"] / res["ours_bf16_ms"]
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/encoder_bench.json", "w"), indent=1)


def _eager(state, tok, extra, heads):
    # the oracle's table helper builds on the CPU; move it once per call like the reference's `.to(self._float_tensor)`
    orig = eo.sinusoid_table
    eo.sinusoid_table = lambda n, d, p=0: orig(n, d, p).to(tok.device)
    try:
        return eo.encoder_forward(state, tok, extra, heads)
    finally:
        eo.sinusoid_table = orig


if __name__ == "__main__":
    main()

"""Text -> mel on the CUDA path, end to end (SURVEY 8f N1 + the hot path): token-level embeddings (torch), FastspeechEncoder
(pd_encoder_forward), frame-level condition assembly (pd_cond_assemble), 4-step ProDiff sampler (pd_sample), for the text
side of BASELINE configs[1]: 64 utterances x 200 tokens -> 64 x 1000 frames x 128 mel.  Reports ms per batch for the whole
chain and for its parts (CUDA events on the launching stream), so the encoder's share of text -> mel is a measured number.
Usage (GPU box): python tools/text_to_mel_bench.py -> gpurun_out/text_to_mel.json"""
import json
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, ".")
import prodiff_b200 as pb                          # noqa: E402
from prodiff_b200 import synth                     # noqa: E402


def cuda_ms(fn, n=10, warm=3):
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
    B, Tph, T, K = 64, 200, 1000, 4
    ecfg = synth.EncoderConfig(64, 256, 4, 9, 2)
    enc = pb.FastspeechEncoder(ecfg.vocab_size, 256, 4, 9, num_heads=2)
    enc.load_state_dict(synth.make_encoder_state(ecfg, 1234), strict=False)
    enc = enc.cuda().eval()
    cfg = synth.DiffNetConfig(128, 256, 20, 256, 1)
    wn = pb.WaveNet(128, 256, 20, 256, 1)
    wn.load_state_dict(synth.make_state(cfg, 1234, "trained"))
    gd = pb.GaussianDiffusion(128, wn, timesteps=4, time_scale=1000, schedule_type="vpsde", max_beta=40, spec_min=[-12.0],
                              spec_max=[0.0]).cuda().eval()
    rng = np.random.default_rng(7)
    tok, _ = synth.make_encoder_inputs(ecfg, [Tph] * B, Tph, 4321)
    mel2ph = torch.from_numpy(np.sort(rng.integers(1, Tph + 1, size=(B, T)), axis=1))
    f0 = torch.from_numpy(rng.uniform(80, 800, size=(B, T)).astype(np.float32))
    dur_embed = torch.nn.Linear(1, 256).cuda()
    pitch = torch.nn.Linear(1, 256).cuda()
    spk = torch.randn(B, 1, 256).cuda()
    teacher = types.SimpleNamespace(with_dur_embed=True, dur_embed=dur_embed, with_lang_embed=False, encoder=enc, pitch_embed=pitch,
                                    with_spk_embed=True, with_gender_embed=False, add_spk_embed=lambda i, mix: mix,
                                    with_voicing_embed=False, with_breath_embed=False)
    tk, m2p, f0d = tok.cuda(), mel2ph.cuda(), f0.cuda()
    with torch.no_grad():
        cond = lambda: pb.forward_condition(teacher, tk, m2p, f0d, spk_mix_embed=spk, validate=False)   # no host sync in the chain
        ac = cond()
        res = dict(workload=f"{B} utt x {Tph} tokens -> {B} x {T} frames x 128 mel, K={K}, DiffNet 20x256, encoder 4x256",
                   encoder_ms=cuda_ms(lambda: enc(tk, None)),
                   forward_condition_ms=cuda_ms(cond),
                   sampler_from_assembled_condition_ms=cuda_ms(lambda: gd(ac, infer_step=K, infer=True, seed=1)),
                   text_to_mel_ms=cuda_ms(lambda: gd(cond(), infer_step=K, infer=True, seed=1)))
    res["frames_per_s_text_to_mel"] = B * T / res["text_to_mel_ms"] * 1e3
    res["encoder_share"] = res["encoder_ms"] / res["text_to_mel_ms"]
    print(json.dumps(res))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/text_to_mel.json", "w"), indent=1)


if __name__ == "__main__":
    main()

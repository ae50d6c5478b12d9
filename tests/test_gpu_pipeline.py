"""Tokens -> waveform for a batch of segments (prodiff_b200.pipeline.Synthesizer) on the CUDA path: a padded, ragged batch
must give every segment the waveform of that segment run alone -- the reference runs one segment per call
(handler/infer/handler.py:133-158, 373-378), so this is what makes batching a drop-in.

Bound: the sampler picks its kernels by problem size (one frame tile per call -> one launch per layer with the residual
added in the epilogue; several -> the persistent stack with the residual added on the tensor core), so a batched and a
single run agree to bf16 rounding, not bit for bit (the vocoder and the encoder ARE bit-identical row by row:
test_gpu_vocoder.py::test_ragged_batch_rows_equal_single_utterance_runs, test_gpu_parity.py).  1e-2 max-abs on the
waveform; an unmasked padded batch misses it by 15 x at the segment ends (0.15, measured)."""
import types

import numpy as np
import pytest
import torch

from oracle.golden_cases import VOCODER_CASES

pytestmark = pytest.mark.gpu


def _teacher(K):
    import prodiff_b200 as pb
    from prodiff_b200 import synth
    ecfg = synth.EncoderConfig(40, 256, 2, 9, 2)
    enc = pb.FastspeechEncoder(ecfg.vocab_size, 256, 2, 9, num_heads=2)
    enc.load_state_dict(synth.make_encoder_state(ecfg, 11), strict=False)
    cfg = synth.DiffNetConfig(128, 256, 3, 256, 1)
    wn = pb.WaveNet(128, 256, 3, 256, 1)
    wn.load_state_dict(synth.make_state(cfg, 12, "trained"))
    gd = pb.GaussianDiffusion(128, wn, timesteps=4, time_scale=1000, schedule_type="vpsde", max_beta=40, spec_min=[-12.0], spec_max=[0.0])
    torch.manual_seed(13)
    pitch = torch.nn.Linear(1, 256)
    t = types.SimpleNamespace(with_dur_embed=False, with_lang_embed=False, encoder=enc.cuda().eval(), pitch_embed=pitch.cuda(),
                              with_spk_embed=False, with_gender_embed=False, with_voicing_embed=False, with_breath_embed=False,
                              diffusion=gd.cuda().eval())
    return t, ecfg, cfg


def test_ragged_batch_of_segments_equals_one_segment_per_call():
    import prodiff_b200 as pb
    from prodiff_b200 import synth
    from prodiff_b200.pipeline import Synthesizer
    K = 3
    teacher, ecfg, cfg = _teacher(K)
    h = VOCODER_CASES["ladder"]["h"]
    gen = pb.Generator(h)
    gen.load_state_dict(synth.make_vocoder_state(h, 14), strict=True)
    gen = gen.cuda().eval()
    syn = Synthesizer(teacher, gen, infer_step=K)
    lens, tph = [64, 41, 9], [20, 13, 4]
    B, T, Tph = len(lens), max(lens), max(tph)
    rng = np.random.default_rng(15)
    tok = torch.zeros(B, Tph, dtype=torch.long)
    m2p = torch.zeros(B, T, dtype=torch.long)
    f0 = torch.zeros(B, T)
    for b in range(B):
        tok[b, :tph[b]] = torch.from_numpy(rng.integers(1, ecfg.vocab_size, size=tph[b]))
        m2p[b, :lens[b]] = torch.from_numpy(np.sort(rng.integers(1, tph[b] + 1, size=lens[b])))
        f0[b, :lens[b]] = torch.from_numpy(rng.uniform(100, 600, size=lens[b]).astype(np.float32))
    _, x_T, noise = synth.make_inputs(cfg, B, T, K, 16)
    upp = gen.upp
    _, _, ri, nz = synth.make_vocoder_inputs(h, B, T, 17)
    wavs = syn(tok.cuda(), m2p.cuda(), f0.cuda(), x_T=x_T.cuda(), noise=noise.cuda(), rand_ini=ri.cuda(), voc_noise=nz.cuda())
    assert [w.shape for w in wavs] == [(n * upp,) for n in lens]
    for b in range(B):
        n = lens[b]
        # (token rows keep their padding: the reference's own encoder treats padding tokens as LayerNorm(0) = bias rows inside
        # the FFN convolution's window, tts_modules.py:262-288, so a token row is only comparable at equal padded length)
        one = syn(tok[b:b + 1].cuda(), m2p[b:b + 1, :n].cuda(), f0[b:b + 1, :n].cuda(), x_T=x_T[b:b + 1, :, :, :n].contiguous().cuda(),
                  noise=noise[:, b:b + 1, :, :, :n].contiguous().cuda(), rand_ini=ri.cuda(), voc_noise=nz[b:b + 1, :n * upp].contiguous().cuda())
        assert torch.isfinite(one[0]).all()
        err = float((wavs[b].cpu() - one[0].cpu()).abs().max())
        assert err <= 1e-2, (b, err)

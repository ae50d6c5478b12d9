"""Parity of the NSF-HiFiGAN generator on the CUDA path against the reference's golden vectors and the oracle.

Tolerance: the reference computes in fp32; the CUDA path uses fp16 operands and fp16 activation tensors with fp32
accumulation, through ~40 convolutions in sequence.  The bound is on the waveform itself (tanh output, |wav| <= 1):
max-abs <= 2e-2 and mean-abs <= 2e-3 with identical rand_ini / SineGen noise (measured: 3-5e-3 / 5e-4).
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import vocoder_oracle as vo
from oracle.golden_cases import VOCODER_CASES
from tests import helpers

pytestmark = pytest.mark.gpu

WAV_MAX_TOL = 2e-2
WAV_MEAN_TOL = 2e-3


def _build(h, state):
    import prodiff_b200 as pb
    gen = pb.Generator(h)
    gen.load_state_dict(state, strict=True)
    return gen.cuda().eval()


def _case(name):
    from prodiff_b200 import synth
    case = VOCODER_CASES[name]
    h = case["h"]
    state = synth.make_vocoder_state(h, case["wseed"])
    c, f0, ri, nz = synth.make_vocoder_inputs(h, case["B"], case["T"], case["iseed"])
    return h, state, c, f0, ri, nz


@pytest.mark.parametrize("name", sorted(VOCODER_CASES))
def test_generator_vs_reference_golden(name):
    from prodiff_b200 import synth
    h, state, c, f0, ri, nz = _case(name)
    g = helpers.load_golden("vocoder", name)
    assert str(g["digest"]) == synth.digest(state, c, f0, ri, nz), "regenerated inputs differ from the fixture's"
    gen = _build(h, state)
    with torch.no_grad():
        wav = gen(c.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda())
    torch.cuda.synchronize()
    upp = int(np.prod(h["upsample_rates"]))
    assert wav.shape == (c.shape[0], 1, c.shape[2] * upp) and wav.dtype == torch.float32
    err = np.abs(wav.cpu().numpy() - g["wav"])
    assert np.isfinite(wav.cpu().numpy()).all()
    assert err.max() <= WAV_MAX_TOL, f"waveform max-abs {err.max():.4g}"
    assert err.mean() <= WAV_MEAN_TOL, f"waveform mean-abs {err.mean():.4g}"


CONV_SHAPES = [  # Cin, Cout, k, dil, B, T: every tile width, both A-operand paths, every kernel size / dilation of the generator
    (128, 512, 7, 1, 2, 300), (256, 256, 11, 5, 1, 700), (128, 128, 7, 3, 2, 500), (64, 64, 3, 1, 2, 400), (64, 64, 11, 5, 1, 1000),
    (32, 32, 3, 1, 2, 777), (32, 32, 11, 5, 2, 640), (16, 16, 7, 5, 1, 2048), (16, 16, 11, 3, 2, 1111), (16, 1, 7, 1, 2, 1000),
    (64, 64, 3, 1, 1, 5), (16, 16, 11, 5, 1, 3), (32, 32, 7, 3, 3, 129),
]


@pytest.mark.parametrize("shape", CONV_SHAPES, ids=lambda s: "cin%d_cout%d_k%d_d%d_B%d_T%d" % s)
def test_conv_kernel_vs_fp32_conv(shape):
    """One convolution through the generator's tcgen05 kernel (pd_voc_debug_conv) against torch's CPU conv1d on the same
    fp16-rounded operands: only the fp16 rounding of the output (2^-11 relative) separates them."""
    from prodiff_b200 import _lib
    lib = _lib.load()
    Cin, Cout, k, dil, B, T = shape
    g = torch.Generator().manual_seed(Cin * 1000 + Cout + k + dil)
    x = torch.randn(B, Cin, T, generator=g).half().float()
    w = (torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5).half().float()
    b = torch.randn(Cout, generator=g)
    ref = torch.nn.functional.conv1d(x.double(), w.double(), b.double(), dilation=dil, padding=(k - 1) // 2 * dil).float()
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    y = torch.full((B, Cout, T), float("nan"), device="cuda")
    n = lib.pd_voc_debug_conv_scratch(B, T, Cin, Cout, k)
    scratch = torch.empty(n + 1024, dtype=torch.uint8, device="cuda")
    off = (-scratch.data_ptr()) % 1024
    _lib.check(lib.pd_voc_debug_conv(xd.data_ptr(), B, T, Cin, wd.data_ptr(), bd.data_ptr(), Cout, k, dil, y.data_ptr(),
                                     scratch.data_ptr() + off, n, torch.cuda.current_stream().cuda_stream), "pd_voc_debug_conv")
    torch.cuda.synchronize()
    err = (y.cpu() - ref).abs()
    assert float(err.max()) <= 2e-3 * max(1.0, float(ref.abs().max())), float(err.max())


PAIR_SHAPES = [  # C, k, dil, B, T: the three fused widths, every kernel size / dilation of the generator, single- and many-tile runs
    (16, 3, 1, 2, 900), (16, 7, 5, 1, 2048), (16, 11, 3, 2, 1111), (16, 11, 5, 1, 3), (32, 3, 1, 2, 777), (32, 7, 3, 1, 1500),
    (32, 11, 5, 2, 640), (64, 3, 1, 2, 400), (64, 7, 3, 1, 1000), (64, 11, 5, 1, 1000), (64, 3, 1, 1, 5), (64, 11, 1, 3, 246),
    (16, 7, 1, 3, 250), (32, 11, 1, 40, 4000), (64, 7, 1, 30, 3000), (16, 3, 1, 64, 8000),
]


@pytest.mark.parametrize("shape", PAIR_SHAPES, ids=lambda s: "c%d_k%d_d%d_B%d_T%d" % s)
def test_fused_resblock_round_vs_fp64_convs(shape):
    """One ResBlock1 round (models.py:62-69: x + c2(leaky_relu(c1(leaky_relu(x))))) through the resident-tile kernel
    (pd_voc_debug_pair) against float64 convolutions over the same fp16-rounded operands and fp16-rounded intermediate: only
    the fp16 rounding of the result separates them.  Covers tiles at both utterance edges (zero padding of BOTH convolutions),
    utterances shorter than the halo, and persistent CTAs walking many tiles."""
    import torch.nn.functional as F
    from prodiff_b200 import _lib
    lib = _lib.load()
    Cc, k, dil, B, T = shape
    h = (k - 1) // 2
    g = torch.Generator().manual_seed(Cc * 1000 + k + dil + T)
    x = torch.randn(B, Cc, T, generator=g)
    w1 = torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5
    w2 = torch.randn(Cc, Cc, k, generator=g) / (Cc * k) ** 0.5
    b1 = torch.randn(Cc, generator=g) * 0.1
    b2 = torch.randn(Cc, generator=g) * 0.1
    xa = F.leaky_relu(x, 0.1).half().double()
    y1 = F.leaky_relu(F.conv1d(xa, w1.half().double(), b1.double(), dilation=dil, padding=h * dil), 0.1).half().double()
    ref = (F.conv1d(y1, w2.half().double(), b2.double(), padding=h) + torch.where(xa >= 0, xa, xa * 10.0)).float()
    d = [t.cuda().contiguous() for t in (x, w1, b1, w2, b2)]
    y = torch.full((B, Cc, T), float("nan"), device="cuda")
    n = lib.pd_voc_debug_pair_scratch(B, T, Cc, k)
    scratch = torch.empty(n + 1024, dtype=torch.uint8, device="cuda")
    off = (-scratch.data_ptr()) % 1024
    _lib.check(lib.pd_voc_debug_pair(d[0].data_ptr(), B, T, Cc, d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr(), k, dil,
                                     y.data_ptr(), scratch.data_ptr() + off, n, torch.cuda.current_stream().cuda_stream), "pd_voc_debug_pair")
    torch.cuda.synchronize()
    err = (y.cpu() - ref).abs()
    assert float(err.max()) <= 2e-3 * max(1.0, float(ref.abs().max())), float(err.max())


def test_fused_rounds_match_the_two_launch_form():
    """The generator with fused ResBlock rounds (default) against the same generator running every round as two convolution
    launches: same operands, same fp16 intermediates, fp32 accumulation in the same tap order."""
    from prodiff_b200 import _lib, synth
    lib = _lib.load()
    h, state, c, f0, ri, nz = _case("full44k_short")
    gen = _build(h, state)
    args = dict(rand_ini=ri.cuda(), noise=nz.cuda())
    prev = lib.pd_set_voc_fused(1)
    try:
        with torch.no_grad():
            fused = gen(c.cuda(), f0.cuda(), **args).cpu()
            lib.pd_set_voc_fused(0)
            plain = gen(c.cuda(), f0.cuda(), **args).cpu()
    finally:
        lib.pd_set_voc_fused(1 if prev != 0 else 0)
    assert float((fused - plain).abs().max()) <= 1e-3, float((fused - plain).abs().max())


def test_batch_rows_equal_single_utterance_runs():
    """Batching the vocoder (what N3 asks for; the reference's SineGen takes one utterance per call, models.py:163): every
    row of a batch must come out exactly as that utterance alone -- tiles never span utterances and the guard rows carry
    each utterance's own zero padding."""
    from prodiff_b200 import synth
    h = VOCODER_CASES["ladder"]["h"]
    state = synth.make_vocoder_state(h, 301)
    c, f0, ri, nz = synth.make_vocoder_inputs(h, 3, 131, 302)
    gen = _build(h, state)
    with torch.no_grad():
        full = gen(c.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda()).cpu()
        for b in range(3):
            one = gen(c[b:b + 1].cuda(), f0[b:b + 1].cuda(), rand_ini=ri.cuda(), noise=nz[b:b + 1].cuda()).cpu()
            assert torch.equal(one[0], full[b]), f"row {b} depends on its batch neighbours"
        ref = vo.generator_forward(state, h, c, f0, ri, nz)
    assert float((full - ref).abs().max()) <= WAV_MAX_TOL


@pytest.mark.parametrize("fused", [1, 0], ids=["fused_rounds", "two_launch_rounds"])
def test_ragged_batch_rows_equal_single_utterance_runs(fused):
    """A padded batch with per-utterance lengths (what batching the reference's one-segment-per-call inferer needs,
    handler/infer/handler.py:373-378): the generator's receptive field (~13 frames per side) would let whatever lies past
    an utterance's end leak into its last frames; with ``lengths`` every activation past the end is held at zero, so row b
    equals the run on utterance b alone -- exactly -- and the tail of the row is zeros.  Garbage is planted in the padding."""
    from prodiff_b200 import synth
    h = VOCODER_CASES["full44k_short"]["h"]
    state = synth.make_vocoder_state(h, 401)
    B, T = 3, 48
    lens = [48, 29, 7]
    c, f0, ri, nz = synth.make_vocoder_inputs(h, B, T, 402)
    for b, n in enumerate(lens):           # the padding holds junk, as a sampler's ragged output may
        c[b, :, n:] = 37.0
        f0[b, n:] = 555.0
    from prodiff_b200 import _lib
    gen = _build(h, state)
    upp = gen.upp
    prev = _lib.load().pd_set_voc_fused(fused)
    try:
        with torch.no_grad():
            full = gen(c.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda(), lengths=lens).cpu()
            for b, n in enumerate(lens):
                one = gen(c[b:b + 1, :, :n].contiguous().cuda(), f0[b:b + 1, :n].contiguous().cuda(), rand_ini=ri.cuda(),
                          noise=nz[b:b + 1, :n * upp].contiguous().cuda()).cpu()
                assert torch.equal(full[b, 0, :n * upp], one[0, 0]), (b, float((full[b, 0, :n * upp] - one[0, 0]).abs().max()))
                assert float(full[b, 0, n * upp:].abs().max()) == 0.0 if n < T else True
            # and without lengths the junk does reach the last frames (the test would be vacuous otherwise)
            plain = gen(c.cuda(), f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda()).cpu()
            assert not torch.equal(plain[1, 0, :lens[1] * upp], full[1, 0, :lens[1] * upp])
    finally:
        _lib.load().pd_set_voc_fused(1 if prev != 0 else 0)


def test_philox_noise_path_and_determinism():
    """Without injected noise the SineGen noise comes from in-kernel Philox: same seed -> same waveform, another seed ->
    another one, and both stay close to the noise-free... to the injected-noise result in the voiced-signal sense (the noise
    term is 0.003 / 0.033 of the source amplitude, models.py:184-187)."""
    from prodiff_b200 import synth
    h = VOCODER_CASES["rate8"]["h"]
    state = synth.make_vocoder_state(h, 311)
    c, f0, ri, nz = synth.make_vocoder_inputs(h, 2, 90, 312)
    gen = _build(h, state)
    with torch.no_grad():
        a = gen(c.cuda(), f0.cuda(), rand_ini=ri.cuda(), seed=5).cpu()
        b = gen(c.cuda(), f0.cuda(), rand_ini=ri.cuda(), seed=5).cpu()
        d = gen(c.cuda(), f0.cuda(), rand_ini=ri.cuda(), seed=6).cpu()
    assert torch.equal(a, b) and not torch.equal(a, d)
    assert torch.isfinite(a).all() and float(a.abs().max()) <= 1.0


def test_handoff_chain_sampler_to_waveform():
    """Sampler -> (fused 2.30259 scale, [B,M,T]) -> generator, against the chain of oracles: GaussianDiffusion with
    vocoder_scale hands its state tensor straight to Generator.forward (nsf_hifigan.py:49-58)."""
    import prodiff_b200 as pb
    from prodiff_b200 import handoff, synth
    from oracle import prodiff_oracle as oracle
    cfg = synth.DiffNetConfig(128, 256, 3, 256, 1)
    dstate = synth.make_state(cfg, 321, "trained")
    B, T, K = 2, 40, 3
    cond, x_T, noise = synth.make_inputs(cfg, B, T, K, 322)
    h = VOCODER_CASES["ladder"]["h"]
    vstate = synth.make_vocoder_state(h, 323)
    _, f0, ri, nz = synth.make_vocoder_inputs(h, B, T, 324)
    wn = pb.WaveNet(128, 256, 3, 256, 1)
    wn.load_state_dict(dstate)
    gd = pb.GaussianDiffusion(128, wn, timesteps=4, time_scale=1000, schedule_type="vpsde", max_beta=40, spec_min=[-12.0], spec_max=[0.0]).cuda().eval()
    gen = _build(h, vstate)
    with torch.no_grad():
        c = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), noise=noise.cuda(), vocoder_scale=handoff.LOG10_TO_LN)
        wav = handoff.spec2wav_torch(gen, c=c, f0=f0.cuda(), rand_ini=ri.cuda(), noise=nz.cuda()).cpu()
        cut = handoff.spec2wav_torch(gen, c=c, f0=f0.cuda(), lengths=[40, 17], rand_ini=ri.cuda(), noise=nz.cuda())
    tables = oracle.diffusion_tables(4, "vpsde", 40)
    mel_ref, _ = oracle.sample(dstate, tables, cond, x_T, noise, K, 4, 1)
    with torch.no_grad():
        ref = vo.spec2wav(vstate, h, mel_ref, f0, ri, nz).view(B, -1)
    upp = int(np.prod(h["upsample_rates"]))
    # lengths: row 0 is full length (nothing changes); row 1 ends at frame 17 and is recomputed as that utterance alone
    # (test_ragged_batch_rows_equal_single_utterance_runs)
    assert wav.shape == (B, T * upp) and cut[1].shape == (17 * upp,) and torch.equal(cut[0].cpu(), wav[0])
    # the sampler's own bf16 error (mel L1 <= 1e-2) rides into the generator input, so the chain bound is looser
    assert float((wav - ref).abs().mean()) <= 2e-2


def test_loud_failures():
    import prodiff_b200 as pb
    from prodiff_b200 import synth, _lib
    h = dict(VOCODER_CASES["rate8"]["h"])
    gen = _build(h, synth.make_vocoder_state(h, 331))
    with pytest.raises(RuntimeError, match="inference-only"):
        gen(torch.zeros(1, 64, 10, device="cuda"), torch.zeros(1, 10, device="cuda"))       # autograd is refused
    with torch.no_grad():
        with pytest.raises(ValueError):
            gen(torch.zeros(1, 128, 10, device="cuda"), torch.zeros(1, 10, device="cuda"))   # 64 mel bins expected
        with pytest.raises(ValueError):
            gen(torch.zeros(1, 64, 10, device="cuda"), torch.zeros(1, 11, device="cuda"))
        with pytest.raises(RuntimeError):
            pb.Generator(h)(torch.zeros(1, 64, 10), torch.zeros(1, 10))                      # CPU tensors: no fallback
        bad = pb.Generator(dict(h, upsample_kernel_sizes=[16, 4]))
        bad.h = dict(h, upsample_kernel_sizes=[15, 4])       # a kernel size the library does not implement (!= 2 * rate)
        with pytest.raises(RuntimeError, match="unsupported generator config"):
            bad.cuda()(torch.zeros(1, 64, 10, device="cuda"), torch.zeros(1, 10, device="cuda"))
    lib = _lib.load()
    assert lib.pd_vocoder_workspace_bytes(None, 1, 1) == 0
    assert lib.pd_vocoder_forward(None, None, None, None, None, 0, None, None, 1, 1, None, 0, None) != 0

"""Parity of the CUDA path (through the module API -> C-ABI) against the reference's golden vectors and the oracle.

Tolerances (BASELINE.json north_star): per-step denoiser output max-abs <= 1e-2 with bf16 operands / fp32
accumulation, final mel mean-abs (L1) <= 1e-2 against the fp32 reference, on identical x_T and noise.  The bound is
stated for a trained-like output scale (std ~0.2, synth out_init="trained"); with the worst-case Kaiming output
projection (std ~0.9) bf16 rounding alone gives ~2.5e-2, so those cases scale the bound by the output
scale, and the fp16-operand mode is held to the strict 1e-2 everywhere.
"""
import numpy as np
import pytest
import torch

from oracle import prodiff_oracle as oracle
from oracle.golden_cases import SAMPLE_CASES, DENOISE_CASES
from tests import helpers

pytestmark = pytest.mark.gpu

STEP_TOL = 1e-2
MEL_L1_TOL = 1e-2


def _build(cfg, state, case=None, operand_dtype="bf16"):
    import prodiff_b200 as pb
    wn = pb.WaveNet(cfg.in_dims, cfg.hidden_size, cfg.residual_layers, cfg.residual_channels, cfg.dilation_cycle_length,
                    operand_dtype=operand_dtype)
    wn.load_state_dict(state, strict=True)
    if case is None:
        return wn.cuda().eval()
    gd = pb.GaussianDiffusion(cfg.in_dims, wn, timesteps=case["timesteps"], time_scale=1000,
                              schedule_type=case["schedule_type"], max_beta=case["max_beta"], spec_min=[-12.0], spec_max=[0.0])
    return gd.cuda().eval()


def _scale(out_init):
    return 1.0 if out_init == "trained" else 4.5


@pytest.mark.parametrize("operand_dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("name", sorted(SAMPLE_CASES))
def test_sample_vs_reference_golden(name, operand_dtype):
    case, cfg, state, cond, x_T, noise = helpers.regen_sample_case(name)
    g = helpers.load_golden("sample", name)
    gd = _build(cfg, state, case, operand_dtype)
    K = case["infer_step"]
    trace = torch.empty(K, case["B"], 1, cfg.in_dims, case["T"], device="cuda")
    with torch.no_grad():
        mel = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), noise=noise.cuda(), x0_trace=trace)
    torch.cuda.synchronize()
    assert mel.shape == (case["B"], case["T"], cfg.in_dims) and mel.dtype == torch.float32
    tol = STEP_TOL * (_scale(case["out_init"]) if operand_dtype == "bf16" else 1.0)
    # step 0 sees identical inputs on both sides: the pure per-step denoiser error
    e0 = np.abs(trace[0].cpu().numpy() - g["x0"][0]).max()
    assert e0 <= tol, f"step-0 denoiser max-abs {e0:.4g} > {tol}"
    e_all = np.abs(trace.cpu().numpy() - g["x0"]).max()
    assert e_all <= 2 * tol, f"chained per-step max-abs {e_all:.4g}"
    l1 = np.abs(mel.cpu().numpy() - g["mel"]).mean()
    assert l1 <= MEL_L1_TOL, f"final mel L1 {l1:.4g}"
    assert np.isfinite(mel.cpu().numpy()).all()


@pytest.mark.parametrize("name", sorted(DENOISE_CASES))
def test_denoise_vs_reference_golden(name):
    case, cfg, state, cond, x_T, steps = helpers.regen_denoise_case(name)
    g = helpers.load_golden("denoise", name)
    for dt in ("bf16", "fp16"):
        wn = _build(cfg, state, None, dt)
        with torch.no_grad():
            out = wn(x_T.cuda(), steps.cuda(), cond.cuda().transpose(1, 2))
        err = np.abs(out.cpu().numpy() - g["out"]).max()
        tol = STEP_TOL * (_scale(case["out_init"]) if dt == "bf16" else 1.0)
        assert out.shape == g["out"].shape
        assert err <= tol, f"{dt}: max-abs {err:.4g} > {tol}"


def test_oracle_on_fresh_seed_config2_shape_slice():
    """Same seeded inputs through the oracle and the CUDA path at a size the oracle finishes in seconds
    (the BASELINE config-2 model, B=2 x T=1000)."""
    from prodiff_b200 import synth
    cfg = synth.DiffNetConfig(128, 256, 20, 256, 1)
    state = synth.make_state(cfg, 1234, "trained")
    cond, x_T, noise = synth.make_inputs(cfg, 2, 1000, 4, 999)
    case = dict(timesteps=4, schedule_type="vpsde", max_beta=40, infer_step=4)
    mel_ref, x0s = helpers.oracle_sample(case, cfg, state, cond, x_T, noise)
    gd = _build(cfg, state, case, "bf16")
    trace = torch.empty(4, 2, 1, 128, 1000, device="cuda")
    with torch.no_grad():
        mel = gd(cond.cuda(), infer_step=4, infer=True, x_T=x_T.cuda(), noise=noise.cuda(), x0_trace=trace)
    assert float((trace[0].cpu() - x0s[0]).abs().max()) <= STEP_TOL
    assert float((mel.cpu() - mel_ref).abs().mean()) <= MEL_L1_TOL


def test_t0_noise_masked_and_philox_path_runs():
    case, cfg, state, cond, x_T, noise = helpers.regen_sample_case("cosine_sched")
    gd = _build(cfg, state, case)
    K = case["infer_step"]
    with torch.no_grad():
        a = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), noise=noise.cuda()).clone()
        n2 = noise.clone()
        n2[-1] += 100.0
        b = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), noise=n2.cuda()).clone()
        assert torch.equal(a, b)                       # noise at t == 0 must not reach the output (prodiff.py:119-121)
        c = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), seed=7).clone()
        d = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), seed=7).clone()
        e = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), seed=8).clone()
    assert torch.equal(c, d) and not torch.equal(c, e)  # in-kernel Philox: deterministic per seed
    assert torch.isfinite(c).all()


def test_philox_noise_is_standard_normal():
    """Distribution check of the in-kernel RNG: with coef1 = coef2 = 0 the update is x = sigma * noise."""
    import prodiff_b200 as pb
    from prodiff_b200 import synth
    cfg = synth.DiffNetConfig(128, 256, 1, 256, 1)
    wn = _build(cfg, synth.make_state(cfg, 5, "trained"))
    eng = wn.engine()
    B, T = 4, 512
    x = torch.zeros(B, 1, 128, T, device="cuda")
    cond = torch.zeros(B, T, 256, device="cuda")
    eng.sample(cond, x, [1], [0.0, 0.0], [0.0, 0.0], [0.0, 0.0], noise=None, seed=123)
    torch.cuda.synchronize()
    v = x.flatten().double().cpu()
    assert abs(float(v.mean())) < 0.01 and abs(float(v.std()) - 1.0) < 0.01
    assert abs(float((v ** 4).mean()) - 3.0) < 0.1
    assert float(v.abs().max()) < 7.0
    # no correlation between neighbouring frames / bins
    xm = x[:, 0].double().cpu()
    assert abs(float((xm[..., 1:] * xm[..., :-1]).mean())) < 0.01
    assert abs(float((xm[:, 1:] * xm[:, :-1]).mean())) < 0.01


def test_dropin_contract():
    """Signatures, state_dict keys and buffers match the reference's; errors are loud."""
    import prodiff_b200 as pb
    from prodiff_b200 import synth
    cfg = synth.DiffNetConfig(128, 256, 2, 256, 1)
    state = synth.make_state(cfg, 3, "trained")
    wn = pb.WaveNet(128, 256, 2, 256, 1)
    assert set(wn.state_dict()) == set(state)
    assert float(wn.output_projection.weight.abs().max()) == 0.0     # reference zero-init, wavenet.py:98
    gd = pb.GaussianDiffusion(128, wn, timesteps=4, time_scale=1000, schedule_type="vpsde", max_beta=40,
                              spec_min=[-12.0], spec_max=[0.0])
    assert [k for k, _ in gd.named_buffers()][:4] == ["timesteps", "timescale", "betas", "alphas_cumprod"]
    with pytest.raises(RuntimeError):
        gd(torch.zeros(1, 8, 256), infer=True)                       # CPU tensor: no fallback
    gd = gd.cuda()
    out = gd(torch.zeros(1, 8, 256, device="cuda"), infer=True)      # default init: bias-only output, still runs
    assert out.shape == (1, 8, 128)
    with pytest.raises(RuntimeError):
        wn(torch.zeros(1, 1, 128, 8, device="cuda"), torch.zeros(1, device="cuda"),
           torch.zeros(1, 256, 8, device="cuda"))                    # grad enabled + trainable params -> refused
    # infer_step is clipped to [1, num_timesteps] like prodiff.py:146
    assert gd(torch.zeros(1, 8, 256, device="cuda"), infer_step=99, infer=True).shape == (1, 8, 128)
    # weights changed in place -> packed cache must refresh
    wn.load_state_dict(state)
    with torch.no_grad():
        o2 = gd(torch.zeros(1, 8, 256, device="cuda"), infer=True, x_T=torch.zeros(1, 1, 128, 8), seed=1)
    assert float((o2 - out).abs().max()) > 1e-3


def test_fused_layer_kernel_equals_unfused_pair():
    """The fused 2-CTA ResidualBlock kernel against the conv+gate / res_proj kernel pair on the same inputs.
    Same operands, same fp32 accumulation per output, so the two paths must agree to bf16 rounding of one layer
    (they differ only in accumulation order across K blocks: the fused kernel staggers its weight walk per cluster)."""
    from prodiff_b200 import _lib, synth
    lib = _lib.load()
    cfg = synth.DiffNetConfig(128, 256, 6, 256, 3)
    state = synth.make_state(cfg, 17, "trained")
    cond, x_T, _ = synth.make_inputs(cfg, 5, 333, 1, 18)     # 5 x 3 = 15 tiles: odd count -> dummy peer tile path
    steps = torch.tensor([0.0, 1.0, 2.0, 3.0, 4.0])
    outs = []
    for fused in (1, 0):
        prev = lib.pd_set_fused(fused)
        try:
            wn = _build(cfg, state)
            with torch.no_grad():
                outs.append(wn(x_T.cuda(), steps.cuda(), cond.cuda().transpose(1, 2)).cpu())
        finally:
            lib.pd_set_fused(1 if prev != 0 else 0)
    with torch.no_grad():
        ref = oracle.wavenet_forward(state, x_T, steps, cond.transpose(1, 2), cfg.dilation_cycle_length)
    assert float((outs[0] - outs[1]).abs().max()) <= 4e-3
    assert float((outs[0] - ref).abs().max()) <= STEP_TOL and float((outs[1] - ref).abs().max()) <= STEP_TOL


def test_plan_cache_is_lru_and_reuses_workspace_blocks():
    """Engine.plan: least-recently-used eviction (a plan in constant use survives), bucketed workspace blocks handed from an
    evicted plan to the next shape, and results unaffected by the reuse."""
    from prodiff_b200 import synth
    from prodiff_b200.engine import workspace_bucket
    cfg = synth.DiffNetConfig(128, 256, 2, 256, 1)
    state = synth.make_state(cfg, 51, "trained")
    wn = _build(cfg, state)
    eng = wn.engine()
    eng.MAX_PLANS = 3
    hot = eng.plan(2, 256, 4)
    ptrs = set()
    for T in (250, 251, 252, 253, 254, 255):
        assert eng.plan(2, 256, 4) is hot                      # touched between every miss: never the LRU entry
        p = eng.plan(2, T, 4)
        ptrs.add(p.block.data_ptr())
        assert p.block.numel() == workspace_bucket(p.workspace.numel() + 1024)
    assert (2, 256, 4) in eng.plans and len(eng.plans) == 3
    assert len(ptrs) <= 3                                       # evicted blocks were reused, not reallocated
    cond, x_T, _ = synth.make_inputs(cfg, 2, 255, 1, 52)
    steps = torch.tensor([1.0, 3.0])
    with torch.no_grad():
        out = wn(x_T.cuda(), steps.cuda(), cond.cuda().transpose(1, 2)).cpu()
        ref = oracle.wavenet_forward(state, x_T, steps, cond.transpose(1, 2), cfg.dilation_cycle_length)
    assert float((out - ref).abs().max()) <= STEP_TOL


@pytest.mark.parametrize("mode", ["sumtail", "tail"])
def test_fused_sampler_tail_equals_separate_kernels(mode):
    """(mode "sumtail": the skip sum runs in the same kernel as well, pd_sumtail.cuh.)
    pd_sample with the fused tail kernel (skip_projection -> output_projection -> q-posterior -> next step's
    input_projection, intermediates in shared memory) against the same loop running those as separate GEMM launches.
    Identical operands, identical K order and identical 16-bit rounding points, so the mel must agree exactly; both
    injected noise and the in-kernel Philox stream are covered, and the per-step x0 trace as well."""
    from prodiff_b200 import _lib, synth
    lib = _lib.load()
    cfg = synth.DiffNetConfig(128, 256, 4, 256, 2)
    state = synth.make_state(cfg, 23, "trained")
    case = dict(timesteps=4, schedule_type="vpsde", max_beta=40)
    B, T, K = 3, 333, 4                                     # ragged last tile per utterance
    cond, x_T, noise = synth.make_inputs(cfg, B, T, K, 24)
    res = []
    for tail, sumtail in ((1, mode == "sumtail"), (0, 0)):
        prev = lib.pd_set_tail(tail)
        prev_st = lib.pd_set_sumtail(1 if sumtail else 0)
        try:
            gd = _build(cfg, state, case)
            with torch.no_grad():
                tr = torch.empty(K, B, 1, 128, T, device="cuda")
                a = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), noise=noise.cuda(), x0_trace=tr).cpu()
                b = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), seed=77).cpu()
                res.append((a, b, tr.cpu()))
        finally:
            lib.pd_set_tail(1 if prev != 0 else 0)
            lib.pd_set_sumtail(1 if prev_st != 0 else 0)
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert torch.equal(res[0][2], res[1][2])
    mel_ref, x0s = helpers.oracle_sample(dict(case, infer_step=K), cfg, state, cond, x_T, noise)
    assert float((res[0][0] - mel_ref).abs().mean()) <= MEL_L1_TOL
    assert float((res[0][2][K - 1] - x0s[K - 1]).abs().max()) <= STEP_TOL * 4   # last step: errors of K - 1 updates ride along


@pytest.mark.parametrize("path", ["tail", "sumtail", "separate", "unfused512", "reflow"])
def test_vocoder_handoff_scale_is_fused_and_bit_identical(path):
    """SURVEY 8f N3 (first sub-step): with ``vocoder_scale`` the last posterior update (denorm_spec kernel for rectified
    flow) stores fl(2.30259 * x), so the call returns the contiguous [B,M,T] generator input of nsf_hifigan.py:50-52 --
    bit-identical to ``2.30259 * mel.transpose(2, 1)`` of the unscaled call on the same x_T / noise / seed, in every
    kernel variant that owns the final write; a later call without the scale is unaffected."""
    import prodiff_b200 as pb
    from prodiff_b200 import _lib, handoff, synth
    lib = _lib.load()
    C = 512 if path == "unfused512" else 256
    cfg = synth.DiffNetConfig(128, 256, 3, C, 2)
    state = synth.make_state(cfg, 31, "trained")
    B, T, K = 2, 300, 3
    cond, x_T, noise = synth.make_inputs(cfg, B, T, K, 32)
    prev = lib.pd_set_tail(0 if path == "separate" else 1)
    prev_st = lib.pd_set_sumtail(1 if path == "sumtail" else 0)
    try:
        if path == "reflow":
            rf = pb.RectifiedFlow(128, _build(cfg, state), time_scale=1000, sampling_algorithm="rk2",
                                  spec_min=[-12.0], spec_max=[0.0]).cuda().eval()
            run = lambda **kw: rf(cond.cuda(), infer_step=2, infer=True, x_start=x_T.cuda(), **kw)
            runs = [run]
        else:
            gd = _build(cfg, state, dict(timesteps=4, schedule_type="vpsde", max_beta=40))
            runs = [lambda **kw: gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), noise=noise.cuda(), **kw),
                    lambda **kw: gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), seed=5, **kw)]
        with torch.no_grad():
            for run in runs:
                mel = run()
                c = run(vocoder_scale=handoff.LOG10_TO_LN)
                again = run()
                assert c.shape == (B, 128, T) and c.is_contiguous()
                assert torch.equal(c, handoff.generator_input(mel)) and torch.equal(again, mel)
                assert float(c.abs().max()) > 0
    finally:
        lib.pd_set_tail(1 if prev != 0 else 0)
        lib.pd_set_sumtail(1 if prev_st != 0 else 0)


@pytest.mark.parametrize("name", ["euler", "rk2", "rk4", "rk5", "pitch64"])
def test_reflow_matches_oracle_and_fixture(name):
    """RectifiedFlow / PitchRectifiedFlow .forward(infer=True) through pd_reflow_sample against the oracle
    (same injected start noise) and the reference-made fixture.  Bar: every velocity (= denoiser output) within max-abs
    1e-2 (x output scale for the Kaiming case), final result L1 within 1e-2 in normalised units, i.e. x (max - min) / 2
    after denorm_spec."""
    import prodiff_b200 as pb
    case, cfg, state, cond, x_start = helpers.regen_reflow_case(name)
    out_ref, vs_ref = helpers.oracle_reflow(case, cfg, state, cond, x_start)
    g = helpers.load_golden("reflow", name)
    wn = _build(cfg, state)
    smin, smax = helpers.reflow_spec_range(case)
    if case["spec"] == "pitch":
        rf = pb.PitchRectifiedFlow(cfg.in_dims, wn, time_scale=case["time_scale"], sampling_algorithm=case["algorithm"],
                                   spec_min=smin[0], spec_max=smax[0]).cuda().eval()
    else:
        rf = pb.RectifiedFlow(cfg.in_dims, wn, time_scale=case["time_scale"], sampling_algorithm=case["algorithm"],
                              spec_min=smin, spec_max=smax).cuda().eval()
    n_eval = len(vs_ref)
    tr = torch.empty(n_eval, case["B"], 1, cfg.in_dims, case["T"], device="cuda")
    with torch.no_grad():
        out = rf(cond.cuda(), infer_step=case["infer_step"], infer=True, x_start=x_start.cuda(), v_trace=tr).cpu()
        out2 = rf(cond.cuda(), infer_step=case["infer_step"], infer=True, x_start=x_start.cuda()).cpu()   # scratch path
    assert out.shape == out_ref.shape and torch.equal(out, out2)
    scale = _scale(case["out_init"])
    v_err = float((tr.cpu() - torch.stack(vs_ref)).abs().max())
    assert v_err <= STEP_TOL * scale, f"velocity max-abs {v_err:.4g}"
    half_range = (max(smax) - min(smin)) / 2
    l1 = float((out - out_ref).abs().mean())
    assert l1 <= MEL_L1_TOL * half_range * scale, f"final L1 {l1:.4g}"
    assert float((out - torch.from_numpy(g["out"])).abs().mean()) <= MEL_L1_TOL * half_range * scale


def test_reflow_inference_is_normalised_and_random_start():
    """``inference`` returns the normalised result like reflow.py:86-101 (denorm_spec is forward's job), and without
    ``x_start`` the start is torch.randn: two calls differ."""
    import prodiff_b200 as pb
    case, cfg, state, cond, x_start = helpers.regen_reflow_case("rk2")
    wn = _build(cfg, state)
    rf = pb.RectifiedFlow(cfg.in_dims, wn, time_scale=1000, sampling_algorithm="rk2", spec_min=[-12.0], spec_max=[0.0]).cuda().eval()
    with torch.no_grad():
        full = rf(cond.cuda(), infer_step=3, infer=True, x_start=x_start.cuda())
        norm = rf.inference(cond.cuda().transpose(1, 2), b=1, infer_step=3, x_start=x_start.cuda())
        assert torch.allclose(rf.denorm_spec(norm), full, atol=1e-5)
        a = rf(cond.cuda(), infer_step=2, infer=True)
        b = rf(cond.cuda(), infer_step=2, infer=True)
    assert a.shape == (1, case["T"], 128) and float((a - b).abs().mean()) > 1e-3


def test_large_batches_are_walked_in_row_blocks():
    """Engine.sample splits batches above FRAMES_PER_CALL into row blocks; utterances are independent, so the blocked
    result must equal the one-call result up to fp32 accumulation order (tiles pair up differently), including the
    injected-noise and x0-trace plumbing of the blocks."""
    from prodiff_b200 import synth
    from prodiff_b200.engine import Engine
    cfg = synth.DiffNetConfig(128, 256, 4, 256, 2)
    state = synth.make_state(cfg, 41, "trained")
    case = dict(timesteps=4, schedule_type="vpsde", max_beta=40)
    B, T, K = 7, 150, 3
    cond, x_T, noise = synth.make_inputs(cfg, B, T, K, 42)
    gd = _build(cfg, state, case)
    outs = []
    old = Engine.FRAMES_PER_CALL
    try:
        for fpc in (old, 300):                      # 300 frames per call -> 2 rows per block -> 4 blocks
            Engine.FRAMES_PER_CALL = fpc
            tr = torch.empty(K, B, 1, 128, T, device="cuda")
            with torch.no_grad():
                mel = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), noise=noise.cuda(), x0_trace=tr).cpu()
                cond2, x2 = cond.clone(), x_T.clone()
                cond2[2], x2[2] = cond2[0], x2[0]                  # rows 0 and 2: same inputs, same position in their blocks
                rnd = gd(cond2.cuda(), infer_step=K, infer=True, x_T=x2.cuda(), seed=9).cpu()
            outs.append((mel, tr.cpu(), rnd))
    finally:
        Engine.FRAMES_PER_CALL = old
    assert float((outs[0][0] - outs[1][0]).abs().max()) <= 2e-2 and float((outs[0][0] - outs[1][0]).abs().mean()) <= 1e-3
    assert float((outs[0][1] - outs[1][1]).abs().max()) <= 2e-2
    # in-kernel noise: every block has its own Philox stream, so identical rows in different blocks get different noise
    assert float((outs[1][2][0] - outs[1][2][2]).abs().mean()) > 1e-2 and torch.isfinite(outs[1][2]).all()


@pytest.mark.parametrize("T", [130, 333])
def test_no_writes_outside_the_caller_tensors(T):
    """The kernels write the caller's state / trace tensors with plain stores guarded by a frame-validity test (the TMA
    stores clip by construction).  Each tensor is carved out of a larger sentinel-filled buffer; after a ProDiff sampling
    call and a rectified-flow call on a ragged length (T not a multiple of the 128-frame tile, nor of 4) every sentinel
    byte must be intact.  (compute-sanitizer is not available on the GPU pool.)"""
    import prodiff_b200 as pb
    from prodiff_b200 import synth
    cfg = synth.DiffNetConfig(128, 256, 3, 256, 2)
    state = synth.make_state(cfg, 51, "trained")
    case = dict(timesteps=4, schedule_type="vpsde", max_beta=40)
    B, K, pad = 3, 3, 4096
    cond, x_T, noise = synth.make_inputs(cfg, B, T, K, 52)

    def carve(shape):
        n = int(np.prod(shape))
        big = torch.full((n + 2 * pad,), -12345.0, device="cuda")
        return big, big[pad:pad + n].view(shape)

    gd = _build(cfg, state, case)
    eng = gd.denoise_fn.engine()
    big_x, x = carve((B, 1, 128, T))
    big_t, tr = carve((K, B, 1, 128, T))
    x.copy_(x_T)
    tabs = gd._tables_on_host()
    with torch.no_grad():
        eng.sample(cond.cuda().contiguous(), x, list(range(K - 1, -1, -1)), *tabs, noise=noise.cuda(), x0_trace=tr)
    torch.cuda.synchronize()
    for big in (big_x, big_t):
        assert bool((big[:pad] == -12345.0).all()) and bool((big[-pad:] == -12345.0).all())
    assert torch.isfinite(x).all() and torch.isfinite(tr).all() and not bool((tr == -12345.0).any())
    # rectified flow: state, velocity trace
    big_x2, x2 = carve((B, 1, 128, T))
    big_v, vt = carve((2 * 4, B, 1, 128, T))
    x2.copy_(noise[0])
    smin = torch.full((1,), -12.0, device="cuda")
    smax = torch.zeros(1, device="cuda")
    with torch.no_grad():
        eng.reflow(cond.cuda().contiguous(), x2, "rk4", 2, 1000.0, smin, smax, vt)
    torch.cuda.synchronize()
    for big in (big_x2, big_v):
        assert bool((big[:pad] == -12345.0).all()) and bool((big[-pad:] == -12345.0).all())
    assert torch.isfinite(x2).all() and not bool((vt == -12345.0).any())


def test_ragged_batch_rows_equal_single_utterance_runs():
    """pd_plan_set_lengths: a padded batch with per-utterance lengths must give every row, up to its length, what that
    utterance gets when sampled alone at T = its length (the reference's offline loop, binarizer/svs.py:193-223, runs them
    one at a time).  Checked against single runs of the CUDA path (fp32 accumulation order is the only difference) and
    against the oracle per utterance, for the ProDiff loop and a rectified-flow integrator; dilations up to 4."""
    import prodiff_b200 as pb
    from prodiff_b200 import synth
    cfg = synth.DiffNetConfig(128, 256, 6, 256, 3)
    state = synth.make_state(cfg, 61, "trained")
    case = dict(timesteps=4, schedule_type="vpsde", max_beta=40, infer_step=3)
    lens = [300, 129, 37, 256, 1]
    B, T, K = len(lens), 300, 3
    cond, x_T, noise = synth.make_inputs(cfg, B, T, K, 62)
    gd = _build(cfg, state, case)
    rf = pb.RectifiedFlow(128, gd.denoise_fn, time_scale=1000, sampling_algorithm="rk2", spec_min=[-12.0], spec_max=[0.0]).cuda().eval()
    with torch.no_grad():
        mel = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), noise=noise.cuda(), lengths=lens).cpu()
        flow = rf(cond.cuda(), infer_step=2, infer=True, x_start=noise[0].cuda(), lengths=lens).cpu()
        padded = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), noise=noise.cuda()).cpu()
    differs = 0.0
    for b, n in enumerate(lens):
        c1, x1, n1 = cond[b:b + 1, :n], x_T[b:b + 1, :, :, :n].contiguous(), noise[:, b:b + 1, :, :, :n].contiguous()
        with torch.no_grad():
            solo = gd(c1.cuda(), infer_step=K, infer=True, x_T=x1.cuda(), noise=n1.cuda()).cpu()
            solo_flow = rf(c1.cuda(), infer_step=2, infer=True, x_start=n1[0].cuda()).cpu()
        d = (mel[b:b + 1, :n] - solo).abs()
        assert float(d.max()) <= 2e-2 and float(d.mean()) <= 1e-3, (b, n, float(d.max()), float(d.mean()))
        d = (flow[b:b + 1, :n] - solo_flow).abs()
        assert float(d.max()) <= 6 * 2e-2 and float(d.mean()) <= 6 * 1e-3, (b, n, float(d.max()), float(d.mean()))
        ref, _ = helpers.oracle_sample(case, cfg, state, c1, x1, n1)
        assert float((mel[b:b + 1, :n] - ref).abs().mean()) <= MEL_L1_TOL
        if n < T:
            differs = max(differs, float((padded[b, :n] - mel[b, :n]).abs().max()))
    assert differs > 1e-3     # without lengths the padded frames do leak into the last frames of short rows


def test_offline_sample_items_matches_single_utterance_loop():
    """offline.sample_items: the batched form of the reference's per-item teacher sampling loop
    (binarizer/svs.py:214-222) returns, per item, the condition, the x_T it drew and an x_0 equal to sampling that item
    alone from that x_T (use_x_T=True; rectified flow, euler, so the run is deterministic given x_T)."""
    import prodiff_b200 as pb
    from prodiff_b200 import offline, synth
    cfg = synth.DiffNetConfig(128, 256, 4, 256, 2)
    state = synth.make_state(cfg, 71, "trained")
    wn = _build(cfg, state)
    rf = pb.RectifiedFlow(128, wn, time_scale=1000, sampling_algorithm="euler", spec_min=[-12.0], spec_max=[0.0]).cuda().eval()
    g = torch.Generator().manual_seed(5)
    lens = [200, 64, 131, 199, 17, 260]
    conds = [torch.randn(n, 256, generator=g) for n in lens]
    gen = torch.Generator(device="cuda").manual_seed(6)
    res = offline.sample_items(rf, conds, infer_step=3, use_x_T=True, generator=gen, frames_per_batch=768)
    assert sorted(res) == list(range(len(lens)))
    for i, n in enumerate(lens):
        it = res[i]
        assert it["condition"].shape == (n, 256) and it["x_T"].shape == (n, 128) and it["x_0"].shape == (n, 128)
        assert np.array_equal(it["condition"], conds[i].numpy())
        x_start = torch.from_numpy(it["x_T"]).t()[None, None].contiguous().cuda()
        with torch.no_grad():
            solo = rf(conds[i][None].cuda(), infer_step=3, infer=True, x_start=x_start).cpu().numpy()[0]
        d = np.abs(solo - it["x_0"])
        assert d.max() <= 6 * 2e-2 and d.mean() <= 6 * 1e-3, (i, n, d.max(), d.mean())
    # as written in the reference (use_x_T=False) the sampler ignores the x_T it is handed and draws its own start
    res2 = offline.sample_items(rf, conds[:2], infer_step=2, use_x_T=False, generator=gen)
    assert res2[0]["x_0"].shape == (lens[0], 128) and np.isfinite(res2[0]["x_0"]).all()


@pytest.mark.parametrize("name", ["full", "frame_mix", "bare"])
def test_condition_assembly_kernel_matches_oracle_and_fixture(name):
    """SURVEY 8f N1 (frame-level half): pd_cond_assemble against the oracle restatement of prodiff_teacher.py:127-149 and
    the reference-made fixture; fp32 elementwise work, bound 1e-5 absolute (one-ulp differences of log / fma order)."""
    import types
    import prodiff_b200 as pb
    from oracle.golden_cases import CONDITION_CASES
    from prodiff_b200 import synth
    case = CONDITION_CASES[name]
    d, ref = helpers.oracle_condition(name)
    g = helpers.load_golden("condition", name)
    cfg = synth.DiffNetConfig(128, 256, 2, 256, 1)
    wn = _build(cfg, synth.make_state(cfg, 5, "trained"))
    cu = lambda t: None if t is None else t.cuda()
    utt, frame = [], None
    for e in (d["spk"], d["gender"]):
        if e is None:
            continue
        if e.shape[1] == case["T"] and case["T"] != 1:
            frame = e
        else:
            utt.append(e[:, 0])
    ac = pb.AssembledCondition(cu(d["encoder_out"]), cu(d["mel2ph"]), cu(d["f0"]), cu(d["pitch_w"]), cu(d["pitch_b"]),
                               [cu(u) for u in utt], cu(frame), [(cu(x), cu(w), cu(b)) for x, w, b in d["variances"]])
    out = ac.materialize(wn.engine()).cpu()
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 1e-5 and float((out - torch.from_numpy(g["out"])).abs().max()) <= 1e-5
    assert ac.lengths() == [int(v) for v in (d["mel2ph"] > 0).sum(1)]
    # the same through conditioning.forward_condition on a stand-in for the reference teacher module
    lin = lambda w, b: types.SimpleNamespace(weight=cu(w), bias=cu(b))
    teacher = types.SimpleNamespace(
        with_dur_embed=False, with_lang_embed=False, encoder=lambda tok, extra: cu(d["encoder_out"]),
        pitch_embed=lin(d["pitch_w"], d["pitch_b"]), with_spk_embed=d["spk"] is not None, with_gender_embed=d["gender"] is not None,
        add_spk_embed=lambda i, mix: mix, add_gender_embed=lambda i, mix: mix,
        with_voicing_embed=case["n_var"] >= 1, with_breath_embed=case["n_var"] >= 2)
    if case["n_var"] >= 1:
        teacher.voicing_embed = lin(d["variances"][0][1], d["variances"][0][2])
    if case["n_var"] >= 2:
        teacher.breath_embed = lin(d["variances"][1][1], d["variances"][1][2])
    ac2 = pb.forward_condition(teacher, torch.ones(case["B"], case["Tph"], dtype=torch.long).cuda(), cu(d["mel2ph"]), cu(d["f0"]),
                               spk_mix_embed=cu(d["spk"]), gender_mix_embed=cu(d["gender"]),
                               voicing=cu(d["variances"][0][0]) if case["n_var"] >= 1 else None,
                               breath=cu(d["variances"][1][0]) if case["n_var"] >= 2 else None)
    assert float((ac2.materialize(wn.engine()).cpu() - ref).abs().max()) <= 1e-5


def test_sampling_from_assembled_condition_equals_dense_condition():
    """GaussianDiffusion / RectifiedFlow accept the AssembledCondition in place of the dense [B,T,H] tensor: the 16-bit
    operand written by pd_cond_assemble is the rounding of the same fp32 values the dense path casts, so the samples are
    identical; with lengths taken from mel2ph the batch is ragged as well."""
    import prodiff_b200 as pb
    from prodiff_b200 import synth
    d, _ = helpers.oracle_condition("full")
    cfg = synth.DiffNetConfig(128, 256, 4, 256, 2)
    state = synth.make_state(cfg, 81, "trained")
    case = dict(timesteps=4, schedule_type="vpsde", max_beta=40)
    gd = _build(cfg, state, case)
    rf = pb.RectifiedFlow(128, gd.denoise_fn, time_scale=1000, sampling_algorithm="euler", spec_min=[-12.0], spec_max=[0.0]).cuda().eval()
    cu = lambda t: t.cuda()
    ac = pb.AssembledCondition(cu(d["encoder_out"]), cu(d["mel2ph"]), cu(d["f0"]), cu(d["pitch_w"]), cu(d["pitch_b"]),
                               [cu(d["spk"][:, 0]), cu(d["gender"][:, 0])], None, [(cu(x), cu(w), cu(b)) for x, w, b in d["variances"]])
    dense = ac.materialize(gd.denoise_fn.engine())
    B, T = dense.shape[:2]
    _, x_T, noise = synth.make_inputs(cfg, B, T, 3, 82)
    with torch.no_grad():
        a = gd(ac, infer_step=3, infer=True, x_T=x_T.cuda(), noise=noise.cuda())
        b = gd(dense, infer_step=3, infer=True, x_T=x_T.cuda(), noise=noise.cuda())
        assert torch.equal(a, b)
        fa = rf(ac, infer_step=2, infer=True, x_start=noise[0].cuda(), lengths=ac.lengths())
        fb = rf(dense, infer_step=2, infer=True, x_start=noise[0].cuda(), lengths=ac.lengths())
        assert torch.equal(fa, fb) and torch.isfinite(fa).all()


ENC_TOL = {"bf16": (2e-2, 3e-3), "fp16": (3e-3, 5e-4)}   # (max-abs, mean-abs) of encoder_out, whose entries are ~N(0,1) after the last LayerNorm


def _build_encoder(cfg, state, operand_dtype="bf16"):
    import prodiff_b200 as pb
    enc = pb.FastspeechEncoder(cfg.vocab_size, cfg.hidden_size, cfg.num_layers, cfg.kernel_size, num_heads=cfg.num_heads,
                               operand_dtype=operand_dtype)
    missing = enc.load_state_dict(state, strict=False)
    assert not missing.unexpected_keys and missing.missing_keys == ["embed_positions._float_tensor"]
    return enc.cuda().eval()


@pytest.mark.parametrize("operand_dtype", ["bf16", "fp16"])
@pytest.mark.parametrize("name", ["base4", "long_noextra", "tiny"])
def test_text_encoder_matches_oracle_and_fixture(name, operand_dtype):
    """SURVEY 8f N1 (token-level half): FastspeechEncoder.forward through pd_encoder_forward against the oracle restatement
    of tts_modules.py:291-330 and the reference-made fixture.  16-bit tensor-core operands for the four GEMMs of each block,
    fp32 LayerNorm / softmax / residual stream: encoder_out (unit scale) within 2e-2 max-abs / 3e-3 mean-abs for bf16 operands
    (measured <= 9.6e-3 / 1.4e-3), 3e-3 / 5e-4 for fp16 (measured 1.3e-3 / 1.6e-4); padding tokens are exact zeros."""
    case, cfg, state, tok, extra = helpers.regen_encoder_case(name)
    ref = helpers.oracle_encoder(case, cfg, state, tok, extra)
    g = torch.from_numpy(helpers.load_golden("encoder", name)["out"])
    enc = _build_encoder(cfg, state, operand_dtype)
    with torch.no_grad():
        out = enc(tok.cuda(), extra.cuda() if case["extra"] else None)
        out2 = enc(tok.cuda(), extra.cuda() if case["extra"] else None)
    assert out.shape == ref.shape and out.dtype == torch.float32 and torch.equal(out, out2)
    out = out.cpu()
    tol_max, tol_mean = ENC_TOL[operand_dtype]
    for want in (ref, g):
        assert float((out - want).abs().max()) <= tol_max, float((out - want).abs().max())
        assert float((out - want).abs().mean()) <= tol_mean
    for b, n in enumerate(case["lengths"]):
        assert not out[b, n:].any()


def test_text_encoder_ragged_rows_and_interior_padding():
    """Rows are independent (each row equals that utterance run alone at its own length: attention masks padded keys, the
    FFN convolution sees LayerNorm(0) = bias at padded positions inside T exactly like the reference), and a PAD token in
    the middle of a row is handled like the reference handles it (masked key, zero output, positions skip it)."""
    from prodiff_b200 import synth
    cfg = synth.EncoderConfig(50, 256, 2, 9, 2)
    state = synth.make_encoder_state(cfg, 131)
    tok, extra = synth.make_encoder_inputs(cfg, [140, 90, 129], 140, 132)
    tok[1, 17] = 0                                                     # interior padding
    case = dict(extra=True, lengths=[140, 90, 129])
    ref = helpers.oracle_encoder(case, cfg, state, tok, extra)
    enc = _build_encoder(cfg, state, "fp16")
    with torch.no_grad():
        out = enc(tok.cuda(), extra.cuda()).cpu()
    assert float((out - ref).abs().max()) <= ENC_TOL["fp16"][0] and not out[1, 17].any() and not out[1, 90:].any()


def test_text_encoder_edge_rows():
    """A one-token utterance, an utterance with no real token at all (the reference's softmax over an all-masked row is NaN;
    here that row is defined as zeros and must not disturb its neighbours) and a long row (700 tokens: 11 key tiles)."""
    from prodiff_b200 import synth
    cfg = synth.EncoderConfig(50, 256, 2, 9, 2)
    state = synth.make_encoder_state(cfg, 161)
    tok, extra = synth.make_encoder_inputs(cfg, [1, 0, 700], 700, 162)
    enc = _build_encoder(cfg, state, "fp16")
    with torch.no_grad():
        out = enc(tok.cuda(), extra.cuda()).cpu()
    assert torch.isfinite(out).all() and not out[1].any() and not out[0, 1:].any()
    for b in (0, 2):                                                    # each real row against the oracle run on it alone
        ref = helpers.oracle_encoder(dict(extra=True), cfg, state, tok[b:b + 1], extra[b:b + 1])
        assert float((out[b:b + 1] - ref).abs().max()) <= ENC_TOL["fp16"][0]


def test_text_encoder_random_shape_sweep():
    """Seeded sweep over batch size, token count (tile boundaries of the 128-row GEMM tiles and the 64-key attention tiles),
    depth, FFN kernel size, head count, operand type and padding patterns, each against the oracle."""
    from prodiff_b200 import synth
    rng = np.random.default_rng(171)
    sizes = [1, 2, 63, 64, 65, 127, 128, 129, 191, 192, 193, 257, 300]
    for i in range(12):
        heads = int(rng.choice([2, 4]))
        cfg = synth.EncoderConfig(int(rng.integers(5, 80)), 256, int(rng.integers(1, 4)), int(rng.choice([1, 3, 5, 9])), heads)
        Tph = int(rng.choice(sizes))
        B = int(rng.integers(1, 6))
        lengths = [Tph] + [int(rng.integers(1, Tph + 1)) for _ in range(B - 1)]
        state = synth.make_encoder_state(cfg, 200 + i)
        tok, extra = synth.make_encoder_inputs(cfg, lengths, Tph, 300 + i)
        use_extra = bool(i % 3)
        dt = "fp16" if i % 2 else "bf16"
        ref = helpers.oracle_encoder(dict(extra=use_extra), cfg, state, tok, extra)
        with torch.no_grad():
            out = _build_encoder(cfg, state, dt)(tok.cuda(), extra.cuda() if use_extra else None).cpu()
        err = float((out - ref).abs().max())
        assert err <= ENC_TOL[dt][0], (i, cfg, B, Tph, lengths, dt, err)
        assert float((out - ref).abs().mean()) <= ENC_TOL[dt][1]


def test_text_encoder_four_heads():
    """head_dim 64 instantiation of the attention kernel (the configured encoder has 2 heads of 128)."""
    from prodiff_b200 import synth
    cfg = synth.EncoderConfig(40, 256, 2, 5, 4)                        # 4 heads, FFN kernel 5
    state = synth.make_encoder_state(cfg, 151)
    tok, extra = synth.make_encoder_inputs(cfg, [70, 33], 70, 152)
    ref = helpers.oracle_encoder(dict(extra=True), cfg, state, tok, extra)
    with torch.no_grad():
        out = _build_encoder(cfg, state, "bf16")(tok.cuda(), extra.cuda()).cpu()
    assert float((out - ref).abs().max()) <= ENC_TOL["bf16"][0] and float((out - ref).abs().mean()) <= ENC_TOL["bf16"][1]


def test_text_to_mel_chain_matches_oracle_chain():
    """The whole front of ProDiffTeacher.forward (prodiff_teacher.py:103-167) on the CUDA path -- dur/lang embeddings (torch),
    FastspeechEncoder (pd_encoder_forward), frame-level condition assembly (pd_cond_assemble), K-step sampler (pd_sample) --
    against the same chain of the oracle restatements on identical x_T / noise: final mel L1 within the north-star 1e-2."""
    import types
    import prodiff_b200 as pb
    from oracle import encoder_oracle as eo
    from prodiff_b200 import synth
    ecfg = synth.EncoderConfig(40, 256, 4, 9, 2)
    estate = synth.make_encoder_state(ecfg, 141)
    B, Tph, T, K = 2, 30, 200, 4
    lens_ph = [30, 22]
    tok, _ = synth.make_encoder_inputs(ecfg, lens_ph, Tph, 142)
    rng = np.random.default_rng(143)
    mel2ph = np.zeros((B, T), dtype=np.int64)
    for b, n_fr in enumerate([200, 170]):
        mel2ph[b, :n_fr] = np.sort(rng.integers(1, lens_ph[b] + 1, size=n_fr))
    mel2ph = torch.from_numpy(mel2ph)
    f0 = torch.from_numpy(rng.uniform(80, 800, size=(B, T)).astype(np.float32))
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    dur_w, dur_b = f32(rng.normal(size=(256, 1)) * 0.05), f32(rng.normal(size=(256,)) * 0.05)
    pitch_w, pitch_b = f32(rng.normal(size=(256, 1)) * 0.3), f32(rng.normal(size=(256,)) * 0.1)
    spk = f32(rng.normal(size=(B, 1, 256)) * 0.3)
    cfg = synth.DiffNetConfig(128, 256, 6, 256, 2)
    state = synth.make_state(cfg, 144, "trained")
    case = dict(timesteps=4, schedule_type="vpsde", max_beta=40, infer_step=K)
    _, x_T, noise = synth.make_inputs(cfg, B, T, K, 145)
    # oracle chain
    with torch.no_grad():
        dur = eo.mel2ph_to_dur(mel2ph, Tph).float()
        extra = torch.nn.functional.linear(dur[:, :, None], dur_w, dur_b)
        enc_ref = eo.encoder_forward(estate, tok, extra, ecfg.num_heads)
        cond_ref = oracle.assemble_condition(enc_ref, mel2ph, f0, pitch_w, pitch_b, spk, None, [])
        mel_ref, _ = helpers.oracle_sample(case, cfg, state, cond_ref, x_T, noise)
    # CUDA chain through the reference-facing entry points
    enc = _build_encoder(ecfg, estate)
    gd = _build(cfg, state, case)
    dur_embed = torch.nn.Linear(1, 256)
    with torch.no_grad():
        dur_embed.weight.copy_(dur_w)
        dur_embed.bias.copy_(dur_b)
    dur_embed = dur_embed.cuda()
    teacher = types.SimpleNamespace(
        with_dur_embed=True, dur_embed=dur_embed, with_lang_embed=False, encoder=enc,
        pitch_embed=types.SimpleNamespace(weight=pitch_w.cuda(), bias=pitch_b.cuda()), with_spk_embed=True, with_gender_embed=False,
        add_spk_embed=lambda i, mix: mix, with_voicing_embed=False, with_breath_embed=False)
    with torch.no_grad():
        ac = pb.forward_condition(teacher, tok.cuda(), mel2ph.cuda(), f0.cuda(), spk_mix_embed=spk.cuda())
        mel = gd(ac, infer_step=K, infer=True, x_T=x_T.cuda(), noise=noise.cuda()).cpu()
    assert float((ac.encoder_out.cpu() - enc_ref).abs().max()) <= ENC_TOL["bf16"][0]
    valid = (mel2ph > 0)[:, :, None].expand_as(mel_ref)
    l1 = float((mel - mel_ref)[valid].abs().mean())
    assert l1 <= MEL_L1_TOL, f"text -> mel L1 {l1:.4g}"


def test_text_encoder_fails_loudly():
    import prodiff_b200 as pb
    with pytest.raises(NotImplementedError):
        pb.FastspeechEncoder(30, 256, 1, 9, rel_pos=True)
    enc = pb.FastspeechEncoder(30, 256, 1, 9).cuda()
    tok = torch.ones(1, 4, dtype=torch.long).cuda()
    with pytest.raises(RuntimeError, match="inference-only"):
        enc(tok)                                                       # training mode
    enc.eval()
    with pytest.raises(RuntimeError, match="inference-only"):
        enc(tok)                                                       # grad mode with trainable parameters
    with torch.no_grad():
        assert enc(tok).shape == (1, 4, 256)
        with pytest.raises(ValueError):
            enc(tok, torch.zeros(1, 5, 256).cuda())
    with pytest.raises(RuntimeError, match="512"):
        with torch.no_grad():
            pb.FastspeechEncoder(30, 512, 1, 9).cuda().eval()(tok)


SHAPE_SWEEP = [
    # (B, T, L, C, cycle, M, K, operand dtype, ragged)      what it stresses
    (1, 1, 2, 256, 1, 128, 1, "bf16", False),               # a single frame, a single step
    (2, 128, 3, 256, 2, 128, 2, "bf16", False),             # T exactly one tile, even tile count
    (3, 129, 2, 256, 5, 128, 2, "bf16", True),              # one frame into the second tile, dilation up to 16, ragged
    (5, 257, 3, 256, 3, 128, 3, "fp16", True),              # odd tile count (dummy peer tile), fp16 operands, ragged
    (2, 200, 2, 256, 1, 64, 2, "bf16", False),              # 64 mel bins: unfused tail, fused layer stack
    (1, 300, 2, 512, 2, 128, 2, "bf16", False),             # 512 channels: unfused conv+gate / res kernels
    (4, 96, 4, 256, 4, 128, 4, "bf16", True),               # T shorter than a tile, dilation 8 > T / 16, ragged
    (7, 40, 2, 256, 2, 128, 2, "fp16", False),              # many short rows
]


@pytest.mark.parametrize("B,T,L,C,cycle,M,K,dt,ragged", SHAPE_SWEEP)
def test_shape_sweep_against_oracle(B, T, L, C, cycle, M, K, dt, ragged):
    """Seeded sweep over shapes the fixtures do not cover (tile-boundary lengths, odd tile counts, the largest dilation,
    64 mel bins, 512 channels, single frames / steps, ragged rows), each against the oracle on identical x_T and noise."""
    from prodiff_b200 import synth
    cfg = synth.DiffNetConfig(M, 256, L, C, cycle)
    state = synth.make_state(cfg, 900 + B * 7 + T, "trained")
    case = dict(timesteps=max(4, K), schedule_type="vpsde", max_beta=40, infer_step=K)
    cond, x_T, noise = synth.make_inputs(cfg, B, T, K, 1000 + T)
    gd = _build(cfg, state, case, dt)
    lens = [max(1, T - (b * (T // 3 + 1)) % T) for b in range(B)] if ragged else None
    tr = torch.empty(K, B, 1, M, T, device="cuda")
    with torch.no_grad():
        mel = gd(cond.cuda(), infer_step=K, infer=True, x_T=x_T.cuda(), noise=noise.cuda(), x0_trace=tr, lengths=lens).cpu()
    tr = tr.cpu()
    assert torch.isfinite(mel).all()
    for b in range(B):
        n = lens[b] if ragged else T
        if ragged:      # the oracle of a ragged row is that utterance alone at its own length
            ref, x0s = helpers.oracle_sample(case, cfg, state, cond[b:b + 1, :n], x_T[b:b + 1, :, :, :n].contiguous(),
                                             noise[:, b:b + 1, :, :, :n].contiguous())
        elif b == 0:
            ref_all, x0s_all = helpers.oracle_sample(case, cfg, state, cond, x_T, noise)
        if not ragged:
            ref, x0s = ref_all[b:b + 1], [v[b:b + 1] for v in x0s_all]
        e0 = float((tr[0, b:b + 1, :, :, :n] - x0s[0]).abs().max())
        l1 = float((mel[b:b + 1, :n] - ref).abs().mean())
        assert e0 <= STEP_TOL and l1 <= MEL_L1_TOL, (b, n, e0, l1)


def test_unsupported_combinations_fail_loudly():
    """No silent fallbacks: per-utterance lengths on a shape the fused layer kernel does not cover, a bad length, and a
    v_trace of the wrong size are errors with a message, not a different code path."""
    import prodiff_b200 as pb
    from prodiff_b200 import synth
    cfg = synth.DiffNetConfig(128, 256, 2, 512, 1)
    gd = _build(cfg, synth.make_state(cfg, 3, "trained"), dict(timesteps=4, schedule_type="vpsde", max_beta=40))
    cond = torch.zeros(2, 64, 256, device="cuda")
    with pytest.raises(RuntimeError, match="fused layer kernel"):
        gd(cond, infer_step=2, infer=True, lengths=[64, 10])
    cfg2 = synth.DiffNetConfig(128, 256, 2, 256, 1)
    gd2 = _build(cfg2, synth.make_state(cfg2, 3, "trained"), dict(timesteps=4, schedule_type="vpsde", max_beta=40))
    with pytest.raises(RuntimeError, match="outside"):
        gd2(cond, infer_step=2, infer=True, lengths=[64, 65])
    with pytest.raises(ValueError):
        gd2(cond, infer_step=2, infer=True, lengths=[64])
    rf = pb.RectifiedFlow(128, gd2.denoise_fn, sampling_algorithm="rk4", spec_min=[-12.0], spec_max=[0.0]).cuda().eval()
    with pytest.raises(ValueError, match="v_trace"):
        rf(cond, infer_step=2, infer=True, v_trace=torch.empty(2, 2, 1, 128, 64, device="cuda"))
    out = gd2(cond, infer_step=2, infer=True, lengths=[64, 10])      # and the plan is usable again afterwards
    assert torch.isfinite(out).all()


def test_full_size_roundtrip_properties():
    """BASELINE config-2 size (64 x 1000, K = 4), properties that need no oracle run:
    determinism per seed, batch-row independence (a row sampled alone equals the same row inside the batch, which is
    what makes batch sharding across GPUs exact), finiteness, and linear response of the posterior step to noise."""
    from prodiff_b200 import synth
    cfg = synth.DiffNetConfig(128, 256, 20, 256, 1)
    state = synth.make_state(cfg, 1234, "trained")
    case = dict(timesteps=4, schedule_type="vpsde", max_beta=40)
    gd = _build(cfg, state, case)
    B, T, K = 64, 1000, 4
    cond, x_T, _ = synth.make_inputs(cfg, B, T, 1, 31)
    cond, x_T = cond.cuda(), x_T.cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    noise = torch.randn(K, B, 1, 128, T, device="cuda", generator=g)
    with torch.no_grad():
        a = gd(cond, infer_step=K, infer=True, x_T=x_T, noise=noise).clone()
        b = gd(cond, infer_step=K, infer=True, x_T=x_T, noise=noise).clone()
        assert torch.equal(a, b) and torch.isfinite(a).all() and a.shape == (B, T, 128)
        rows = [0, 37, 63]
        sub = gd(cond[rows], infer_step=K, infer=True, x_T=x_T[rows], noise=noise[:, rows].contiguous())
        # tiles pair up with different neighbours / clusters when the batch changes, so the fp32 accumulation ORDER (not
        # the operands) differs; single bf16 rounding flips then propagate through 20 layers x 4 steps
        assert float((sub - a[rows]).abs().mean()) <= 1e-3 and float((sub - a[rows]).abs().max()) <= 2e-2
        # x_{t-1} is affine in the injected noise of that step: check on the last noisy step (t = 1)
        n2 = noise.clone()
        n2[K - 2] += 1.0
        c = gd(cond[:2], infer_step=K, infer=True, x_T=x_T[:2], noise=noise[:, :2].contiguous())
        d = gd(cond[:2], infer_step=K, infer=True, x_T=x_T[:2], noise=n2[:, :2].contiguous())
    assert float((c - d).abs().mean()) > 1e-3

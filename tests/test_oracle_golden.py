"""The oracle restatement against the known-answer vectors the reference produced (CPU, no GPU).

The fixtures were written by oracle/make_golden.py running the unmodified reference; this is the
pin that lets the oracle stand in for the reference on the GPU box.  Tolerance is 2e-5 absolute, not
bit-exact, only because mkldnn may pick a different conv algorithm / thread split on another host;
bit-exactness against the live reference in one process is asserted in test_oracle_vs_reference.py.
"""
import numpy as np
import pytest
import torch

from oracle import prodiff_oracle as oracle
from oracle.golden_cases import SAMPLE_CASES, DENOISE_CASES, REFLOW_CASES, CONDITION_CASES, ENCODER_CASES
from prodiff_b200 import synth
from tests import helpers

TOL = 2e-5


@pytest.mark.parametrize("name", sorted(SAMPLE_CASES))
def test_sample_matches_reference_fixture(name):
    case, cfg, state, cond, x_T, noise = helpers.regen_sample_case(name)
    g = helpers.load_golden("sample", name)
    assert str(g["digest"]) == synth.digest(state, cond, x_T, noise), "regenerated inputs differ from the fixture's"
    mel, x0s = helpers.oracle_sample(case, cfg, state, cond, x_T, noise)
    assert mel.shape == (case["B"], case["T"], cfg.in_dims)
    x0 = torch.stack(x0s).numpy()
    assert x0.shape == g["x0"].shape
    assert np.abs(x0 - g["x0"]).max() <= TOL
    assert np.abs(mel.numpy() - g["mel"]).max() <= TOL
    # outputs are not degenerate (H1): the denoiser output must carry signal, not just its bias
    assert np.abs(g["x0"]).max() > 0.1


@pytest.mark.parametrize("name", sorted(SAMPLE_CASES))
def test_tables_match_reference_fixture(name):
    case = SAMPLE_CASES[name]
    g = helpers.load_golden("sample", name)
    tb = oracle.diffusion_tables(case["timesteps"], case["schedule_type"], case["max_beta"])
    for k_fix, k_tab in (("coef1", "posterior_mean_coef1"), ("coef2", "posterior_mean_coef2"),
                         ("logvar", "posterior_log_variance_clipped")):
        assert tb[k_tab].shape[0] == case["timesteps"] + 1
        np.testing.assert_array_equal(tb[k_tab].numpy(), g[k_fix])


@pytest.mark.parametrize("name", sorted(DENOISE_CASES))
def test_denoise_matches_reference_fixture(name):
    case, cfg, state, cond, x_T, steps = helpers.regen_denoise_case(name)
    g = helpers.load_golden("denoise", name)
    assert str(g["digest"]) == synth.digest(state, cond, x_T)
    with torch.no_grad():
        out = oracle.wavenet_forward(state, x_T, steps, cond.transpose(1, 2), cfg.dilation_cycle_length)
    assert out.shape == (case["B"], 1, cfg.in_dims, case["T"])
    assert np.abs(out.numpy() - g["out"]).max() <= TOL


@pytest.mark.parametrize("name", sorted(REFLOW_CASES))
def test_reflow_matches_reference_fixture(name):
    """RectifiedFlow / PitchRectifiedFlow .forward(infer=True) of the reference (reflow.py:32-107, :110-143): every
    velocity in call order and the denormalised result.  The denorm multiplies errors by (max - min) / 2 <= 8."""
    case, cfg, state, cond, x_start = helpers.regen_reflow_case(name)
    g = helpers.load_golden("reflow", name)
    assert str(g["digest"]) == synth.digest(state, cond, x_start), "regenerated inputs differ from the fixture's"
    out, vs = helpers.oracle_reflow(case, cfg, state, cond, x_start)
    stages = {"euler": 1, "rk2": 2, "rk4": 4, "rk5": 6}[case["algorithm"]]
    assert len(vs) == case["infer_step"] * stages and g["v"].shape[0] == len(vs)
    assert np.abs(torch.stack(vs).numpy() - g["v"]).max() <= TOL
    assert out.shape == g["out"].shape and np.abs(out.numpy() - g["out"]).max() <= 8 * TOL
    assert np.abs(g["v"]).max() > 0.1


def test_reflow_unknown_algorithm_is_euler():
    """reflow.py:90-95: ``dict.get(name, self.sample_euler)``."""
    case, cfg, state, cond, x_start = helpers.regen_reflow_case("euler")
    a, _ = helpers.oracle_reflow(case, cfg, state, cond, x_start)
    b, _ = helpers.oracle_reflow(dict(case, algorithm="heun?"), cfg, state, cond, x_start)
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", sorted(CONDITION_CASES))
def test_condition_assembly_matches_reference_fixture(name):
    """Frame-level half of ProDiffTeacher.forward_condition (prodiff_teacher.py:127-149): the fixture holds what the
    reference method returned for the seeded inputs."""
    d, out = helpers.oracle_condition(name)
    g = helpers.load_golden("condition", name)
    assert out.shape == g["out"].shape and np.abs(out.numpy() - g["out"]).max() <= TOL
    pad = d["mel2ph"].numpy() == 0
    assert not pad.any() or np.abs(out.numpy()[pad]).max() == 0.0      # non-padding mask, :148-149


def test_default_schedule_values():
    """SURVEY §8 a1/a2 probed values for the default config (vpsde, max_beta=40, timesteps=4)."""
    tb = oracle.diffusion_tables(4, "vpsde", 40)
    np.testing.assert_allclose(tb["betas"].numpy(), [.5587, .9105, .9819, .9963, .9993], atol=5e-5)
    np.testing.assert_allclose(tb["posterior_mean_coef1"].numpy(), [1, .6297, .1952, .0267, .0016], atol=5e-5)
    np.testing.assert_allclose(tb["posterior_mean_coef2"].numpy(), [0, .1740, .1294, .0606, .0273], atol=5e-5)


def test_unknown_schedule_raises():
    with pytest.raises(NotImplementedError):
        oracle.get_noise_schedule_list("nope", 4)


def test_t0_noise_is_masked():
    """prodiff.py:119-121: the noise drawn at t == 0 must not reach the output."""
    case, cfg, state, cond, x_T, noise = helpers.regen_sample_case("cosine_sched")
    mel_a, _ = helpers.oracle_sample(case, cfg, state, cond, x_T, noise)
    noise2 = noise.clone()
    noise2[-1] += 100.0
    mel_b, _ = helpers.oracle_sample(case, cfg, state, cond, x_T, noise2)
    assert torch.equal(mel_a, mel_b)


@pytest.mark.parametrize("name", sorted(ENCODER_CASES))
def test_encoder_matches_reference_fixture(name):
    """FastspeechEncoder.forward of the reference (tts_modules.py:291-330; SURVEY 8f N1, token-level half)."""
    case, cfg, state, tok, extra = helpers.regen_encoder_case(name)
    g = helpers.load_golden("encoder", name)
    assert str(g["digest"]) == synth.digest(state, tok.float(), extra), "regenerated inputs differ from the fixture's"
    out = helpers.oracle_encoder(case, cfg, state, tok, extra)
    assert out.shape == g["out"].shape == (len(case["lengths"]), case["Tph"], cfg.hidden_size)
    assert np.abs(out.numpy() - g["out"]).max() <= TOL
    for b, n in enumerate(case["lengths"]):           # padding tokens come out as exact zeros (FFTBlocks.forward :269-279)
        assert not out[b, n:].any() and out[b, :n].abs().max() > 0.5


def test_encoder_positions_and_durations():
    """make_positions over ~padding_mask (tts_utils.py:6-18, tts_modules.py:322) and mel2ph_to_dur (tts_modules.py:223-229)."""
    from oracle import encoder_oracle as eo
    tok = torch.tensor([[5, 7, 0, 9, 0], [0, 0, 0, 0, 0]])
    assert eo.make_positions(~tok.eq(0)).tolist() == [[1, 2, 0, 3, 0], [0, 0, 0, 0, 0]]
    tab = eo.sinusoid_table(8, 6)
    assert not tab[0].any() and torch.allclose(tab[1, :3], torch.sin(torch.tensor([1.0, 1e-2, 1e-4])), atol=1e-6)
    mel2ph = torch.tensor([[1, 1, 2, 4, 4, 4, 0, 0]])
    assert eo.mel2ph_to_dur(mel2ph, 4).tolist() == [[2, 1, 0, 3]]


@pytest.mark.parametrize("name", ["ladder", "rate8", "full44k_short"])
def test_vocoder_matches_reference_fixture(name):
    """oracle/vocoder_oracle.py against what the unmodified reference Generator produced (tests/golden/vocoder_*.npz)."""
    from oracle import vocoder_oracle as vo
    from oracle.golden_cases import VOCODER_CASES
    from prodiff_b200 import synth
    case = VOCODER_CASES[name]
    h = case["h"]
    state = synth.make_vocoder_state(h, case["wseed"])
    c, f0, ri, nz = synth.make_vocoder_inputs(h, case["B"], case["T"], case["iseed"])
    g = helpers.load_golden("vocoder", name)
    assert str(g["digest"]) == synth.digest(state, c, f0, ri, nz)
    with torch.no_grad():
        wav = vo.generator_forward(state, h, c, f0, ri, nz)
    assert np.array_equal(wav.numpy(), g["wav"])


def test_vocoder_source_phase_is_invariant_to_whole_turns():
    """The CUDA path computes SineGen's phase as P[f] + (j + 1) * rad[f] mod 1 (pd_voc.cuh, k_voc_phase) instead of the
    reference's sample-level float64 cumsum with whole-turn corrections (models.py:146-165).  Check that closed form
    against the oracle's op-for-op restatement on the CPU, including unvoiced stretches and the rand_ini quirk of frame 0."""
    from oracle import vocoder_oracle as vo
    from prodiff_b200 import synth
    h = synth.VOCODER_44K
    _, f0, ri, _ = synth.make_vocoder_inputs(h, 2, 60, 77)
    upp, sr = vo.upp_of(h), float(h["sampling_rate"])
    fn = f0.unsqueeze(-1) * torch.arange(1, 10).reshape(1, 1, -1)
    ref = vo.f02sine(fn.clone(), upp, sr, ri)
    rad = (fn / sr).fmod(1.)
    rad[:, 0, :] += ri
    P = torch.cumsum(rad.double() * upp, 1) - rad.double() * upp            # exclusive frame-level sum
    j = torch.arange(1, upp + 1, dtype=torch.double).reshape(1, 1, upp, 1)
    ph = (P.unsqueeze(2) + j * rad.double().unsqueeze(2)).reshape(2, 60 * upp, 9)
    mine = torch.sin((ph - ph.floor()) * 2 * np.pi).float()
    assert float((mine - ref).abs().max()) < 1e-5

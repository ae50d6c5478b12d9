"""Bit-exact check of the oracle against the live reference (build container only).

Skipped wherever /root/reference is absent (e.g. the GPU box): nothing that runs there may read it.
"""
import os
import sys

import numpy as np
import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "modules", "diffusion")),
                                reason="reference not mounted")


@pytest.fixture(scope="module")
def mg():
    import oracle.make_golden as mg
    mg.load_reference()
    yield mg
    if REF in sys.path:
        sys.path.remove(REF)


@pytest.mark.parametrize("name", ["dil4_ragged", "kaiming_short", "cosine_sched"])
def test_sample_bit_exact(mg, name):
    from oracle.golden_cases import SAMPLE_CASES
    from tests import helpers
    case, cfg, state, cond, x_T, noise = helpers.regen_sample_case(name)
    ref = mg.run_reference_sample(SAMPLE_CASES[name])
    mel, x0s = helpers.oracle_sample(case, cfg, state, cond, x_T, noise)
    assert np.array_equal(torch.stack(x0s).numpy(), ref["x0"])
    assert np.array_equal(mel.contiguous().numpy(), ref["mel"])


@pytest.mark.parametrize("name", ["fwd_int_steps", "fwd_float_steps", "fwd_in64"])
def test_denoise_bit_exact(mg, name):
    from oracle import prodiff_oracle as oracle
    from oracle.golden_cases import DENOISE_CASES
    from tests import helpers
    case, cfg, state, cond, x_T, steps = helpers.regen_denoise_case(name)
    ref = mg.run_reference_denoise(DENOISE_CASES[name])
    with torch.no_grad():
        out = oracle.wavenet_forward(state, x_T, steps, cond.transpose(1, 2), cfg.dilation_cycle_length)
    assert np.array_equal(out.numpy(), ref["out"])


@pytest.mark.parametrize("name", ["euler", "rk2", "rk4", "rk5", "pitch64"])
def test_reflow_bit_exact(mg, name):
    from oracle.golden_cases import REFLOW_CASES
    from tests import helpers
    case, cfg, state, cond, x_start = helpers.regen_reflow_case(name)
    ref = mg.run_reference_reflow(REFLOW_CASES[name])
    out, vs = helpers.oracle_reflow(case, cfg, state, cond, x_start)
    assert np.array_equal(torch.stack(vs).numpy(), ref["v"])
    assert np.array_equal(out.contiguous().numpy(), ref["out"])


@pytest.mark.parametrize("name", ["full", "frame_mix", "bare"])
def test_condition_assembly_bit_exact(mg, name):
    from oracle.golden_cases import CONDITION_CASES
    from tests import helpers
    ref = mg.run_reference_condition(CONDITION_CASES[name])
    _, out = helpers.oracle_condition(name)
    assert np.array_equal(out.numpy(), ref["out"])


@pytest.mark.parametrize("name", ["base4", "long_noextra", "tiny"])
def test_encoder_bit_exact(mg, name):
    from oracle.golden_cases import ENCODER_CASES
    from tests import helpers
    ref = mg.run_reference_encoder(ENCODER_CASES[name])
    out = helpers.oracle_encoder(*helpers.regen_encoder_case(name))
    assert np.array_equal(out.contiguous().numpy(), ref["out"])


def test_reference_init_distribution_matches_synth(mg):
    """synth.make_state mirrors the reference's init *distributions* (not its torch RNG stream)."""
    from prodiff_b200 import synth
    WaveNet, _ = mg.load_reference()
    torch.manual_seed(0)
    ref = WaveNet(128, 256, 4, 256, 1).state_dict()
    mine = synth.make_state(synth.DiffNetConfig(128, 256, 4, 256, 1), 0, out_init="zero")
    assert set(ref) == set(mine)
    for k in ref:
        assert ref[k].shape == mine[k].shape, k
        if ref[k].numel() >= 4096:
            assert abs(float(ref[k].std()) - float(mine[k].std())) <= 0.08 * float(ref[k].std()) + 1e-8, k


@pytest.mark.parametrize("name", ["ladder", "rate8", "full44k_short"])
def test_vocoder_bit_exact(mg, name):
    """oracle/vocoder_oracle.py against the live reference Generator (models.py:209-297), weight norm folded as load_model does."""
    from oracle import vocoder_oracle as vo
    from oracle.golden_cases import VOCODER_CASES
    from prodiff_b200 import synth
    case = VOCODER_CASES[name]
    h = case["h"]
    ref = mg.run_reference_vocoder(case)
    state = synth.make_vocoder_state(h, case["wseed"])
    c, f0, ri, nz = synth.make_vocoder_inputs(h, case["B"], case["T"], case["iseed"])
    with torch.no_grad():
        wav = vo.generator_forward(state, h, c, f0, ri, nz)
    assert np.array_equal(wav.numpy(), ref["wav"])


def test_vocoder_dropin_loads_a_weight_normed_reference_checkpoint(mg):
    """load_model (models.py:21-36) builds Generator(h), loads the checkpoint's weight-normed state and removes the norm.  The
    drop-in class must accept that very state dict (weight_g / weight_v keys), end up with the reference's plain weights and
    expose the reference's parameter names."""
    import prodiff_b200 as pb
    from oracle.golden_cases import VOCODER_CASES
    h = VOCODER_CASES["rate8"]["h"]
    Generator, AttrDict = mg.load_reference_generator()
    torch.manual_seed(5)
    ref = Generator(AttrDict(h))
    for p_ in ref.parameters():
        p_.data.normal_(0, 0.3)
    ckpt = {k: v.clone() for k, v in ref.state_dict().items()}          # weight_g / weight_v form
    assert any(k.endswith("weight_g") for k in ckpt)
    ours = pb.Generator(AttrDict(h))
    ours.load_state_dict(ckpt)
    ours.remove_weight_norm()
    ref.remove_weight_norm()
    plain = ref.state_dict()
    assert sorted(ours.state_dict().keys()) == sorted(plain.keys())   # (remove_weight_norm re-registers weight after bias)
    for k, v in plain.items():
        assert torch.allclose(ours.state_dict()[k], v, rtol=1e-6, atol=1e-7), k
    assert ours.upp == ref.upp

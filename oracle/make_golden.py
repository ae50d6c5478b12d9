"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

Run in the build container (the reference cannot travel to the GPU box):

    python oracle/make_golden.py            # writes tests/golden/<case>.npz

The reference draws x_T with ``torch.rand`` (prodiff.py:147) and per-step noise with
``torch.randn`` (prodiff.py:118); to feed it the injected tensors without editing it, the two
functions are patched for the duration of one ``forward`` call.  Per-step
x0 predictions are captured with a forward hook on the reference ``WaveNet``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("PRODIFF_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


def load_reference():
    if not os.path.isdir(REF):
        raise SystemExit(f"reference not found at {REF}")
    sys.path.insert(0, REF)
    from modules.decoder.wavenet import WaveNet            # noqa
    from modules.diffusion.prodiff import GaussianDiffusion  # noqa
    return WaveNet, GaussianDiffusion


class injected_randomness:
    """Patch torch.rand / torch.randn to pop pre-generated tensors (shape-checked)."""

    def __init__(self, x_T, noise):
        self.queue_rand = [x_T]
        self.queue_randn = list(noise)

    def __enter__(self):
        self._rand, self._randn = torch.rand, torch.randn

        def rand(*shape, **kw):
            t = self.queue_rand.pop(0)
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            return t.clone()

        def randn(shape, **kw):
            t = self.queue_randn.pop(0)
            assert tuple(t.shape) == tuple(shape), (t.shape, shape)
            return t.clone()

        torch.rand, torch.randn = rand, randn
        return self

    def __exit__(self, *a):
        torch.rand, torch.randn = self._rand, self._randn
        assert not self.queue_rand and not self.queue_randn, "reference consumed less randomness than injected"


def run_reference_sample(case):
    from prodiff_b200 import synth
    WaveNet, GaussianDiffusion = load_reference()
    cfg = case["cfg"]
    state = synth.make_state(cfg, case["wseed"], case["out_init"])
    cond, x_T, noise = synth.make_inputs(cfg, case["B"], case["T"], case["infer_step"], case["iseed"])
    wn = WaveNet(cfg.in_dims, cfg.hidden_size, cfg.residual_layers, cfg.residual_channels, cfg.dilation_cycle_length)
    wn.load_state_dict(state, strict=True)
    gd = GaussianDiffusion(cfg.in_dims, wn, timesteps=case["timesteps"], time_scale=1000,
                           schedule_type=case["schedule_type"], max_beta=case["max_beta"],
                           spec_min=[-12.0], spec_max=[0.0]).eval()
    x0s = []
    h = wn.register_forward_hook(lambda m, i, o: x0s.append(o.detach().clone()))
    with torch.no_grad(), injected_randomness(x_T, noise):
        mel = gd(cond, infer_step=case["infer_step"], infer=True)
    h.remove()
    return dict(mel=mel.contiguous().numpy(), x0=torch.stack(x0s).numpy(),
                digest=np.array(synth.digest(state, cond, x_T, noise)),
                coef1=gd.posterior_mean_coef1.numpy(), coef2=gd.posterior_mean_coef2.numpy(),
                logvar=gd.posterior_log_variance_clipped.numpy())


def run_reference_denoise(case):
    from prodiff_b200 import synth
    WaveNet, _ = load_reference()
    cfg = case["cfg"]
    state = synth.make_state(cfg, case["wseed"], case["out_init"])
    cond, x_T, _ = synth.make_inputs(cfg, case["B"], case["T"], 1, case["iseed"])
    wn = WaveNet(cfg.in_dims, cfg.hidden_size, cfg.residual_layers, cfg.residual_channels, cfg.dilation_cycle_length)
    wn.load_state_dict(state, strict=True)
    steps = torch.tensor(case["steps"])
    with torch.no_grad():
        out = wn(x_T, steps, cond.transpose(1, 2))
    return dict(out=out.numpy(), digest=np.array(synth.digest(state, cond, x_T)))


def run_reference_reflow(case):
    """RectifiedFlow / PitchRectifiedFlow .forward(infer=True) of the unmodified reference on injected start noise."""
    from prodiff_b200 import synth
    from oracle.golden_cases import reflow_spec_range
    WaveNet, _ = load_reference()
    from modules.diffusion.reflow import RectifiedFlow, PitchRectifiedFlow  # noqa
    cfg = case["cfg"]
    state = synth.make_state(cfg, case["wseed"], case["out_init"])
    cond, _, noise = synth.make_inputs(cfg, case["B"], case["T"], 1, case["iseed"])
    wn = WaveNet(cfg.in_dims, cfg.hidden_size, cfg.residual_layers, cfg.residual_channels, cfg.dilation_cycle_length)
    wn.load_state_dict(state, strict=True)
    smin, smax = reflow_spec_range(case)
    if case["spec"] == "pitch":
        rf = PitchRectifiedFlow(cfg.in_dims, wn, time_scale=case["time_scale"], sampling_algorithm=case["algorithm"],
                                spec_min=smin[0], spec_max=smax[0]).eval()
    else:
        rf = RectifiedFlow(cfg.in_dims, wn, time_scale=case["time_scale"], num_features=1,
                           sampling_algorithm=case["algorithm"], spec_min=smin, spec_max=smax).eval()
    vs = []
    h = wn.register_forward_hook(lambda m, i, o: vs.append(o.detach().clone()))
    real_randn = torch.randn

    def randn(*shape, **kw):
        assert tuple(shape) == tuple(noise[0].shape), (shape, noise[0].shape)
        return noise[0].clone()

    torch.randn = randn
    try:
        with torch.no_grad():
            out = rf(cond, infer_step=case["infer_step"], infer=True)
    finally:
        torch.randn = real_randn
        h.remove()
    return dict(out=out.contiguous().numpy(), v=torch.stack(vs).numpy(), digest=np.array(synth.digest(state, cond, noise[0])))


def run_reference_condition(case):
    """The reference's own ``ProDiffTeacher.forward_condition`` (prodiff_teacher.py:103-149), unmodified, run on a stub
    ``self`` whose encoder returns the injected ``encoder_out`` -- the frame-level half is what the fixture pins."""
    import types
    from oracle.golden_cases import make_condition_inputs
    load_reference()
    sys.modules.setdefault("chardet", types.ModuleType("chardet"))      # utils/__init__.py imports it; unused here
    from modules.svs.prodiff_teacher import ProDiffTeacher              # noqa
    d = make_condition_inputs(case)
    H = case["H"]

    def linear(w, b):
        m = torch.nn.Linear(1, H)
        with torch.no_grad():
            m.weight.copy_(w)
            m.bias.copy_(b)
        return m

    stub = types.SimpleNamespace(
        with_dur_embed=True, dur_embed=lambda dur: torch.zeros(1), with_lang_embed=False,
        encoder=lambda txt_tokens, extra_embed: d["encoder_out"].clone(),
        pitch_embed=linear(d["pitch_w"], d["pitch_b"]),
        with_spk_embed=case["spk"], with_gender_embed=case["gender"],
        with_voicing_embed=case["n_var"] >= 1, with_breath_embed=case["n_var"] >= 2)
    if case["n_var"] >= 1:
        stub.voicing_embed = linear(d["variances"][0][1], d["variances"][0][2])
    if case["n_var"] >= 2:
        stub.breath_embed = linear(d["variances"][1][1], d["variances"][1][2])
    for name in ("add_pitch", "add_spk_embed", "add_gender_embed"):
        setattr(stub, name, types.MethodType(getattr(ProDiffTeacher, name), stub))
    txt = torch.ones(case["B"], case["Tph"], dtype=torch.long)
    with torch.no_grad():
        out = ProDiffTeacher.forward_condition(
            stub, txt, d["mel2ph"], d["f0"], spk_mix_embed=d["spk"], gender_mix_embed=d["gender"],
            voicing=d["variances"][0][0] if case["n_var"] >= 1 else None,
            breath=d["variances"][1][0] if case["n_var"] >= 2 else None)
    return dict(out=out.numpy())


def run_reference_encoder(case):
    """The reference's own ``FastspeechEncoder`` (tts_modules.py:291-330), unmodified, in eval mode on seeded weights.
    Parameters are frozen (requires_grad False, as an inference deployment would have them): ATen picks the Conv1d backend
    from that flag, and the oracle's plain tensors take the same one, which is what makes bit-exact comparison possible."""
    import types
    from prodiff_b200 import synth
    load_reference()
    sys.modules.setdefault("chardet", types.ModuleType("chardet"))      # utils/__init__.py imports it; unused here
    from modules.fastspeech.tts_modules import FastspeechEncoder       # noqa
    cfg = case["cfg"]
    state = synth.make_encoder_state(cfg, case["wseed"])
    tok, extra = synth.make_encoder_inputs(cfg, case["lengths"], case["Tph"], case["iseed"])
    enc = FastspeechEncoder(cfg.vocab_size, cfg.hidden_size, cfg.num_layers, cfg.kernel_size, dropout=0.1,
                            num_heads=cfg.num_heads).eval()
    missing = enc.load_state_dict(state, strict=False)
    assert not missing.unexpected_keys and missing.missing_keys == ["embed_positions._float_tensor"], missing
    for p in enc.parameters():
        p.requires_grad_(False)
    with torch.no_grad():
        out = enc(tok, extra if case["extra"] else None)
    return dict(out=out.contiguous().numpy(), digest=np.array(synth.digest(state, tok.float(), extra)))


def load_reference_generator():
    """The reference's NSF-HiFiGAN ``Generator`` class (models.py:209).  nsf_hifigan/utils.py imports matplotlib at module
    level (absent here, unused by the generator): stub it."""
    import types
    load_reference()
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        sys.modules["matplotlib"] = mpl
    from modules.nsf_hifigan.models import Generator      # noqa
    from modules.nsf_hifigan.env import AttrDict           # noqa
    return Generator, AttrDict


def build_reference_generator(h, state):
    """Construct the unmodified reference generator, fold its weight norm (what load_model does, models.py:33) and load the
    seeded plain weights."""
    Generator, AttrDict = load_reference_generator()
    _print = __builtins__["print"] if isinstance(__builtins__, dict) else __builtins__.print
    gen = Generator(AttrDict(h)).eval()
    gen.remove_weight_norm()
    missing = gen.load_state_dict(state, strict=True)
    for p in gen.parameters():
        p.requires_grad_(False)
    return gen


class injected_sinegen_randomness:
    """Patch torch.rand / torch.randn_like for one Generator.forward: SineGen draws rand_ini with torch.rand(1, dim, device=...)
    (models.py:143) and the additive noise with torch.randn_like(sine_waves) (:186)."""

    def __init__(self, rand_ini, noise):
        self.rand_ini, self.noise = rand_ini, noise

    def __enter__(self):
        self._rand, self._randn_like = torch.rand, torch.randn_like

        def rand(*shape, **kw):
            assert tuple(shape) == tuple(self.rand_ini.shape), shape
            return self.rand_ini.clone()

        def randn_like(t, **kw):
            assert tuple(t.shape) == tuple(self.noise.shape), (t.shape, self.noise.shape)
            return self.noise.clone()

        torch.rand, torch.randn_like = rand, randn_like
        return self

    def __exit__(self, *a):
        torch.rand, torch.randn_like = self._rand, self._randn_like


def run_reference_vocoder(case):
    from prodiff_b200 import synth
    h = case["h"]
    state = synth.make_vocoder_state(h, case["wseed"])
    c, f0, rand_ini, noise = synth.make_vocoder_inputs(h, case["B"], case["T"], case["iseed"])
    gen = build_reference_generator(h, state)
    with torch.no_grad(), injected_sinegen_randomness(rand_ini, noise):
        wav = gen(c, f0)
    return dict(wav=wav.contiguous().numpy(), digest=np.array(synth.digest(state, c, f0, rand_ini, noise)))


def main():
    from oracle.golden_cases import SAMPLE_CASES, DENOISE_CASES, REFLOW_CASES, CONDITION_CASES, ENCODER_CASES, VOCODER_CASES
    if "--vocoder-only" in sys.argv:
        SAMPLE_CASES, DENOISE_CASES, REFLOW_CASES, CONDITION_CASES, ENCODER_CASES = {}, {}, {}, {}, {}
    if "--encoder-only" in sys.argv:
        SAMPLE_CASES, DENOISE_CASES, REFLOW_CASES, CONDITION_CASES, VOCODER_CASES = {}, {}, {}, {}, {}
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, case in SAMPLE_CASES.items():
        r = run_reference_sample(case)
        np.savez(os.path.join(out_dir, f"sample_{name}.npz"), **r)
        print(f"sample_{name}: mel {r['mel'].shape} x0 {r['x0'].shape} |x0|max {np.abs(r['x0']).max():.3f}")
    for name, case in DENOISE_CASES.items():
        r = run_reference_denoise(case)
        np.savez(os.path.join(out_dir, f"denoise_{name}.npz"), **r)
        print(f"denoise_{name}: out {r['out'].shape} |out|max {np.abs(r['out']).max():.3f}")
    for name, case in REFLOW_CASES.items():
        r = run_reference_reflow(case)
        np.savez(os.path.join(out_dir, f"reflow_{name}.npz"), **r)
        print(f"reflow_{name}: out {r['out'].shape} v {r['v'].shape} |v|max {np.abs(r['v']).max():.3f}")
    for name, case in CONDITION_CASES.items():
        r = run_reference_condition(case)
        np.savez(os.path.join(out_dir, f"condition_{name}.npz"), **r)
        print(f"condition_{name}: out {r['out'].shape} |out|max {np.abs(r['out']).max():.3f}")
    for name, case in ENCODER_CASES.items():
        r = run_reference_encoder(case)
        np.savez(os.path.join(out_dir, f"encoder_{name}.npz"), **r)
        print(f"encoder_{name}: out {r['out'].shape} |out|max {np.abs(r['out']).max():.3f}")
    for name, case in VOCODER_CASES.items():
        r = run_reference_vocoder(case)
        np.savez_compressed(os.path.join(out_dir, f"vocoder_{name}.npz"), **r)
        print(f"vocoder_{name}: wav {r['wav'].shape} |wav|max {np.abs(r['wav']).max():.3f} std {r['wav'].std():.3f}")


if __name__ == "__main__":
    main()

"""Known-answer cases for the sampling hot path (shared by make_golden.py and tests/).

Each case is regenerated from seeds by ``prodiff_b200.synth``; the ``.npz`` under
``tests/golden/`` stores only what the *reference* produced (per-step x0 predictions and the
final mel) plus a sha256 of the regenerated weights/inputs.
"""
from prodiff_b200.synth import DiffNetConfig, EncoderConfig

# name -> dict(cfg, timesteps, schedule_type, max_beta, infer_step, B, T, out_init, wseed, iseed)
SAMPLE_CASES = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case, full-size DiffNet.
    "cfg1_full": dict(cfg=DiffNetConfig(128, 256, 20, 256, 1), timesteps=4, schedule_type="vpsde", max_beta=40,
                      infer_step=4, B=1, T=400, out_init="trained", wseed=1234, iseed=4321),
    # dilation cycle 4 (dilations 1,2,4,8), ragged T (not a multiple of the 128-frame tile), B>1
    "dil4_ragged": dict(cfg=DiffNetConfig(128, 256, 8, 256, 4), timesteps=4, schedule_type="vpsde", max_beta=40,
                        infer_step=2, B=2, T=150, out_init="trained", wseed=11, iseed=12),
    # 512 residual channels (BASELINE config 4 shape), 100-step teacher schedule, first 3 reverse steps
    "c512_t100": dict(cfg=DiffNetConfig(128, 256, 4, 512, 4), timesteps=100, schedule_type="vpsde", max_beta=0.06,
                      infer_step=3, B=1, T=200, out_init="trained", wseed=21, iseed=22),
    # worst-case output scale (Kaiming output_projection), tiny T (shorter than every dilation halo sum)
    "kaiming_short": dict(cfg=DiffNetConfig(128, 256, 6, 256, 3), timesteps=4, schedule_type="vpsde", max_beta=40,
                          infer_step=4, B=3, T=7, out_init="kaiming", wseed=31, iseed=32),
    # other schedules the constructor accepts (prodiff.py:28-44)
    "linear_sched": dict(cfg=DiffNetConfig(128, 256, 2, 256, 1), timesteps=8, schedule_type="linear", max_beta=0.06,
                         infer_step=8, B=1, T=64, out_init="trained", wseed=41, iseed=42),
    "cosine_sched": dict(cfg=DiffNetConfig(128, 256, 2, 256, 2), timesteps=6, schedule_type="cosine", max_beta=0.06,
                         infer_step=6, B=1, T=33, out_init="trained", wseed=51, iseed=52),
}

# Denoiser-only cases: WaveNet.forward with per-row (possibly float) diffusion steps, as the training
# branch (prodiff.py:141-143) and the rectified-flow sampler (reflow.py:50) call it.
DENOISE_CASES = {
    "fwd_int_steps": dict(cfg=DiffNetConfig(128, 256, 4, 256, 2), B=3, T=130, steps=[0, 3, 1], out_init="trained",
                          wseed=61, iseed=62),
    "fwd_float_steps": dict(cfg=DiffNetConfig(128, 256, 3, 256, 1), B=2, T=257, steps=[250.0, 999.5],
                            out_init="kaiming", wseed=71, iseed=72),
    # PitchPredictor-shaped denoiser (in_dims = 64 repeat bins)
    "fwd_in64": dict(cfg=DiffNetConfig(64, 256, 2, 256, 1), B=1, T=100, steps=[2], out_init="trained",
                     wseed=81, iseed=82),
}

# Rectified-flow sampler over the same denoiser (reflow.py:48-107).  x_start ~ N(0,1) is injected in place
# of the torch.randn at reflow.py:88.  spec_min / spec_max: "bins" = a per-bin ramp, else the one-element lists the configs use.
REFLOW_CASES = {
    "euler": dict(cfg=DiffNetConfig(128, 256, 4, 256, 2), algorithm="euler", infer_step=4, time_scale=1000, B=2, T=150,
                  spec="bins", out_init="trained", wseed=91, iseed=92),
    "rk2": dict(cfg=DiffNetConfig(128, 256, 3, 256, 1), algorithm="rk2", infer_step=3, time_scale=1000, B=1, T=129,
                spec="scalar", out_init="trained", wseed=93, iseed=94),
    "rk4": dict(cfg=DiffNetConfig(128, 256, 3, 256, 3), algorithm="rk4", infer_step=2, time_scale=1000, B=2, T=77,
                spec="scalar", out_init="kaiming", wseed=95, iseed=96),
    "rk5": dict(cfg=DiffNetConfig(128, 256, 2, 256, 1), algorithm="rk5", infer_step=2, time_scale=1000, B=1, T=64,
                spec="bins", out_init="trained", wseed=97, iseed=98),
    # PitchRectifiedFlow shape: 64 repeat bins, denorm = mean over bins + clamp (reflow.py:110-143), default euler
    "pitch64": dict(cfg=DiffNetConfig(64, 256, 2, 256, 1), algorithm="euler", infer_step=5, time_scale=1000, B=2, T=100,
                    spec="pitch", out_init="trained", wseed=99, iseed=100),
}


def reflow_spec_range(case):
    """(spec_min, spec_max) lists for a REFLOW case."""
    M = case["cfg"].in_dims
    if case["spec"] == "bins":
        return [-12.0 + 0.03 * i for i in range(M)], [0.0 + 0.01 * i for i in range(M)]
    if case["spec"] == "pitch":
        return [-8.0], [8.0]
    return [-12.0], [0.0]


# Frame-level half of ProDiffTeacher.forward_condition (prodiff_teacher.py:127-149): inputs are drawn by
# make_condition_inputs below, the fixture stores what the reference method returned for them.
CONDITION_CASES = {
    # speaker + gender rows, both variance embeddings, trailing padding of different lengths
    "full": dict(B=3, T=150, Tph=23, H=256, spk=True, gender=True, n_var=2, frame_mix=False, seed=111),
    # speaker mix given per frame ([B,T,H]) as the web handler does, no gender, one variance embedding
    "frame_mix": dict(B=2, T=77, Tph=9, H=256, spk=True, gender=False, n_var=1, frame_mix=True, seed=112),
    # nothing optional
    "bare": dict(B=1, T=40, Tph=5, H=256, spk=False, gender=False, n_var=0, frame_mix=False, seed=113),
}


def make_condition_inputs(case):
    """Seeded inputs of a CONDITION case (numpy PCG64 like prodiff_b200.synth, reproducible anywhere)."""
    import numpy as np
    import torch
    rng = np.random.default_rng(case["seed"])
    B, T, Tph, H = case["B"], case["T"], case["Tph"], case["H"]
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    d = dict(encoder_out=f32(rng.normal(size=(B, Tph, H))), pitch_w=f32(rng.normal(size=(H, 1)) * 0.3), pitch_b=f32(rng.normal(size=(H,)) * 0.1))
    mel2ph = np.zeros((B, T), dtype=np.int64)
    for b in range(B):
        n = T - (b * 17) % max(1, T // 2)                       # valid frames of row b; the rest is padding (0)
        mel2ph[b, :n] = np.sort(rng.integers(1, Tph + 1, size=n))
    d["mel2ph"] = torch.from_numpy(mel2ph)
    d["f0"] = f32(np.where(rng.random((B, T)) < 0.2, 0.0, rng.uniform(80, 800, size=(B, T))))   # unvoiced frames have f0 = 0
    d["spk"] = (f32(rng.normal(size=(B, T if case["frame_mix"] else 1, H))) if case["spk"] else None)
    d["gender"] = f32(rng.normal(size=(B, 1, H))) if case["gender"] else None
    d["variances"] = [(f32(rng.random((B, T))), f32(rng.normal(size=(H, 1)) * 0.3), f32(rng.normal(size=(H,)) * 0.1))
                      for _ in range(case["n_var"])]
    return d


# FastSpeech text encoder (tts_modules.py:291-330; SURVEY §8f N1, token-level half): FastspeechEncoder.forward(txt_tokens,
# extra_embed) in eval mode.  ``lengths`` = non-padding tokens per row (trailing PAD), extra = whether extra_embed is given.
ENCODER_CASES = {
    # the configured encoder (base_config.yaml:106-116: 4 layers, 2 heads, kernel 9), ragged rows
    "base4": dict(cfg=EncoderConfig(64, 256, 4, 9, 2), Tph=50, lengths=[50, 37, 44], extra=True, wseed=121, iseed=122),
    # rows longer than one 128-token tile, a one-token row, no extra_embed
    "long_noextra": dict(cfg=EncoderConfig(40, 256, 2, 9, 2), Tph=200, lengths=[200, 1], extra=False, wseed=123, iseed=124),
    # shorter than the conv kernel
    "tiny": dict(cfg=EncoderConfig(30, 256, 1, 9, 2), Tph=5, lengths=[5, 3], extra=True, wseed=125, iseed=126),
}


# NSF-HiFiGAN generator (modules/nsf_hifigan/models.py:269-290).  ``h`` = the config.json fields the reference
# constructor reads; rand_ini and the SineGen noise are injected in place of torch.rand / torch.randn_like (models.py:143,186).
# B = 1 everywhere: the reference's SineGen only takes one utterance per call (models.py:163).
VOCODER_CASES = {
    # every tile width of the conv kernel: stages of 128 / 64 / 32 / 16 channels, rate-2 up-sampling (kernel 4)
    "ladder": dict(h=dict(resblock="1", num_mels=128, sampling_rate=44100, upsample_rates=[2, 2, 2, 2],
                          upsample_kernel_sizes=[4, 4, 4, 4], upsample_initial_channel=256, resblock_kernel_sizes=[3, 7, 11],
                          resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]]), B=1, T=150, wseed=201, iseed=202),
    # rate-8 up-sampling (kernel 16) into 64 and 32 channels, two resblock kernels, 64 mel bins
    "rate8": dict(h=dict(resblock="1", num_mels=64, sampling_rate=22050, upsample_rates=[8, 2], upsample_kernel_sizes=[16, 4],
                         upsample_initial_channel=128, resblock_kernel_sizes=[3, 11],
                         resblock_dilation_sizes=[[1, 3, 5], [1, 2, 4]]), B=1, T=70, wseed=203, iseed=204),
    # the 44.1 kHz configuration (512 samples per frame), one short utterance; a single resblock kernel exercises the
    # no-running-sum path
    "full44k_short": dict(h=dict(resblock="1", num_mels=128, sampling_rate=44100, upsample_rates=[8, 8, 2, 2, 2],
                                 upsample_kernel_sizes=[16, 16, 4, 4, 4], upsample_initial_channel=512,
                                 resblock_kernel_sizes=[7], resblock_dilation_sizes=[[1, 3, 5]]), B=1, T=9, wseed=205, iseed=206),
}

"""prodiff_b200 -- B200-native implementation of the ProDiff mel sampling hot path.

Public surface (mirrors the reference's module API for this path):

    WaveNet(in_dims, hidden_size, residual_layers, residual_channels, dilation_cycle_length).forward(spec, step, cond)
    GaussianDiffusion(out_dims, denoise_fn, timesteps, time_scale, betas, schedule_type, max_beta, spec_min, spec_max)
        .forward(cond, src_spec=None, gt_spec=None, infer_step=4, infer=False)
    RectifiedFlow(out_dims, denoise_fn, time_scale, num_features, sampling_algorithm, spec_min, spec_max)
        .forward(cond, gt_spec=None, infer_step=20, infer=True)          (PitchRectifiedFlow likewise)
    FastspeechEncoder(vocab_size, hidden_size, num_layers, kernel_size, dropout, num_heads, rel_pos)
        .forward(txt_tokens, extra_embed=None)                           (the step before the sampler)
    Generator(h).forward(x, f0)                                          (the NSF-HiFiGAN generator after it)
    pipeline.Synthesizer(teacher, generator)(txt_tokens, mel2ph, f0)     (tokens -> waveforms for a padded batch of segments)

``install()`` swaps these classes into the reference's own modules so that ``ProDiffTeacher``, the SVS inferer
(``handler/infer/handler.py``), ``main.py infer`` and the vocoder hand-off run unchanged on top of them.
"""
from .wavenet import WaveNet
from .diffusion import GaussianDiffusion
from .reflow import RectifiedFlow, PitchRectifiedFlow
from .conditioning import AssembledCondition, forward_condition
from .encoder import FastspeechEncoder
from .vocoder import Generator
from .synth import DiffNetConfig

__version__ = "0.1.0"


def install():
    """Replace the reference classes in-place (call before ``modules.svs.prodiff_teacher`` is imported)."""
    import importlib
    wn = importlib.import_module("modules.decoder.wavenet")
    gd = importlib.import_module("modules.diffusion.prodiff")
    wn.WaveNet = WaveNet
    gd.GaussianDiffusion = GaussianDiffusion
    rf = importlib.import_module("modules.diffusion.reflow")
    rf.RectifiedFlow = RectifiedFlow
    rf.PitchRectifiedFlow = PitchRectifiedFlow
    tm = importlib.import_module("modules.fastspeech.tts_modules")   # prodiff_teacher.py:8 imports the name from here
    tm.FastspeechEncoder = FastspeechEncoder
    try:   # load_model (models.py:21-36) then builds, loads (weight-normed keys) and "un-norms" the replacement class
        voc = importlib.import_module("modules.nsf_hifigan.models")
        voc.Generator = Generator
    except ImportError:   # the reference's nsf_hifigan package needs matplotlib at import time (utils.py:1-3)
        pass
    return wn, gd


__all__ = ["WaveNet", "GaussianDiffusion", "RectifiedFlow", "PitchRectifiedFlow", "FastspeechEncoder", "Generator", "AssembledCondition", "forward_condition",
           "DiffNetConfig", "install"]

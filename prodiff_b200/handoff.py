"""Vocoder hand-off (first sub-step): the sampler's last kernel writes the vocoder generator's input.

The reference hands the sampler's ``mel_out [B,T,M]`` to ``NsfHifiGAN.spec2wav_torch``
(`/root/reference/handler/infer/handler.py:156-158,349`), which undoes the sampler's final transpose and rescales the
mel from log10 to natural log before calling the generator (`/root/reference/component/vocoder/nsf_hifigan.py:49-58`)::

    c = mel.transpose(2, 1)      # [B, M, T]  -- the physical layout of the sampler state
    c = 2.30259 * c
    y = self.model(c, f0).view(-1)

``[B,M,T]`` is exactly how the sampler state lies in HBM, so ``GaussianDiffusion.forward(..., vocoder_scale=LOG10_TO_LN)``
(and ``RectifiedFlow.forward`` likewise) has the last posterior update (the denorm_spec kernel for rectified flow) store
``fl(2.30259 * x)`` directly (``pd_plan_set_output_scale``): no transpose view, no extra elementwise kernel, no extra pass
over the mel, and the tensor is bit-identical to the reference's ``c``.

Second sub-step: the generator itself.  ``prodiff_b200.Generator`` (vocoder.py) is a drop-in for
``modules/nsf_hifigan/models.py:Generator`` on the CUDA path, batched; ``spec2wav_torch`` below is the whole hand-off
(`nsf_hifigan.py:49-58`) for a batch: mel (or the already scaled ``c``) and f0 in, one waveform row per utterance out.
"""
from __future__ import annotations

import torch

LOG10_TO_LN = 2.30259   # the literal of nsf_hifigan.py:52 (not math.log(10))


def generator_input(mel: torch.Tensor, scale: float = LOG10_TO_LN) -> torch.Tensor:
    """What nsf_hifigan.py:50-52 computes from a [B,T,M] mel -- for callers that only have the unscaled result."""
    return scale * mel.transpose(2, 1)


@torch.no_grad()
def spec2wav_from_generator_input(vocoder, c: torch.Tensor, **kwargs) -> torch.Tensor:
    """The tail of ``spec2wav_torch`` (nsf_hifigan.py:53-58) for an input that already is the generator's ``c``
    (as returned by ``forward(..., vocoder_scale=LOG10_TO_LN)``): ``vocoder.model(c, f0).view(-1)``."""
    f0 = kwargs.get("f0")
    return (vocoder.model(c, f0) if f0 is not None else vocoder.model(c)).view(-1)


@torch.no_grad()
def spec2wav_torch(generator, mel: torch.Tensor = None, f0: torch.Tensor = None, *, c: torch.Tensor = None, lengths=None, **kwargs):
    """``NsfHifiGAN.spec2wav_torch`` (nsf_hifigan.py:49-58) over a batch with ``prodiff_b200.Generator``: either ``mel``
    [B,T,M] (log10; scaled and transposed here like :50-52) or ``c`` [B,M,T] (the sampler's ``vocoder_scale`` output).
    Returns [B, T * upp] -- or, with ``lengths`` (frames per utterance of a padded batch), a list of 1-D waveforms of
    ``lengths[b] * upp`` samples, each exactly what the call on that utterance alone returns (the generator holds every
    activation past an utterance's own end at zero; the reference flattens its single utterance with ``.view(-1)``)."""
    if (mel is None) == (c is None):
        raise ValueError("pass exactly one of mel [B,T,M] or c [B,M,T]")
    if c is None:
        c = LOG10_TO_LN * mel.transpose(2, 1)
    wav = generator(c, f0, lengths=lengths, **kwargs)[:, 0]
    if lengths is None:
        return wav
    return [wav[b, :int(n) * generator.upp] for b, n in enumerate(lengths)]


__all__ = ["LOG10_TO_LN", "generator_input", "spec2wav_from_generator_input", "spec2wav_torch"]

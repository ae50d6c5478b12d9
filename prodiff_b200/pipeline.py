"""Tokens -> waveform for a BATCH of segments: the callers either side of the hot path, tied together.

The reference's SVS inferer runs one segment per call (`/root/reference/handler/infer/handler.py:373-378`):
``InferHandler.run_model`` (:133-149) = ``ProDiffTeacher.forward(..., infer=True)`` (`modules/svs/prodiff_teacher.py:147-168`:
``forward_condition`` then ``diffusion(condition, infer=True)``), followed by ``run_vocoder`` (:156-158) =
``NsfHifiGAN.spec2wav_torch(mel, f0=f0)`` (`component/vocoder/nsf_hifigan.py:49-58`).  Segments are independent, so the
B200-first form takes all segments of a project as one padded batch:

    condition assembly (pd_cond_assemble, fed by the FastSpeech encoder, pd_encoder_forward)
      -> K-step sampler on the ragged batch (pd_sample with per-utterance lengths: every row comes out as that segment alone)
      -> generator input straight from the last posterior update (vocoder_scale fused, no transpose / scale pass)
      -> NSF-HiFiGAN generator on the batch (pd_vocoder_forward) -> one waveform per segment, cut to its own length.

``Synthesizer`` holds the teacher (the reference ``ProDiffTeacher`` built after ``prodiff_b200.install()``, or anything
with its attributes and a ``diffusion`` member) and a ``prodiff_b200.Generator``.  Nothing here computes: every step is one
of the library's entry points, and off a B200 the first of them raises.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .conditioning import forward_condition
from .handoff import LOG10_TO_LN, spec2wav_torch


class Synthesizer:
    def __init__(self, teacher, generator, infer_step: int = 4):
        self.teacher = teacher
        self.generator = generator
        self.infer_step = int(infer_step)

    @torch.no_grad()
    def mel(self, txt_tokens, mel2ph, f0, *, lengths: Optional[Sequence[int]] = None, validate: bool = True, vocoder_scale=None,
            **kw):
        """``InferHandler.run_model`` for a batch: the arguments of ``ProDiffTeacher.forward`` (conditioning keywords
        ``lang_seq, spk_embed_id, spk_mix_embed, gender_embed_id, gender_mix_embed, voicing, breath``) plus the sampler's
        keyword-only extras (``x_T, noise, seed``).  ``lengths``: valid frames per segment (default: the non-padding
        frames of ``mel2ph``); returns the mel ``[B,T,M]`` -- or, with ``vocoder_scale``, the generator input ``[B,M,T]``."""
        cond_kw = {k: kw.pop(k) for k in ("lang_seq", "spk_embed_id", "spk_mix_embed", "gender_embed_id", "gender_mix_embed",
                                          "voicing", "breath") if k in kw}
        ac = forward_condition(self.teacher, txt_tokens, mel2ph, f0, validate=validate, **cond_kw)
        if lengths is None and validate:
            lengths = ac.lengths()
        if lengths is not None and all(int(n) == ac.shape[1] for n in lengths):
            lengths = None   # a full batch needs no ragged handling
        return self.teacher.diffusion(ac, infer_step=self.infer_step, infer=True, lengths=lengths, vocoder_scale=vocoder_scale, **kw), lengths

    @torch.no_grad()
    def __call__(self, txt_tokens, mel2ph, f0, *, lengths: Optional[Sequence[int]] = None, validate: bool = True,
                 rand_ini=None, voc_noise=None, voc_seed: Optional[int] = None, **kw) -> List[torch.Tensor]:
        """Tokens -> one 1-D waveform per segment (``lengths[b] * upp`` samples), what ``InferHandler.infer`` (:335-349)
        produces segment by segment.  ``rand_ini`` / ``voc_noise`` / ``voc_seed`` go to the generator's SineGen."""
        c, lengths = self.mel(txt_tokens, mel2ph, f0, lengths=lengths, validate=validate, vocoder_scale=LOG10_TO_LN, **kw)
        gen_kw = {k: v for k, v in (("rand_ini", rand_ini), ("noise", voc_noise), ("seed", voc_seed)) if v is not None}
        if lengths is None:
            lengths = [c.shape[2]] * c.shape[0]
        return spec2wav_torch(self.generator, c=c, f0=f0, lengths=lengths, **gen_kw)


__all__ = ["Synthesizer"]

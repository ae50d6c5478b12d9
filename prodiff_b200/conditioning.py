"""Frame-level half of ``ProDiffTeacher.forward_condition`` (`/root/reference/modules/svs/prodiff_teacher.py:103-149`,
SURVEY §8f N1) fused into the sampler's condition operand.

The reference materialises the condition ``[B,T,H]`` in fp32 (gather by ``mel2ph``, pitch / speaker / gender / variance
embeddings, non-padding mask), hands it to the sampler, which casts it and projects it for every layer.  Here
``forward_condition`` runs the reference module's own token-level part (duration / language embeddings and the FastSpeech
encoder stay PyTorch -- 64 x ~200 tokens) and returns an ``AssembledCondition``: the small inputs of the frame-level part.
``GaussianDiffusion.forward`` / ``RectifiedFlow.forward`` accept it in place of the dense tensor and hand it to
``pd_cond_assemble``, which writes the 16-bit operand of the condition-projection GEMM directly; the dense fp32 tensor
(65 MB at 64 x 1000 x 256) is only produced on request (``materialize()``, e.g. for the offline binarizer that stores it).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch


class AssembledCondition:
    """Inputs of the frame-level condition assembly (all on one CUDA device, fp32 unless noted).

    encoder_out [B,Tph,H]; mel2ph [B,T] int64 (0 = padding); f0 [B,T]; pitch_w [H], pitch_b [H];
    utt: up to two [B,H] rows added per utterance (speaker, gender); frame: optional [B,T,H] added per frame (a frame-wise
    speaker / gender mix); variances: up to two (x [B,T], w [H], b [H]) Linear(1,H) embeddings (voicing, breath).
    """

    def __init__(self, encoder_out, mel2ph, f0, pitch_w, pitch_b, utt: Sequence[torch.Tensor] = (), frame=None,
                 variances: Sequence[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = (), validate: bool = True):
        dev = encoder_out.device
        f = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
        self.encoder_out = f(encoder_out)
        self.mel2ph = mel2ph.detach().to(device=dev, dtype=torch.int64).contiguous()
        self.f0 = f(f0)
        B, _, H = self.encoder_out.shape
        T = self.mel2ph.shape[1]
        if self.mel2ph.shape != (B, T) or self.f0.shape != (B, T):
            raise ValueError("mel2ph and f0 must both be [B, T]")
        # (a device -> host read: it makes the host wait for the encoder before it can queue the sampler's launches;
        # validate=False skips it -- the kernel itself treats an out-of-range index as padding and never reads through it)
        if validate and (int(self.mel2ph.max()) > self.encoder_out.shape[1] or int(self.mel2ph.min()) < 0):
            raise ValueError("mel2ph indexes outside the encoder output")
        self.pitch_w, self.pitch_b = f(pitch_w).reshape(H), f(pitch_b).reshape(H)
        self.utt = [f(u).reshape(B, H) for u in utt]
        self.frame = None if frame is None else f(frame).reshape(B, T, H)
        self.variances = [(f(x).reshape(B, T), f(w).reshape(H), f(b).reshape(H)) for x, w, b in variances]
        if len(self.utt) > 2 or len(self.variances) > 2:
            raise ValueError("at most two utterance-level and two variance embeddings")
        self.shape = torch.Size((B, T, H))
        self.device = dev
        self.dtype = torch.float32

    def lengths(self) -> List[int]:
        """Valid frames per utterance (count of non-padding frames; the reference pads at the end)."""
        return [int(v) for v in (self.mel2ph > 0).sum(dim=1).tolist()]

    def rows(self, s: int, e: int) -> "AssembledCondition":
        out = object.__new__(AssembledCondition)
        out.__dict__.update(self.__dict__)
        out.encoder_out, out.mel2ph, out.f0 = self.encoder_out[s:e], self.mel2ph[s:e], self.f0[s:e]
        out.utt = [u[s:e] for u in self.utt]
        out.frame = None if self.frame is None else self.frame[s:e]
        out.variances = [(x[s:e], w, b) for x, w, b in self.variances]
        out.shape = torch.Size((e - s,) + tuple(self.shape[1:]))
        return out

    def materialize(self, engine) -> torch.Tensor:
        """The dense fp32 condition [B,T,H] ``forward_condition`` returns (prodiff_teacher.py:149), computed by the kernel."""
        out = torch.empty(tuple(self.shape), dtype=torch.float32, device=self.device)
        engine.assemble(self, out)
        return out


def mel2ph_to_dur(mel2ph: torch.Tensor, T_txt: int) -> torch.Tensor:
    """Frames per token: tts_modules.py:223-229 with max_dur=None (an integer histogram of mel2ph, index 0 = padding dropped)."""
    dur = mel2ph.new_zeros(mel2ph.shape[0], T_txt + 1).scatter_add(1, mel2ph, torch.ones_like(mel2ph))
    return dur[:, 1:]


def forward_condition(teacher, txt_tokens, mel2ph, f0, lang_seq=None, spk_embed_id=None, spk_mix_embed=None,
                      gender_embed_id=None, gender_mix_embed=None, voicing=None, breath=None, *,
                      validate: bool = True) -> AssembledCondition:
    """Same arguments as ``ProDiffTeacher.forward_condition`` (prodiff_teacher.py:103-110), ``teacher`` being the reference
    module (or anything with its attributes).  Token-level part as in :111-126, frame-level part deferred to the kernel."""
    extra_embed = None
    if teacher.with_dur_embed:
        dur = mel2ph_to_dur(mel2ph, txt_tokens.shape[1]).float()                # :113
        extra_embed = teacher.dur_embed(dur[:, :, None])
    if teacher.with_lang_embed:
        assert lang_seq is not None, "use_lang_embed is True, lang_seq is required"
        lang_embed = teacher.lang_embed(lang_seq)
        extra_embed = lang_embed if extra_embed is None else extra_embed + lang_embed
    encoder_out = teacher.encoder(txt_tokens, extra_embed)
    utt, frame = [], None
    B, T = mel2ph.shape

    def place(e):
        nonlocal frame
        if e.dim() == 3 and e.shape[1] == T and T != 1:
            frame = e if frame is None else frame + e
        else:
            utt.append(e.reshape(B, -1))

    if teacher.with_spk_embed:
        place(teacher.add_spk_embed(spk_embed_id, spk_mix_embed))
    if teacher.with_gender_embed:
        place(teacher.add_gender_embed(gender_embed_id, gender_mix_embed))
    variances = []
    if teacher.with_voicing_embed:
        variances.append((voicing, teacher.voicing_embed.weight, teacher.voicing_embed.bias))
    if teacher.with_breath_embed:
        variances.append((breath, teacher.breath_embed.weight, teacher.breath_embed.bias))
    return AssembledCondition(encoder_out, mel2ph, f0, teacher.pitch_embed.weight, teacher.pitch_embed.bias, utt, frame, variances,
                              validate=validate)


__all__ = ["AssembledCondition", "forward_condition", "mel2ph_to_dur"]

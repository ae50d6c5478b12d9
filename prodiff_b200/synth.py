"""Deterministic synthetic DiffNet weights and sampler inputs (numpy PCG64, fp32).

There is no network and no checkpoint in the reference repo, so every test, the bench and the
golden-vector generator build their inputs here, from seeds only.  The distributions mirror the
reference's construction (`/root/reference/modules/decoder/wavenet.py:41-44,75-98`):

* every ``Conv1d`` weight: Kaiming-normal, std = sqrt(2 / (C_in * k));  bias U(+-1/sqrt(fan_in))
* every ``nn.Linear``: weight and bias U(+-1/sqrt(fan_in))  (torch default)
* final ``output_projection.weight``: the reference zero-initialises it (wavenet.py:98), which makes
  the freshly built denoiser output its bias only and every numerical comparison degenerate
  (SURVEY.md §7.3 H1).  ``out_init`` selects what to put there instead:
  ``"trained"`` = N(0, 0.02) (output std ~0.2, trained-like, the default for parity tests),
  ``"kaiming"`` = Kaiming-normal (output std ~0.9, worst case), ``"zero"`` = the reference's init.

The streams are numpy's, not torch's, so the same seed gives the same tensors on every box.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass, asdict
from typing import Dict

import numpy as np
import torch


@dataclass(frozen=True)
class DiffNetConfig:
    """Constructor arguments of the reference ``WaveNet`` (wavenet.py:75)."""
    in_dims: int = 128
    hidden_size: int = 256
    residual_layers: int = 20
    residual_channels: int = 256
    dilation_cycle_length: int = 1

    def dict(self):
        return asdict(self)


def _kaiming_normal(rng, shape, fan_in):
    return (rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)


def _uniform(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def make_state(cfg: DiffNetConfig, seed: int = 1234, out_init: str = "trained") -> Dict[str, torch.Tensor]:
    """A full ``WaveNet.state_dict()`` (same keys and shapes as the reference) from a seed."""
    rng = np.random.default_rng(seed)
    M, H, L, C = cfg.in_dims, cfg.hidden_size, cfg.residual_layers, cfg.residual_channels
    st: Dict[str, np.ndarray] = {}

    def conv(name, cout, cin, k):
        st[name + ".weight"] = _kaiming_normal(rng, (cout, cin, k), cin * k)
        st[name + ".bias"] = _uniform(rng, (cout,), 1.0 / np.sqrt(cin * k))

    def linear(name, cout, cin):
        st[name + ".weight"] = _uniform(rng, (cout, cin), 1.0 / np.sqrt(cin))
        st[name + ".bias"] = _uniform(rng, (cout,), 1.0 / np.sqrt(cin))

    conv("input_projection", C, M, 1)
    linear("mlp.0", 4 * C, C)
    linear("mlp.2", C, 4 * C)
    for i in range(L):
        p = f"residual_layers.{i}."
        conv(p + "dilated_conv", 2 * C, C, 3)
        linear(p + "diffusion_projection", C, C)
        conv(p + "conditioner_projection", 2 * C, H, 1)
        conv(p + "output_projection", 2 * C, C, 1)
    conv("skip_projection", C, C, 1)
    conv("output_projection", M, C, 1)
    if out_init == "trained":
        st["output_projection.weight"] = (rng.standard_normal((M, C, 1)) * 0.02).astype(np.float32)
    elif out_init == "zero":
        st["output_projection.weight"] = np.zeros((M, C, 1), np.float32)
    elif out_init != "kaiming":
        raise ValueError(f"unknown out_init {out_init!r}")
    return {k: torch.from_numpy(v) for k, v in st.items()}


def make_inputs(cfg: DiffNetConfig, B: int, T: int, K: int, seed: int = 4321):
    """cond [B,T,H] ~ N(0,1); x_T [B,1,M,T] ~ U[0,1) (the reference's start, prodiff.py:147);
    noise [K,B,1,M,T] ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    cond = rng.standard_normal((B, T, cfg.hidden_size)).astype(np.float32)
    x_T = rng.random((B, 1, cfg.in_dims, T)).astype(np.float32)
    noise = rng.standard_normal((K, B, 1, cfg.in_dims, T)).astype(np.float32)
    return torch.from_numpy(cond), torch.from_numpy(x_T), torch.from_numpy(noise)


def digest(*tensors) -> str:
    """sha256 over raw fp32 bytes; pins that regenerated inputs equal the ones a fixture was made from."""
    h = hashlib.sha256()
    for t in tensors:
        if isinstance(t, dict):
            for k in sorted(t):
                h.update(k.encode())
                h.update(t[k].contiguous().numpy().tobytes())
        else:
            h.update(t.contiguous().numpy().tobytes())
    return h.hexdigest()


# ---------------------------------------------------------------------------------------------------------------------
# FastSpeech text encoder (token-level half): constructor arguments of the reference ``FastspeechEncoder``
# (`/root/reference/modules/fastspeech/tts_modules.py:291-303`) and seeded weights with its init distributions
# (embedding N(0, H^-0.5) with a zero padding row, `common_layers.py:63-68`; xavier-uniform attention / ffn_2 weights,
# `:86-91,225-233`; torch-default Conv1d init for ffn_1).  LayerNorm affine parameters and the zero-initialised biases get a
# small perturbation ("trained-like"), otherwise they would not be exercised at all.
@dataclass(frozen=True)
class EncoderConfig:
    vocab_size: int = 64
    hidden_size: int = 256
    num_layers: int = 4
    kernel_size: int = 9
    num_heads: int = 2

    def dict(self):
        return asdict(self)


def make_encoder_state(cfg: EncoderConfig, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """A full ``FastspeechEncoder.state_dict()`` (same keys and shapes as the reference, minus the one-element
    ``embed_positions._float_tensor`` device marker) from a seed."""
    rng = np.random.default_rng(seed)
    H, k = cfg.hidden_size, cfg.kernel_size
    st: Dict[str, np.ndarray] = {}
    xavier = lambda shape: _uniform(rng, shape, np.sqrt(6.0 / (shape[0] + shape[1])))
    for i in range(cfg.num_layers):
        p = f"layers.{i}.op."
        for ln in ("layer_norm1", "layer_norm2"):
            st[p + ln + ".weight"] = (1.0 + 0.1 * rng.standard_normal(H)).astype(np.float32)
            st[p + ln + ".bias"] = (0.1 * rng.standard_normal(H)).astype(np.float32)
        st[p + "self_attn.in_proj_weight"] = xavier((3 * H, H))
        st[p + "self_attn.out_proj.weight"] = xavier((H, H))
        st[p + "ffn.ffn_1.weight"] = _uniform(rng, (4 * H, H, k), 1.0 / np.sqrt(H * k))
        st[p + "ffn.ffn_1.bias"] = _uniform(rng, (4 * H,), 1.0 / np.sqrt(H * k))
        st[p + "ffn.ffn_2.weight"] = xavier((H, 4 * H))
        st[p + "ffn.ffn_2.bias"] = (0.02 * rng.standard_normal(H)).astype(np.float32)
    st["layer_norm.weight"] = (1.0 + 0.1 * rng.standard_normal(H)).astype(np.float32)
    st["layer_norm.bias"] = (0.1 * rng.standard_normal(H)).astype(np.float32)
    emb = (rng.standard_normal((cfg.vocab_size, H)) * H ** -0.5).astype(np.float32)
    emb[0] = 0
    st["embed_tokens.weight"] = emb
    return {k_: torch.from_numpy(v) for k_, v in st.items()}


def make_encoder_inputs(cfg: EncoderConfig, lengths, Tph: int, seed: int = 4321):
    """txt_tokens [B,Tph] int64 (row b has lengths[b] non-padding tokens, then PAD = 0) and extra_embed [B,Tph,H]
    (what dur_embed + lang_embed add, prodiff_teacher.py:112-120; small, like Linear(1,H) of a duration)."""
    rng = np.random.default_rng(seed)
    B = len(lengths)
    tok = np.zeros((B, Tph), dtype=np.int64)
    for b, n in enumerate(lengths):
        tok[b, :n] = rng.integers(1, cfg.vocab_size, size=n)
    extra = (0.3 * rng.standard_normal((B, Tph, cfg.hidden_size))).astype(np.float32)
    return torch.from_numpy(tok), torch.from_numpy(extra)

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


# ---------------------------------------------------------------------------------------------------------------------
# NSF-HiFiGAN generator: the fields of the checkpoint's config.json that ``Generator.__init__`` reads
# (`/root/reference/modules/nsf_hifigan/models.py:210-267`) and seeded weights under the reference's ``state_dict`` keys
# after ``remove_weight_norm()``.  The reference initialises its convs with N(0, 0.01) (`nsf_hifigan/utils.py:6-9`), which
# leaves a freshly built generator with activations ~1e-3 and a waveform that is all bias; the synthetic state uses
# fan-in-scaled weights instead so that every stage carries O(1) signal and the output tanh is driven over its range.
VOCODER_44K = dict(resblock="1", num_mels=128, sampling_rate=44100, upsample_rates=[8, 8, 2, 2, 2],
                   upsample_kernel_sizes=[16, 16, 4, 4, 4], upsample_initial_channel=512,
                   resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]])


def make_vocoder_state(h: dict, seed: int = 1234) -> Dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    st: Dict[str, np.ndarray] = {}
    nrm = lambda shape, std: (rng.standard_normal(shape) * std).astype(np.float32)
    c0, rates, ksz = h["upsample_initial_channel"], h["upsample_rates"], h["upsample_kernel_sizes"]
    st["conv_pre.weight"] = nrm((c0, h["num_mels"], 7), 1.0 / np.sqrt(h["num_mels"] * 7))
    st["conv_pre.bias"] = nrm((c0,), 0.1)
    ch = c0
    for i, (u, k) in enumerate(zip(rates, ksz)):
        st[f"ups.{i}.weight"] = nrm((ch, ch // 2, k), 1.0 / np.sqrt(ch * k / u))      # k / u taps reach each output sample
        st[f"ups.{i}.bias"] = nrm((ch // 2,), 0.1)
        ch //= 2
        sk = 2 * int(np.prod(rates[i + 1:])) if i + 1 < len(rates) else 1
        st[f"noise_convs.{i}.weight"] = nrm((ch, 1, sk), 1.0 / np.sqrt(sk))
        st[f"noise_convs.{i}.bias"] = nrm((ch,), 0.1)
        for j, rk in enumerate(h["resblock_kernel_sizes"]):
            p = f"resblocks.{i * len(h['resblock_kernel_sizes']) + j}"
            for d in range(3):
                for cv in ("convs1", "convs2"):
                    st[f"{p}.{cv}.{d}.weight"] = nrm((ch, ch, rk), 0.6 / np.sqrt(ch * rk))
                    st[f"{p}.{cv}.{d}.bias"] = nrm((ch,), 0.05)
    st["conv_post.weight"] = nrm((1, ch, 7), 1.0 / np.sqrt(ch * 7))
    st["conv_post.bias"] = nrm((1,), 0.1)
    st["m_source.l_linear.weight"] = nrm((1, 9), 1.0 / 3.0 / 0.1)     # sine amplitudes are 0.1: keep the tanh input O(1)
    st["m_source.l_linear.bias"] = nrm((1,), 0.1)
    return {k_: torch.from_numpy(v) for k_, v in st.items()}


def make_vocoder_inputs(h: dict, lengths_or_B, T: int, seed: int = 4321):
    """c [B, num_mels, T] (the hand-off tensor: natural-log mel, roughly N(-5, 2)), f0 [B, T] Hz with unvoiced (0) stretches,
    rand_ini [1, 9] (entry 0 zero, models.py:143-144) and noise [B, T * upp, 9] ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    B = int(lengths_or_B)
    upp = int(np.prod(h["upsample_rates"]))
    c = (rng.standard_normal((B, h["num_mels"], T)) * 2.0 - 5.0).astype(np.float32)
    f0 = (220.0 * 2.0 ** (rng.uniform(-1.0, 1.5, size=(B, 1)) + 0.3 * np.sin(np.arange(T)[None, :] / 9.0 + rng.uniform(0, 6, size=(B, 1))))).astype(np.float32)
    voiced = (np.sin(np.arange(T)[None, :] / 5.0 + rng.uniform(0, 6, size=(B, 1))) > -0.4)
    f0 = np.where(voiced, f0, 0.0).astype(np.float32)
    rand_ini = rng.uniform(0, 1, size=(1, 9)).astype(np.float32)
    rand_ini[:, 0] = 0
    noise = rng.standard_normal((B, T * upp, 9)).astype(np.float32)
    return torch.from_numpy(c), torch.from_numpy(f0), torch.from_numpy(rand_ini), torch.from_numpy(noise)

"""CPU fp32 oracle for the FastSpeech text encoder (token-level half).  TEST INFRASTRUCTURE ONLY.

Restates, op for op, what ``FastspeechEncoder.forward(txt_tokens, extra_embed)`` executes in eval mode:

    /root/reference/modules/fastspeech/tts_modules.py:291-330   FastspeechEncoder (embedding, positions)
    /root/reference/modules/fastspeech/tts_modules.py:262-288   FFTBlocks.forward (mask, layers, last LayerNorm)
    /root/reference/modules/commons/common_layers.py:625-673    EncSALayer (pre-LN self-attention + conv FFN)
    /root/reference/modules/commons/common_layers.py:242-287    MultiheadAttention.forward -> the
                                                                F.multi_head_attention_forward branch (need_weights=True,
                                                                bias=False, key_padding_mask only)
    /root/reference/modules/commons/common_layers.py:542-583    TransformerFFNLayer (Conv1d k, SAME; * k**-0.5; gelu; Linear)
    /root/reference/modules/commons/common_layers.py:94-149     SinusoidalPositionalEmbedding
    /root/reference/utils/tts_utils.py:6-18                     make_positions

as a plain function over a ``state`` dict with the reference's own ``state_dict`` keys.  Same rules as
``prodiff_oracle.py``: only ``tests/``, ``smoke()`` and bench CPU legs may import it.

Pinning: the reference has no tests or vectors for the encoder either, so this file is pinned against the live module
imported from /root/reference (``tests/test_oracle_vs_reference.py``: bit-identical) and ``oracle/make_golden.py`` ran
the unmodified reference class to produce ``tests/golden/encoder_*.npz``.

The attention arithmetic lives in PyTorch (``torch.nn.functional.multi_head_attention_forward``, torch 2.11.0 here; the
reference pins no torch version): packed in-projection ``linear(x, in_proj_weight)``, ``q * sqrt(1 / head_dim)``,
``baddbmm(-inf key-padding mask, q, k^T)``, ``softmax``, ``bmm`` with v, ``out_proj``.  That sequence is restated below.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]
PAD_ID = 0


def num_layers(state: State) -> int:
    n = 0
    while f"layers.{n}.op.layer_norm1.weight" in state:
        n += 1
    return n


def sinusoid_table(num_embeddings: int, dim: int, padding_idx: Optional[int] = PAD_ID) -> torch.Tensor:
    """common_layers.py:111-128."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, dtype=torch.float) * -e)
    e = torch.arange(num_embeddings, dtype=torch.float).unsqueeze(1) * e.unsqueeze(0)
    e = torch.cat([torch.sin(e), torch.cos(e)], dim=1).view(num_embeddings, -1)
    if dim % 2 == 1:
        e = torch.cat([e, torch.zeros(num_embeddings, 1)], dim=1)
    if padding_idx is not None:
        e[padding_idx, :] = 0
    return e


def make_positions(nonpad: torch.Tensor, padding_idx: int = PAD_ID) -> torch.Tensor:
    """tts_utils.py:6-18 applied to ``~padding_mask`` (tts_modules.py:322): 1, 2, ... over non-padding tokens, 0 on padding."""
    mask = nonpad.ne(padding_idx).int()
    return (torch.cumsum(mask, dim=1).type_as(mask) * mask).long() + padding_idx


def self_attention(x_tbc: torch.Tensor, in_proj_weight: torch.Tensor, out_proj_weight: torch.Tensor,
                   key_padding_mask: torch.Tensor, num_heads: int) -> torch.Tensor:
    """MultiheadAttention(c, heads, self_attention=True, bias=False) on [T,B,C] with a [B,T] bool key padding mask."""
    tgt_len, bsz, embed_dim = x_tbc.shape
    head_dim = embed_dim // num_heads
    proj = F.linear(x_tbc, in_proj_weight)
    proj = proj.unflatten(-1, (3, embed_dim)).unsqueeze(0).transpose(0, -2).squeeze(-2).contiguous()
    q, k, v = proj[0], proj[1], proj[2]
    q = q.view(tgt_len, bsz * num_heads, head_dim).transpose(0, 1)
    k = k.view(tgt_len, bsz * num_heads, head_dim).transpose(0, 1)
    v = v.view(tgt_len, bsz * num_heads, head_dim).transpose(0, 1)
    mask = torch.zeros(key_padding_mask.shape, dtype=x_tbc.dtype, device=x_tbc.device).masked_fill_(key_padding_mask, float("-inf"))
    mask = mask.view(bsz, 1, 1, tgt_len).expand(-1, num_heads, -1, -1).reshape(bsz * num_heads, 1, tgt_len)
    q_scaled = q * math.sqrt(1.0 / float(head_dim))
    w = torch.baddbmm(mask, q_scaled, k.transpose(-2, -1))
    w = F.softmax(w, dim=-1)
    o = torch.bmm(w, v)
    o = o.transpose(0, 1).contiguous().view(tgt_len * bsz, embed_dim)
    o = F.linear(o, out_proj_weight, None)
    return o.view(tgt_len, bsz, embed_dim)


def enc_sa_layer(state: State, i: int, x: torch.Tensor, padding_mask: torch.Tensor, num_heads: int,
                 trace: Optional[dict] = None) -> torch.Tensor:
    """EncSALayer.forward in eval mode (common_layers.py:649-673), x [T,B,C]."""
    p = f"layers.{i}.op."
    C = x.shape[-1]
    keep = (1 - padding_mask.float()).transpose(0, 1)[..., None]
    residual = x
    x = F.layer_norm(x, (C,), state[p + "layer_norm1.weight"], state[p + "layer_norm1.bias"], 1e-5)
    x = self_attention(x, state[p + "self_attn.in_proj_weight"], state[p + "self_attn.out_proj.weight"], padding_mask, num_heads)
    x = residual + x
    x = x * keep
    if trace is not None:
        trace[f"attn{i}"] = x.transpose(0, 1).clone()
    residual = x
    x = F.layer_norm(x, (C,), state[p + "layer_norm2.weight"], state[p + "layer_norm2.bias"], 1e-5)
    w1 = state[p + "ffn.ffn_1.weight"]
    ks = w1.shape[2]
    x = F.conv1d(x.permute(1, 2, 0), w1, state[p + "ffn.ffn_1.bias"], padding=ks // 2).permute(2, 0, 1)
    x = x * ks ** -0.5
    x = F.gelu(x)
    x = F.linear(x, state[p + "ffn.ffn_2.weight"], state[p + "ffn.ffn_2.bias"])
    x = residual + x
    x = x * keep
    return x


def encoder_forward(state: State, txt_tokens: torch.Tensor, extra_embed: Optional[torch.Tensor] = None, num_heads: int = 2,
                    trace: Optional[dict] = None) -> torch.Tensor:
    """FastspeechEncoder.forward (tts_modules.py:305-312): txt_tokens [B,T] int64, extra_embed [B,T,C] -> [B,T,C]."""
    W = state["embed_tokens.weight"]
    C = W.shape[1]
    padding_mask = txt_tokens.eq(PAD_ID)
    x = math.sqrt(C) * F.embedding(txt_tokens, W, PAD_ID)                       # :316
    if extra_embed is not None:
        x = x + extra_embed
    pos = make_positions(~padding_mask)
    table = sinusoid_table(max(2000, PAD_ID + 1 + txt_tokens.shape[1]), C)       # init_size = DEFAULT_MAX_TARGET_POSITIONS
    x = x + table.index_select(0, pos.view(-1)).view(txt_tokens.shape[0], txt_tokens.shape[1], -1)
    keep_tb = 1 - padding_mask.transpose(0, 1).float()[:, :, None]               # FFTBlocks.forward :269
    x = x.transpose(0, 1) * keep_tb
    if trace is not None:
        trace["embed"] = x.transpose(0, 1).clone()
    for i in range(num_layers(state)):
        x = enc_sa_layer(state, i, x, padding_mask, num_heads, trace) * keep_tb
        if trace is not None:
            trace[f"layer{i}"] = x.transpose(0, 1).clone()
    x = F.layer_norm(x, (C,), state["layer_norm.weight"], state["layer_norm.bias"], 1e-5) * keep_tb
    return x.transpose(0, 1)


def mel2ph_to_dur(mel2ph: torch.Tensor, T_txt: int) -> torch.Tensor:
    """tts_modules.py:223-229 (max_dur=None)."""
    B = mel2ph.shape[0]
    dur = mel2ph.new_zeros(B, T_txt + 1).scatter_add(1, mel2ph, torch.ones_like(mel2ph))
    return dur[:, 1:]

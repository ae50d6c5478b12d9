"""Drop-in for the reference text encoder ``FastspeechEncoder``
(`/root/reference/modules/fastspeech/tts_modules.py:291-330`; SURVEY §8f N1, token-level half).

Same constructor, same ``forward(txt_tokens, extra_embed=None)``, same ``state_dict`` keys and initialisation, so
``ProDiffTeacher.__init__`` (`prodiff_teacher.py:14-22`) and ``load_ckpt`` work unchanged.  The sub-modules are parameter
containers only: ``forward`` hands the tensors to the C-ABI library (``pd_encoder_forward``), which runs the embedding,
the four pre-LN self-attention + conv-FFN blocks and the last LayerNorm as sm_100a kernels (tcgen05 GEMMs for the four
dense contractions of each block, fp32 LayerNorm / softmax / residual stream).  Inference only (dropout is the identity in
eval mode, which is all the sampling path uses); no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List

import torch
import torch.nn as nn

from . import _lib

PAD_ID = 0


class _Positions(nn.Module):
    """Keeps the reference's one-element ``embed_positions._float_tensor`` buffer (a device marker) in the state_dict."""

    def __init__(self):
        super().__init__()
        self.register_buffer("_float_tensor", torch.zeros(1))   # (the reference leaves it uninitialised: torch.FloatTensor(1))


class _SelfAttention(nn.Module):
    """Parameters of MultiheadAttention(c, heads, self_attention=True, bias=False) (common_layers.py:172-233)."""

    def __init__(self, c):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * c, c))
        self.out_proj = nn.Linear(c, c, bias=False)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj.weight)


class _FFN(nn.Module):
    """Parameters of TransformerFFNLayer(c, 4c, kernel_size, padding='SAME') (common_layers.py:542-555)."""

    def __init__(self, c, kernel_size):
        super().__init__()
        self.ffn_1 = nn.Conv1d(c, 4 * c, kernel_size, padding=kernel_size // 2)
        self.ffn_2 = nn.Linear(4 * c, c)
        nn.init.xavier_uniform_(self.ffn_2.weight)
        nn.init.constant_(self.ffn_2.bias, 0.)


class _EncSALayer(nn.Module):
    def __init__(self, c, kernel_size):
        super().__init__()
        self.layer_norm1 = nn.LayerNorm(c)
        self.self_attn = _SelfAttention(c)
        self.layer_norm2 = nn.LayerNorm(c)
        self.ffn = _FFN(c, kernel_size)


class _Layer(nn.Module):
    def __init__(self, c, kernel_size):
        super().__init__()
        self.op = _EncSALayer(c, kernel_size)


def encoder_state_key_order(num_layers: int) -> List[str]:
    """The order pd_encoder_create expects (include/prodiff_b200.h)."""
    keys = ["embed_tokens.weight", "layer_norm.weight", "layer_norm.bias"]
    for i in range(num_layers):
        p = f"layers.{i}.op."
        keys += [p + "layer_norm1.weight", p + "layer_norm1.bias", p + "self_attn.in_proj_weight", p + "self_attn.out_proj.weight",
                 p + "layer_norm2.weight", p + "layer_norm2.bias", p + "ffn.ffn_1.weight", p + "ffn.ffn_1.bias",
                 p + "ffn.ffn_2.weight", p + "ffn.ffn_2.bias"]
    return keys


class _EncoderEngine:
    """Packed weights + handle of one encoder on one device (rebuilt when the parameters change)."""

    def __init__(self, cfg: dict, state: dict, device: torch.device, operand_dtype: str):
        if device.type != "cuda":
            raise RuntimeError("prodiff_b200.FastspeechEncoder runs on a CUDA (sm_100a) device only -- there is no CPU fallback")
        self.lib = _lib.load()
        self.device = device
        self.cfg = _lib.pd_encoder_config(cfg["vocab_size"], cfg["hidden_size"], cfg["num_layers"], cfg["kernel_size"],
                                          cfg["num_heads"], 1 if operand_dtype == "fp16" else 0)
        keys = encoder_state_key_order(cfg["num_layers"])
        with torch.cuda.device(device):
            tensors = [state[k].detach().to(device=device, dtype=torch.float32).contiguous() for k in keys]
            nbytes = self.lib.pd_encoder_packed_bytes(C.byref(self.cfg))
            if nbytes == 0:
                raise RuntimeError(f"prodiff_b200: {self.lib.pd_last_error().decode()}")
            self.packed = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            base = (self.packed.data_ptr() + 1023) // 1024 * 1024
            ptrs = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
            handle = C.c_void_p()
            _lib.check(self.lib.pd_encoder_create(C.byref(self.cfg), ptrs, len(tensors), base, nbytes,
                                                  torch.cuda.current_stream(device).cuda_stream, C.byref(handle)),
                       "pd_encoder_create")
            torch.cuda.current_stream(device).synchronize()   # the fp32 sources may be temporaries
        self.handle = handle

    def forward(self, tok: torch.Tensor, extra) -> torch.Tensor:
        B, T = tok.shape
        H = self.cfg.hidden_size
        with torch.cuda.device(self.device):
            out = torch.empty(B, T, H, dtype=torch.float32, device=self.device)
            nbytes = self.lib.pd_encoder_workspace_bytes(self.handle, B, T)
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)   # stream-ordered reuse by the allocator
            base = (ws.data_ptr() + 1023) // 1024 * 1024
            _lib.check(self.lib.pd_encoder_forward(self.handle, tok.data_ptr(), extra.data_ptr() if extra is not None else None,
                                                   out.data_ptr(), B, T, base, nbytes,
                                                   torch.cuda.current_stream(self.device).cuda_stream), "pd_encoder_forward")
        return out

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.pd_encoder_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class FastspeechEncoder(nn.Module):
    def __init__(self, vocab_size, hidden_size, num_layers, kernel_size, dropout=0.1, num_heads=2, rel_pos=False,
                 operand_dtype: str = "bf16"):
        super().__init__()
        if rel_pos:
            raise NotImplementedError("prodiff_b200.FastspeechEncoder: rel_pos=True (RelPositionalEncoding) is not supported "
                                      "(the configured encoder uses rel_pos: false, base_config.yaml:116)")
        if operand_dtype not in ("bf16", "fp16"):
            raise ValueError("operand_dtype must be 'bf16' or 'fp16'")
        self.cfg = dict(vocab_size=int(vocab_size), hidden_size=int(hidden_size), num_layers=int(num_layers),
                        kernel_size=int(kernel_size), num_heads=int(num_heads))
        self.operand_dtype = operand_dtype
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.dropout = dropout
        self.embed_scale = math.sqrt(hidden_size)
        self.padding_idx = PAD_ID
        self.embed_tokens = nn.Embedding(vocab_size, hidden_size, padding_idx=PAD_ID)      # common_layers.py:63-68
        nn.init.normal_(self.embed_tokens.weight, mean=0, std=hidden_size ** -0.5)
        nn.init.constant_(self.embed_tokens.weight[PAD_ID], 0)
        self.embed_positions = _Positions()
        self.layers = nn.ModuleList([_Layer(hidden_size, kernel_size) for _ in range(num_layers)])
        self.layer_norm = nn.LayerNorm(hidden_size)
        self._engine = None
        self._engine_key = None

    @classmethod
    def from_reference(cls, ref_module: nn.Module, operand_dtype: str = "bf16") -> "FastspeechEncoder":
        """Build from an instance of the reference's own ``FastspeechEncoder`` (shapes are read off its state)."""
        sd = ref_module.state_dict()
        V, H = sd["embed_tokens.weight"].shape
        L = len(ref_module.layers)
        k = sd["layers.0.op.ffn.ffn_1.weight"].shape[2]
        heads = ref_module.layers[0].op.self_attn.num_heads
        new = cls(V, H, L, k, num_heads=heads, operand_dtype=operand_dtype)
        new.load_state_dict(sd)
        return new.to(sd["embed_tokens.weight"].device)

    def _state_signature(self):
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in self.parameters())

    def engine(self) -> _EncoderEngine:
        key = self._state_signature()
        if self._engine is None or key != self._engine_key:
            self._engine = _EncoderEngine(self.cfg, dict(self.state_dict()), self.embed_tokens.weight.device, self.operand_dtype)
            self._engine_key = key
        return self._engine

    def forward(self, txt_tokens, extra_embed=None):
        """
        :param txt_tokens: [B, T] int64, 0 = padding
        :param extra_embed: optional [B, T, H] fp32
        :return: [B, T, H] fp32, zeros at padding tokens
        """
        if self.training:
            raise RuntimeError("prodiff_b200.FastspeechEncoder is inference-only: call .eval() (dropout is not implemented)")
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                        (extra_embed is not None and extra_embed.requires_grad)):
            raise RuntimeError("prodiff_b200.FastspeechEncoder is inference-only; call it under torch.no_grad(), or train "
                               "with the reference module -- state_dict keys are identical")
        if txt_tokens.dim() != 2:
            raise ValueError(f"txt_tokens must be [B, T], got {tuple(txt_tokens.shape)}")
        dev = self.embed_tokens.weight.device
        tok = txt_tokens.detach().to(device=dev, dtype=torch.int64).contiguous()
        extra = None
        if extra_embed is not None:
            if tuple(extra_embed.shape) != (tok.shape[0], tok.shape[1], self.cfg["hidden_size"]):
                raise ValueError(f"extra_embed must be {(tok.shape[0], tok.shape[1], self.cfg['hidden_size'])}, got {tuple(extra_embed.shape)}")
            extra = extra_embed.detach().to(device=dev, dtype=torch.float32).contiguous()
        return self.engine().forward(tok, extra)


__all__ = ["FastspeechEncoder", "encoder_state_key_order"]

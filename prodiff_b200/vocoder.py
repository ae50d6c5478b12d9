"""Drop-in for the reference NSF-HiFiGAN generator: ``Generator``
(`/root/reference/modules/nsf_hifigan/models.py:209-297`) as called by ``NsfHifiGAN.spec2wav_torch``
(`/root/reference/component/vocoder/nsf_hifigan.py:49-58`).

Same constructor argument (``h``: the fields of the checkpoint's ``config.json``), same ``forward(x, f0)`` signature and
result shape ``[B, 1, T * upp]``, same parameter names as the reference has after ``remove_weight_norm()`` -- and the
weight-normed form a checkpoint stores (``*.weight_g`` / ``*.weight_v``, the keys ``load_model`` loads at models.py:30-33)
is folded on ``load_state_dict``, with ``remove_weight_norm()`` kept as a no-op so ``load_model`` runs unchanged.  The
parameters live in ordinary ``nn.Conv1d`` / ``nn.ConvTranspose1d`` containers that are never *called*: ``forward`` hands
the tensors to the C-ABI library (``pd_vocoder_*``), which runs every convolution as a tcgen05 implicit GEMM
(``csrc/pd_voc.cuh``) and the harmonic source as two small kernels.  No CPU or PyTorch fallback: off a B200 the call raises.

Batches: the reference's ``SineGen`` is written for one utterance per call (models.py:163,174-179); here a batch is the
row-wise application (one ``rand_ini`` shared by the rows), which is what batching the vocoder behind the batched sampler
needs.  Extra keyword-only arguments inject ``rand_ini`` / ``noise`` (the parity tests need identical randomness).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .engine import _aligned_bytes, _require_b200, workspace_bucket

LRELU_SLOPE = 0.1


def _get(h, key):
    return h[key] if isinstance(h, dict) else getattr(h, key)


def state_key_order(h) -> list:
    """The order pd_vocoder_create expects (include/prodiff_b200.h)."""
    n_up, n_k = len(_get(h, "upsample_rates")), len(_get(h, "resblock_kernel_sizes"))
    keys = ["conv_pre.weight", "conv_pre.bias"]
    for i in range(n_up):
        keys += [f"ups.{i}.weight", f"ups.{i}.bias", f"noise_convs.{i}.weight", f"noise_convs.{i}.bias"]
        for j in range(n_k):
            p = f"resblocks.{i * n_k + j}"
            for cv in ("convs1", "convs2"):
                for d in range(3):
                    keys += [f"{p}.{cv}.{d}.weight", f"{p}.{cv}.{d}.bias"]
    keys += ["conv_post.weight", "conv_post.bias", "m_source.l_linear.weight", "m_source.l_linear.bias"]
    return keys


def fold_weight_norm(state: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """``*.weight_g`` / ``*.weight_v`` pairs (torch.nn.utils.weight_norm, dim 0) -> ``*.weight`` = g * v / ||v||, the value
    ``remove_weight_norm`` leaves in the module (models.py:292-297)."""
    out = {}
    for k, v in state.items():
        if k.endswith(".weight_g"):
            base = k[:-len("_g")]
            wv = state[base + "_v"]
            norm = wv.reshape(wv.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (wv.dim() - 1)))
            out[base] = wv * (v / norm)
        elif k.endswith(".weight_v"):
            continue
        else:
            out[k] = v
    return out


class _ResBlock1(nn.Module):
    """Parameter container mirroring models.py:39-60."""

    def __init__(self, channels, kernel_size, dilation):
        super().__init__()
        self.convs1 = nn.ModuleList([nn.Conv1d(channels, channels, kernel_size, 1, dilation=d) for d in dilation])
        self.convs2 = nn.ModuleList([nn.Conv1d(channels, channels, kernel_size, 1) for _ in dilation])


class _Source(nn.Module):
    """Parameter container for SourceModuleHnNSF.l_linear (models.py:208)."""

    def __init__(self):
        super().__init__()
        self.l_linear = nn.Linear(9, 1)


class Generator(nn.Module):
    def __init__(self, h):
        super().__init__()
        self.h = h
        g = lambda k: _get(h, k)
        if str(g("resblock")) != "1":
            raise NotImplementedError("prodiff_b200.Generator implements resblock '1' (ResBlock1, models.py:39) only")
        self.num_kernels = len(g("resblock_kernel_sizes"))
        self.num_upsamples = len(g("upsample_rates"))
        # attribute order = the reference's registration order (models.py:215-262), so state_dict() lists the same keys in
        # the same order
        self.m_source = _Source()
        self.noise_convs = nn.ModuleList()
        c0 = g("upsample_initial_channel")
        self.conv_pre = nn.Conv1d(g("num_mels"), c0, 7, 1, padding=3)
        self.ups = nn.ModuleList()
        resblocks = []
        rates = list(g("upsample_rates"))
        ch = c0
        for i, (u, k) in enumerate(zip(rates, g("upsample_kernel_sizes"))):
            self.ups.append(nn.ConvTranspose1d(ch, ch // 2, k, u, padding=(k - u) // 2))
            ch //= 2
            if i + 1 < len(rates):
                stride_f0 = int(np.prod(rates[i + 1:]))
                self.noise_convs.append(nn.Conv1d(1, ch, kernel_size=stride_f0 * 2, stride=stride_f0, padding=stride_f0 // 2))
            else:
                self.noise_convs.append(nn.Conv1d(1, ch, kernel_size=1))
            for k2, d in zip(g("resblock_kernel_sizes"), g("resblock_dilation_sizes")):
                resblocks.append(_ResBlock1(ch, k2, d))
        self.resblocks = nn.ModuleList(resblocks)
        self.conv_post = nn.Conv1d(ch, 1, 7, 1, padding=3)
        for m in list(self.ups) + [self.conv_post]:
            m.weight.data.normal_(0.0, 0.01)             # init_weights (nsf_hifigan/utils.py:6-9), as models.py:264-265
        for rb in self.resblocks:
            for m in list(rb.convs1) + list(rb.convs2):
                m.weight.data.normal_(0.0, 0.01)
        self.upp = int(np.prod(rates))
        self._handle = None
        self._key = None
        self._packed = None
        self._ws = None

    # ------------------------------------------------------------------ reference-compatible loading
    def remove_weight_norm(self):
        """No-op: weight norm is folded when the state is loaded (models.py:292-297 on the reference)."""
        return None

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        return super().load_state_dict(fold_weight_norm(dict(state_dict)), strict=strict, **kw)

    @classmethod
    def from_reference(cls, ref_module: nn.Module) -> "Generator":
        """Adopt a built reference ``Generator`` (before or after its ``remove_weight_norm()``)."""
        new = cls(ref_module.h)
        new.load_state_dict(ref_module.state_dict())
        return new.to(ref_module.conv_post.bias.device).eval()

    # ------------------------------------------------------------------ packed-weight cache
    def _signature(self):
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in self.parameters())

    def _config(self) -> "_lib.pd_vocoder_config":
        g = lambda k: _get(self.h, k)
        cfg = _lib.pd_vocoder_config()
        cfg.num_mels = int(g("num_mels"))
        cfg.upsample_initial_channel = int(g("upsample_initial_channel"))
        rates, ks = list(g("upsample_rates")), list(g("upsample_kernel_sizes"))
        rk, rd = list(g("resblock_kernel_sizes")), list(g("resblock_dilation_sizes"))
        if len(rates) > 8 or len(rk) > 4 or any(len(d) != 3 for d in rd):
            raise ValueError("prodiff_b200.Generator: at most 8 up-sampling stages, 4 resblock kernels, 3 dilations each")
        cfg.n_upsamples = len(rates)
        for i, (u, k) in enumerate(zip(rates, ks)):
            cfg.upsample_rates[i], cfg.upsample_kernel_sizes[i] = int(u), int(k)
        cfg.n_resblock_kernels = len(rk)
        for j, (k, d) in enumerate(zip(rk, rd)):
            cfg.resblock_kernel_sizes[j] = int(k)
            for q in range(3):
                cfg.resblock_dilation_sizes[j][q] = int(d[q])
        cfg.resblock_type = 1
        cfg.sampling_rate = int(g("sampling_rate"))
        return cfg

    def _engine(self):
        key = self._signature()
        if self._handle is not None and key == self._key:
            return self._handle
        device = self.conv_post.bias.device
        _require_b200(device)
        lib = _lib.load()
        if self._handle is not None:
            lib.pd_vocoder_destroy(self._handle)
            self._handle = None
        cfg = self._config()
        sd = self.state_dict()
        with torch.cuda.device(device):
            tensors = [sd[k].detach().to(device=device, dtype=torch.float32).contiguous() for k in state_key_order(self.h)]
            nbytes = lib.pd_vocoder_packed_bytes(C.byref(cfg))
            if nbytes == 0:
                raise RuntimeError(f"prodiff_b200: unsupported generator config: {lib.pd_last_error().decode()}")
            self._packed = _aligned_bytes(nbytes, device)
            ptrs = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
            hnd = C.c_void_p()
            stream = torch.cuda.current_stream(device)
            _lib.check(lib.pd_vocoder_create(C.byref(cfg), ptrs, len(tensors), self._packed.data_ptr(), nbytes, stream.cuda_stream,
                                             C.byref(hnd)), "pd_vocoder_create")
            stream.synchronize()     # the pack kernels read `tensors` asynchronously
        self._handle, self._key = hnd, key
        return hnd

    def _workspace(self, nbytes: int, device) -> torch.Tensor:
        need = workspace_bucket(nbytes + 1024)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        off = (-self._ws.data_ptr()) % 1024
        return self._ws[off:off + nbytes]

    # ------------------------------------------------------------------ forward (models.py:269-290)
    def forward(self, x, f0, *, rand_ini: Optional[torch.Tensor] = None, noise: Optional[torch.Tensor] = None,
                seed: Optional[int] = None, lengths=None):
        """x [B, num_mels, T] fp32 (the hand-off tensor ``2.30259 * mel.transpose(2, 1)``), f0 [B, T] -> [B, 1, T * upp].
        ``lengths`` (B ints <= T, keyword-only): a padded batch -- samples [0, lengths[b] * upp) of row b are exactly what the
        call on that utterance alone gives (every activation past an utterance's own end is held at zero), the rest zeros."""
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("prodiff_b200.Generator is inference-only; call it under torch.no_grad()")
        B, M, T = x.shape
        if M != int(_get(self.h, "num_mels")) or tuple(f0.shape) != (B, T):
            raise ValueError(f"x must be [B,{_get(self.h, 'num_mels')},T] and f0 [B,T]; got {tuple(x.shape)}, {tuple(f0.shape)}")
        hnd = self._engine()
        lib = _lib.load()
        device = x.device
        x = x.detach().to(dtype=torch.float32).contiguous()
        f0 = f0.detach().to(device=device, dtype=torch.float32).contiguous()
        if rand_ini is None:
            rand_ini = torch.rand(1, 9, device=device)    # models.py:143-144
            rand_ini[:, 0] = 0
        rand_ini = rand_ini.detach().to(device=device, dtype=torch.float32).contiguous()
        if rand_ini.numel() != 9:
            raise ValueError("rand_ini must have 9 entries (fundamental + 8 overtones)")
        if noise is not None:
            noise = noise.detach().to(device=device, dtype=torch.float32).contiguous()
            if tuple(noise.shape) != (B, T * self.upp, 9):
                raise ValueError(f"noise must be {(B, T * self.upp, 9)}, got {tuple(noise.shape)}")
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        len_dev = None
        if lengths is not None:
            ln = [int(v) for v in lengths]
            if len(ln) != B or any(v < 0 or v > T for v in ln):
                raise ValueError(f"lengths must be {B} integers in [0, {T}]")
            len_dev = torch.tensor(ln, dtype=torch.int32, device=device)
        wave = torch.empty(B, 1, T * self.upp, dtype=torch.float32, device=device)
        with torch.cuda.device(device):
            nbytes = lib.pd_vocoder_workspace_bytes(hnd, B, T)
            ws = self._workspace(nbytes, device)
            _lib.check(lib.pd_vocoder_forward(hnd, x.data_ptr(), f0.data_ptr(), rand_ini.data_ptr(),
                                              noise.data_ptr() if noise is not None else None, C.c_uint64(seed),
                                              len_dev.data_ptr() if len_dev is not None else None, wave.data_ptr(), B, T,
                                              ws.data_ptr(), nbytes, torch.cuda.current_stream(device).cuda_stream), "pd_vocoder_forward")
        return wave

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.load().pd_vocoder_destroy(self._handle)
                self._handle = None
        except Exception:
            pass


__all__ = ["Generator", "state_key_order", "fold_weight_norm"]

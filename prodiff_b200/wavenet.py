"""Drop-in for the reference DiffNet denoiser ``WaveNet`` (`/root/reference/modules/decoder/wavenet.py:74-123`).

Same constructor, same ``forward(spec, diffusion_step, cond)`` signature, same parameter names/shapes (so
``utils/ckpt_utils.load_ckpt`` and ``load_state_dict`` work unchanged), same initialisation.  The parameters live in
ordinary ``nn.Conv1d`` / ``nn.Linear`` containers, but those containers are never *called*: ``forward`` hands the
tensors to the C-ABI library, which runs the whole network as hand-written sm_100a kernels.  No CPU or PyTorch
fallback exists -- off a B200 the call raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .engine import Engine, state_key_order


def _kaiming_conv1d(*args, **kwargs):
    layer = nn.Conv1d(*args, **kwargs)  # parameter container; init as wavenet.py:41-44
    nn.init.kaiming_normal_(layer.weight)
    return layer


class Mish(nn.Module):
    """Placeholder that keeps ``mlp`` indices 0/2 (and so the state_dict keys) identical to the reference."""

    def forward(self, x):  # pragma: no cover - never called on the product path
        raise RuntimeError("prodiff_b200: the step MLP runs inside the CUDA library")


class ResidualBlock(nn.Module):
    """Parameter container mirroring wavenet.py:52-58."""

    def __init__(self, encoder_hidden, residual_channels, dilation):
        super().__init__()
        self.dilation = dilation
        self.dilated_conv = _kaiming_conv1d(residual_channels, 2 * residual_channels, 3, padding=dilation, dilation=dilation)
        self.diffusion_projection = nn.Linear(residual_channels, residual_channels)
        self.conditioner_projection = _kaiming_conv1d(encoder_hidden, 2 * residual_channels, 1)
        self.output_projection = _kaiming_conv1d(residual_channels, 2 * residual_channels, 1)


class WaveNet(nn.Module):
    def __init__(self, in_dims, hidden_size, residual_layers, residual_channels, dilation_cycle_length,
                 operand_dtype: str = "bf16"):
        super().__init__()
        if operand_dtype not in ("bf16", "fp16"):
            raise ValueError("operand_dtype must be 'bf16' or 'fp16'")
        self.cfg = dict(in_dims=int(in_dims), hidden_size=int(hidden_size), residual_layers=int(residual_layers),
                        residual_channels=int(residual_channels), dilation_cycle_length=int(dilation_cycle_length))
        self.operand_dtype = operand_dtype
        C = residual_channels
        self.input_projection = _kaiming_conv1d(in_dims, C, 1)
        self.mlp = nn.Sequential(nn.Linear(C, C * 4), Mish(), nn.Linear(C * 4, C))
        self.residual_layers = nn.ModuleList([
            ResidualBlock(hidden_size, C, 2 ** (i % dilation_cycle_length)) for i in range(residual_layers)])
        self.skip_projection = _kaiming_conv1d(C, C, 1)
        self.output_projection = _kaiming_conv1d(C, in_dims, 1)
        nn.init.zeros_(self.output_projection.weight)  # wavenet.py:98
        self._engine = None
        self._engine_key = None

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_reference(cls, ref_module: nn.Module, operand_dtype: str = "bf16") -> "WaveNet":
        """Build from an instance of the reference's own ``WaveNet`` (shapes and dilations are read off it)."""
        sd = ref_module.state_dict()
        C, M = sd["input_projection.weight"].shape[:2]
        H = sd["residual_layers.0.conditioner_projection.weight"].shape[1]
        L = len(ref_module.residual_layers)
        dils = [int(b.dilated_conv.dilation[0]) for b in ref_module.residual_layers]
        cycle = next((i for i in range(1, L) if dils[i] == 1), L)
        assert all(d == 2 ** (i % cycle) for i, d in enumerate(dils)), "unexpected dilation pattern"
        new = cls(M, H, L, C, cycle, operand_dtype=operand_dtype)
        new.load_state_dict(sd)
        return new.to(sd["input_projection.weight"].device)

    # ------------------------------------------------------------------ packed-weight cache
    def _state_signature(self):
        # invalidate on load_state_dict / .to() / optimizer steps: data_ptr and the in-place version counter
        return tuple((p.data_ptr(), p._version, str(p.device)) for p in self.parameters())

    def engine(self) -> Engine:
        key = self._state_signature()
        if self._engine is None or key != self._engine_key:
            device = self.input_projection.weight.device
            sd = {k: v for k, v in self.state_dict().items()}
            self._engine = Engine(self.cfg, sd, device, self.operand_dtype)
            self._engine_key = key
        return self._engine

    # ------------------------------------------------------------------ forward
    def forward(self, spec, diffusion_step, cond):
        """
        :param spec: [B, 1, M, T] fp32
        :param diffusion_step: [B] (int64 from the sampler, float from rectified flow)
        :param cond: [B, H, T] fp32 (usually the transposed view of a [B, T, H] tensor)
        :return: [B, 1, M, T] fp32
        """
        if torch.is_grad_enabled() and (spec.requires_grad or cond.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            raise RuntimeError("prodiff_b200.WaveNet is inference-only (training is outside the accelerated path); "
                               "call it under torch.no_grad(), or train with the reference module -- state_dict keys "
                               "are identical")
        B, one, M, T = spec.shape
        if one != 1 or M != self.cfg["in_dims"]:
            raise ValueError(f"spec must be [B,1,{self.cfg['in_dims']},T], got {tuple(spec.shape)}")
        if cond.shape != (B, self.cfg["hidden_size"], T):
            raise ValueError(f"cond must be [B,{self.cfg['hidden_size']},T]={B, self.cfg['hidden_size'], T}, got {tuple(cond.shape)}")
        eng = self.engine()
        spec = spec.detach().to(dtype=torch.float32).contiguous()
        cond_bth = cond.detach().to(dtype=torch.float32).transpose(1, 2).contiguous()  # no copy for the sampler's view
        step = diffusion_step.detach().reshape(B).to(device=spec.device, dtype=torch.float32).contiguous()
        return eng.denoise(spec, step, cond_bth)


__all__ = ["WaveNet", "ResidualBlock", "Mish", "state_key_order"]

"""Drop-ins for the reference rectified-flow samplers ``RectifiedFlow`` / ``PitchRectifiedFlow``
(`/root/reference/modules/diffusion/reflow.py:5-143`) over the same DiffNet denoiser.

Constructor arguments, buffers and method names are the reference's.  ``forward(infer=True)`` hands the whole ODE loop
(euler / rk2 / rk4 / rk5, reflow.py:48-101) and ``denorm_spec`` to the C-ABI library (``pd_reflow_sample``): condition
projection once, one denoiser evaluation per stage, stage combinations and the final update as sm_100a kernels.
Keyword-only extras: ``x_start`` replaces the ``torch.randn`` start (reflow.py:88) so two implementations can be compared
on identical randomness; ``v_trace`` receives every velocity in call order.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .wavenet import WaveNet


class RectifiedFlow(nn.Module):
    def __init__(self, out_dims, denoise_fn, time_scale=1000, num_features=1, sampling_algorithm="euler",
                 spec_min=None, spec_max=None):
        super().__init__()
        if not isinstance(denoise_fn, WaveNet):
            if hasattr(denoise_fn, "residual_layers") and hasattr(denoise_fn, "input_projection"):
                denoise_fn = WaveNet.from_reference(denoise_fn)
            else:
                raise TypeError("denoise_fn must be a prodiff_b200.WaveNet (or the reference WaveNet)")
        if num_features != 1:
            # the DiffNet denoiser itself only takes spec[:, 0] (wavenet.py:108); the reference never builds F > 1 with it
            raise ValueError("prodiff_b200.RectifiedFlow supports num_features=1 (the only value the reference uses)")
        self.velocity_fn = denoise_fn
        self.out_dims = out_dims
        self.num_features = num_features
        self.sampling_algorithm = sampling_algorithm
        self.t_start = 0.
        self.time_scale = time_scale
        spec_min = torch.FloatTensor(spec_min)[None, None, :out_dims].transpose(-3, -2)
        spec_max = torch.FloatTensor(spec_max)[None, None, :out_dims].transpose(-3, -2)
        self.register_buffer("spec_min", spec_min, persistent=False)
        self.register_buffer("spec_max", spec_max, persistent=False)

    # ------------------------------------------------------------------ training-side helpers (reflow.py:25-44)
    def p_losses(self, x_end, t, cond):
        x_start = torch.randn_like(x_end)
        x_t = x_start + t[:, None, None, None] * (x_end - x_start)
        v_pred = self.velocity_fn(x_t, t * self.time_scale, cond)   # the CUDA denoiser refuses autograd (see WaveNet)
        return v_pred, x_end - x_start

    def forward(self, cond, gt_spec=None, infer_step=20, infer=True, *, x_start: Optional[torch.Tensor] = None,
                v_trace: Optional[torch.Tensor] = None, lengths=None, vocoder_scale: Optional[float] = None):
        """reflow.py:32-46.  ``vocoder_scale``: as in GaussianDiffusion.forward -- the denorm_spec kernel also applies the
        vocoder's log10 -> ln factor and the contiguous [B,M,T] generator input is returned."""
        b, device = cond.shape[0], cond.device
        if not infer:
            spec = self.norm_spec(gt_spec).transpose(-2, -1)
            if self.num_features == 1:
                spec = spec[:, None, :, :]
            t = self.t_start + (1.0 - self.t_start) * torch.rand((b,), device=device)
            v_pred, v_gt = self.p_losses(spec, t, cond=cond.transpose(1, 2))
            return v_pred, v_gt, t
        # denorm_spec (reflow.py:106-107) runs as the last kernel of the sequence, on the state, before the view is taken
        x = self._run(cond.transpose(1, 2) if isinstance(cond, torch.Tensor) else cond, b, infer_step, device, x_start, v_trace,
                      denorm=True, lengths=lengths, out_scale=1.0 if vocoder_scale is None else float(vocoder_scale))
        if vocoder_scale is not None:
            return x.transpose(1, 2)   # back to the physical [B,M,T] layout (a contiguous view)
        return self._finish(x)

    @torch.no_grad()
    def inference(self, cond, b=1, infer_step=20, device=None, *, x_start=None, v_trace=None):
        """reflow.py:86-101: cond [B,H,T]; returns the normalised [B,T,M] result (no denorm_spec), like the reference."""
        return self._run(cond, b, infer_step, device, x_start, v_trace, denorm=False)

    @torch.no_grad()
    def _run(self, cond, b, infer_step, device, x_start, v_trace, denorm, lengths=None, out_scale=1.0):
        dense = isinstance(cond, torch.Tensor)   # else a conditioning.AssembledCondition ([B,T,H] inputs, assembled by the kernels)
        device = cond.device if device is None else device
        T = cond.shape[2] if dense else cond.shape[1]
        if x_start is None:
            x = torch.randn(b, self.num_features, self.out_dims, T, device=device)
        else:
            x = x_start.detach().to(device=device, dtype=torch.float32).clone().contiguous()
            if x.shape != (b, 1, self.out_dims, T):
                raise ValueError(f"x_start must be {(b, 1, self.out_dims, T)}, got {tuple(x.shape)}")
        n = max(1, int(infer_step))
        stages = self.velocity_fn.engine().REFLOW_ALGORITHMS.get(self.sampling_algorithm, (0, 1))[1]
        if v_trace is not None and (v_trace.shape != (n * stages, b, 1, self.out_dims, T) or not v_trace.is_contiguous()
                                    or v_trace.dtype != torch.float32 or v_trace.device != x.device):
            raise ValueError(f"v_trace must be a contiguous fp32 {(n * stages, b, 1, self.out_dims, T)} tensor on cond's device")
        cond_bth = cond.detach().transpose(1, 2).to(dtype=torch.float32).contiguous() if dense else cond
        smin = self.spec_min.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        smax = self.spec_max.detach().to(device=device, dtype=torch.float32).reshape(-1).contiguous()
        if not denorm:
            smin = smax = None
        self.velocity_fn.engine().reflow(cond_bth, x, self.sampling_algorithm, n, float(self.time_scale), smin, smax, v_trace,
                                         lengths=lengths, out_scale=out_scale)
        return x.transpose(2, 3).squeeze(1)

    def _finish(self, x):
        return x

    def norm_spec(self, x):
        return (x - self.spec_min) / (self.spec_max - self.spec_min) * 2 - 1

    def denorm_spec(self, x):
        return (x + 1) / 2 * (self.spec_max - self.spec_min) + self.spec_min


class PitchRectifiedFlow(RectifiedFlow):
    """reflow.py:110-143: scalar pitch repeated over ``repeat_bins`` bins; denorm = base denorm, mean over bins, clamp."""

    def __init__(self, repeat_bins, denoise_fn, time_scale=1000, sampling_algorithm="euler", spec_min=-8.0, spec_max=8.0,
                 clamp_min=-12.0, clamp_max=12.0):
        self.clamp_min = clamp_min
        self.clamp_max = clamp_max
        self.repeat_bins = repeat_bins
        super().__init__(out_dims=repeat_bins, denoise_fn=denoise_fn, time_scale=time_scale, num_features=1,
                         sampling_algorithm=sampling_algorithm, spec_min=[spec_min], spec_max=[spec_max])

    def norm_spec(self, x):
        x = x.clamp(self.clamp_min, self.clamp_max)
        return super().norm_spec(x.unsqueeze(-1).repeat([1, 1, self.repeat_bins]))

    def _finish(self, x):
        # the base denorm already ran inside pd_reflow_sample; what remains of reflow.py:143 is the bin mean and the clamp
        return x.mean(dim=-1).clamp(self.clamp_min, self.clamp_max)

    def denorm_spec(self, x):
        return super().denorm_spec(x).mean(dim=-1).clamp(self.clamp_min, self.clamp_max)


__all__ = ["RectifiedFlow", "PitchRectifiedFlow"]

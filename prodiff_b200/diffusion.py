"""Drop-in for the reference sampler ``GaussianDiffusion`` (`/root/reference/modules/diffusion/prodiff.py:48-159`).

Constructor, persistent buffers (names, order, values) and method signatures are the reference's.  The inference
branch of ``forward`` hands the whole K-step reverse loop to the C-ABI library (``pd_sample``): condition projection
once, K denoiser evaluations and K fused posterior updates, all sm_100a kernels.  Extra keyword-only arguments allow
injecting ``x_T`` / per-step noise (the parity tests need identical randomness), tracing per-step x0 predictions, and
``lengths`` (per-utterance frame counts of a padded batch: each row is then computed exactly as that utterance alone).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
from torch import nn

from . import schedule
from .wavenet import WaveNet


def extract(a, t, x_shape):
    b = t.shape[0]
    return a.gather(-1, t).reshape(b, *((1,) * (len(x_shape) - 1)))


class GaussianDiffusion(nn.Module):
    def __init__(self, out_dims, denoise_fn, timesteps=1000, time_scale=1, betas=None, schedule_type="vpsde",
                 max_beta=0.02, spec_min=None, spec_max=None):
        super().__init__()
        if not isinstance(denoise_fn, WaveNet):
            # an instance of the reference's own WaveNet (same attribute layout) is adopted, anything else is an error
            if hasattr(denoise_fn, "residual_layers") and hasattr(denoise_fn, "input_projection"):
                denoise_fn = WaveNet.from_reference(denoise_fn)
            else:
                raise TypeError("denoise_fn must be a prodiff_b200.WaveNet (or the reference WaveNet)")
        self.denoise_fn = denoise_fn
        self.mel_bins = out_dims

        if betas is not None:
            betas = betas.detach().cpu().numpy() if isinstance(betas, torch.Tensor) else betas
        else:
            betas = schedule.get_noise_schedule_list(schedule_mode=schedule_type, timesteps=timesteps + 1,
                                                     min_beta=0.1, max_beta=max_beta, s=0.008)
        self.time_scale = time_scale
        self.num_timesteps = int(timesteps)
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)
        self.register_buffer("timesteps", f32(self.num_timesteps))
        self.register_buffer("timescale", f32(self.time_scale))
        for name, table in schedule.make_tables(betas).items():
            self.register_buffer(name, f32(table))
        spec_min = torch.FloatTensor(spec_min)[None, None, :out_dims].transpose(-3, -2)
        spec_max = torch.FloatTensor(spec_max)[None, None, :out_dims].transpose(-3, -2)
        self.register_buffer("spec_min", spec_min, persistent=False)
        self.register_buffer("spec_max", spec_max, persistent=False)
        self._host_tables = None

    # ------------------------------------------------------------------ small public helpers (prodiff.py:106-133)
    def q_posterior(self, x_0, x_t, t):
        mean = extract(self.posterior_mean_coef1, t, x_t.shape) * x_0 + extract(self.posterior_mean_coef2, t, x_t.shape) * x_t
        return mean, extract(self.posterior_variance, t, x_t.shape), extract(self.posterior_log_variance_clipped, t, x_t.shape)

    def q_posterior_sample(self, x_0, x_t, t, noise: Optional[torch.Tensor] = None):
        b = x_0.shape[0]
        mean, _, log_var = self.q_posterior(x_0=x_0, x_t=x_t, t=t)
        if noise is None:
            noise = torch.randn(x_0.shape, device=x_0.device)
        nonzero_mask = (1 - (t == 0).float()).reshape(b, *((1,) * (len(x_0.shape) - 1)))
        return mean + nonzero_mask * (0.5 * log_var).exp() * noise

    @torch.no_grad()
    def p_sample(self, x_t, t, cond, noise: Optional[torch.Tensor] = None):
        """One reverse step with per-row ``t`` (the batched loop in ``forward`` does not go through here)."""
        x_0_pred = self.denoise_fn(x_t, t, cond)
        return self.q_posterior_sample(x_0=x_0_pred, x_t=x_t, t=t, noise=noise)

    def q_sample(self, x_0, t, x_T=None):
        if x_T is None:
            x_T = torch.randn_like(x_0)
        return (extract(self.sqrt_alphas_cumprod, t, x_0.shape) * x_0 +
                extract(self.sqrt_one_minus_alphas_cumprod, t, x_0.shape) * x_T)

    def norm_spec(self, x):
        return x

    def denorm_spec(self, x):
        return x

    # ------------------------------------------------------------------ forward (prodiff.py:136-153)
    def _tables_on_host(self):
        # three 1-D tables, fetched once (a device->host copy per call would serialise the stream)
        key = (self.posterior_mean_coef1.data_ptr(), self.posterior_mean_coef1._version)
        if self._host_tables is None or self._host_tables[0] != key:
            self._host_tables = (key, self.posterior_mean_coef1.detach().cpu().tolist(),
                                 self.posterior_mean_coef2.detach().cpu().tolist(),
                                 self.posterior_log_variance_clipped.detach().cpu().tolist())
        return self._host_tables[1:]

    def forward(self, cond, src_spec=None, gt_spec=None, infer_step=4, infer=False, *, x_T=None, noise=None,
                seed: Optional[int] = None, x0_trace: Optional[torch.Tensor] = None, lengths=None,
                vocoder_scale: Optional[float] = None):
        """prodiff.py:136-153.  ``vocoder_scale`` (keyword-only, inference): return the generator input of
        ``NsfHifiGAN.spec2wav_torch`` (nsf_hifigan.py:50-52) instead of the mel -- a contiguous [B,M,T] tensor equal to
        ``vocoder_scale * mel.transpose(2, 1)``, produced by the last posterior update itself (handoff.py)."""
        b, device = cond.shape[0], cond.device
        if not infer:
            # training branch: one denoiser call at a random step; the denoiser itself refuses autograd (see WaveNet)
            x_0 = self.norm_spec(gt_spec)
            t = torch.randint(0, self.num_timesteps + 1, (b,), device=device).long()
            x_t = self.q_sample(x_0, t=t, x_T=src_spec)
            x_0_pred = self.denoise_fn(x_t, t, cond.transpose(1, 2))
            return x_0_pred, x_0
        infer_step = int(np.clip(infer_step, 1, self.num_timesteps))
        T = cond.shape[1]
        if x_T is None:
            x = torch.rand(b, 1, self.mel_bins, T, device=device)  # uniform start, as prodiff.py:147
        else:
            x = x_T.detach().to(device=device, dtype=torch.float32).clone().contiguous()
            if x.shape != (b, 1, self.mel_bins, T):
                raise ValueError(f"x_T must be {(b, 1, self.mel_bins, T)}, got {tuple(x.shape)}")
        if noise is not None:
            noise = noise.detach().to(device=device, dtype=torch.float32).contiguous()
            if noise.shape != (infer_step, b, 1, self.mel_bins, T):
                raise ValueError(f"noise must be {(infer_step, b, 1, self.mel_bins, T)}, got {tuple(noise.shape)}")
        if x0_trace is not None and (x0_trace.shape != (infer_step, b, 1, self.mel_bins, T) or not x0_trace.is_contiguous()
                                     or x0_trace.dtype != torch.float32 or x0_trace.device != device):
            raise ValueError("x0_trace must be a contiguous fp32 [K,B,1,M,T] tensor on cond's device")
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        c1, c2, lv = self._tables_on_host()
        # a conditioning.AssembledCondition (frame-level assembly deferred to the kernels) passes through as it is
        cond_bth = cond.detach().to(dtype=torch.float32).contiguous() if isinstance(cond, torch.Tensor) else cond
        t_list = list(range(infer_step - 1, -1, -1))
        self.denoise_fn.engine().sample(cond_bth, x, t_list, c1, c2, lv, noise=noise, seed=seed, x0_trace=x0_trace,
                                        lengths=lengths, out_scale=1.0 if vocoder_scale is None else float(vocoder_scale))
        if vocoder_scale is not None:
            return x[:, 0]   # denorm_spec is the identity here (prodiff.py:158), so scaling first changes nothing
        x = x[:, 0].transpose(1, 2)
        return self.denorm_spec(x)


__all__ = ["GaussianDiffusion", "extract"]

"""CPU fp32 oracle for the ProDiff sampling hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, op for op, the two reference files on the hot path

    /root/reference/modules/decoder/wavenet.py      (DiffNet denoiser)
    /root/reference/modules/diffusion/prodiff.py    (GaussianDiffusion sampler)

as plain functions over a ``state`` dict of fp32 tensors (the reference's own
``state_dict`` keys).  It exists so that the CUDA path can be checked on a GPU box where
``/root/reference`` is not mounted.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it; the product package
``prodiff_b200`` never does (tests/test_no_oracle_in_product.py enforces that).

Pinning: the reference ships no tests, golden vectors or checkpoints for this path
(SURVEY.md §4, §8c), so the oracle is pinned against the *live reference modules* imported
from /root/reference in the build container: ``tests/test_oracle_vs_reference.py`` demands
bit-identical outputs, and ``oracle/make_golden.py`` (committed) ran the reference itself to
produce the known-answer vectors under ``tests/golden/`` that travel to the GPU box.

The only deliberate difference from the reference: ``x_T`` and the per-step Gaussian noise are
*injected* tensors instead of ``torch.rand`` / ``torch.randn`` calls (prodiff.py:118,147), so two
implementations can be compared on identical randomness.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# noise schedules and posterior tables  (prodiff.py:18-46, 49-104)
# --------------------------------------------------------------------------------------
def vpsde_beta_t(t, T, min_beta, max_beta):
    """prodiff.py:18-20"""
    t_coef = (2 * t - 1) / (T ** 2)
    return 1.0 - np.exp(-min_beta / T - 0.5 * (max_beta - min_beta) * t_coef)


def logsnr_schedule_cosine(t, *, logsnr_min, logsnr_max):
    """prodiff.py:22-25"""
    b = np.arctan(np.exp(-0.5 * logsnr_max))
    a = np.arctan(np.exp(-0.5 * logsnr_min)) - b
    return -2.0 * np.log(np.tan(a * t + b))


def get_noise_schedule_list(schedule_mode, timesteps, min_beta=0.0, max_beta=0.01, s=0.008):
    """prodiff.py:27-46 (float64 numpy, same formulas, same error on unknown mode)."""
    if schedule_mode == "linear":
        return np.linspace(1e-4, max_beta, timesteps)
    if schedule_mode == "cosine":
        steps = timesteps + 1
        x = np.linspace(0, steps, steps)
        ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = 1 - (ac[1:] / ac[:-1])
        return np.clip(betas, a_min=0, a_max=0.999)
    if schedule_mode == "vpsde":
        return np.array([vpsde_beta_t(t, timesteps, min_beta, max_beta) for t in range(1, timesteps + 1)])
    if schedule_mode == "logsnr":
        return np.array([logsnr_schedule_cosine(t / timesteps, logsnr_min=-20.0, logsnr_max=20.0)
                         for t in range(1, timesteps + 1)])
    raise NotImplementedError


def diffusion_tables(timesteps=1000, schedule_type="vpsde", max_beta=0.02, betas=None) -> Dict[str, torch.Tensor]:
    """The persistent buffers GaussianDiffusion.__init__ registers (prodiff.py:54-99).

    All arithmetic in float64 numpy, cast to fp32 at the end, exactly as the reference does.
    Every table has ``timesteps + 1`` entries.
    """
    if betas is None:
        betas = get_noise_schedule_list(schedule_mode=schedule_type, timesteps=timesteps + 1,
                                        min_beta=0.1, max_beta=max_beta, s=0.008)
    betas = np.asarray(betas, dtype=np.float64)
    alphas = 1.0 - betas
    alphas_cumprod = np.cumprod(alphas, axis=0)
    alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
    posterior_variance = betas * (1.0 - alphas_cumprod_prev) / (1.0 - alphas_cumprod)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    return {
        "betas": f32(betas),
        "alphas_cumprod": f32(alphas_cumprod),
        "alphas_cumprod_prev": f32(alphas_cumprod_prev),
        "sqrt_alphas_cumprod": f32(np.sqrt(alphas_cumprod)),
        "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - alphas_cumprod)),
        "log_one_minus_alphas_cumprod": f32(np.log(1.0 - alphas_cumprod)),
        "sqrt_recip_alphas_cumprod": f32(np.sqrt(1.0 / alphas_cumprod)),
        "sqrt_recipm1_alphas_cumprod": f32(np.sqrt(1.0 / alphas_cumprod - 1)),
        "posterior_variance": f32(posterior_variance),
        "posterior_log_variance_clipped": f32(np.log(np.maximum(posterior_variance, 1e-20))),
        "posterior_mean_coef1": f32(betas * np.sqrt(alphas_cumprod_prev) / (1.0 - alphas_cumprod)),
        "posterior_mean_coef2": f32((1.0 - alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - alphas_cumprod)),
    }


# --------------------------------------------------------------------------------------
# DiffNet denoiser (wavenet.py)
# --------------------------------------------------------------------------------------
def sinusoidal_pos_emb(x: torch.Tensor, dim: int) -> torch.Tensor:
    """wavenet.py:31-38.  x: [B] (int64 or float) -> [B, dim]."""
    half_dim = dim // 2
    emb = math.log(10000) / (half_dim - 1)
    emb = torch.exp(torch.arange(half_dim, device=x.device) * -emb)
    emb = x[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def mish(x: torch.Tensor) -> torch.Tensor:
    """wavenet.py:22-24"""
    return x * torch.tanh(F.softplus(x))


def num_layers(state: State) -> int:
    n = 0
    while f"residual_layers.{n}.dilated_conv.weight" in state:
        n += 1
    return n


def residual_block(state: State, i: int, dilation: int, x, conditioner, diffusion_step, trace: Optional[dict] = None):
    """wavenet.py:60-72"""
    p = f"residual_layers.{i}."
    d = F.linear(diffusion_step, state[p + "diffusion_projection.weight"],
                 state[p + "diffusion_projection.bias"]).unsqueeze(-1)
    c = F.conv1d(conditioner, state[p + "conditioner_projection.weight"], state[p + "conditioner_projection.bias"])
    y = x + d
    y = F.conv1d(y, state[p + "dilated_conv.weight"], state[p + "dilated_conv.bias"],
                 padding=dilation, dilation=dilation) + c
    gate, filt = torch.chunk(y, 2, dim=1)
    y = torch.sigmoid(gate) * torch.tanh(filt)
    if trace is not None:
        trace.setdefault("cond_proj", []).append(c)
        trace.setdefault("gated", []).append(y)
    y = F.conv1d(y, state[p + "output_projection.weight"], state[p + "output_projection.bias"])
    residual, skip = torch.chunk(y, 2, dim=1)
    return (x + residual) / math.sqrt(2.0), skip


def wavenet_forward(state: State, spec, diffusion_step, cond, dilation_cycle_length: int,
                    trace: Optional[dict] = None):
    """wavenet.py:100-123.  spec [B,1,M,T], diffusion_step [B], cond [B,H,T] -> [B,1,M,T]."""
    L = num_layers(state)
    C = state["input_projection.weight"].shape[0]
    x = spec[:, 0]
    x = F.conv1d(x, state["input_projection.weight"], state["input_projection.bias"])
    x = F.relu(x)
    e = sinusoidal_pos_emb(diffusion_step, C)
    e = F.linear(e, state["mlp.0.weight"], state["mlp.0.bias"])
    e = mish(e)
    e = F.linear(e, state["mlp.2.weight"], state["mlp.2.bias"])
    skip = []
    if trace is not None:
        trace["x_in"] = x
        trace["x"] = []
    for i in range(L):
        x, s = residual_block(state, i, 2 ** (i % dilation_cycle_length), x, cond, e, trace)
        skip.append(s)
        if trace is not None:
            trace["x"].append(x)
    x = torch.sum(torch.stack(skip), dim=0) / math.sqrt(L)
    x = F.conv1d(x, state["skip_projection.weight"], state["skip_projection.bias"])
    x = F.relu(x)
    x = F.conv1d(x, state["output_projection.weight"], state["output_projection.bias"])
    return x[:, None, :, :]


# --------------------------------------------------------------------------------------
# sampler (prodiff.py:13-16, 106-153)
# --------------------------------------------------------------------------------------
def extract(a, t, x_shape):
    """prodiff.py:13-16"""
    b, *_ = t.shape
    out = a.gather(-1, t)
    return out.reshape(b, *((1,) * (len(x_shape) - 1)))


def q_posterior_sample(tables, x_0, x_t, t, noise):
    """prodiff.py:106-121 with the ``torch.randn`` at :118 replaced by ``noise``."""
    b = x_0.shape[0]
    mean = (extract(tables["posterior_mean_coef1"], t, x_t.shape) * x_0 +
            extract(tables["posterior_mean_coef2"], t, x_t.shape) * x_t)
    log_var = extract(tables["posterior_log_variance_clipped"], t, x_t.shape)
    nonzero_mask = (1 - (t == 0).float()).reshape(b, *((1,) * (len(x_0.shape) - 1)))
    return mean + nonzero_mask * (0.5 * log_var).exp() * noise


def q_sample(tables, x_0, t, x_T):
    """prodiff.py:128-133"""
    return (extract(tables["sqrt_alphas_cumprod"], t, x_0.shape) * x_0 +
            extract(tables["sqrt_one_minus_alphas_cumprod"], t, x_0.shape) * x_T)


@torch.no_grad()
def sample(state: State, tables, cond, x_T, noise, infer_step: int, num_timesteps: int,
           dilation_cycle_length: int) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """GaussianDiffusion.forward(infer=True) (prodiff.py:136-138,145-153).

    cond [B,T,H]; x_T [B,1,M,T] replaces torch.rand (:147); noise [K,B,1,M,T] replaces the
    per-step torch.randn (:118), noise[j] being consumed by the j-th loop iteration
    (t = K-1-j), including the masked one at t == 0.
    Returns (mel [B,T,M] as the same transposed view the reference returns, [x0_pred per step]).
    """
    b = cond.shape[0]
    cond = cond.transpose(1, 2)
    infer_step = int(np.clip(infer_step, 1, num_timesteps))
    x = x_T
    x0s = []
    for j, i in enumerate(range(infer_step - 1, -1, -1)):
        t = torch.full((b,), i, dtype=torch.long)
        x_0_pred = wavenet_forward(state, x, t, cond, dilation_cycle_length)
        x0s.append(x_0_pred)
        x = q_posterior_sample(tables, x_0_pred, x, t, noise[j])
    return x[:, 0].transpose(1, 2), x0s


# --------------------------------------------------------------------------------------
# rectified-flow sampler over the same denoiser
#     /root/reference/modules/diffusion/reflow.py:48-107
# --------------------------------------------------------------------------------------
REFLOW_ALGORITHMS = ("euler", "rk2", "rk4", "rk5")


@torch.no_grad()
def reflow_sample(state: State, cond, x_start, infer_step: int, algorithm: str, time_scale,
                  spec_min, spec_max, dilation_cycle_length: int) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """RectifiedFlow.forward(infer=True) = inference + denorm_spec (reflow.py:32-46, 86-107).

    cond [B,T,H]; x_start [B,1,M,T] replaces the torch.randn at :88; ``spec_min`` / ``spec_max`` are the
    per-bin lists the constructor takes (:20-23).  The four integrators follow sample_euler / sample_rk2 /
    sample_rk4 / sample_rk5 (:48-84) expression by expression, including the fp32 tensor ``dts`` the loop
    multiplies the step index with (:97-99) and the in-place ``x +=``.  An unknown algorithm falls back to
    euler as the reference's ``dict.get`` does (:95).
    Returns (mel [B,T,M] after denorm_spec, [velocity of every denoiser evaluation, in call order]).
    """
    cond = cond.transpose(1, 2)
    M = x_start.shape[2]
    vs: List[torch.Tensor] = []

    def velocity_fn(x, t):
        v = wavenet_forward(state, x, t, cond, dilation_cycle_length)
        vs.append(v)
        return v

    def sample_euler(x, t, dt):
        x += velocity_fn(x, time_scale * t) * dt
        return x

    def sample_rk2(x, t, dt):
        k_1 = velocity_fn(x, time_scale * t)
        k_2 = velocity_fn(x + 0.5 * k_1 * dt, time_scale * (t + 0.5 * dt))
        x += k_2 * dt
        return x

    def sample_rk4(x, t, dt):
        k_1 = velocity_fn(x, time_scale * t)
        k_2 = velocity_fn(x + 0.5 * k_1 * dt, time_scale * (t + 0.5 * dt))
        k_3 = velocity_fn(x + 0.5 * k_2 * dt, time_scale * (t + 0.5 * dt))
        k_4 = velocity_fn(x + k_3 * dt, time_scale * (t + dt))
        x += (k_1 + 2 * k_2 + 2 * k_3 + k_4) * dt / 6
        return x

    def sample_rk5(x, t, dt):
        k_1 = velocity_fn(x, time_scale * t)
        k_2 = velocity_fn(x + 0.25 * k_1 * dt, time_scale * (t + 0.25 * dt))
        k_3 = velocity_fn(x + 0.125 * (k_2 + k_1) * dt, time_scale * (t + 0.25 * dt))
        k_4 = velocity_fn(x + 0.5 * (-k_2 + 2 * k_3) * dt, time_scale * (t + 0.5 * dt))
        k_5 = velocity_fn(x + 0.0625 * (3 * k_1 + 9 * k_4) * dt, time_scale * (t + 0.75 * dt))
        k_6 = velocity_fn(x + (-3 * k_1 + 2 * k_2 + 12 * k_3 - 12 * k_4 + 8 * k_5) * dt / 7, time_scale * (t + dt))
        x += (7 * k_1 + 32 * k_3 + 12 * k_4 + 32 * k_5 + 7 * k_6) * dt / 90
        return x

    fn = {"euler": sample_euler, "rk2": sample_rk2, "rk4": sample_rk4, "rk5": sample_rk5}.get(algorithm, sample_euler)
    x = x_start.clone()
    dt = 1.0 / max(1, infer_step)
    dts = torch.tensor([dt]).to(x)
    for i in range(infer_step):
        x = fn(x, i * dts, dt)
    x = x.float()
    x = x.transpose(2, 3).squeeze(1)                                  # [B,1,M,T] -> [B,T,M]
    smin = torch.FloatTensor(spec_min)[None, None, :M].transpose(-3, -2)
    smax = torch.FloatTensor(spec_max)[None, None, :M].transpose(-3, -2)
    return (x + 1) / 2 * (smax - smin) + smin, vs


# --------------------------------------------------------------------------------------
# frame-level half of the condition assembly
#     /root/reference/modules/svs/prodiff_teacher.py:98-101 (add_pitch), :127-149 (forward_condition after the encoder)
# --------------------------------------------------------------------------------------
@torch.no_grad()
def assemble_condition(encoder_out, mel2ph, f0, pitch_w, pitch_b, spk_embed=None, gender_embed=None,
                       variances=()) -> torch.Tensor:
    """``forward_condition`` from the length regulator on (prodiff_teacher.py:127-149).

    encoder_out [B,Tph,H] (what ``self.encoder`` returned), mel2ph [B,T] int64 (0 = padding), f0 [B,T];
    pitch_w [H,1] / pitch_b [H] = ``pitch_embed`` (a Linear(1,H)); spk_embed / gender_embed: what add_spk_embed /
    add_gender_embed return ([B,1,H] looked-up rows or a caller-supplied mix that may be [B,T,H]); variances: sequence of
    (x [B,T], weight [H,1], bias [H]) for the voicing / breath embeddings, summed through torch.stack(...).sum(-1) like
    :146-147.  Returns the condition [B,T,H].
    """
    condition = F.pad(encoder_out, [0, 0, 1, 0])                                  # :128
    mel2ph_ = mel2ph[..., None].repeat([1, 1, encoder_out.shape[-1]])             # :129
    condition = torch.gather(condition, 1, mel2ph_)                               # :130
    f0_mel = (1 + f0 / 700).log()                                                 # :99
    condition += F.linear(f0_mel[:, :, None], pitch_w, pitch_b)                   # :100, :132
    if spk_embed is not None:
        condition += spk_embed                                                    # :135
    if gender_embed is not None:
        condition += gender_embed                                                 # :137
    variance_embeds = [F.linear(x[:, :, None], w, b) for x, w, b in variances]    # :141-145
    if len(variance_embeds) > 0:
        condition += torch.stack(variance_embeds, dim=-1).sum(-1)                 # :147
    nonpadding = (mel2ph > 0).float()[:, :, None]                                 # :148
    return condition * nonpadding                                                 # :149

"""CPU fp32 oracle for the NSF-HiFiGAN generator (the step after the sampler).  TEST INFRASTRUCTURE ONLY.

Restates, op for op, what ``Generator.forward(x, f0)`` executes once ``remove_weight_norm()`` has folded the weight norm
(what ``load_model`` leaves, models.py:21-36):

    /root/reference/modules/nsf_hifigan/models.py:269-290   Generator.forward
    /root/reference/modules/nsf_hifigan/models.py:62-69     ResBlock1.forward
    /root/reference/modules/nsf_hifigan/models.py:132-188   SineGen._f02uv / _f02sine / forward
    /root/reference/modules/nsf_hifigan/models.py:212-215   SourceModuleHnNSF.forward
    /root/reference/modules/nsf_hifigan/utils.py:12-13      get_padding
    /root/reference/component/vocoder/nsf_hifigan.py:49-58  the hand-off (c = 2.30259 * mel.transpose(2, 1))

(one utterance per call: the reference's SineGen is written for batch size 1, models.py:163,174-179; a batch here is
defined as the row-wise application, the same rand_ini for every row) as plain functions over a ``state`` dict with the reference's own ``state_dict`` keys (after remove_weight_norm) and a
config ``h`` (dict with the fields of the checkpoint's config.json).  The two random draws of ``SineGen`` -- ``torch.rand``
for the initial phases (:143) and ``torch.randn_like`` for the additive noise (:186) -- are arguments here.

Same rules as ``prodiff_oracle.py``: only ``tests/``, ``smoke()`` and bench CPU legs may import it.  Pinning: the
reference has no tests, vectors or checkpoints for the vocoder either, so this file is pinned against the live reference
class (``tests/test_oracle_vs_reference.py``: bit-identical, with ``matplotlib`` stubbed for ``nsf_hifigan/utils.py:1-3``)
and ``oracle/make_golden.py`` ran the unmodified reference to produce ``tests/golden/vocoder_*.npz``.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]
LRELU_SLOPE = 0.1           # models.py:18
HARMONIC_NUM = 8            # models.py:219
SINE_AMP, NOISE_STD = 0.1, 0.003   # SineGen defaults (models.py:120-121), used by SourceModuleHnNSF (:194-202)


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    return int((kernel_size * dilation - dilation) / 2)   # utils.py:12-13


def upp_of(h) -> int:
    return int(np.prod(h["upsample_rates"]))               # models.py:267


def f02sine(f0_values: torch.Tensor, upp: int, sampling_rate: float, rand_ini: torch.Tensor) -> torch.Tensor:
    """models.py:138-170 for fp32 input.  f0_values [B, T, dim]; rand_ini [1, dim] with rand_ini[:, 0] == 0."""
    rad_values = (f0_values / sampling_rate).fmod(1.)
    rad_values[:, 0, :] += rand_ini
    tmp_over_one = torch.cumsum(rad_values.double(), 1)
    tmp_over_one = tmp_over_one.float()
    tmp_over_one *= upp
    tmp_over_one = F.interpolate(tmp_over_one.transpose(2, 1), scale_factor=upp, mode='linear', align_corners=True).transpose(2, 1)
    rad_values = F.interpolate(rad_values.transpose(2, 1), scale_factor=upp, mode='nearest').transpose(2, 1)
    tmp_over_one = tmp_over_one.fmod(1.)
    diff = F.conv2d(tmp_over_one.unsqueeze(1), torch.FloatTensor([[[[-1.], [1.]]]]), stride=(1, 1), padding=0, dilation=(1, 1)).squeeze(1)
    cumsum_shift = (diff < 0).double()
    # the reference writes torch.zeros((1, 1, dim)) here (:163), i.e. its SineGen -- and with it Generator.forward -- only
    # accepts one utterance per call (docstring :174-179); the batch size below is the only generalisation in this file
    cumsum_shift = torch.cat((torch.zeros((f0_values.shape[0], 1, f0_values.shape[2]), dtype=torch.double), cumsum_shift), dim=1)
    sines = torch.sin(torch.cumsum(rad_values.double() + cumsum_shift, dim=1) * 2 * np.pi)
    return sines.float()


def sine_gen(f0: torch.Tensor, upp: int, sampling_rate: float, rand_ini: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """SineGen.forward (models.py:172-188).  f0 [B, T]; noise [B, T * upp, 9] standing in for torch.randn_like (:186)."""
    dim = HARMONIC_NUM + 1
    f0 = f0.unsqueeze(-1)
    fn = torch.multiply(f0, torch.arange(1, dim + 1).reshape((1, 1, -1)))
    sine_waves = f02sine(fn, upp, sampling_rate, rand_ini) * SINE_AMP
    uv = (f0 > 0).float()
    uv = F.interpolate(uv.transpose(2, 1), scale_factor=upp, mode='nearest').transpose(2, 1)
    noise_amp = uv * NOISE_STD + (1 - uv) * SINE_AMP / 3
    return sine_waves * uv + noise_amp * noise


def har_source(state: State, h, f0, rand_ini, noise) -> torch.Tensor:
    """SourceModuleHnNSF.forward (models.py:212-215) then the transpose of Generator.forward :270 -> [B, 1, T * upp]."""
    sine_wavs = sine_gen(f0, upp_of(h), float(h["sampling_rate"]), rand_ini, noise)
    sine_merge = torch.tanh(F.linear(sine_wavs, state["m_source.l_linear.weight"], state["m_source.l_linear.bias"]))
    return sine_merge.transpose(1, 2)


def resblock1(state: State, prefix: str, x: torch.Tensor, k: int, dilations) -> torch.Tensor:
    """ResBlock1.forward (models.py:62-69)."""
    for i, d in enumerate(dilations):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, state[f"{prefix}.convs1.{i}.weight"], state[f"{prefix}.convs1.{i}.bias"], dilation=d, padding=get_padding(k, d))
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, state[f"{prefix}.convs2.{i}.weight"], state[f"{prefix}.convs2.{i}.bias"], dilation=1, padding=get_padding(k, 1))
        x = xt + x
    return x


def generator_forward(state: State, h, x: torch.Tensor, f0: torch.Tensor, rand_ini: torch.Tensor, noise: torch.Tensor, trace=None) -> torch.Tensor:
    """Generator.forward (models.py:269-290).  x [B, num_mels, T] (the hand-off tensor c), f0 [B, T] -> [B, 1, T * upp]."""
    rates, ksz = h["upsample_rates"], h["upsample_kernel_sizes"]
    rk, rd = h["resblock_kernel_sizes"], h["resblock_dilation_sizes"]
    num_kernels = len(rk)
    src = har_source(state, h, f0, rand_ini, noise)
    if trace is not None:
        trace["har_source"] = src
    x = F.conv1d(x, state["conv_pre.weight"], state["conv_pre.bias"], padding=3)
    for i, (u, k) in enumerate(zip(rates, ksz)):
        x = F.leaky_relu(x, LRELU_SLOPE)
        x = F.conv_transpose1d(x, state[f"ups.{i}.weight"], state[f"ups.{i}.bias"], stride=u, padding=(k - u) // 2)
        if i + 1 < len(rates):
            stride_f0 = int(np.prod(rates[i + 1:]))
            x_source = F.conv1d(src, state[f"noise_convs.{i}.weight"], state[f"noise_convs.{i}.bias"], stride=stride_f0, padding=stride_f0 // 2)
        else:
            x_source = F.conv1d(src, state[f"noise_convs.{i}.weight"], state[f"noise_convs.{i}.bias"])
        x = x + x_source
        xs = None
        for j in range(num_kernels):
            r = resblock1(state, f"resblocks.{i * num_kernels + j}", x, rk[j], rd[j])
            xs = r if xs is None else xs + r
        x = xs / num_kernels
        if trace is not None:
            trace[f"stage{i}"] = x
    x = F.leaky_relu(x)
    x = F.conv1d(x, state["conv_post.weight"], state["conv_post.bias"], padding=3)
    return torch.tanh(x)


def spec2wav(state: State, h, mel: torch.Tensor, f0: torch.Tensor, rand_ini, noise) -> torch.Tensor:
    """NsfHifiGAN.spec2wav_torch (nsf_hifigan.py:49-58): mel [B, T, bins] (log10) -> flat waveform."""
    c = 2.30259 * mel.transpose(2, 1)
    return generator_forward(state, h, c, f0, rand_ini, noise).view(-1)

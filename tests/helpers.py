"""Shared helpers for the test-suite (fixture loading, case regeneration)."""
import os

import numpy as np
import torch

from oracle import prodiff_oracle as oracle
from oracle import encoder_oracle
from oracle.golden_cases import (SAMPLE_CASES, DENOISE_CASES, REFLOW_CASES, CONDITION_CASES, ENCODER_CASES,
                                 reflow_spec_range, make_condition_inputs)
from prodiff_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(kind, name):
    return np.load(os.path.join(GOLDEN, f"{kind}_{name}.npz"))


def regen_sample_case(name):
    case = SAMPLE_CASES[name]
    cfg = case["cfg"]
    state = synth.make_state(cfg, case["wseed"], case["out_init"])
    cond, x_T, noise = synth.make_inputs(cfg, case["B"], case["T"], case["infer_step"], case["iseed"])
    return case, cfg, state, cond, x_T, noise


def regen_denoise_case(name):
    case = DENOISE_CASES[name]
    cfg = case["cfg"]
    state = synth.make_state(cfg, case["wseed"], case["out_init"])
    cond, x_T, _ = synth.make_inputs(cfg, case["B"], case["T"], 1, case["iseed"])
    return case, cfg, state, cond, x_T, torch.tensor(case["steps"])


def oracle_sample(case, cfg, state, cond, x_T, noise):
    tables = oracle.diffusion_tables(case["timesteps"], case["schedule_type"], case["max_beta"])
    return oracle.sample(state, tables, cond, x_T, noise, case["infer_step"], case["timesteps"],
                         cfg.dilation_cycle_length)


def regen_reflow_case(name):
    case = REFLOW_CASES[name]
    cfg = case["cfg"]
    state = synth.make_state(cfg, case["wseed"], case["out_init"])
    cond, _, noise = synth.make_inputs(cfg, case["B"], case["T"], 1, case["iseed"])
    return case, cfg, state, cond, noise[0]


def oracle_reflow(case, cfg, state, cond, x_start):
    """(denormalised output, velocities) of the oracle; the pitch flavour applies reflow.py:143 on top."""
    smin, smax = reflow_spec_range(case)
    out, vs = oracle.reflow_sample(state, cond, x_start, case["infer_step"], case["algorithm"], case["time_scale"],
                                   smin, smax, cfg.dilation_cycle_length)
    if case["spec"] == "pitch":
        out = out.mean(dim=-1).clamp(-12.0, 12.0)
    return out, vs


def oracle_condition(name):
    """(inputs dict, oracle condition [B,T,H]) of a CONDITION case."""
    d = make_condition_inputs(CONDITION_CASES[name])
    out = oracle.assemble_condition(d["encoder_out"], d["mel2ph"], d["f0"], d["pitch_w"], d["pitch_b"], d["spk"], d["gender"],
                                    d["variances"])
    return d, out


def regen_encoder_case(name):
    """(case, cfg, state, txt_tokens, extra_embed or None) of an ENCODER case; the digest always covers extra_embed."""
    case = ENCODER_CASES[name]
    cfg = case["cfg"]
    state = synth.make_encoder_state(cfg, case["wseed"])
    tok, extra = synth.make_encoder_inputs(cfg, case["lengths"], case["Tph"], case["iseed"])
    return case, cfg, state, tok, extra


def oracle_encoder(case, cfg, state, tok, extra, trace=None):
    with torch.no_grad():
        return encoder_oracle.encoder_forward(state, tok, extra if case["extra"] else None, cfg.num_heads, trace)

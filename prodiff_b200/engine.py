"""Host-side owner of the C-ABI handles: packed weights (pd_model) and per-shape plans (pd_plan).

PyTorch is plumbing here: it owns device memory (packed-weight block and scratch come from the caching allocator as
uint8 tensors) and the current stream; every kernel is launched by the C library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib

STATE_HEAD = ["input_projection.weight", "input_projection.bias", "mlp.0.weight", "mlp.0.bias", "mlp.2.weight",
              "mlp.2.bias", "skip_projection.weight", "skip_projection.bias", "output_projection.weight",
              "output_projection.bias"]
STATE_LAYER = ["dilated_conv.weight", "dilated_conv.bias", "diffusion_projection.weight", "diffusion_projection.bias",
               "conditioner_projection.weight", "conditioner_projection.bias", "output_projection.weight",
               "output_projection.bias"]


def state_key_order(residual_layers: int):
    """The order pd_model_create expects (include/prodiff_b200.h)."""
    keys = list(STATE_HEAD)
    for i in range(residual_layers):
        keys += [f"residual_layers.{i}.{k}" for k in STATE_LAYER]
    return keys


def _require_b200(device: torch.device) -> None:
    if device.type != "cuda":
        raise RuntimeError(f"prodiff_b200 runs only on a CUDA sm_100 device (got a tensor on '{device}'); "
                           "there is no CPU fallback")
    major, _ = torch.cuda.get_device_capability(device)
    if major != 10:
        raise RuntimeError(f"prodiff_b200 needs compute capability 10.x (B200), found {major}.x; no fallback path")


def _aligned_bytes(nbytes: int, device) -> torch.Tensor:
    buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    off = (-buf.data_ptr()) % 1024
    return buf[off:off + nbytes]


def row_chunks(B: int, T: int, frames_per_call: int):
    """Row blocks [(start, end)] for one sampling call over B utterances of T frames.

    The layer stack keeps its activation ping-pong buffers (2 x frames x C x 2 B) resident in the 126 MB L2 between layers
    only up to about 64 k frames per call (measured: 0.73 of peak at 64 x 1000 against 0.63 at 256 x 1000), and the hoisted
    condition projection costs frames x L x 2C x 2 B of workspace (21 GB at 1024 x 1000), so large batches are walked in
    nearly equal blocks of at most ``frames_per_call`` frames.  A batch up to 1.5 x that size stays one call."""
    rows = max(1, frames_per_call // max(1, T))
    if B <= rows + rows // 2:
        return [(0, B)]
    n = -(-B // rows)
    base, extra = divmod(B, n)
    out, s = [], 0
    for i in range(n):
        e = s + base + (1 if i < extra else 0)
        out.append((s, e))
        s = e
    return out


def workspace_bucket(nbytes: int) -> int:
    """Allocation size for a plan workspace of ``nbytes``: the next of eight steps per power of two (<= 12.5 % slack), at
    least 1 MiB.  Plans of nearby shapes (online inference: every new T) then share a handful of block sizes, so an evicted
    plan's block fits the next plan and the caching allocator is not left with multi-GB holes of slightly different sizes."""
    n = max(int(nbytes), 1 << 20)
    step = 1 << max(n.bit_length() - 4, 0)
    return -(-n // step) * step


class Plan:
    def __init__(self, engine: "Engine", B: int, T: int, entries: int):
        lib = engine.lib
        self.engine, self.B, self.T, self.entries = engine, B, T, entries
        nbytes = lib.pd_workspace_bytes(engine.handle, B, T, entries)
        if nbytes == 0:
            raise RuntimeError(f"prodiff_b200: bad plan shape B={B} T={T} entries={entries}")
        self.block = engine.take_block(nbytes)        # uint8 tensor of a bucketed size, returned to the pool on eviction
        off = (-self.block.data_ptr()) % 1024
        self.workspace = self.block[off:off + nbytes]
        h = C.c_void_p()
        _lib.check(lib.pd_plan_create(engine.handle, B, T, entries, self.workspace.data_ptr(), nbytes, C.byref(h)),
                   "pd_plan_create")
        self.handle = h

    def buffer(self, name: str) -> torch.Tensor:
        """uint8 view of a named scratch buffer (tests only)."""
        off, n = C.c_size_t(), C.c_size_t()
        _lib.check(self.engine.lib.pd_plan_buffer(self.handle, name.encode(), C.byref(off), C.byref(n)), "pd_plan_buffer")
        return self.workspace[off.value:off.value + n.value]

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.engine.lib.pd_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class Engine:
    """Packed weights of one DiffNet on one device + a small cache of plans."""

    MAX_PLANS = 8
    FRAMES_PER_CALL = 65536   # see row_chunks

    def __init__(self, cfg: dict, state: Dict[str, torch.Tensor], device: torch.device, operand_dtype: str = "bf16"):
        _require_b200(device)
        self.lib = _lib.load()
        self.device = device
        self.operand_dtype = operand_dtype
        dt = {"bf16": _lib.PD_BF16, "fp16": _lib.PD_FP16}[operand_dtype]
        self.cfg = _lib.pd_config(cfg["in_dims"], cfg["hidden_size"], cfg["residual_layers"], cfg["residual_channels"],
                                  cfg["dilation_cycle_length"], dt)
        keys = state_key_order(cfg["residual_layers"])
        with torch.cuda.device(device):
            tensors = [state[k].detach().to(device=device, dtype=torch.float32).contiguous() for k in keys]
            nbytes = self.lib.pd_packed_bytes(C.byref(self.cfg))
            if nbytes == 0:
                raise RuntimeError(f"prodiff_b200: unsupported DiffNet config: {self.lib.pd_last_error().decode()}")
            self.packed = _aligned_bytes(nbytes, device)
            ptrs = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
            h = C.c_void_p()
            stream = torch.cuda.current_stream(device).cuda_stream
            _lib.check(self.lib.pd_model_create(C.byref(self.cfg), ptrs, len(tensors), self.packed.data_ptr(), nbytes,
                                                stream, C.byref(h)), "pd_model_create")
            # the pack kernels read `tensors` asynchronously: keep them alive until the stream has run them
            torch.cuda.current_stream(device).synchronize()
        self.handle = h
        self.plans: Dict[Tuple[int, int, int], Plan] = {}   # insertion order = recency (LRU first)
        self.pool: list = []                                # workspace blocks of evicted plans

    def take_block(self, nbytes: int) -> torch.Tensor:
        """A workspace block of at least nbytes + 1024 bytes: the smallest pooled one that fits, else a new allocation of
        the bucketed size.  Reuse is stream-ordered like every other use of the plan (the engine launches on the caller's
        current stream)."""
        need = workspace_bucket(nbytes + 1024)
        fit = [i for i, b in enumerate(self.pool) if b.numel() >= need]
        if fit:
            return self.pool.pop(min(fit, key=lambda i: self.pool[i].numel()))
        self.pool.clear()   # nothing pooled fits: let the allocator reuse those bytes for the new block
        return torch.empty(need, dtype=torch.uint8, device=self.device)

    def plan(self, B: int, T: int, entries: int) -> Plan:
        """Least-recently-used cache of MAX_PLANS plans keyed by (B, T, entries); an evicted plan hands its workspace
        block to the pool, so a new shape usually costs only the ~30 tensor-map encodes of pd_plan_create."""
        key = (B, T, entries)
        p = self.plans.pop(key, None)
        if p is None:
            while len(self.plans) >= self.MAX_PLANS:
                old = self.plans.pop(next(iter(self.plans)))    # first key = least recently used
                self.lib.pd_plan_destroy(old.handle)
                old.handle = None
                self.pool.append(old.block)
                old.block = old.workspace = None
            with torch.cuda.device(self.device):
                p = Plan(self, B, T, entries)
        self.plans[key] = p   # (re)insert at the most-recently-used end
        return p

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def assemble(self, ac, cond_out: Optional[torch.Tensor] = None, plan: Optional[Plan] = None, entries: int = 1) -> None:
        """pd_cond_assemble: frame-level condition assembly + condition projection of all layers for the rows of ``ac`` (a
        conditioning.AssembledCondition), into ``plan`` (default: a plan of its shape); ``cond_out`` receives the dense
        fp32 condition when given."""
        B, T, _ = ac.shape
        p = plan if plan is not None else self.plan(B, T, entries)
        vp = C.c_void_p
        nv = len(ac.variances)
        arr = lambda k: (vp * 2)(*[ac.variances[j][k].data_ptr() if j < nv else None for j in range(2)])
        with torch.cuda.device(self.device):
            _lib.check(self.lib.pd_cond_assemble(
                p.handle, ac.encoder_out.data_ptr(), int(ac.encoder_out.shape[1]), ac.mel2ph.data_ptr(), ac.f0.data_ptr(),
                ac.pitch_w.data_ptr(), ac.pitch_b.data_ptr(),
                ac.utt[0].data_ptr() if len(ac.utt) > 0 else None, ac.utt[1].data_ptr() if len(ac.utt) > 1 else None,
                nv, arr(0), arr(1), arr(2), ac.frame.data_ptr() if ac.frame is not None else None,
                cond_out.data_ptr() if cond_out is not None else None, self._stream()), "pd_cond_assemble")

    def _cond_arg(self, p: Plan, cond, s: int, e: int):
        """The ``cond`` argument of pd_sample / pd_reflow_sample for rows [s, e): a pointer to the dense fp32 tensor, or
        NULL after assembling + projecting an AssembledCondition into the plan."""
        if isinstance(cond, torch.Tensor):
            return cond[s:e].data_ptr()
        self.assemble(cond.rows(s, e), None, plan=p)
        return None

    def _set_lengths(self, p: Plan, lengths, s: int, e: int) -> None:
        """Install (or clear) the per-utterance valid lengths of rows [s, e) on a plan (pd_plan_set_lengths)."""
        if lengths is None:
            _lib.check(self.lib.pd_plan_set_lengths(p.handle, None, self._stream()), "pd_plan_set_lengths")
            return
        arr = (C.c_int32 * (e - s))(*[int(v) for v in lengths[s:e]])
        _lib.check(self.lib.pd_plan_set_lengths(p.handle, arr, self._stream()), "pd_plan_set_lengths")

    def denoise(self, spec: torch.Tensor, step: torch.Tensor, cond_bth: torch.Tensor) -> torch.Tensor:
        """spec [B,1,M,T] fp32, step [B] fp32, cond_bth [B,T,H] fp32 contiguous -> [B,1,M,T] fp32."""
        B, _, M, T = spec.shape
        p = self.plan(B, T, B)
        out = torch.empty_like(spec)
        with torch.cuda.device(self.device):
            self._set_lengths(p, None, 0, B)
            _lib.check(self.lib.pd_cond_project(p.handle, cond_bth.data_ptr(), self._stream()), "pd_cond_project")
            _lib.check(self.lib.pd_denoise(p.handle, spec.data_ptr(), step.data_ptr(), out.data_ptr(), self._stream()),
                       "pd_denoise")
        return out

    def sample(self, cond_bth: torch.Tensor, x: torch.Tensor, t_list: Sequence[int], coef1, coef2, logvar,
               noise: Optional[torch.Tensor] = None, seed: int = 0, x0_trace: Optional[torch.Tensor] = None,
               lengths: Optional[Sequence[int]] = None, out_scale: float = 1.0) -> torch.Tensor:
        """The whole reverse loop, in place on x [B,1,M,T] (in: x_T, out: x_0).  ``lengths`` (B ints <= T): ragged batch,
        row b is computed exactly as an utterance of lengths[b] frames; frames past it are unspecified.  ``out_scale``:
        the last posterior update stores out_scale * x_0 (vocoder hand-off, pd_plan_set_output_scale)."""
        B, _, M, T = x.shape
        if lengths is not None and len(lengths) != B:
            raise ValueError(f"lengths must have {B} entries")
        K = len(t_list)
        tl = (C.c_int32 * K)(*[int(t) for t in t_list])
        n = len(coef1)
        fa = lambda a: (C.c_float * n)(*[float(v) for v in a])
        c1, c2, lv = fa(coef1), fa(coef2), fa(logvar)
        chunks = row_chunks(B, T, self.FRAMES_PER_CALL)
        with torch.cuda.device(self.device):
            for ci, (s, e) in enumerate(chunks):
                whole = len(chunks) == 1
                # [K, B, ...] tensors are not contiguous per row block: stage them (parity/debug paths only)
                nz = None if noise is None else (noise if whole else noise[:, s:e].contiguous())
                tr = None if x0_trace is None else (x0_trace if whole else torch.empty_like(x0_trace[:, s:e]))
                p = self.plan(e - s, T, K)
                self._set_lengths(p, lengths, s, e)
                _lib.check(self.lib.pd_plan_set_output_scale(p.handle, float(out_scale)), "pd_plan_set_output_scale")
                # every block draws from its own Philox stream (the kernel counts frames from 0 inside a call)
                sd = (seed + ci * 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
                _lib.check(self.lib.pd_sample(p.handle, self._cond_arg(p, cond_bth, s, e), x[s:e].data_ptr(),
                                              nz.data_ptr() if nz is not None else None, C.c_uint64(sd),
                                              K, tl, c1, c2, lv, n, tr.data_ptr() if tr is not None else None, self._stream()),
                           "pd_sample")
                if tr is not None and not whole:
                    x0_trace[:, s:e].copy_(tr)
                if nz is not None and not whole:
                    nz.record_stream(torch.cuda.current_stream(self.device))
        return x

    REFLOW_ALGORITHMS = {"euler": (0, 1), "rk2": (1, 2), "rk4": (2, 4), "rk5": (3, 6)}   # name -> (C-ABI id, stages)

    def reflow(self, cond_bth: torch.Tensor, x: torch.Tensor, algorithm: str, n_steps: int, time_scale: float,
               spec_min: Optional[torch.Tensor], spec_max: Optional[torch.Tensor],
               v_trace: Optional[torch.Tensor] = None, lengths: Optional[Sequence[int]] = None,
               out_scale: float = 1.0) -> torch.Tensor:
        """The rectified-flow ODE loop (pd_reflow_sample), in place on x [B,1,M,T] (in: N(0,1) start, out: denormalised,
        times ``out_scale`` -- applied by the denorm_spec kernel, so it needs spec_min / spec_max)."""
        if out_scale != 1.0 and spec_min is None:
            raise ValueError("out_scale is applied by the denorm_spec kernel: pass spec_min / spec_max")
        B, _, M, T = x.shape
        algo, stages = self.REFLOW_ALGORITHMS.get(algorithm, self.REFLOW_ALGORITHMS["euler"])   # reflow.py:95 falls back to euler
        chunks = row_chunks(B, T, self.FRAMES_PER_CALL)
        with torch.cuda.device(self.device):
            for s, e in chunks:
                whole = len(chunks) == 1
                p = self.plan(e - s, T, 6)
                self._set_lengths(p, lengths, s, e)
                _lib.check(self.lib.pd_plan_set_output_scale(p.handle, float(out_scale)), "pd_plan_set_output_scale")
                tr = None if v_trace is None else (v_trace if whole else torch.empty_like(v_trace[:, s:e]))
                scratch = None if tr is not None else torch.empty(stages, e - s, 1, M, T, dtype=torch.float32, device=self.device)
                _lib.check(self.lib.pd_reflow_sample(
                    p.handle, self._cond_arg(p, cond_bth, s, e), x[s:e].data_ptr(), algo, int(n_steps), float(time_scale),
                    (tr if tr is not None else scratch).data_ptr(),
                    spec_min.data_ptr() if spec_min is not None else None, spec_max.data_ptr() if spec_max is not None else None,
                    int(spec_min.numel()) if spec_min is not None else 0,
                    tr.data_ptr() if tr is not None else None, self._stream()), "pd_reflow_sample")
                if tr is not None and not whole:
                    v_trace[:, s:e].copy_(tr)
                if scratch is not None:
                    scratch.record_stream(torch.cuda.current_stream(self.device))
        return x

    def __del__(self):
        try:
            self.plans.clear()
            if getattr(self, "handle", None):
                self.lib.pd_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

"""ctypes binding of the C-ABI library (include/prodiff_b200.h).

The shared object is built in-tree by ``prodiff_b200/csrc/build.py`` (``__graft_entry__.build()`` calls it).  There is
no fallback: if the library is missing or a call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PRODIFF_B200_LIB: developer hook to A/B two builds of the same ABI inside one GPU session
LIB_PATH = os.environ.get("PRODIFF_B200_LIB") or os.path.join(_HERE, "csrc", "libprodiff_b200.so")

PD_BF16, PD_FP16 = 0, 1

# every symbol include/prodiff_b200.h declares (tests/test_abi.py checks the header against this list)
SYMBOLS = [
    "pd_state_count", "pd_last_error", "pd_abi_version", "pd_packed_bytes", "pd_model_create", "pd_model_destroy",
    "pd_workspace_bytes", "pd_plan_create", "pd_plan_destroy", "pd_cond_project", "pd_denoise", "pd_sample",
    "pd_launch_count", "pd_plan_buffer", "pd_debug_gemm", "pd_profile_begin", "pd_profile_end", "pd_debug_shift", "pd_debug_set_trace", "pd_set_fused", "pd_set_stack", "pd_set_tail", "pd_set_sumtail", "pd_set_voc_fused", "pd_reflow_sample", "pd_plan_set_lengths", "pd_cond_assemble",
    "pd_plan_set_output_scale",
    "pd_encoder_state_count", "pd_encoder_packed_bytes", "pd_encoder_create", "pd_encoder_destroy",
    "pd_encoder_workspace_bytes", "pd_encoder_forward",
    "pd_vocoder_state_count", "pd_vocoder_packed_bytes", "pd_vocoder_create", "pd_vocoder_destroy", "pd_vocoder_upp",
    "pd_vocoder_workspace_bytes", "pd_vocoder_forward", "pd_voc_debug_conv_scratch", "pd_voc_debug_conv",
    "pd_voc_debug_pair_scratch", "pd_voc_debug_pair",
]


class pd_vocoder_config(C.Structure):
    _fields_ = [("num_mels", C.c_int32), ("upsample_initial_channel", C.c_int32), ("n_upsamples", C.c_int32),
                ("upsample_rates", C.c_int32 * 8), ("upsample_kernel_sizes", C.c_int32 * 8),
                ("n_resblock_kernels", C.c_int32), ("resblock_kernel_sizes", C.c_int32 * 4),
                ("resblock_dilation_sizes", (C.c_int32 * 3) * 4), ("resblock_type", C.c_int32), ("sampling_rate", C.c_int32)]


class pd_encoder_config(C.Structure):
    _fields_ = [("vocab_size", C.c_int32), ("hidden_size", C.c_int32), ("num_layers", C.c_int32),
                ("kernel_size", C.c_int32), ("num_heads", C.c_int32), ("dtype", C.c_int32)]


class pd_config(C.Structure):
    _fields_ = [("in_dims", C.c_int32), ("hidden_size", C.c_int32), ("residual_layers", C.c_int32),
                ("residual_channels", C.c_int32), ("dilation_cycle_length", C.c_int32), ("dtype", C.c_int32)]


_lib = None


def load() -> C.CDLL:
    """Load the library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"prodiff_b200: CUDA library not built ({LIB_PATH} missing). Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `python prodiff_b200/csrc/build.py`. There is no CPU/PyTorch fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    vp, cp, u64, sz = C.c_void_p, C.c_char_p, C.c_uint64, C.c_size_t
    P = C.POINTER
    lib.pd_state_count.argtypes = [P(pd_config)]; lib.pd_state_count.restype = C.c_int
    lib.pd_last_error.argtypes = []; lib.pd_last_error.restype = cp
    lib.pd_abi_version.argtypes = []; lib.pd_abi_version.restype = C.c_int
    lib.pd_packed_bytes.argtypes = [P(pd_config)]; lib.pd_packed_bytes.restype = sz
    lib.pd_model_create.argtypes = [P(pd_config), P(vp), C.c_int, vp, sz, vp, P(vp)]; lib.pd_model_create.restype = C.c_int
    lib.pd_model_destroy.argtypes = [vp]; lib.pd_model_destroy.restype = None
    lib.pd_workspace_bytes.argtypes = [vp, C.c_int, C.c_int, C.c_int]; lib.pd_workspace_bytes.restype = sz
    lib.pd_plan_create.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, sz, P(vp)]; lib.pd_plan_create.restype = C.c_int
    lib.pd_plan_destroy.argtypes = [vp]; lib.pd_plan_destroy.restype = None
    lib.pd_cond_project.argtypes = [vp, vp, vp]; lib.pd_cond_project.restype = C.c_int
    lib.pd_denoise.argtypes = [vp, vp, vp, vp, vp]; lib.pd_denoise.restype = C.c_int
    lib.pd_sample.argtypes = [vp, vp, vp, vp, u64, C.c_int, P(C.c_int32), P(C.c_float), P(C.c_float), P(C.c_float),
                              C.c_int, vp, vp]
    lib.pd_sample.restype = C.c_int
    lib.pd_reflow_sample.argtypes = [vp, vp, vp, C.c_int, C.c_int, C.c_float, vp, vp, vp, C.c_int, vp, vp]
    lib.pd_reflow_sample.restype = C.c_int
    lib.pd_plan_set_lengths.argtypes = [vp, P(C.c_int32), vp]; lib.pd_plan_set_lengths.restype = C.c_int
    lib.pd_plan_set_output_scale.argtypes = [vp, C.c_float]; lib.pd_plan_set_output_scale.restype = C.c_int
    lib.pd_cond_assemble.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, P(vp), P(vp), P(vp), vp, vp, vp]
    lib.pd_cond_assemble.restype = C.c_int
    lib.pd_encoder_state_count.argtypes = [P(pd_encoder_config)]; lib.pd_encoder_state_count.restype = C.c_int
    lib.pd_encoder_packed_bytes.argtypes = [P(pd_encoder_config)]; lib.pd_encoder_packed_bytes.restype = sz
    lib.pd_encoder_create.argtypes = [P(pd_encoder_config), P(vp), C.c_int, vp, sz, vp, P(vp)]; lib.pd_encoder_create.restype = C.c_int
    lib.pd_encoder_destroy.argtypes = [vp]; lib.pd_encoder_destroy.restype = None
    lib.pd_encoder_workspace_bytes.argtypes = [vp, C.c_int, C.c_int]; lib.pd_encoder_workspace_bytes.restype = sz
    lib.pd_encoder_forward.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp, sz, vp]; lib.pd_encoder_forward.restype = C.c_int
    lib.pd_vocoder_state_count.argtypes = [P(pd_vocoder_config)]; lib.pd_vocoder_state_count.restype = C.c_int
    lib.pd_vocoder_packed_bytes.argtypes = [P(pd_vocoder_config)]; lib.pd_vocoder_packed_bytes.restype = sz
    lib.pd_vocoder_create.argtypes = [P(pd_vocoder_config), P(vp), C.c_int, vp, sz, vp, P(vp)]; lib.pd_vocoder_create.restype = C.c_int
    lib.pd_vocoder_destroy.argtypes = [vp]; lib.pd_vocoder_destroy.restype = None
    lib.pd_vocoder_upp.argtypes = [vp]; lib.pd_vocoder_upp.restype = C.c_int
    lib.pd_vocoder_workspace_bytes.argtypes = [vp, C.c_int, C.c_int]; lib.pd_vocoder_workspace_bytes.restype = sz
    lib.pd_vocoder_forward.argtypes = [vp, vp, vp, vp, vp, u64, vp, vp, C.c_int, C.c_int, vp, sz, vp]; lib.pd_vocoder_forward.restype = C.c_int
    lib.pd_voc_debug_conv_scratch.argtypes = [C.c_int] * 5; lib.pd_voc_debug_conv_scratch.restype = sz
    lib.pd_voc_debug_conv.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, sz, vp]
    lib.pd_voc_debug_conv.restype = C.c_int
    lib.pd_voc_debug_pair_scratch.argtypes = [C.c_int] * 4; lib.pd_voc_debug_pair_scratch.restype = sz
    lib.pd_voc_debug_pair.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, sz, vp]
    lib.pd_voc_debug_pair.restype = C.c_int
    lib.pd_launch_count.argtypes = [C.c_int]; lib.pd_launch_count.restype = C.c_longlong
    lib.pd_plan_buffer.argtypes = [vp, cp, P(sz), P(sz)]; lib.pd_plan_buffer.restype = C.c_int
    lib.pd_debug_gemm.argtypes = [C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, vp, vp, vp]
    lib.pd_debug_gemm.restype = C.c_int
    lib.pd_debug_shift.argtypes = [C.c_int, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]; lib.pd_debug_shift.restype = C.c_int
    lib.pd_debug_set_trace.argtypes = [vp]; lib.pd_debug_set_trace.restype = C.c_int
    lib.pd_set_fused.argtypes = [C.c_int]; lib.pd_set_fused.restype = C.c_int
    lib.pd_set_stack.argtypes = [C.c_int]; lib.pd_set_stack.restype = C.c_int
    lib.pd_set_tail.argtypes = [C.c_int]; lib.pd_set_tail.restype = C.c_int
    lib.pd_set_sumtail.argtypes = [C.c_int]; lib.pd_set_sumtail.restype = C.c_int
    lib.pd_set_voc_fused.argtypes = [C.c_int]; lib.pd_set_voc_fused.restype = C.c_int
    lib.pd_profile_begin.argtypes = []; lib.pd_profile_begin.restype = C.c_int
    lib.pd_profile_end.argtypes = [P(C.c_float), P(C.c_int), C.c_int]; lib.pd_profile_end.restype = C.c_int
    if lib.pd_abi_version() != 1:
        raise RuntimeError(f"prodiff_b200: ABI version mismatch ({lib.pd_abi_version()} != 1); rebuild the library")
    _lib = lib
    return lib


KIND_NAMES = ["cond_proj", "input_proj", "dilated_conv+gate", "res_proj", "skip_sum", "skip_proj", "output_proj+posterior",
              "debug", "residual_layer(fused)", "sampler_tail(fused)", "skip_sum+tail(fused)", "encoder_gemm", "encoder_ffn_conv", "vocoder_conv"]
N_KINDS = len(KIND_NAMES)   # == PD_N_KINDS (include/prodiff_b200.h); debug trace buffers hold N_KINDS x 1024 int64


def profile_begin() -> None:
    check(load().pd_profile_begin(), "pd_profile_begin")


def profile_end() -> dict:
    """{kind name: (total ms, launches)} for every GEMM launched since profile_begin()."""
    ms = (C.c_float * N_KINDS)()
    cnt = (C.c_int * N_KINDS)()
    check(load().pd_profile_end(ms, cnt, N_KINDS), "pd_profile_end")
    return {KIND_NAMES[i]: (float(ms[i]), int(cnt[i])) for i in range(N_KINDS) if cnt[i]}


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"prodiff_b200: {what} failed: {load().pd_last_error().decode()}")

"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/prodiff_b200.h declares.
No compute call is made here (there is no GPU); argument validation that does not touch the device is exercised."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "prodiff_b200.h")


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from prodiff_b200 import _lib
    return _lib.load()


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pd_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    from prodiff_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 15
    for s in declared:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(_lib.SYMBOLS) == declared, "python binding list and header disagree"


def test_no_torch_or_cxx_types_in_header():
    src = open(HEADER).read()
    assert 'extern "C"' in src
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)   # declarations only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code and "Tensor" not in code


def test_library_is_sm100a_tcgen05_tma():
    """SASS evidence that the shipped library is the Blackwell-native path (B200_PROFILING.md table)."""
    from prodiff_b200 import _lib
    r = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in r.stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in r.stdout, mnemonic
    # The legacy mma.sync path (HMMA) is allowed in exactly one place: the text encoder's attention kernel, whose
    # 200 x 200 x 128 per-head problems are too small for a tcgen05 / TMEM pipeline.  Every dense contraction
    # of the denoiser, the sampler and the encoder's projections / FFN must stay on tcgen05.
    func = None
    for line in r.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            func = m.group(1)
        elif re.search(r"(?<![A-Z])HMMA\.", line):        # not UTCHMMA (tcgen05)
            assert func is not None and "k_enc_attention_f16" in func, f"legacy tensor-core instruction in {func}"


def test_config_validation_and_sizes(lib):
    from prodiff_b200 import _lib
    ok = _lib.pd_config(128, 256, 20, 256, 1, 0)
    assert lib.pd_state_count(C.byref(ok)) == 10 + 8 * 20
    n = lib.pd_packed_bytes(C.byref(ok))
    # 16-bit operands: 15.1 M weights + fp32 copies for the step path + biases
    assert 50e6 < n < 110e6
    bad = _lib.pd_config(100, 256, 20, 256, 1, 0)
    assert lib.pd_packed_bytes(C.byref(bad)) == 0
    assert b"in_dims" in lib.pd_last_error()
    bad = _lib.pd_config(128, 256, 20, 300, 1, 0)
    assert lib.pd_packed_bytes(C.byref(bad)) == 0
    assert b"residual_channels" in lib.pd_last_error()
    assert lib.pd_workspace_bytes(None, 1, 1, 1) == 0
    assert lib.pd_abi_version() == 1


def test_model_create_rejects_bad_arguments_without_gpu(lib):
    from prodiff_b200 import _lib
    cfg = _lib.pd_config(128, 256, 2, 256, 1, 0)
    h = C.c_void_p()
    ptrs = (C.c_void_p * 3)()
    assert lib.pd_model_create(C.byref(cfg), ptrs, 3, None, 0, None, C.byref(h)) != 0
    assert b"expected 26 state tensors" in lib.pd_last_error()
    assert not h.value


def test_encoder_config_validation_without_gpu(lib):
    """pd_encoder_* argument checks that run before anything touches a device (include/prodiff_b200.h, text encoder)."""
    from prodiff_b200 import _lib
    ok = _lib.pd_encoder_config(64, 256, 4, 9, 2, 0)
    assert lib.pd_encoder_state_count(C.byref(ok)) == 3 + 10 * 4
    n = lib.pd_encoder_packed_bytes(C.byref(ok))
    # 4 layers x (3 + 1 + 36 + 4) x 65536 16-bit weights + fp32 LayerNorm / bias / embedding tables
    assert 4 * 44 * 65536 * 2 < n < 4 * 44 * 65536 * 2 + 1e6
    for bad, word in ((_lib.pd_encoder_config(64, 512, 4, 9, 2, 0), b"hidden_size"), (_lib.pd_encoder_config(64, 256, 4, 9, 3, 0), b"num_heads"),
                      (_lib.pd_encoder_config(64, 256, 4, 8, 2, 0), b"kernel_size"), (_lib.pd_encoder_config(64, 256, 0, 9, 2, 0), b"num_layers"),
                      (_lib.pd_encoder_config(1, 256, 4, 9, 2, 0), b"vocab_size"), (_lib.pd_encoder_config(64, 256, 4, 9, 2, 7), b"dtype")):
        assert lib.pd_encoder_packed_bytes(C.byref(bad)) == 0 and word in lib.pd_last_error(), word
    h = C.c_void_p()
    ptrs = (C.c_void_p * 3)()
    assert lib.pd_encoder_create(C.byref(ok), ptrs, 3, None, 0, None, C.byref(h)) != 0
    assert b"expected 43 encoder state tensors" in lib.pd_last_error() and not h.value
    assert lib.pd_encoder_workspace_bytes(None, 1, 1) == 0
    assert lib.pd_encoder_forward(None, None, None, None, 1, 1, None, 0, None) != 0 and b"null argument" in lib.pd_last_error()


def test_vocoder_config_validation_without_gpu(lib):
    """pd_vocoder_* argument checks that run before anything touches a device (include/prodiff_b200.h, NSF-HiFiGAN generator)."""
    from prodiff_b200 import _lib

    def cfg(rates=(8, 8, 2, 2, 2), ks=(16, 16, 4, 4, 4), c0=512, mels=128, rk=(3, 7, 11), rd=((1, 3, 5),) * 3, rb=1):
        c = _lib.pd_vocoder_config()
        c.num_mels, c.upsample_initial_channel, c.n_upsamples, c.n_resblock_kernels = mels, c0, len(rates), len(rk)
        for i, (u, k) in enumerate(zip(rates, ks)):
            c.upsample_rates[i], c.upsample_kernel_sizes[i] = u, k
        for j, (k, d) in enumerate(zip(rk, rd)):
            c.resblock_kernel_sizes[j] = k
            for q in range(3):
                c.resblock_dilation_sizes[j][q] = d[q]
        c.resblock_type, c.sampling_rate = rb, 44100
        return c

    ok = cfg()
    assert lib.pd_vocoder_state_count(C.byref(ok)) == 2 + 5 * (4 + 3 * 12) + 4
    n = lib.pd_vocoder_packed_bytes(C.byref(ok))
    assert 20e6 < n < 40e6        # ~14 M weights as fp16 (transposed convs carry one third zeros) + biases
    for bad, word in ((cfg(ks=(15, 16, 4, 4, 4)), b"kernel"), (cfg(c0=500), b"upsample_initial_channel"), (cfg(mels=80), b"num_mels"),
                      (cfg(rb=2), b"resblock"), (cfg(rk=(3, 8, 11)), b"kernel size"), (cfg(rd=((1, 3, 5), (1, 3, 12), (1, 3, 5))), b"guard"),
                      (cfg(rates=(8, 8, 2, 2, 2, 2), ks=(16, 16, 4, 4, 4, 4)), b"channels")):
        assert lib.pd_vocoder_packed_bytes(C.byref(bad)) == 0 and word in lib.pd_last_error(), (word, lib.pd_last_error())
    h = C.c_void_p()
    ptrs = (C.c_void_p * 3)()
    assert lib.pd_vocoder_create(C.byref(ok), ptrs, 3, None, 0, None, C.byref(h)) != 0
    assert b"expected 206 state tensors" in lib.pd_last_error() and not h.value
    assert lib.pd_vocoder_workspace_bytes(None, 1, 1) == 0 and lib.pd_vocoder_upp(None) == 0
    assert lib.pd_vocoder_forward(None, None, None, None, None, 0, None, None, 1, 1, None, 0, None) != 0 and b"null argument" in lib.pd_last_error()


def test_missing_library_fails_loudly(monkeypatch):
    from prodiff_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libprodiff_b200.so")
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        _lib.load()

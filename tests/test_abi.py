"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol include/b200_diffusion.h
declares, host-only helpers work, and compute entry points refuse to run without a B200 (no fallback)."""
import ctypes as C

import pytest
import torch

from diffusers_b200 import _lib, ops, packing


def test_library_loads_and_exports_header_symbols():
    lib = _lib.lib()
    names = _lib.exported_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert lib.b200_version() >= 100


def test_host_helpers():
    lib = _lib.lib()
    assert lib.b200_conv_gemm_packed_k(3, 320, 0) == 9 * 320
    assert lib.b200_conv_gemm_packed_k(3, 4, 0) == 9 * 64
    assert lib.b200_conv_gemm_packed_k(1, 1280, 640) == 1920
    assert lib.b200_conv_gemm_packed_k(3, 96, 40) == 9 * (128 + 64)
    assert lib.b200_conv_gemm_pick_tile_n(2048, 10240, 1) == 256
    assert lib.b200_conv_gemm_pick_tile_n(128, 192, 1) == 64
    assert lib.b200_conv_gemm_pick_tile_n(32768, 4, 0) == 32
    assert lib.b200_group_norm_workspace_bytes(2, 16384, 32) > 0


def test_packed_k_matches_python_packer():
    for ks, c0, c1 in [(3, 320, 0), (1, 64, 0), (3, 8, 0), (3, 1280, 640), (1, 3072, 12288), (3, 100, 28)]:
        assert _lib.lib().b200_conv_gemm_packed_k(ks, c0, c1) == packing.packed_k(ks, c0, c1)


def test_argument_validation_reports_errors():
    lib = _lib.lib()
    a = _lib.ConvGemmArgs()
    rc = lib.b200_conv_gemm(C.byref(a), None)
    assert rc == -1
    assert b"null" in lib.b200_last_error()
    g = _lib.GroupNormArgs()
    assert lib.b200_group_norm(C.byref(g), None) == -1
    at = _lib.AttentionArgs()
    assert lib.b200_attention(C.byref(at), None) == -1
    assert lib.b200_euler_step(None, None, None, 0, 1.0, 0.5, 0, None) == -1


def test_no_cpu_fallback():
    x = torch.zeros(128, 64, dtype=torch.bfloat16)
    w = torch.zeros(64, 64, dtype=torch.bfloat16)
    with pytest.raises(ops.B200Error):
        ops.linear(x, w, 64)
    with pytest.raises(ops.B200Error):
        ops.layer_norm(x, eps=1e-5)
    with pytest.raises(ops.B200Error):
        ops.attention(x.view(1, 128, 64), x.view(1, 128, 64), x.view(1, 128, 64), heads=1, head_dim=64)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")
def test_init_fails_loudly_without_gpu():
    with pytest.raises(ops.B200Error):
        _lib.init(0)

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


def test_plain_c_client(tmp_path):
    """tests/c/abi_client.c: gcc + dlopen only (no torch, no CUDA headers) against the header and the built library;
    also checks that the ctypes mirror of b200_conv_gemm_args has the C struct's size."""
    import os
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "abi_client")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-o", exe, os.path.join(here, "c", "abi_client.c"), "-ldl"], check=True)
    p = subprocess.run([exe, _lib.library_path()], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "abi client ok" in p.stdout
    size = int(p.stdout.strip().rsplit("=", 1)[1])
    assert size == C.sizeof(_lib.ConvGemmArgs)


def test_weight_prefetch_plan_records_and_replays():
    """ops.WeightPrefetchPlan: first forward records the launch order, later forwards hand every launch the NEXT weight
    (wrapping around to the first), a changed order re-records and never hands out a stale pointer."""
    class W:  # stands in for a packed weight tensor
        def __init__(self, ptr, n):
            self.ptr, self.n = ptr, n

        def data_ptr(self):
            return self.ptr

        def numel(self):
            return self.n

        def element_size(self):
            return 2

    ws = [W(0x1000 * (i + 1), 100 * (i + 1)) for i in range(4)]
    plan = ops.WeightPrefetchPlan()
    with ops.weight_prefetch(plan):
        assert [plan._step(w) for w in ws] == [None] * 4  # nothing known yet
    assert plan.seq == [(w.ptr, 2 * w.n) for w in ws]
    with ops.weight_prefetch(plan):
        nxt = [plan._step(w) for w in ws]
    assert nxt == [(ws[1].ptr, 400), (ws[2].ptr, 600), (ws[3].ptr, 800), (ws[0].ptr, 200)]
    with ops.weight_prefetch(plan):  # a different order: hints stop at the first mismatch, the new order is recorded
        order = [ws[0], ws[2], ws[1]]
        got = [plan._step(w) for w in order]
    assert got[0] == (ws[1].ptr, 400) and got[1] is None and got[2] is None
    assert plan.seq == [(w.ptr, 2 * w.n) for w in order]
    big = W(0x9000, ops.WeightPrefetchPlan.MAX_BYTES)  # 2 bytes per element: twice the cap
    with ops.weight_prefetch(ops.WeightPrefetchPlan()) as p2:
        p2._step(big)
    assert p2.seq == [(0x9000, ops.WeightPrefetchPlan.MAX_BYTES)]
    assert ops._PLAN is None


HELPERS = {"b200_version", "b200_last_error", "b200_init", "b200_num_sms", "b200_conv_gemm_packed_k", "b200_conv_gemm_pick_tile_n",
           "b200_group_norm_workspace_bytes", "b200_conv_gemm_row_stats_parts", "b200_group_norm_launches",
           "b200_attention_workspace_bytes"}


def test_every_compute_entry_point_rejects_null_operands():
    """All 17 compute entry points validate their arguments before the first CUDA call: with every pointer NULL and every
    size 0 they must return a negative code and leave a message - on any box, GPU or not - never crash or launch."""
    lib = _lib.lib()
    compute = [n for n in _lib.exported_symbols() if n not in HELPERS]
    assert len(compute) >= 17
    for name in compute:
        fn = getattr(lib, name)
        assert fn.argtypes is not None, f"{name}: no ctypes signature declared in _lib.py"
        args = []
        for t in fn.argtypes:
            is_ptr = t in (C.c_void_p, C.c_char_p) or isinstance(t, type(C.POINTER(C.c_void_p)))
            args.append(None if is_ptr else (t(0.0) if t is C.c_float else t(0)))
        rc = fn(*args)
        assert rc < 0, f"{name} accepted null operands (rc={rc})"
        assert len(lib.b200_last_error()) > 0, name

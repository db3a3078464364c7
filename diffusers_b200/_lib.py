"""ctypes binding of libb200diff.so (the C-ABI declared in include/b200_diffusion.h).

There is no fallback: if the shared library is missing or fails to initialise, every op raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libb200diff.so")

DTYPE_BF16, DTYPE_FP16 = 0, 1
ACT_NONE, ACT_SILU, ACT_GELU_ERF, ACT_GELU_TANH, ACT_QUICK_GELU = 0, 1, 2, 3, 4


class B200Error(RuntimeError):
    pass


class ConvGemmArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p * 2), ("c", C.c_int32 * 2), ("ldx", C.c_int32 * 2),
        ("batch", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("ksize", C.c_int32), ("stride", C.c_int32),
        ("w", C.c_void_p), ("N", C.c_int32),
        ("bias", C.c_void_p), ("act", C.c_int32), ("geglu", C.c_int32),
        ("gate", C.c_void_p), ("rowvec", C.c_void_p), ("ld_gate", C.c_int32), ("ld_rowvec", C.c_int32),
        ("rows_per_group", C.c_int32),
        ("residual", C.c_void_p), ("ldr", C.c_int32),
        ("y", C.c_void_p), ("ldy", C.c_int32),
        ("dtype", C.c_int32), ("tile_n", C.c_int32), ("out_fp32", C.c_int32), ("cluster_m", C.c_int32),
        ("debug_timestamps", C.c_void_p), ("prefetch", C.c_void_p), ("prefetch_bytes", C.c_int64),
        ("row_stats_out", C.c_void_p), ("ln_stats", C.c_void_p), ("ln_parts", C.c_int32), ("ln_eps", C.c_float),
        ("up2x_parity", C.c_int32), ("pad_after_only", C.c_int32),
        ("qk_cols", C.c_int32), ("qk_head_dim", C.c_int32), ("qk_norm_w", C.c_void_p), ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
        ("rope_ld", C.c_int32), ("rope_row0", C.c_int32), ("qk_eps", C.c_float),
        ("y_peers", C.c_void_p * 8), ("y_block_cols", C.c_int32),
    ]


class AttentionArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
        ("batch", C.c_int32), ("heads", C.c_int32), ("sq", C.c_int32), ("sk", C.c_int32), ("head_dim", C.c_int32),
        ("q_row_stride", C.c_int64), ("q_batch_stride", C.c_int64),
        ("k_row_stride", C.c_int64), ("k_batch_stride", C.c_int64),
        ("v_row_stride", C.c_int64), ("v_batch_stride", C.c_int64),
        ("o_row_stride", C.c_int64), ("o_batch_stride", C.c_int64),
        ("scale", C.c_float), ("dtype", C.c_int32), ("nq_override", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("o_seg_rows", C.c_int32), ("o_seg", C.c_void_p * 8),
    ]


class GroupNormArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p * 2), ("c", C.c_int32 * 2), ("ldx", C.c_int32 * 2),
        ("batch", C.c_int32), ("hw", C.c_int32), ("groups", C.c_int32), ("eps", C.c_float),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("act", C.c_int32),
        ("y", C.c_void_p), ("ldy", C.c_int32),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
        ("dtype", C.c_int32),
    ]


class LayerNormArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32), ("eps", C.c_float),
        ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("scale", C.c_void_p), ("shift", C.c_void_p), ("ld_mod", C.c_int32), ("rows_per_group", C.c_int32),
        ("y", C.c_void_p), ("ldy", C.c_int32), ("dtype", C.c_int32), ("rms", C.c_int32),
    ]


class TextAttentionArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
        ("batch", C.c_int32), ("heads", C.c_int32), ("sq", C.c_int32), ("sk", C.c_int32),
        ("q_row_stride", C.c_int64), ("q_batch_stride", C.c_int64),
        ("k_row_stride", C.c_int64), ("k_batch_stride", C.c_int64),
        ("v_row_stride", C.c_int64), ("v_batch_stride", C.c_int64),
        ("o_row_stride", C.c_int64), ("o_batch_stride", C.c_int64),
        ("scale", C.c_float), ("causal", C.c_int32), ("bias", C.c_void_p), ("dtype", C.c_int32),
    ]


class QkNormRopeArgs(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("ld", C.c_int64),
        ("rows", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32), ("k_off", C.c_int32),
        ("txt_rows", C.c_int32), ("seq", C.c_int32),
        ("wq", C.c_void_p), ("wk", C.c_void_p), ("wq_txt", C.c_void_p), ("wk_txt", C.c_void_p),
        ("cos_table", C.c_void_p), ("sin_table", C.c_void_p), ("eps", C.c_float), ("dtype", C.c_int32),
        ("txt_period", C.c_int32),
    ]


class SmallLinearArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ldx", C.c_int32), ("M", C.c_int32), ("K", C.c_int32),
        ("w", C.c_void_p), ("N", C.c_int32), ("bias", C.c_void_p),
        ("act_in", C.c_int32), ("act_out", C.c_int32),
        ("addend", C.c_void_p), ("ld_add", C.c_int32),
        ("y", C.c_void_p), ("ldy", C.c_int32), ("dtype", C.c_int32),
    ]


_lib = None
_inited_devices = set()


def library_path():
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(
                f"{LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU / PyTorch fallback for the sm_100a kernels)")
        _lib = C.CDLL(LIB_PATH)
        _lib.b200_last_error.restype = C.c_char_p
        _lib.b200_conv_gemm_packed_k.restype = C.c_int64
        _lib.b200_conv_gemm_packed_k.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        _lib.b200_conv_gemm_pick_tile_n.restype = C.c_int32
        _lib.b200_conv_gemm_pick_tile_n.argtypes = [C.c_int64, C.c_int32, C.c_int32]
        _lib.b200_conv_gemm_row_stats_parts.restype = C.c_int32
        _lib.b200_conv_gemm_row_stats_parts.argtypes = [C.c_void_p]
        _lib.b200_group_norm_launches.restype = C.c_int32
        _lib.b200_group_norm_launches.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32]
        _lib.b200_attention_workspace_bytes.restype = C.c_int64
        _lib.b200_attention_workspace_bytes.argtypes = [C.c_int32] * 5
        _lib.b200_group_norm_workspace_bytes.restype = C.c_int64
        _lib.b200_group_norm_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        I32, I64, F32, VP = C.c_int32, C.c_int64, C.c_float, C.c_void_p
        _lib.b200_nchw_to_nhwc.argtypes = [VP, VP, I32, I32, I32, I32, I32, VP]
        _lib.b200_nhwc_to_nchw.argtypes = [VP, I32, VP, I32, I32, I32, I32, VP]
        _lib.b200_upsample_nearest2x.argtypes = [VP, I32, VP, I32, I32, I32, I32, I32, I32, VP]
        _lib.b200_timestep_embedding.argtypes = [VP, I32, VP, I32, I32, I32, F32, F32, F32, I32, VP]
        _lib.b200_euler_step.argtypes = [VP, VP, VP, I64, F32, F32, I32, VP]
        _lib.b200_scale.argtypes = [VP, VP, I64, F32, I32, VP]
        _lib.b200_cfg_euler_step.argtypes = [VP, I32, VP, VP, I32, I32, I32, I32, F32, I32, F32, F32, I32, VP]
        _lib.b200_flow_match_step.argtypes = [VP, VP, VP, I64, F32, F32, I32, VP]
        _lib.b200_linear_step.argtypes = [VP, VP, VP, VP, VP, I64, F32, F32, F32, F32, I32, VP]
        _lib.b200_ddpm_step.argtypes = [VP, VP, VP, VP, I64, F32, F32, F32, F32, F32, I32, F32, I32, VP]
        _lib.b200_softmax_rows.argtypes = [VP, I64, VP, I64, I32, I32, F32, I32, VP]
        _lib.b200_transpose_16.argtypes = [VP, I64, VP, I64, I32, I32, VP]
        _lib.b200_peer_alloc.argtypes = [I64, C.POINTER(VP), C.c_char_p]
        _lib.b200_peer_open.argtypes = [C.c_char_p, C.POINTER(VP)]
        _lib.b200_peer_close.argtypes = [VP]
        _lib.b200_peer_free.argtypes = [VP]
        _lib.b200_peer_barrier.argtypes = [C.POINTER(VP), VP, I32, I32, VP]
        _lib.b200_text_attention.argtypes = [VP, VP]
        for fn in ("b200_conv_gemm", "b200_attention", "b200_group_norm", "b200_layer_norm", "b200_small_linear",
                   "b200_qk_norm_rope"):
            getattr(_lib, fn).argtypes = [VP, VP]
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().b200_last_error()
        raise B200Error(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def init(device_index):
    """b200_init once per device; raises if the device is not a B200 or the driver lacks TMA support."""
    if device_index not in _inited_devices:
        check(lib().b200_init(C.c_int(device_index)), "b200_init")
        _inited_devices.add(device_index)


def exported_symbols():
    """Symbols the header promises (used by the CPU-side ABI test)."""
    hdr = os.path.join(os.path.dirname(_HERE), "include", "b200_diffusion.h")
    import re
    names = []
    with open(hdr) as f:
        for m in re.finditer(r"^\s*(?:const\s+char\*|int(?:32_t|64_t)?)\s+(b200_\w+)\s*\(", f.read(), re.M):
            names.append(m.group(1))
    return names

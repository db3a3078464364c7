"""Tensor-level wrappers over the C-ABI.  Every function enqueues on torch's current CUDA stream and
returns immediately; none of them has a PyTorch fallback."""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_GELU_ERF, ACT_GELU_TANH, ACT_NONE, ACT_SILU, DTYPE_BF16, DTYPE_FP16,  # noqa: F401
                   B200Error)

_LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)


def launches():
    return _LAUNCHES


def _count(n=1):
    global _LAUNCHES
    _LAUNCHES += n


def _dtype_code(t):
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    if t.dtype == torch.float16:
        return DTYPE_FP16
    raise B200Error(f"unsupported dtype {t.dtype}: the sm_100a kernels take bf16 or fp16")


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_cuda(t, name):
    if not t.is_cuda:
        raise B200Error(f"{name} must be a CUDA tensor (no CPU fallback)")
    _lib.init(t.device.index if t.device.index is not None else torch.cuda.current_device())


def conv_gemm(x, w, N, *, batch, H, W, ksize=1, stride=1, x2=None, bias=None, act=ACT_NONE, geglu=False,
              gate=None, rowvec=None, rows_per_group=0, residual=None, out=None, tile_n=0):
    """y = epilogue(conv/linear(x [, x2]))  -  see b200_conv_gemm in include/b200_diffusion.h.

    x, x2: NHWC activations given as 2-D [batch*H*W, C] (or any shape whose last dim is C, contiguous rows).
    w: packed [N, Kp] (packing.pack_conv_weight / pack_linear_weight / pack_geglu)."""
    _need_cuda(x, "x")
    c0 = x.shape[-1]
    c1 = x2.shape[-1] if x2 is not None else 0
    Ho, Wo = (H, W) if stride == 1 else (H // 2, W // 2)
    n_out = N // 2 if geglu else N
    if out is None:
        out = torch.empty((batch * Ho * Wo, n_out), dtype=x.dtype, device=x.device)
    a = _lib.ConvGemmArgs()
    a.x[0] = x.data_ptr()
    a.x[1] = _ptr(x2)
    a.c[0], a.c[1] = c0, c1
    a.ldx[0] = x.stride(-2)
    a.ldx[1] = x2.stride(-2) if x2 is not None else 0
    a.batch, a.H, a.W = batch, H, W
    a.ksize, a.stride = ksize, stride
    a.w, a.N = w.data_ptr(), N
    a.bias = _ptr(bias)
    a.act, a.geglu = act, 1 if geglu else 0
    a.gate, a.rowvec = _ptr(gate), _ptr(rowvec)
    a.ld_gate = gate.stride(-2) if gate is not None else 0
    a.ld_rowvec = rowvec.stride(-2) if rowvec is not None else 0
    a.rows_per_group = rows_per_group
    a.residual = _ptr(residual)
    a.ldr = residual.stride(-2) if residual is not None else 0
    a.y, a.ldy = out.data_ptr(), out.stride(-2)
    a.dtype = _dtype_code(x)
    a.tile_n = tile_n
    _lib.check(_lib.lib().b200_conv_gemm(C.byref(a), _stream()), "b200_conv_gemm")
    _count()
    return out


def linear(x, w, N, *, bias=None, act=ACT_NONE, geglu=False, gate=None, rowvec=None, rows_per_group=0,
           residual=None, x2=None, out=None, tile_n=0):
    """nn.Linear on token rows: x [rows, K] (row stride arbitrary multiple of 8)."""
    rows = x.shape[0]
    return conv_gemm(x, w, N, batch=1, H=1, W=rows, ksize=1, stride=1, x2=x2, bias=bias, act=act, geglu=geglu,
                     gate=gate, rowvec=rowvec, rows_per_group=rows_per_group, residual=residual, out=out,
                     tile_n=tile_n)


def pick_tile_n(M, N, geglu=False):
    return int(_lib.lib().b200_conv_gemm_pick_tile_n(M, N, 1 if geglu else 0))

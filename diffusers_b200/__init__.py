"""diffusers_b200 - sm_100a (B200) kernels and drop-in model shells for the diffusers denoising hot path.

Layout (tier: one hot path):
  csrc/        hand-written CUDA for sm_100a + the C-ABI (libb200diff.so, include/b200_diffusion.h)
  _lib.py      ctypes binding (no fallback: raises when the library is missing)
  ops.py       tensor-level wrappers over the C-ABI (pointers, strides, current CUDA stream)
  packing.py   weight re-layout (OIHW -> K-major tap/channel packing, fused QKV, GEGLU interleave)
"""
__version__ = "0.1.0"

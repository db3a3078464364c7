"""Tensor-level wrappers over the C-ABI.  Every function enqueues on torch's current CUDA stream and
returns immediately; none of them has a PyTorch fallback."""
import ctypes as C

import torch

from . import _lib
from ._lib import (ACT_GELU_ERF, ACT_GELU_TANH, ACT_NONE, ACT_QUICK_GELU, ACT_SILU, DTYPE_BF16, DTYPE_FP16,  # noqa: F401
                   B200Error)

_LAUNCHES = 0  # kernels launched through this module (bench.py reports it as gpu_launches)
_PROFILE = None  # bench.py: list collecting (start_event, end_event, algorithmic_flops, kernel, shape) per conv_gemm / attention launch


def launches():
    return _LAUNCHES


class WeightPrefetchPlan:
    """Launch-order trace of the packed weights one model forward feeds to b200_conv_gemm.

    SDXL's 5 GB of weights never survive in the 126 MB L2 from one forward to the next, so a GEMM would open with
    DRAM-latency-bound weight fetches.  Once a forward has been traced, every launch hands the kernel the weights of
    the launch that FOLLOWS it (b200_conv_gemm_args.prefetch) and the kernel pulls them into L2 in the background.
    A forward whose launch order differs from the trace simply re-records (the hint is never needed for
    correctness)."""
    MAX_BYTES = 48 << 20  # per launch; the head of a bigger weight is what its first tiles read

    def __init__(self):
        self.seq = []
        self._rec = None
        self._i = 0

    def _begin(self):
        self._rec = []
        self._i = 0

    def _step(self, w):
        key = (w.data_ptr(), min(w.numel() * w.element_size(), self.MAX_BYTES))
        self._rec.append(key)
        i = self._i
        self._i += 1
        if i < len(self.seq) and self.seq[i] == key:
            return self.seq[(i + 1) % len(self.seq)]
        return None

    def _end(self):
        if self._rec != self.seq:
            self.seq = self._rec
        self._rec = None


_PLAN = None


class weight_prefetch:
    """`with ops.weight_prefetch(plan):` around one model forward."""

    def __init__(self, plan):
        self.plan = plan

    def __enter__(self):
        global _PLAN
        self._prev = _PLAN
        _PLAN = self.plan
        self.plan._begin()
        return self.plan

    def __exit__(self, *exc):
        global _PLAN
        self.plan._end()
        _PLAN = self._prev
        return False


def prefetching_forward(fn):
    """Decorator for a model's forward/decode: runs it under the model's own WeightPrefetchPlan."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        plan = self.__dict__.get("_weight_prefetch_plan")
        if plan is None:
            plan = self.__dict__["_weight_prefetch_plan"] = WeightPrefetchPlan()
        if _PLAN is plan:  # re-entrant call (e.g. sub-batched decode)
            return fn(self, *args, **kwargs)
        with weight_prefetch(plan):
            return fn(self, *args, **kwargs)

    return wrapped


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
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        # every launch goes to torch.cuda.current_stream(): a stream of the CURRENT device
        raise B200Error(f"{name} lives on cuda:{idx} but the current device is cuda:{torch.cuda.current_device()}: "
                        f"wrap the call in `with torch.cuda.device({idx}):`")
    _lib.init(idx)


class QkRope:
    """Per-head RMSNorm + rotary embedding of the q / k columns of a fused QKV projection, folded into the GEMM epilogue
    (b200_conv_gemm_args.qk_*): `w` 16-bit [2, head_dim] (norm_q.weight, norm_k.weight), `cos` / `sin` fp32
    [head_dim / 2, positions] (rope_tables_transposed), `row0` the position of the launch's first row, `cols` = q + k columns."""
    __slots__ = ("w", "cos", "sin", "row0", "cols", "head_dim", "eps")

    def __init__(self, w, cos, sin, row0, cols, head_dim, eps=1e-6):
        self.w, self.cos, self.sin, self.row0, self.cols, self.head_dim, self.eps = w, cos, sin, row0, cols, head_dim, eps


def rope_tables_transposed(cos, sin):
    """FluxPosEmbed's repeat-interleaved fp32 [positions, head_dim] tables -> the [head_dim / 2, positions] form the fused epilogue reads."""
    return cos[:, 0::2].t().contiguous(), sin[:, 0::2].t().contiguous()


class FoldedLayerNorm:
    """A LayerNorm folded into the Linear that consumes it (b200_conv_gemm_args.ln_*): the row statistics written by the GEMM
    that produced the activations (`row_stats=True`) and eps.  The weight / bias of that Linear come from
    packing.fold_layer_norm."""
    __slots__ = ("stats", "eps")

    def __init__(self, stats, eps):
        self.stats, self.eps = stats, eps


def conv_gemm(x, w, N, *, batch, H, W, ksize=1, stride=1, x2=None, bias=None, act=ACT_NONE, geglu=False,
              gate=None, rowvec=None, rows_per_group=0, residual=None, out=None, tile_n=0, out_fp32=False, cluster_m=0, debug_timestamps=None,
              row_stats=False, ln=None, pad_after_only=False, qk_rope=None, out_blocks=None):
    """y = epilogue(conv/linear(x [, x2]))  -  see b200_conv_gemm in include/b200_diffusion.h.

    x, x2: NHWC activations given as 2-D [batch*H*W, C] (or any shape whose last dim is C, contiguous rows).
    w: packed [N, Kp] (packing.pack_conv_weight / pack_linear_weight / pack_geglu).
    row_stats=True: also returns fp32 [M, parts, 2] partial (sum, sum of squares) of the rounded outputs of each row -> (y, stats).
    ln: a FoldedLayerNorm - w / bias come from packing.fold_layer_norm and the epilogue scales each row by its rstd."""
    _need_cuda(x, "x")
    c0 = x.shape[-1]
    c1 = x2.shape[-1] if x2 is not None else 0
    Ho, Wo = (H, W) if stride == 1 else (H // 2, W // 2)
    n_out = N // 2 if geglu else N
    if out_blocks is not None:
        # context parallelism: column block j of the output goes to out_blocks[j] ([rows, N / len] views of the owners' buffers)
        nb = len(out_blocks)
        if out is not None or nb < 1 or nb > 8 or N % nb or any(tuple(t.shape) != (batch * Ho * Wo, N // nb) or t.stride(0) != out_blocks[0].stride(0) or
                                                                t.stride(1) != 1 for t in out_blocks):
            raise B200Error("conv_gemm: out_blocks must be 1..8 [rows, N / blocks] views with one row stride (and no `out`)")
        out = out_blocks[0]
    elif out is None:
        out = torch.empty((batch * Ho * Wo, n_out), dtype=torch.float32 if out_fp32 else x.dtype, device=x.device)
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
    a.out_fp32 = 1 if out_fp32 else 0
    a.cluster_m = cluster_m
    a.debug_timestamps = _ptr(debug_timestamps)
    a.pad_after_only = 1 if pad_after_only else 0
    if out_blocks is not None:
        a.y_block_cols = N // len(out_blocks)
        for j, t in enumerate(out_blocks):
            a.y_peers[j] = t.data_ptr()
    stats = None
    if row_stats:
        parts = int(_lib.lib().b200_conv_gemm_row_stats_parts(C.byref(a)))
        stats = torch.empty((batch * Ho * Wo, parts, 2), dtype=torch.float32, device=x.device)
        a.row_stats_out = stats.data_ptr()
    if ln is not None:
        a.ln_stats, a.ln_parts, a.ln_eps = ln.stats.data_ptr(), ln.stats.shape[1], ln.eps
    if qk_rope is not None:
        q = qk_rope
        if q.cos.dtype != torch.float32 or q.cos.shape[0] * 2 != q.head_dim or not q.cos.is_contiguous() or q.cos.shape != q.sin.shape or not q.sin.is_contiguous():
            raise B200Error("qk_rope: cos / sin must be contiguous fp32 [head_dim / 2, positions] tables")
        if q.w.dtype != x.dtype or q.w.numel() != 2 * q.head_dim:
            raise B200Error("qk_rope: w must be [2, head_dim] in the activation dtype")
        a.qk_cols, a.qk_head_dim, a.qk_norm_w = q.cols, q.head_dim, q.w.data_ptr()
        a.rope_cos, a.rope_sin, a.rope_ld, a.rope_row0, a.qk_eps = q.cos.data_ptr(), q.sin.data_ptr(), q.cos.shape[1], q.row0, q.eps
    if _PLAN is not None:
        nxt = _PLAN._step(w)
        if nxt is not None:
            a.prefetch, a.prefetch_bytes = nxt
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().b200_conv_gemm(C.byref(a), _stream()), "b200_conv_gemm")
        e1.record()
        M_, K_ = batch * Ho * Wo, ksize * ksize * (c0 + c1)
        _PROFILE.append((e0, e1, 2.0 * M_ * N * K_, "conv_gemm", (M_, N, K_)))
    else:
        _lib.check(_lib.lib().b200_conv_gemm(C.byref(a), _stream()), "b200_conv_gemm")
    _count()
    if out_blocks is not None:
        return None
    return (out, stats) if row_stats else out


def linear(x, w, N, *, bias=None, act=ACT_NONE, geglu=False, gate=None, rowvec=None, rows_per_group=0,
           residual=None, x2=None, out=None, tile_n=0, out_fp32=False, cluster_m=0, debug_timestamps=None, row_stats=False, ln=None, qk_rope=None,
           out_blocks=None):
    """nn.Linear on token rows: x [rows, K] (row stride arbitrary multiple of 8)."""
    rows = x.shape[0]
    return conv_gemm(x, w, N, batch=1, H=1, W=rows, ksize=1, stride=1, x2=x2, bias=bias, act=act, geglu=geglu,
                     gate=gate, rowvec=rowvec, rows_per_group=rows_per_group, residual=residual, out=out,
                     tile_n=tile_n, out_fp32=out_fp32, cluster_m=cluster_m, debug_timestamps=debug_timestamps, row_stats=row_stats, ln=ln,
                     qk_rope=qk_rope, out_blocks=out_blocks)


def upsample2x_conv(x, w4, N, *, batch, H, W, bias=None, act=ACT_NONE, out=None):
    """conv3x3(nearest2x(x)) without the upsampled tensor: x [batch*H*W, C] NHWC -> [batch*2H*2W, N]; w4 = packing.pack_upsample_conv(w).
    Four launches (one per output parity class), each a 2 x 2 convolution over the low-resolution input: 4/9 of the FLOPs."""
    _need_cuda(x, "x")
    if out is None:
        out = torch.empty((batch * 4 * H * W, N), dtype=x.dtype, device=x.device)
    for par in range(4):
        a = _lib.ConvGemmArgs()
        a.x[0], a.c[0], a.ldx[0] = x.data_ptr(), x.shape[-1], x.stride(-2)
        a.batch, a.H, a.W, a.ksize, a.stride = batch, H, W, 2, 1
        a.w, a.N, a.bias, a.act = w4[par].data_ptr(), N, _ptr(bias), act
        a.y, a.ldy, a.dtype = out.data_ptr(), out.stride(-2), _dtype_code(x)
        a.up2x_parity = par + 1
        if _PLAN is not None:
            nxt = _PLAN._step(w4[par])
            if nxt is not None:
                a.prefetch, a.prefetch_bytes = nxt
        if _PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(_lib.lib().b200_conv_gemm(C.byref(a), _stream()), "b200_conv_gemm (upsample parity)")
            e1.record()
            _PROFILE.append((e0, e1, 2.0 * batch * H * W * N * 4 * x.shape[-1], "conv_gemm", (batch * H * W, N, 4 * x.shape[-1])))
        else:
            _lib.check(_lib.lib().b200_conv_gemm(C.byref(a), _stream()), "b200_conv_gemm (upsample parity)")
        _count()
    return out


def pick_tile_n(M, N, geglu=False):
    return int(_lib.lib().b200_conv_gemm_pick_tile_n(M, N, 1 if geglu else 0))


_ATTN_WS = {}


def _attention_workspace(device, B, heads, Sq, Sk, head_dim):
    """Scratch of the tail split (b200_attention_args.workspace): one buffer per (device, stream), zeroed once - the kernel
    leaves its ticket words zero again."""
    need = int(_lib.lib().b200_attention_workspace_bytes(B, heads, Sq, Sk, head_dim))
    if need == 0:
        return None
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    ws = _ATTN_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 16 << 20), dtype=torch.uint8, device=device)
        _ATTN_WS[key] = ws
    return ws


def attention(q, k, v, *, heads, head_dim, scale=None, out=None, nq=0, o_seg=None, o_seg_rows=0):
    """softmax(q k^T * scale) v.  q [B, Sq, heads*head_dim-wide rows], k/v [B, Sk, ...]: 3-D views whose last
    dim starts at this tensor's first head (row stride / batch stride taken from the view, so slices of a fused
    QKV buffer work).  Returns [B, Sq, heads*head_dim].
    o_seg (context parallelism): list of 2-D [o_seg_rows, heads*head_dim]-wide views (same row stride), one per
    o_seg_rows query rows; row r is stored into o_seg[r // o_seg_rows][r % o_seg_rows] and nothing is returned."""
    _need_cuda(q, "q")
    B, Sq = q.shape[0], q.shape[1]
    Sk = k.shape[1]
    a = _lib.AttentionArgs()
    if o_seg is not None:
        if B != 1 or o_seg_rows <= 0 or len(o_seg) * o_seg_rows < Sq or len(o_seg) > 8:
            raise B200Error("attention: o_seg needs batch 1 and ceil(Sq / o_seg_rows) <= 8 segments")
        if any(t.stride(0) != o_seg[0].stride(0) or t.stride(1) != 1 for t in o_seg):
            raise B200Error("attention: o_seg views must share one row stride")
        out = o_seg[0].unsqueeze(0)
        a.o_seg_rows = o_seg_rows
        for i, t in enumerate(o_seg):
            a.o_seg[i] = t.data_ptr()
    elif out is None:
        out = torch.empty((B, Sq, heads * head_dim), dtype=q.dtype, device=q.device)
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.batch, a.heads, a.sq, a.sk, a.head_dim = B, heads, Sq, Sk, head_dim
    a.q_row_stride, a.q_batch_stride = q.stride(1), q.stride(0)
    a.k_row_stride, a.k_batch_stride = k.stride(1), k.stride(0)
    a.v_row_stride, a.v_batch_stride = v.stride(1), v.stride(0)
    a.o_row_stride, a.o_batch_stride = out.stride(1), out.stride(0)
    a.scale = float(scale) if scale is not None else 0.0
    a.dtype = _dtype_code(q)
    a.nq_override = nq
    ws = _attention_workspace(q.device, B, heads, Sq, Sk, head_dim)
    if ws is not None:
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    if _PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().b200_attention(C.byref(a), _stream()), "b200_attention")
        e1.record()
        _PROFILE.append((e0, e1, 4.0 * B * heads * Sq * Sk * head_dim, "attention", (B * heads, Sq, Sk, head_dim)))
    else:
        _lib.check(_lib.lib().b200_attention(C.byref(a), _stream()), "b200_attention")
    _count()
    return None if o_seg is not None else out


def text_attention(q, k, v, *, heads, scale, causal=False, bias=None, out=None):
    """Attention of the text encoders (head_dim 64, <= 512 keys): q/k/v [B, S, heads*64] views, optional causal mask and
    fp32 bias [heads, Sq, Sk] added to the scaled scores.  Returns [B, Sq, heads*64]."""
    _need_cuda(q, "q")
    B, Sq = q.shape[0], q.shape[1]
    Sk = k.shape[1]
    if out is None:
        out = torch.empty((B, Sq, heads * 64), dtype=q.dtype, device=q.device)
    a = _lib.TextAttentionArgs()
    a.q, a.k, a.v, a.o = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.batch, a.heads, a.sq, a.sk = B, heads, Sq, Sk
    a.q_row_stride, a.q_batch_stride = q.stride(1), q.stride(0)
    a.k_row_stride, a.k_batch_stride = k.stride(1), k.stride(0)
    a.v_row_stride, a.v_batch_stride = v.stride(1), v.stride(0)
    a.o_row_stride, a.o_batch_stride = out.stride(1), out.stride(0)
    a.scale, a.causal = float(scale), 1 if causal else 0
    if bias is not None:
        if bias.dtype != torch.float32 or tuple(bias.shape) != (heads, Sq, Sk) or not bias.is_contiguous():
            raise B200Error("text_attention: bias must be a contiguous fp32 [heads, Sq, Sk] tensor")
        a.bias = bias.data_ptr()
    a.dtype = _dtype_code(q)
    _lib.check(_lib.lib().b200_text_attention(C.byref(a), _stream()), "b200_text_attention")
    _count()
    return out


_GN_WS = {}


def _gn_workspace(device, batch, hw, groups):
    need = int(_lib.lib().b200_group_norm_workspace_bytes(batch, hw, groups))
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    ws = _GN_WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=device)  # zero-initialised once (counters)
        _GN_WS[key] = ws
    return ws


def group_norm(x, *, batch, hw, groups, eps, gamma=None, beta=None, silu=False, x2=None, out=None):
    """GroupNorm(+SiLU) over NHWC rows x [batch*hw, C0] (and x2 [batch*hw, C1] concatenated along channels)."""
    _need_cuda(x, "x")
    c0 = x.shape[-1]
    c1 = x2.shape[-1] if x2 is not None else 0
    if out is None:
        out = torch.empty((batch * hw, c0 + c1), dtype=x.dtype, device=x.device)
    ws = _gn_workspace(x.device, batch, hw, groups)
    a = _lib.GroupNormArgs()
    a.x[0], a.x[1] = x.data_ptr(), _ptr(x2)
    a.c[0], a.c[1] = c0, c1
    a.ldx[0], a.ldx[1] = x.stride(-2), (x2.stride(-2) if x2 is not None else 0)
    a.batch, a.hw, a.groups, a.eps = batch, hw, groups, eps
    a.gamma, a.beta = _ptr(gamma), _ptr(beta)
    a.act = ACT_SILU if silu else ACT_NONE
    a.y, a.ldy = out.data_ptr(), out.stride(-2)
    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
    a.dtype = _dtype_code(x)
    _lib.check(_lib.lib().b200_group_norm(C.byref(a), _stream()), "b200_group_norm")
    _count(int(_lib.lib().b200_group_norm_launches(hw, c0 + c1, groups, c0, 1 if x2 is not None else 0)))
    return out


def layer_norm(x, *, eps, gamma=None, beta=None, scale=None, shift=None, rows_per_group=0, out=None, rms=False):
    """LayerNorm over rows of x [rows, C]; optional AdaLN modulation y*(1+scale[g])+shift[g].  rms=True: RMSNorm (no mean)."""
    _need_cuda(x, "x")
    rows, cols = x.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=x.dtype, device=x.device)
    a = _lib.LayerNormArgs()
    a.x, a.ldx, a.rows, a.cols, a.eps = x.data_ptr(), x.stride(0), rows, cols, eps
    a.gamma, a.beta = _ptr(gamma), _ptr(beta)
    a.scale, a.shift = _ptr(scale), _ptr(shift)
    mod = scale if scale is not None else shift
    a.ld_mod = mod.stride(-2) if mod is not None else 0
    a.rows_per_group = rows_per_group
    a.y, a.ldy = out.data_ptr(), out.stride(0)
    a.dtype = _dtype_code(x)
    a.rms = 1 if rms else 0
    _lib.check(_lib.lib().b200_layer_norm(C.byref(a), _stream()), "b200_layer_norm")
    _count()
    return out


def small_linear(x, w, *, bias=None, act_in=ACT_NONE, act_out=ACT_NONE, addend=None, out=None):
    """Linear for M <= 8 rows (weight-bandwidth bound): act_out(act_in(x) @ w.T + bias) (+ addend)."""
    _need_cuda(x, "x")
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=x.dtype, device=x.device)
    for m0 in range(0, M, 8):
        mm = min(8, M - m0)
        a = _lib.SmallLinearArgs()
        xs = x[m0:m0 + mm]
        a.x, a.ldx, a.M, a.K = xs.data_ptr(), x.stride(0), mm, K
        a.w, a.N, a.bias = w.data_ptr(), N, _ptr(bias)
        a.act_in, a.act_out = act_in, act_out
        if addend is not None:
            a.addend, a.ld_add = addend[m0:m0 + mm].data_ptr(), addend.stride(0)
        a.y, a.ldy = out[m0:m0 + mm].data_ptr(), out.stride(0)
        a.dtype = _dtype_code(x)
        _lib.check(_lib.lib().b200_small_linear(C.byref(a), _stream()), "b200_small_linear")
        _count()
    return out


def nchw_to_nhwc(x, c_pad=None, out=None):
    """[B, C, H, W] -> [B*H*W, c_pad] (zero-padded channels)."""
    _need_cuda(x, "x")
    B, Cc, H, W = x.shape
    ld = c_pad or Cc
    x = x.contiguous()
    if out is None:
        out = torch.empty((B * H * W, ld), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().b200_nchw_to_nhwc(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), B, Cc, H * W, ld,
                                            _dtype_code(x), _stream()), "b200_nchw_to_nhwc")
    _count()
    return out


def nhwc_to_nchw(x, *, batch, C_out, H, W, out=None):
    """[B*H*W, ld] (first C_out channels) -> [B, C_out, H, W]."""
    _need_cuda(x, "x")
    if out is None:
        out = torch.empty((batch, C_out, H, W), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().b200_nhwc_to_nchw(C.c_void_p(x.data_ptr()), x.stride(-2), C.c_void_p(out.data_ptr()), batch,
                                            C_out, H * W, _dtype_code(x), _stream()), "b200_nhwc_to_nchw")
    _count()
    return out


def upsample_nearest2x(x, *, batch, H, W, out=None):
    _need_cuda(x, "x")
    Cc = x.shape[-1]
    if out is None:
        out = torch.empty((batch * 4 * H * W, Cc), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().b200_upsample_nearest2x(C.c_void_p(x.data_ptr()), x.stride(-2), C.c_void_p(out.data_ptr()),
                                                  out.stride(-2), batch, H, W, Cc, _dtype_code(x), _stream()),
               "b200_upsample_nearest2x")
    _count()
    return out


def timestep_embedding(t, dim, *, dtype, flip_sin_to_cos, downscale_freq_shift, scale=1.0, max_period=10000.0,
                       out=None):
    """t fp32 [n] (device) -> [n, dim] in `dtype`."""
    _need_cuda(t, "t")
    assert t.dtype == torch.float32 and t.dim() == 1
    n = t.shape[0]
    if out is None:
        out = torch.empty((n, dim), dtype=dtype, device=t.device)
    _lib.check(_lib.lib().b200_timestep_embedding(C.c_void_p(t.data_ptr()), n, C.c_void_p(out.data_ptr()),
                                                  out.stride(0), dim, 1 if flip_sin_to_cos else 0,
                                                  C.c_float(downscale_freq_shift), C.c_float(scale),
                                                  C.c_float(max_period), _dtype_code(out), _stream()),
               "b200_timestep_embedding")
    _count()
    return out


def euler_step(model_output, sample, sigma, sigma_next, out=None):
    _need_cuda(sample, "sample")
    model_output = model_output.contiguous()
    sample = sample.contiguous()
    if sample.dtype != model_output.dtype:
        sample = sample.to(model_output.dtype)
    if out is None:
        out = torch.empty_like(model_output)
    _lib.check(_lib.lib().b200_euler_step(C.c_void_p(model_output.data_ptr()), C.c_void_p(sample.data_ptr()),
                                          C.c_void_p(out.data_ptr()), C.c_int64(sample.numel()), C.c_float(sigma),
                                          C.c_float(sigma_next), _dtype_code(model_output), _stream()),
               "b200_euler_step")
    _count()
    return out


def scale_div(x, divisor, out=None):
    _need_cuda(x, "x")
    x = x.contiguous()
    if out is None:
        out = torch.empty_like(x)
    _lib.check(_lib.lib().b200_scale(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), C.c_int64(x.numel()),
                                     C.c_float(divisor), _dtype_code(x), _stream()), "b200_scale")
    _count()
    return out


def cfg_euler_step(eps_nhwc, latents_nchw, next_in_nhwc, *, guidance_scale, do_cfg, sigma, sigma_next):
    """In place on `latents_nchw` [B, C, H, W]; writes the next UNet input into next_in_nhwc [(2)B*HW, ld]."""
    _need_cuda(latents_nchw, "latents")
    B, Cc, H, W = latents_nchw.shape
    _lib.check(_lib.lib().b200_cfg_euler_step(C.c_void_p(eps_nhwc.data_ptr()), eps_nhwc.stride(-2),
                                              C.c_void_p(latents_nchw.data_ptr()),
                                              C.c_void_p(next_in_nhwc.data_ptr()), next_in_nhwc.stride(-2), B, Cc,
                                              H * W, C.c_float(guidance_scale), 1 if do_cfg else 0, C.c_float(sigma),
                                              C.c_float(sigma_next), _dtype_code(latents_nchw), _stream()),
               "b200_cfg_euler_step")
    _count()


def flow_match_step(model_output, sample, sigma, sigma_next, out=None):
    _need_cuda(sample, "sample")
    model_output = model_output.contiguous()
    sample = sample.contiguous()
    if sample.dtype != model_output.dtype:
        sample = sample.to(model_output.dtype)
    if out is None:
        out = torch.empty_like(model_output)
    _lib.check(_lib.lib().b200_flow_match_step(C.c_void_p(model_output.data_ptr()), C.c_void_p(sample.data_ptr()),
                                               C.c_void_p(out.data_ptr()), C.c_int64(sample.numel()), C.c_float(sigma),
                                               C.c_float(sigma_next), _dtype_code(model_output), _stream()),
               "b200_flow_match_step")
    _count()
    return out


def linear_step(sample, m0=None, m1=None, noise=None, *, a=1.0, b=0.0, c=0.0, s=0.0, out=None):
    """prev = a*sample + b*m0 + c*m1 + s*noise in fp32 with one rounding (b200_linear_step): the update of the DDIM,
    Euler-ancestral and DPM-Solver++ steppers."""
    _need_cuda(sample, "sample")
    sample = sample.contiguous()
    dt = sample.dtype
    prep = lambda t: None if t is None else t.to(dt).contiguous()  # noqa: E731
    m0, m1, noise = prep(m0), prep(m1), prep(noise)
    if out is None:
        out = torch.empty_like(sample)
    _lib.check(_lib.lib().b200_linear_step(C.c_void_p(sample.data_ptr()), C.c_void_p(_ptr(m0)), C.c_void_p(_ptr(m1)), C.c_void_p(_ptr(noise)),
                                           C.c_void_p(out.data_ptr()), C.c_int64(sample.numel()), C.c_float(a), C.c_float(b), C.c_float(c),
                                           C.c_float(s), _dtype_code(sample), _stream()), "b200_linear_step")
    _count()
    return out


def softmax_rows(s, scale, dtype, out=None):
    """softmax(s * scale) over the last dim: s fp32 [rows, cols] -> 16-bit [rows, cols]."""
    _need_cuda(s, "s")
    rows, cols = s.shape
    if out is None:
        out = torch.empty((rows, cols), dtype=dtype, device=s.device)
    _lib.check(_lib.lib().b200_softmax_rows(C.c_void_p(s.data_ptr()), s.stride(0), C.c_void_p(out.data_ptr()),
                                            out.stride(0), rows, cols, C.c_float(scale), _dtype_code(out), _stream()),
               "b200_softmax_rows")
    _count()
    return out


def transpose_16(x, out=None):
    """[R, C] 16-bit -> [C, R]."""
    _need_cuda(x, "x")
    R_, C_ = x.shape
    if out is None:
        out = torch.empty((C_, R_), dtype=x.dtype, device=x.device)
    _lib.check(_lib.lib().b200_transpose_16(C.c_void_p(x.data_ptr()), x.stride(0), C.c_void_p(out.data_ptr()),
                                            out.stride(0), R_, C_, _stream()), "b200_transpose_16")
    _count()
    return out


def attention_unfused(q, k, v, *, scale):
    """Single-head attention for head dims the fused kernel does not cover (AutoencoderKL: 512).
    q [Sq, D], k/v [Sk, D] (row strides arbitrary multiples of 8) -> [Sq, D].

    Both GEMMs hand an activation to b200_conv_gemm as its "weight" operand, whose rows the kernel reads with a stride of
    rup(K, 64) elements (the packed-weight layout): K = D for Q K^T and K = Sk for P V.  Operands whose K is not a multiple
    of 64 (e.g. an 800x800 image: Sk = 10000) are therefore staged in zero-padded buffers of that stride."""
    Sq, D = q.shape
    Sk = k.shape[0]
    if Sk % 8 or D % 8:
        raise B200Error(f"attention_unfused: Sk={Sk} and D={D} must be multiples of 8 (16-byte rows for TMA)")
    Dp, Skp = (D + 63) // 64 * 64, (Sk + 63) // 64 * 64
    if k.stride(0) == Dp and k.stride(1) == 1:
        kc = k
    else:
        kc = torch.zeros((Sk, Dp), dtype=k.dtype, device=k.device) if Dp != D else torch.empty((Sk, Dp), dtype=k.dtype, device=k.device)
        kc[:, :D] = k
    s = linear(q, kc, Sk, out_fp32=True)
    p = softmax_rows(s, scale, q.dtype)
    # [D, Sk] K-major "weights" for the P @ V GEMM, row stride rup(Sk, 64)
    vt = torch.zeros((D, Skp), dtype=v.dtype, device=v.device) if Skp != Sk else torch.empty((D, Skp), dtype=v.dtype, device=v.device)
    transpose_16(v, out=vt[:, :Sk])
    return linear(p, vt, D)


def qk_norm_rope(qkv, *, heads, head_dim, k_off, seq, txt_rows=0, txt_period=0, wq=None, wk=None, wq_txt=None, wk_txt=None,
                 cos=None, sin=None, eps=1e-6):
    """In place on qkv [rows, ld]: RMSNorm(q), RMSNorm(k) per head (+ weight), then rotary embedding."""
    _need_cuda(qkv, "qkv")
    a = _lib.QkNormRopeArgs()
    a.qkv, a.ld = qkv.data_ptr(), qkv.stride(0)
    a.rows, a.heads, a.head_dim, a.k_off = qkv.shape[0], heads, head_dim, k_off
    a.txt_rows, a.seq, a.txt_period = txt_rows, seq, txt_period
    a.wq, a.wk, a.wq_txt, a.wk_txt = _ptr(wq), _ptr(wk), _ptr(wq_txt), _ptr(wk_txt)
    a.cos_table, a.sin_table = _ptr(cos), _ptr(sin)
    a.eps = eps
    a.dtype = _dtype_code(qkv)
    _lib.check(_lib.lib().b200_qk_norm_rope(C.byref(a), _stream()), "b200_qk_norm_rope")
    _count()
    return qkv


def ddpm_step(model_output, sample, noise, *, sqrt_beta_prod, sqrt_alpha_prod, c0, c1, sigma, clip, clip_range, out=None):
    _need_cuda(sample, "sample")
    model_output, sample = model_output.contiguous(), sample.contiguous()
    if out is None:
        out = torch.empty_like(model_output)
    _lib.check(_lib.lib().b200_ddpm_step(C.c_void_p(model_output.data_ptr()), C.c_void_p(sample.data_ptr()),
                                         C.c_void_p(_ptr(noise.contiguous()) if noise is not None else None),
                                         C.c_void_p(out.data_ptr()), C.c_int64(sample.numel()), C.c_float(sqrt_beta_prod),
                                         C.c_float(sqrt_alpha_prod), C.c_float(c0), C.c_float(c1), C.c_float(sigma),
                                         1 if clip else 0, C.c_float(clip_range), _dtype_code(model_output), _stream()),
               "b200_ddpm_step")
    _count()
    return out

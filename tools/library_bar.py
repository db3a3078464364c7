"""The "existing sm_100 kernel" bar: this repo's kernels next to the library kernels the reference's CUDA-eager
path would call on the same B200 for the same shapes - cuBLASLt (F.linear), cuDNN convolution (F.conv2d, channels_last),
SDPA cuDNN / flash backends (models/attention_dispatch.py:3742, :3453), flash_attn 2.8 (FA2 kernels recompiled for sm_100),
torch's GroupNorm / LayerNorm kernels.

    python tools/library_bar.py [--json gpurun_out/library_bar.json]

Timing: 16 launches of one op inside one CUDA graph over ROTATING operand sets (weights of one op never sit in L2 for the
next one, as inside a real forward), CUDA events around one replay, best of 3.  Never run under a profiler.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from diffusers_b200 import ops, packing

DEV, DT = "cuda", torch.bfloat16
G = torch.Generator(device=DEV).manual_seed(0)
N_LAUNCH = 16


def rnd(*s, sc=1.0):
    return (torch.randn(*s, generator=G, device=DEV) * sc).to(DT)


def timed(fn, n=N_LAUNCH):
    """fn(i) enqueues launch i; returns microseconds per launch inside a CUDA graph."""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(n):
            fn(i)
    gr.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, 1000.0 * e0.elapsed_time(e1) / n)
    return best


ROWS = []


def row(op, shape, flops, ours_us, libs, note=""):
    r = dict(op=op, shape=shape, ours_us=round(ours_us, 2), ours_tflops=round(flops / ours_us / 1e6, 1) if flops else None)
    for name, us in libs.items():
        r[name + "_us"] = None if us is None else round(us, 2)
        if us is not None and flops:
            r[name + "_tflops"] = round(flops / us / 1e6, 1)
    best = min([u for u in libs.values() if u is not None], default=None)
    r["ours_vs_best_library"] = None if best is None else round(best / ours_us, 3)  # > 1: this repo's kernel is faster
    if note:
        r["note"] = note
    ROWS.append(r)
    print(json.dumps(r), flush=True)


def safe(fn):
    try:
        return timed(fn)
    except Exception as e:  # noqa: BLE001
        print(f"# library call failed: {type(e).__name__}: {str(e)[:200]}", flush=True)
        return None


def gemm_rows():
    R = 8  # rotating copies
    for name, M, N, K, res, geglu in (("linear+bias+residual (attn out-proj)", 2048, 1280, 1280, True, False),
                                      ("linear (QKV)", 2048, 3840, 1280, False, False),
                                      ("linear+bias+GEGLU (FF in)", 2048, 10240, 1280, False, True),
                                      ("linear+bias+residual (FF out)", 2048, 1280, 5120, True, False),
                                      ("linear+bias+residual (64x64 level)", 8192, 640, 640, True, False),
                                      ("linear+bias (Flux QKV)", 4608, 9216, 3072, False, False),
                                      ("linear+bias (Flux MLP in, gelu-tanh)", 4608, 12288, 3072, False, False)):
        xs = [rnd(M, K) for _ in range(R)]
        ws = [rnd(N, K, sc=K ** -0.5) for _ in range(R)]
        b = rnd(N)
        rs = [rnd(M, N if not geglu else N // 2) for _ in range(R)] if res else None
        if geglu:
            tn = ops.pick_tile_n(M, N, True)
            packed = [packing.pack_geglu(w, b, tn) for w in ws]
            ours = timed(lambda i: ops.linear(xs[i % R], packed[i % R][0], N, bias=packed[i % R][1], geglu=True, tile_n=tn))

            def lib(i):
                h = F.linear(xs[i % R], ws[i % R], b)
                v, g = h.chunk(2, dim=-1)
                return v * F.gelu(g)
            libs = dict(cublaslt_plus_eager_geglu=safe(lib), cublaslt_gemm_only=safe(lambda i: F.linear(xs[i % R], ws[i % R], b)))
        else:
            pw = [packing.pack_linear_weight(w) for w in ws]
            ours = timed(lambda i: ops.linear(xs[i % R], pw[i % R], N, bias=b, residual=rs[i % R] if res else None))
            libs = dict(cublaslt_gemm_only=safe(lambda i: F.linear(xs[i % R], ws[i % R], b)))
            if res:
                libs["cublaslt_plus_residual_add"] = safe(lambda i: F.linear(xs[i % R], ws[i % R], b) + rs[i % R])
        row(name, f"{M}x{N}x{K}", 2.0 * M * N * K, ours, libs)
        del xs, ws, rs
        torch.cuda.empty_cache()


def ksweep_rows():
    """Fixed cost per launch vs cost per K chunk: the one-wave 2048 x 1280 GEMM at growing K (intercept = launch + prologue +
    first-load latency + epilogue, slope = main loop), with and without the epilogue operands, beside cuBLASLt."""
    R = 8
    M, N = 2048, 1280
    for K in (64, 320, 640, 1280, 2560, 5120):
        xs = [rnd(M, K) for _ in range(R)]
        ws = [rnd(N, K, sc=K ** -0.5) for _ in range(R)]
        pw = [packing.pack_linear_weight(w) for w in ws]
        b = rnd(N)
        rs = [rnd(M, N) for _ in range(R)]
        plain = timed(lambda i: ops.linear(xs[i % R], pw[i % R], N))
        bias = timed(lambda i: ops.linear(xs[i % R], pw[i % R], N, bias=b))
        full = timed(lambda i: ops.linear(xs[i % R], pw[i % R], N, bias=b, residual=rs[i % R]))
        libs = dict(cublaslt_no_bias=safe(lambda i: F.linear(xs[i % R], ws[i % R])), cublaslt_bias=safe(lambda i: F.linear(xs[i % R], ws[i % R], b)))
        row("ksweep linear+bias+residual", f"{M}x{N}x{K}", 2.0 * M * N * K, full, libs, note=f"ours without operands {plain:.2f} us, bias only {bias:.2f} us")
        del xs, ws, pw, rs
        torch.cuda.empty_cache()


def lnfold_rows():
    """The folded-LayerNorm variants of the three dominant transformer GEMMs next to their plain forms and to the LayerNorm
    kernel they replace (this repo only: the reference has no such fusion)."""
    R = 8
    M, Cc = 2048, 1280
    xs = [rnd(M, Cc) for _ in range(R)]
    rs = [rnd(M, Cc) for _ in range(R)]
    b = rnd(Cc)
    gam, bet = rnd(Cc), rnd(Cc)
    wo = [packing.pack_linear_weight(rnd(Cc, Cc, sc=Cc ** -0.5)) for _ in range(R)]
    plain = timed(lambda i: ops.linear(xs[i % R], wo[i % R], Cc, bias=b, residual=rs[i % R]))
    with_stats = timed(lambda i: ops.linear(xs[i % R], wo[i % R], Cc, bias=b, residual=rs[i % R], row_stats=True))
    ln = timed(lambda i: ops.layer_norm(xs[i % R], eps=1e-5, gamma=gam, beta=bet))
    _, st = ops.linear(xs[0], wo[0], Cc, bias=b, residual=rs[0], row_stats=True)
    for name, N, geglu in (("QKV 2048x3840x1280", 3840, False), ("attn2.to_q 2048x1280x1280", 1280, False), ("FF in + GEGLU 2048x10240x1280", 10240, True)):
        ws = [rnd(N, Cc, sc=Cc ** -0.5) for _ in range(R)]
        bn = rnd(N)
        if geglu:
            tn = ops.pick_tile_n(M, N, True)
            pk = [packing.pack_geglu(w, bn, tn) for w in ws]
            f_plain = lambda i: ops.linear(xs[i % R], pk[i % R][0], N, bias=pk[i % R][1], geglu=True, tile_n=tn)  # noqa: E731
            f_fold = lambda i: ops.linear(xs[i % R], pk[i % R][0], N, bias=pk[i % R][1], geglu=True, tile_n=tn, ln=ops.FoldedLayerNorm(st, 1e-5))  # noqa: E731
        else:
            pw = [packing.pack_linear_weight(w) for w in ws]
            f_plain = lambda i: ops.linear(xs[i % R], pw[i % R], N, bias=bn)  # noqa: E731
            f_fold = lambda i: ops.linear(xs[i % R], pw[i % R], N, bias=bn, ln=ops.FoldedLayerNorm(st, 1e-5))  # noqa: E731
        r = dict(op="folded LayerNorm: " + name, plain_gemm_us=round(timed(f_plain), 2), folded_gemm_us=round(timed(f_fold), 2),
                 layer_norm_kernel_us=round(ln, 2), producer_plain_us=round(plain, 2), producer_with_row_stats_us=round(with_stats, 2))
        r["saved_us_per_layer"] = round(r["plain_gemm_us"] + r["layer_norm_kernel_us"] + r["producer_plain_us"] - r["folded_gemm_us"] - r["producer_with_row_stats_us"], 2)
        ROWS.append(r)
        print(json.dumps(r), flush=True)


def conv_rows():
    R = 4
    for name, B, C, N, H in (("conv3x3 320->320 @128^2 (SDXL)", 2, 320, 320, 128), ("conv3x3 640->640 @64^2 (SDXL)", 2, 640, 640, 64),
                             ("conv3x3 1280->1280 @32^2 (SDXL)", 2, 1280, 1280, 32), ("conv3x3 128->128 @1024^2 (VAE)", 1, 128, 128, 1024),
                             ("conv3x3 256->256 @512^2 (VAE)", 1, 256, 256, 512), ("conv3x3 512->512 @256^2 (VAE)", 1, 512, 512, 256)):
        xs = [rnd(B, H, H, C) for _ in range(R)]  # NHWC storage
        ws = [rnd(N, C, 3, 3, sc=(9 * C) ** -0.5) for _ in range(R)]
        b = rnd(N)
        pw = [packing.pack_conv_weight(w) for w in ws]
        ours = timed(lambda i: ops.conv_gemm(xs[i % R].view(-1, C), pw[i % R], N, batch=B, H=H, W=H, ksize=3, bias=b))
        xcl = [x.permute(0, 3, 1, 2) for x in xs]  # NCHW view of channels_last memory: cuDNN's NHWC kernels
        wcl = [w.contiguous(memory_format=torch.channels_last) for w in ws]
        lib = safe(lambda i: F.conv2d(xcl[i % R], wcl[i % R], b, padding=1))
        row(name, f"B{B} {C}->{N} {H}x{H}", 2.0 * B * H * H * N * 9 * C, ours, dict(cudnn_channels_last=lib))
        del xs, ws, pw, xcl, wcl
        torch.cuda.empty_cache()


def attention_rows():
    from torch.nn.attention import SDPBackend, sdpa_kernel
    try:
        from flash_attn import flash_attn_func
    except Exception:  # noqa: BLE001
        flash_attn_func = None
    R = 4
    for name, B, H, Sq, Sk, D in (("self-attention 4096 tokens, 10 heads, d64 (SDXL)", 2, 10, 4096, 4096, 64),
                                  ("self-attention 1024 tokens, 20 heads, d64 (SDXL)", 2, 20, 1024, 1024, 64),
                                  ("cross-attention 4096x77, d64 (SDXL)", 2, 10, 4096, 77, 64),
                                  ("cross-attention 1024x77, d64 (SDXL)", 2, 20, 1024, 77, 64),
                                  ("joint attention 4608 tokens, 24 heads, d128 (Flux)", 1, 24, 4608, 4608, 128)):
        qs = [rnd(B, Sq, H * D) for _ in range(R)]
        ks = [rnd(B, Sk, H * D) for _ in range(R)]
        vs = [rnd(B, Sk, H * D) for _ in range(R)]
        ours = timed(lambda i: ops.attention(qs[i % R], ks[i % R], vs[i % R], heads=H, head_dim=D))
        bshd = lambda t, S: t.view(B, S, H, D)  # noqa: E731
        bhsd = lambda t, S: t.view(B, S, H, D).transpose(1, 2)  # noqa: E731  (what AttnProcessor2_0 feeds SDPA)

        def sdpa(backend):
            def f(i):
                with sdpa_kernel(backend):
                    return F.scaled_dot_product_attention(bhsd(qs[i % R], Sq), bhsd(ks[i % R], Sk), bhsd(vs[i % R], Sk))
            return f
        libs = dict(sdpa_cudnn=safe(sdpa(SDPBackend.CUDNN_ATTENTION)), sdpa_flash=safe(sdpa(SDPBackend.FLASH_ATTENTION)),
                    sdpa_default=safe(lambda i: F.scaled_dot_product_attention(bhsd(qs[i % R], Sq), bhsd(ks[i % R], Sk), bhsd(vs[i % R], Sk))))
        if flash_attn_func is not None:
            libs["flash_attn_2_8"] = safe(lambda i: flash_attn_func(bshd(qs[i % R], Sq), bshd(ks[i % R], Sk), bshd(vs[i % R], Sk)))
        row(name, f"B{B} H{H} {Sq}x{Sk} d{D}", 4.0 * B * H * Sq * Sk * D, ours, libs)


def norm_rows():
    R = 8
    for name, B, HW, C in (("GroupNorm(32)+SiLU 320ch @128^2", 2, 16384, 320), ("GroupNorm(32)+SiLU 640ch @64^2", 2, 4096, 640),
                           ("GroupNorm(32)+SiLU 1280ch @64^2", 2, 4096, 1280), ("GroupNorm(32)+SiLU 1280ch @32^2", 2, 1024, 1280),
                           ("GroupNorm(32)+SiLU 2560ch @32^2", 2, 1024, 2560), ("GroupNorm(32)+SiLU 128ch @1024^2 (VAE)", 1, 1 << 20, 128)):
        xs = [rnd(B * HW, C) for _ in range(R)]
        gam, bet = rnd(C), rnd(C)
        ours = timed(lambda i: ops.group_norm(xs[i % R], batch=B, hw=HW, groups=32, eps=1e-5, gamma=gam, beta=bet, silu=True))
        side = int(HW ** 0.5)
        xcl = [x.view(B, side, side, C).permute(0, 3, 1, 2) for x in xs]
        lib = safe(lambda i: F.silu(F.group_norm(xcl[i % R], 32, gam, bet, 1e-5)))
        r = dict(op=name, shape=f"B{B} {HW}x{C}", ours_us=round(ours, 2), ours_gbs=round(2 * B * HW * C * 2 / ours / 1e3, 1),
                 torch_group_norm_plus_silu_us=None if lib is None else round(lib, 2),
                 ours_vs_best_library=None if lib is None else round(lib / ours, 3))
        ROWS.append(r)
        print(json.dumps(r), flush=True)
        del xs, xcl
        torch.cuda.empty_cache()
    for name, rows_, C in (("LayerNorm 1280 (SDXL 32^2 level)", 2048, 1280), ("LayerNorm 640 (SDXL 64^2 level)", 8192, 640),
                           ("LayerNorm 3072 no affine (Flux)", 4608, 3072)):
        xs = [rnd(rows_, C) for _ in range(R)]
        gam, bet = rnd(C), rnd(C)
        ours = timed(lambda i: ops.layer_norm(xs[i % R], eps=1e-5, gamma=gam, beta=bet))
        lib = safe(lambda i: F.layer_norm(xs[i % R], (C,), gam, bet, 1e-5))
        r = dict(op=name, shape=f"{rows_}x{C}", ours_us=round(ours, 2), ours_gbs=round(2 * rows_ * C * 2 / ours / 1e3, 1),
                 torch_layer_norm_us=None if lib is None else round(lib, 2), ours_vs_best_library=None if lib is None else round(lib / ours, 3))
        ROWS.append(r)
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--only", default="gemm,conv,attention,norm,lnfold")
    a = ap.parse_args()
    torch.backends.cudnn.benchmark = True  # let cuDNN pick its best convolution algorithm: the fair bar
    print(f"# {torch.cuda.get_device_name(0)}; torch {torch.__version__}; cudnn {torch.backends.cudnn.version()}", flush=True)
    with torch.no_grad():
        for part in a.only.split(","):
            dict(gemm=gemm_rows, conv=conv_rows, attention=attention_rows, norm=norm_rows, lnfold=lnfold_rows, ksweep=ksweep_rows)[part]()
    if a.json:
        with open(a.json, "w") as f:
            json.dump(dict(device=torch.cuda.get_device_name(0), torch=torch.__version__, rows=ROWS), f, indent=1)

#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): images/sec, SDXL 1024x1024, 50 Euler steps, CFG, bf16, batch 1 per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload sdxl|flux|vae]

A "step" is one full pass of the hot path over one batch: one image = 50 x (CFG-batched UNet forward + fused
CFG/Euler step) + VAE decode to a (1,3,1024,1024) tensor (output_type="pt").  Prints ONE JSON line (rank 0).
  value : images/s, embeddings already resident in HBM when the timed region starts (seeded latent draw included)
  e2e   : same metric through the public pipeline call with HOST (pinned) embeddings: H2D of the embeddings and
          D2H of the finished image inside the timed region
  roofline : the dominant kernel (conv_gemm_kernel: every Linear and Conv of the UNet) timed launch by launch with
          CUDA events on the launching stream in a separate eager pass of one UNet forward; achieved = algorithmic
          FLOPs of those launches / their summed duration, peak = measured sustained bf16 GEMM (MEASURED_PEAKS.json)
  cpu_baseline : the oracle port (the reference's op sequence in torch on the host cores) on a bounded sample
--impl reference times that CPU port as the reference arm.  Multi-GPU: one process per GPU (torchrun), each rank
samples its own image (weak scaling), one all-gather of the decoded images per step, max-over-ranks device timing.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

UNET_FLOP_PER_SAMPLE = 6.7612e12      # BASELINE.md §3 (FlopCounterMode on the reference, meta device)
UNET_GEMM_FLOP_PER_SAMPLE = (3325.3 + 1623.1 + 1028.9) * 1e9  # addmm + conv + mm: what conv_gemm_kernel executes
VAE_FLOP_PER_IMAGE = 10.470e12
FLUX_FLOP_PER_FORWARD = 74.385e12
IMAGE_FLOP = 2 * 50 * UNET_FLOP_PER_SAMPLE + VAE_FLOP_PER_IMAGE  # 686.6 TFLOP
SDXL_SCHED = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", timestep_spacing="leading", steps_offset=1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tflops=d.get("bf16_tflops_sustained", d["bf16_tflops"]), burst=d["bf16_tflops"], source="measured")
    return dict(hbm_gbs=6650.0, tflops=1400.0, burst=1590.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                pass

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None, reasons=sorted(reasons), samples=len(sm))


def synthetic_embeds(batch, dtype, pin):
    g = torch.Generator().manual_seed(1)
    mk = lambda *s: torch.randn(*s, generator=g).to(dtype)  # noqa: E731
    t = dict(prompt_embeds=mk(batch, 77, 2048), negative_prompt_embeds=mk(batch, 77, 2048), pooled_prompt_embeds=mk(batch, 1280),
             negative_pooled_prompt_embeds=mk(batch, 1280))
    if pin:
        t = {k: v.pin_memory() for k, v in t.items()}
    return t


# ----------------------------------------------------------------------------------------------- reference arm
MID_BLOCK_SHARE = 797.3 / 6761.2  # per-block FLOPs of one SDXL UNet sample-forward (GF)


def cpu_reference(args, full=True):
    """The oracle port (the reference's op sequence on torch CPU kernels, all host threads) on a BOUNDED sample:
    the SDXL UNet mid block (ResnetBlock2D + 10-layer Transformer2DModel + ResnetBlock2D at 32x32, CFG batch 2) =
    11.79 % of a UNet forward's FLOPs; t_forward = t_mid / 0.1179.  dtype = whichever of bf16 / fp32 this host's
    GEMM kernels run faster (bf16 is fast only with AMX / AVX512-BF16), stated in `sample`."""
    from diffusers_b200 import specs
    from oracle import blocks as Bk
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    t0 = time.time()

    def probe(dt):
        # the op mix of the block (3x3 conv, token GEMM, attention) at reduced size; a GEMM alone is misleading: hosts
        # with a fast bf16 GEMM can still run bf16 convolutions / attention several times slower than fp32
        import torch.nn.functional as F
        xc, wc = torch.randn(2, 320, 32, 32).to(dt), torch.randn(320, 320, 3, 3).to(dt)
        a, b = torch.randn(2048, 1280).to(dt), torch.randn(1280, 1280).to(dt)
        q = torch.randn(2, 20, 1024, 64).to(dt)

        def once():
            F.conv2d(xc, wc, padding=1)
            a @ b
            F.scaled_dot_product_attention(q, q, q)

        once()
        t = time.perf_counter()
        for _ in range(2):
            once()
        return (time.perf_counter() - t) / 2

    dt = torch.bfloat16 if probe(torch.bfloat16) < probe(torch.float32) else torch.float32
    spec = {k[len("mid_block."):]: v for k, v in specs.unet2d_condition_params(specs.SDXL_UNET_CONFIG).items() if k.startswith("mid_block.")}
    sd = {"mid_block." + k: v for k, v in specs.random_state_dict(spec, seed=0, dtype=dt).items()}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 1280, 32, 32, generator=g).to(dt)
    emb = torch.randn(2, 1280, generator=g).to(dt)
    ehs = torch.randn(2, 77, 2048, generator=g).to(dt)
    fwd = lambda: Bk.unet_mid_block_2d_cross_attn(sd, "mid_block", x, emb, ehs, heads=20, groups=32, eps=1e-5, use_linear_projection=True)  # noqa: E731
    with torch.no_grad():
        n_w, n_t = (max(1, min(args.warmup, 2)), max(1, min(args.steps, 5))) if full else (1, 2)
        for _ in range(n_w):
            fwd()
        ts = []
        for _ in range(n_t):
            a = time.perf_counter()
            fwd()
            ts.append(time.perf_counter() - a)
            if sum(ts) > 40:
                break
    t_mid = sum(ts) / len(ts)
    t_unet = t_mid / MID_BLOCK_SHARE
    t_vae = t_unet * VAE_FLOP_PER_IMAGE / (2 * UNET_FLOP_PER_SAMPLE)  # same FLOP rate assumed for the decode
    ips = 1.0 / (50 * t_unet + t_vae)
    sample = (f"{len(ts)} timed passes of the SDXL UNet mid block (B=2, 32x32, {str(dt).split('.')[-1]}; {t_mid:.2f} s each = "
              f"{100 * MID_BLOCK_SHARE:.2f}% of a CFG-batched UNet forward by FLOPs -> {t_unet:.1f} s/forward); "
              f"images/s = 1/(50*t_forward + t_vae), t_vae scaled by FLOPs")
    return dict(value=ips, unit="images/s", cores=cores, kind="port", sample=sample, t_unet_b2_s=t_unet, setup_s=round(time.time() - t0, 1))


def sdxl_workload(args):
    """config.workload of the headline benchmark: both arms (B200 and reference) report the same string."""
    return f"sdxl_unet_1024_{args.denoise_steps}step_cfg{args.guidance_scale}_b{args.batch}_per_gpu+vae_decode"


def run_reference(args, rank, world):
    if rank != 0:
        return
    cb = cpu_reference(args, full=True)
    line = dict(metric="images/sec @ SDXL 1024^2 50-step", value=cb["value"], unit="images/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1000.0 / cb["value"], higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="bf16", data="synthetic", impl="reference",
                config=dict(workload=sdxl_workload(args), global_batch=args.batch, parallelism="host cores",
                            note="reference op sequence on host cores (oracle port), bounded sample"),
                cpu_baseline={k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                e2e=dict(value=cb["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- B200 arm
def gemm_roofline(unet, B2, pk):
    """Eager pass of one UNet forward with CUDA events around every conv_gemm launch."""
    from diffusers_b200 import ops
    dev, dt = unet.device, unet.dtype
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(B2, 4, 128, 128, generator=g, device=dev).to(dt)
    ehs = torch.randn(B2, 77, 2048, generator=g, device=dev).to(dt)
    added = dict(text_embeds=torch.randn(B2, 1280, generator=g, device=dev).to(dt),
                 time_ids=torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * B2, device=dev).to(dt))
    t = torch.tensor(981.0, device=dev)
    was = unet.use_cuda_graph
    unet.enable_cuda_graph(False)
    for _ in range(2):
        unet(x, t, ehs, added_cond_kwargs=added, return_dict=False)
    torch.cuda.synchronize()
    ops._PROFILE = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    unet(x, t, ehs, added_cond_kwargs=added, return_dict=False)
    e1.record()
    torch.cuda.synchronize()
    prof, ops._PROFILE = ops._PROFILE, None
    unet.use_cuda_graph = was
    ms = [a.elapsed_time(b) for a, b, *_ in prof]
    flops = sum(p[2] for p in prof)
    total_ms = sum(ms)
    ach = flops / (total_ms * 1e-3) / 1e12
    # DRAM traffic of the kernel's most frequent launch (linear 2048x1280x1280 + bias + residual, 192 per forward) from the
    # committed `ncu --set full` capture; its algorithmic bytes are x 5.24 MB + W 3.28 MB + residual 5.24 MB read (the
    # 5.24 MB output stays in L2) = 13.77 MB, i.e. no re-reads reach HBM
    traffic, traffic_note = None, None
    tp = os.path.join(ROOT, "profiles", "r1_conv_gemm_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            cap = json.load(f)["launches"][0]
        traffic = cap["dram_bytes_read"] + cap["dram_bytes_write"]
        traffic_note = f"bytes per launch of '{cap['launch']}' (ncu --set full capture, profiles/r1_ncu_full_hot_kernels.txt); algorithmic 13.77e6"
    return dict(bound="tensor", achieved=round(ach, 1), peak=pk["tflops"], unit="TFLOP/s", frac=round(ach / pk["tflops"], 4), traffic=traffic,
                traffic_note=traffic_note,
                kernel="conv_gemm_kernel", launches_per_forward=len(prof), algorithmic_flop_per_forward=flops,
                avg_launch_us=round(1000 * total_ms / len(prof), 2), kernel_ms_per_forward=round(total_ms, 3),
                forward_ms_eager=round(e0.elapsed_time(e1), 3), share_of_forward=round(total_ms / e0.elapsed_time(e1), 3),
                peak_source=pk["source"] + " (bf16_tflops_sustained: kernel timed inside a long step)")


def hbm_kernel_rooflines(dev, dt, pk):
    """The HBM-bound kernels of the path at their dominant SDXL shapes: 16 launches inside one CUDA graph
    over 8 rotating buffers (168 MB > L2, so the reads really come from HBM), algorithmic bytes = one read + one write of
    the activation, against the measured copy bandwidth."""
    from diffusers_b200 import ops
    g = torch.Generator(device=dev).manual_seed(0)
    out = []

    def timed(fn, n=16):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for i in range(n):
                fn(i)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return 1000.0 * e0.elapsed_time(e1) / n

    xs = [torch.randn(2 * 16384, 320, generator=g, device=dev).to(dt) for _ in range(8)]
    ys = [torch.empty_like(x) for x in xs]
    gam, bet = torch.randn(320, generator=g, device=dev).to(dt), torch.randn(320, generator=g, device=dev).to(dt)
    us = timed(lambda i: ops.group_norm(xs[i % 8], batch=2, hw=16384, groups=32, eps=1e-5, gamma=gam, beta=bet, silu=True, out=ys[i % 8]))
    nbytes = 2 * xs[0].numel() * 2
    out.append(dict(kernel="group_norm_stats+apply (+SiLU)", shape="2x16384x320", algorithmic_bytes=nbytes, us=round(us, 2),
                    achieved_gbs=round(nbytes / us / 1e3, 1), frac=round(nbytes / us / 1e3 / pk["hbm_gbs"], 4),
                    note="two launches; the statistics pass re-reads the input (3 passes over 21 MB for 2 algorithmic)"))
    xl = [torch.randn(2048, 1280, generator=g, device=dev).to(dt) for _ in range(32)]
    yl = [torch.empty_like(x) for x in xl]
    lg, lb = torch.randn(1280, generator=g, device=dev).to(dt), torch.randn(1280, generator=g, device=dev).to(dt)
    us = timed(lambda i: ops.layer_norm(xl[i % 32], eps=1e-5, gamma=lg, beta=lb, out=yl[i % 32]), n=32)
    nbytes = 2 * xl[0].numel() * 2
    out.append(dict(kernel="layer_norm", shape="2048x1280", algorithmic_bytes=nbytes, us=round(us, 2), achieved_gbs=round(nbytes / us / 1e3, 1),
                    frac=round(nbytes / us / 1e3 / pk["hbm_gbs"], 4), note="10 MB per launch: launch-latency bound"))
    return out


def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    from diffusers_b200 import ops, parallel, specs
    from diffusers_b200.autoencoder_kl import AutoencoderKL
    from diffusers_b200.pipelines import StableDiffusionXLPipeline
    from diffusers_b200.schedulers import EulerDiscreteScheduler
    from diffusers_b200.unet_2d_condition import UNet2DConditionModel

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dt = torch.bfloat16
    pk = peaks()
    unet = UNet2DConditionModel.random_init(seed=0, dtype=dt, device=dev)
    vae = AutoencoderKL.random_init(seed=0, dtype=dt, device=dev)
    pipe = StableDiffusionXLPipeline(vae, unet, EulerDiscreteScheduler(**SDXL_SCHED))
    B = args.batch
    host = synthetic_embeds(B, dt, pin=True)
    resident = {k: v.to(dev) for k, v in host.items()}
    call = dict(height=1024, width=1024, num_inference_steps=args.denoise_steps, guidance_scale=args.guidance_scale, output_type="pt")

    def one_image(emb, seed, to_host):
        img = pipe(generator=torch.Generator(device=dev).manual_seed(seed), **emb, **call).images
        if world > 1:
            img = parallel.all_gather_batch(img, B * world) if B * world > 1 else img
        if to_host:
            out = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
            out.copy_(img, non_blocking=True)
            return out
        return img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(emb_fn, to_host):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.launches()
        e0.record()
        for i in range(args.steps):
            one_image(emb_fn(), 1000 + i, to_host)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, ops.launches() - n0

    for i in range(args.warmup):
        one_image(resident, i, False)
    with ClockSampler(local_rank) as cs:
        ms, launches = timed(lambda: resident, False)
    clocks = cs.summary()
    # end to end: host-pinned embeddings in, image back on the host, through the public pipeline call
    one_image({k: v.to(dev, non_blocking=True) for k, v in host.items()}, 0, True)
    ms_e2e, _ = timed(lambda: {k: v.to(dev, non_blocking=True) for k, v in host.items()}, True)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = B * (world if world > 1 and B * world > 1 else 1) * 3 * 1024 * 1024 * 2

    n_img = args.steps * B * world
    value = n_img / (ms * 1e-3)
    line = None
    if rank == 0:
        roof = gemm_roofline(unet, 2 * B, pk)
        flops_per_img = 2 * args.denoise_steps * UNET_FLOP_PER_SAMPLE + VAE_FLOP_PER_IMAGE
        cb = cpu_reference(args, full=False) if world == 1 and not args.no_cpu_baseline else None
        line = dict(metric="images/sec @ SDXL 1024^2 50-step", value=round(value, 4), unit="images/s", n_gpus=world, steps=args.steps,
                    warmup=args.warmup, ms_per_step=round(ms / args.steps, 2), higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="bf16", data="synthetic (random-init weights, N(0,1) text embeddings, seeded latents)",
                    config=dict(workload=sdxl_workload(args),
                                global_batch=B * world, parallelism=f"dp{world}", cfg_batched=True,
                                l2="per-step working set >> L2: 5.1 GB of weights stream from HBM every UNet forward",
                                model="UNet2DConditionModel SDXL-base config (2.567 B params) + AutoencoderKL SDXL decoder"),
                    e2e=dict(value=round(n_img / (ms_e2e * 1e-3), 4), unit="images/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
                    gpu_launches=launches, clocks=clocks, roofline=roof,
                    model_flops=dict(tflop_per_image=round(flops_per_img / 1e12, 1),
                                     achieved_tflops_per_gpu=round(value / world * flops_per_img / 1e12, 1),
                                     frac_of_peak=round(value / world * flops_per_img / 1e12 / pk["tflops"], 4)))
        if cb is not None:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        try:
            line["hbm_kernels"] = hbm_kernel_rooflines(dev, dt, pk)
        except Exception as e:  # noqa: BLE001  (a side measurement must never cost the headline line)
            line["hbm_kernels"] = f"failed: {e}"
        print(json.dumps(line), flush=True)


def run_flux(args, rank, world, local_rank):
    """Secondary headline: latents/sec, FluxTransformer2DModel (Flux.1-dev shape), 1024^2 (4096 image + 512 text
    tokens), 28 FlowMatch steps, guidance 3.5, output_type='latent' (BASELINE.json configs[2])."""
    from diffusers_b200 import ops, specs
    from diffusers_b200.pipelines import FluxPipeline
    from diffusers_b200.schedulers import FlowMatchEulerDiscreteScheduler
    from diffusers_b200.transformer_flux import FluxTransformer2DModel
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dt = torch.bfloat16
    pk = peaks()
    t0 = time.time()
    tr = FluxTransformer2DModel.random_init(seed=0, dtype=dt, device=dev)
    init_s = time.time() - t0

    class _V:
        config = type("C", (), dict(block_out_channels=(128, 256, 512, 512)))()

    sk = dict(shift=3.0, use_dynamic_shifting=True, base_shift=0.5, max_shift=1.15, base_image_seq_len=256, max_image_seq_len=4096)
    pipe = FluxPipeline(FlowMatchEulerDiscreteScheduler(**sk), _V(), tr)
    g = torch.Generator().manual_seed(1)
    host = dict(prompt_embeds=torch.randn(1, 512, 4096, generator=g).to(dt).pin_memory(),
                pooled_prompt_embeds=torch.randn(1, 768, generator=g).to(dt).pin_memory())
    call = dict(height=1024, width=1024, num_inference_steps=args.denoise_steps if args.denoise_steps != 50 else 28,
                guidance_scale=3.5, output_type="latent")
    nsteps = call["num_inference_steps"]

    def one(emb, seed, to_host):
        lat = pipe(generator=torch.Generator(device=dev).manual_seed(seed), **emb, **call).images
        if to_host:
            out = torch.empty(lat.shape, dtype=lat.dtype, pin_memory=True)
            out.copy_(lat, non_blocking=True)
            return out
        return lat

    res = {k: v.to(dev) for k, v in host.items()}
    for i in range(args.warmup):
        one(res, i, False)
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as cs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.launches()
        e0.record()
        for i in range(args.steps):
            one(res, 100 + i, False)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = ops.launches() - n0
    e0.record()
    for i in range(args.steps):
        one({k: v.to(dev, non_blocking=True) for k, v in host.items()}, 200 + i, True)
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1)
    value = args.steps / (ms * 1e-3)
    fl = nsteps * FLUX_FLOP_PER_FORWARD
    line = dict(metric="latents/sec @ Flux.1-dev-shape 1024^2 28-step", value=round(value, 4), unit="latents/s", n_gpus=1,
                steps=args.steps, warmup=args.warmup, ms_per_step=round(ms / args.steps, 1), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="bf16", data="synthetic (random-init weights, N(0,1) embeddings)",
                config=dict(workload=f"flux_dev_shape_1024_{nsteps}step_g3.5_b1", model="FluxTransformer2DModel 11.90 B params",
                            l2="23.8 GB of weights stream from HBM every forward", init_s=round(init_s, 1)),
                e2e=dict(value=round(args.steps / (ms2 * 1e-3), 4), unit="latents/s",
                         h2d_bytes_per_step=sum(v.numel() * 2 for v in host.values()), d2h_bytes_per_step=4096 * 64 * 2),
                gpu_launches=launches, clocks=cs.summary(),
                model_flops=dict(tflop_per_latent=round(fl / 1e12, 1), achieved_tflops=round(value * fl / 1e12, 1),
                                 frac_of_peak=round(value * fl / 1e12 / pk["tflops"], 4), peak=pk["tflops"]))
    print(json.dumps(line), flush=True)


def run_vae(args, rank, world, local_rank):
    """BASELINE.json configs[4]: AutoencoderKL.decode (SDXL decoder, 49.5 M params) of z (B,4,128,128) -> (B,3,1024,1024),
    batch sweep 1 / 8 / 64 - the convolution-roofline view of the path (10.47 TFLOP per image)."""
    from diffusers_b200 import ops
    from diffusers_b200.autoencoder_kl import AutoencoderKL
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dt = torch.bfloat16
    pk = peaks()
    vae = AutoencoderKL.random_init(seed=0, dtype=dt, device=dev)
    sweep, line = {}, None
    with ClockSampler(local_rank) as cs:
        for B in (1, 8, 64):
            g = torch.Generator().manual_seed(B)
            zh = torch.randn(B, 4, 128, 128, generator=g).to(dt).pin_memory()
            z = zh.to(dev)
            for _ in range(max(1, args.warmup if B < 64 else 1)):
                vae.decode(z, return_dict=False)
            torch.cuda.synchronize()
            n = max(1, args.steps if B < 64 else min(args.steps, 2))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = ops.launches()
            e0.record()
            for _ in range(n):
                img = vae.decode(z, return_dict=False)[0]
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / n
            launches = (ops.launches() - n0) // n
            host_out = torch.empty(img.shape, dtype=img.dtype, pin_memory=True)
            e0.record()
            for _ in range(n):
                host_out.copy_(vae.decode(zh.to(dev, non_blocking=True), return_dict=False)[0], non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / n
            sweep[B] = dict(images_per_s=round(B / (ms * 1e-3), 2), ms_per_batch=round(ms, 2), e2e_images_per_s=round(B / (ms2 * 1e-3), 2),
                            tflops=round(B * VAE_FLOP_PER_IMAGE / (ms * 1e-3) / 1e12, 1), launches=launches)
            del img, host_out, z
    best = max(sweep, key=lambda b: sweep[b]["images_per_s"])
    v = sweep[best]
    line = dict(metric="images/sec @ AutoencoderKL.decode 1024^2 (SDXL VAE)", value=v["images_per_s"], unit="images/s", n_gpus=1, steps=args.steps,
                warmup=args.warmup, ms_per_step=v["ms_per_batch"], higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16",
                data="synthetic (random-init weights, N(0,1) latents)",
                config=dict(workload=f"sdxl_vae_decode_1024_b{best}", sweep={str(k): s for k, s in sweep.items()},
                            l2="activations of one image (268 MB per 128-channel 1024^2 tensor) >> L2"),
                e2e=dict(value=v["e2e_images_per_s"], unit="images/s", h2d_bytes_per_step=best * 4 * 128 * 128 * 2,
                         d2h_bytes_per_step=best * 3 * 1024 * 1024 * 2),
                gpu_launches=v["launches"], clocks=cs.summary(),
                model_flops=dict(tflop_per_image=round(VAE_FLOP_PER_IMAGE / 1e12, 2), achieved_tflops=v["tflops"],
                                 frac_of_peak=round(v["tflops"] / pk["tflops"], 4), peak=pk["tflops"]))
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="sdxl", choices=["sdxl", "flux", "vae"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="images per GPU per step")
    ap.add_argument("--denoise-steps", type=int, default=50)
    ap.add_argument("--guidance-scale", type=float, default=7.5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        if args.workload == "flux":
            run_flux(args, rank, world, local_rank)
        elif args.workload == "vae":
            run_vae(args, rank, world, local_rank)
        else:
            run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

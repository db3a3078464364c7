"""Where does one SDXL image (1024^2, 50 steps, CFG) spend its time?  Device-timed pieces of the public pipeline call."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from diffusers_b200.autoencoder_kl import AutoencoderKL
from diffusers_b200.pipelines import StableDiffusionXLPipeline
from diffusers_b200.schedulers import EulerDiscreteScheduler
from diffusers_b200.unet_2d_condition import UNet2DConditionModel

dev = torch.device("cuda", 0)
dt = torch.bfloat16
unet = UNet2DConditionModel.random_init(seed=0, dtype=dt, device=dev)
vae = AutoencoderKL.random_init(seed=0, dtype=dt, device=dev)
pipe = StableDiffusionXLPipeline(vae, unet, EulerDiscreteScheduler(**bench.SDXL_SCHED))
emb = {k: v.to(dev) for k, v in bench.synthetic_embeds(1, dt, pin=False).items()}


def timed(fn, n=2):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    t_host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, 1000 * t_host


call = dict(height=1024, width=1024, guidance_scale=7.5)
gen = lambda: torch.Generator(device=dev).manual_seed(0)  # noqa: E731
full, full_h = timed(lambda: pipe(generator=gen(), **emb, **call, num_inference_steps=50, output_type="pt"))
lat50, lat50_h = timed(lambda: pipe(generator=gen(), **emb, **call, num_inference_steps=50, output_type="latent"))
lat10, lat10_h = timed(lambda: pipe(generator=gen(), **emb, **call, num_inference_steps=10, output_type="latent"))
z = torch.randn(1, 4, 128, 128, device=dev).to(dt)
dec, dec_h = timed(lambda: vae.decode(z, return_dict=False))
st = pipe._graph
rep, rep_h = timed(lambda: st["graph"].replay(), n=10)
print(f"full image           {full:8.1f} ms (host issue {full_h:7.1f} ms)")
print(f"50 steps, latent out {lat50:8.1f} ms (host issue {lat50_h:7.1f} ms)")
print(f"10 steps, latent out {lat10:8.1f} ms (host issue {lat10_h:7.1f} ms) -> per step {(lat50 - lat10) / 40:.2f} ms, fixed {lat10 - 10 * (lat50 - lat10) / 40:.1f} ms")
print(f"vae decode           {dec:8.1f} ms (host issue {dec_h:7.1f} ms)")
print(f"unet graph replay    {rep:8.2f} ms (host issue {rep_h:7.2f} ms)")

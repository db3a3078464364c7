"""ORACLE (test infrastructure): the reference sampling loops restated over the oracle models (CPU).

Reference: pipelines/stable_diffusion_xl/pipeline_stable_diffusion_xl.py:1100-1300 (denoise loop :1197-1255, CFG
:1224-1225, decode :1285-1287, postprocess image_processor.py:738), pipelines/flux/pipeline_flux.py:809-962,
pipelines/ddpm/pipeline_ddpm.py:55-139.
"""
import numpy as np
import torch

from . import flux as oflux
from . import schedulers as osched
from . import unet as ounet
from . import vae as ovae


def sdxl_sample(unet_sd, unet_cfg, sched, latents0, prompt_embeds, negative_prompt_embeds, pooled, negative_pooled,
                time_ids, num_inference_steps, guidance_scale, vae_sd=None, vae_cfg=None, return_all=False):
    """latents0: the seeded N(0,1) draw (prepare_latents multiplies by init_noise_sigma)."""
    sched.set_timesteps(num_inference_steps)
    lat = latents0 * sched.init_noise_sigma
    do_cfg = guidance_scale > 1
    pe = torch.cat([negative_prompt_embeds, prompt_embeds], 0) if do_cfg else prompt_embeds
    te = torch.cat([negative_pooled, pooled], 0) if do_cfg else pooled
    tid = torch.cat([time_ids, time_ids], 0) if do_cfg else time_ids
    tid = tid.repeat(latents0.shape[0], 1)
    trace = []
    for t in sched.timesteps:
        inp = torch.cat([lat] * 2) if do_cfg else lat
        inp = sched.scale_model_input(inp)
        eps = ounet.unet2d_condition_forward(unet_sd, unet_cfg, inp, t, pe, dict(text_embeds=te, time_ids=tid))
        if do_cfg:
            u, c = eps.chunk(2)
            eps = u + guidance_scale * (c - u)
        lat = sched.step(eps, lat)
        if return_all:
            trace.append(lat.clone())
    if vae_sd is None:
        return (lat, trace) if return_all else lat
    img = ovae.vae_decode(vae_sd, vae_cfg, lat / vae_cfg["scaling_factor"])
    img = (img * 0.5 + 0.5).clamp(0, 1)
    return (img, lat, trace) if return_all else img


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15):
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


def flux_sample(sd, cfg, sched, packed_latents, prompt_embeds, pooled, img_ids, txt_ids, num_inference_steps, guidance_scale):
    sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
    mu = calculate_shift(packed_latents.shape[1])
    sched.set_timesteps(num_inference_steps, sigmas=sigmas, mu=mu)
    lat = packed_latents
    guidance = torch.full([1], guidance_scale, dtype=torch.float32).expand(lat.shape[0]) if cfg.get("guidance_embeds") else None
    for t in sched.timesteps:
        timestep = t.expand(lat.shape[0]).to(lat.dtype)
        v = oflux.flux_forward(sd, cfg, lat, prompt_embeds, pooled, timestep / 1000, img_ids, txt_ids, guidance)
        lat = sched.step(v, lat)
    return lat


def ddpm_sample(sd, cfg, sched, image0, num_inference_steps, generator):
    sched.set_timesteps(num_inference_steps)
    image = image0
    for t in sched.timesteps:
        eps = ounet.unet2d_forward(sd, cfg, image, t)
        image = sched.step(eps, t, image, generator=generator)
    return (image / 2 + 0.5).clamp(0, 1)

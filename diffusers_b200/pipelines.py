"""Sampling loops for the hot path when the reference package is not installed (the GPU box).

`StableDiffusionXLPipeline` / `FluxPipeline` here take the same component objects and the same
embedding-level call arguments as the reference pipelines (pipelines/stable_diffusion_xl/
pipeline_stable_diffusion_xl.py:823-1308, pipelines/flux/pipeline_flux.py:600-970) and run the same sequence:
prepare latents -> set timesteps -> [scale -> denoiser -> CFG -> scheduler.step] x N -> VAE decode -> postprocess.
Text encoders / tokenizers are out of scope: prompts enter as embeddings, which is also how
the north-star benchmark feeds them.

`fused=True` (default) keeps the loop state on the device in the kernels' native layout: one CUDA-graph replay
of the NHWC UNet forward plus ONE fused CFG + Euler + next-input kernel per step, no per-step layout changes.
`fused=False` drives the drop-in `unet.forward` / `scheduler.step` API exactly like the reference loop.
"""
import os

import numpy as np
import torch

from . import ops


def randn_tensor(shape, generator=None, device=None, dtype=None):
    """utils/torch_utils.py:183-233: a CPU generator draws on the CPU and the tensor is moved afterwards, so seeds
    reproduce across devices; a list of generators seeds every sample separately."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    if generator is not None:
        gen_type = generator[0].device.type if isinstance(generator, list) else generator.device.type
        if gen_type != device.type and gen_type == "cpu":
            rand_device = torch.device("cpu")
        elif gen_type != device.type and gen_type == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {gen_type}.")
    if isinstance(generator, list) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, list):
        one = (1,) + tuple(shape[1:])
        lat = [torch.randn(one, generator=generator[i], device=rand_device, dtype=dtype) for i in range(shape[0])]
        return torch.cat(lat, dim=0).to(device)
    return torch.randn(tuple(shape), generator=generator, device=rand_device, dtype=dtype).to(device)


def postprocess_pt(image):
    """VaeImageProcessor.postprocess(output_type='pt') (image_processor.py:738, denormalize :222)."""
    return (image * 0.5 + 0.5).clamp(0, 1)


class PipelineOutput:
    def __init__(self, images):
        self.images = images


class StableDiffusionXLPipeline:
    @classmethod
    def from_pretrained(cls, path, torch_dtype=torch.bfloat16, device="cuda", variant=None, text_encoders=False):
        """A reference `StableDiffusionXLPipeline.save_pretrained` directory -> this pipeline (unet/, vae/, scheduler/; with
        text_encoders=True also text_encoder/ and text_encoder_2/ for `encode_prompt` from token ids - the tokenizers stay with
        transformers; otherwise call with prompt_embeds).  checkpoint.py, SURVEY.md 8f N1."""
        from . import checkpoint
        from .autoencoder_kl import AutoencoderKL
        from .schedulers import (DDIMScheduler, DPMSolverMultistepScheduler, EulerAncestralDiscreteScheduler, EulerDiscreteScheduler,
                                 UniPCMultistepScheduler)
        from .unet_2d_condition import UNet2DConditionModel
        steppers = (EulerDiscreteScheduler, DDIMScheduler, EulerAncestralDiscreteScheduler, DPMSolverMultistepScheduler, UniPCMultistepScheduler)
        c = checkpoint.load_pipeline_components(path, "StableDiffusionXLPipeline",
                                                dict(unet=UNet2DConditionModel, vae=AutoencoderKL, scheduler=steppers),
                                                torch_dtype=torch_dtype, device=device, variant=variant)
        te = {}
        if text_encoders:
            from .text_encoders import CLIPTextModel, CLIPTextModelWithProjection
            te = dict(text_encoder=CLIPTextModel.from_pretrained(path, subfolder="text_encoder", variant=variant, torch_dtype=torch_dtype, device=device),
                      text_encoder_2=CLIPTextModelWithProjection.from_pretrained(path, subfolder="text_encoder_2", variant=variant, torch_dtype=torch_dtype,
                                                                                 device=device))
        return cls(c["vae"], c["unet"], c["scheduler"], **te)

    def __init__(self, vae, unet, scheduler, text_encoder=None, text_encoder_2=None, force_zeros_for_empty_prompt=True):
        self.vae, self.unet, self.scheduler = vae, unet, scheduler
        self.text_encoder, self.text_encoder_2 = text_encoder, text_encoder_2  # text_encoders.CLIPTextModel / ...WithProjection (optional)
        self.force_zeros_for_empty_prompt = force_zeros_for_empty_prompt
        self.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1) if vae is not None else 8
        self.default_sample_size = unet.config.sample_size
        self._graph = None

    @torch.no_grad()
    def encode_prompt(self, text_input_ids, text_input_ids_2, negative_input_ids=None, negative_input_ids_2=None, do_classifier_free_guidance=True):
        """StableDiffusionXLPipeline.encode_prompt (pipeline_stable_diffusion_xl.py:283-470) from TOKEN IDS (the tokenizers are host-side
        string processing and stay with transformers): per encoder `hidden_states[-2]`, concatenated on the feature axis; the pooled
        embedding is `[0]` of the last (projection) encoder; without negative ids the negative embeddings are zeros
        (force_zeros_for_empty_prompt).  -> (prompt_embeds, negative_prompt_embeds, pooled_prompt_embeds, negative_pooled_prompt_embeds)"""
        if self.text_encoder is None or self.text_encoder_2 is None:
            raise ValueError("encode_prompt needs text_encoder and text_encoder_2 (diffusers_b200.text_encoders)")

        def run(ids_pair):
            embeds, pooled = [], None
            for ids, enc in zip(ids_pair, (self.text_encoder, self.text_encoder_2)):
                out = enc(ids.to(self.device), output_hidden_states=True)
                if out[0].ndim == 2:
                    pooled = out[0]
                embeds.append(out.hidden_states[-2])
            return torch.cat(embeds, dim=-1), pooled

        pe, pooled = run((text_input_ids, text_input_ids_2))
        npe = npooled = None
        if do_classifier_free_guidance:
            if negative_input_ids is None and self.force_zeros_for_empty_prompt:
                npe, npooled = torch.zeros_like(pe), torch.zeros_like(pooled)
            else:
                if negative_input_ids is None or negative_input_ids_2 is None:
                    raise ValueError("negative token ids for both encoders are needed when force_zeros_for_empty_prompt is False")
                npe, npooled = run((negative_input_ids, negative_input_ids_2))
        return pe, npe, pooled, npooled

    @property
    def device(self):
        return self.unet.device

    def prepare_latents(self, batch_size, num_channels, height, width, dtype, device, generator, latents=None):
        shape = (batch_size, num_channels, int(height) // self.vae_scale_factor, int(width) // self.vae_scale_factor)
        if latents is None:
            latents = randn_tensor(shape, generator=generator, device=device, dtype=dtype)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma  # 0-dim fp32 CPU tensor * 16-bit tensor -> 16-bit

    # ------------------------------------------------------------------ per-shape graph of the NHWC UNet forward
    def _get_graph(self, B2, H, W, kv, added):
        unet = self.unet
        key = (B2, H, W, tuple(kv.shape), tuple(added["text_embeds"].shape))
        if self._graph is not None and self._graph["key"] == key:
            return self._graph
        dev, dt = unet.device, unet.dtype
        st = dict(key=key,
                  x_in=torch.zeros((B2 * H * W, unet.in_pad), dtype=dt, device=dev),
                  t=torch.zeros((), dtype=torch.float32, device=dev),
                  kv=torch.empty_like(kv),
                  added=dict(text_embeds=torch.empty_like(added["text_embeds"]), time_ids=torch.empty_like(added["time_ids"])))

        def run():
            return unet._forward_nhwc(st["x_in"], B2, H, W, st["t"], st["kv"], st["added"])

        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run()  # warm-up outside capture (workspaces, smem attributes, module load)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        n0 = ops.launches()
        with torch.cuda.graph(g):
            st["eps"] = run()
        st["graph"] = g
        st["launches"] = ops.launches() - n0
        self._graph = st
        return st

    @torch.no_grad()
    def __call__(self, prompt_embeds, negative_prompt_embeds=None, pooled_prompt_embeds=None,
                 negative_pooled_prompt_embeds=None, height=None, width=None, num_inference_steps=50,
                 guidance_scale=5.0, generator=None, latents=None, output_type="pt", original_size=None,
                 crops_coords_top_left=(0, 0), target_size=None, return_dict=True, fused=True, guidance_rescale=0.0):
        """guidance_rescale > 0 (`rescale_noise_cfg`, pipeline_stable_diffusion_xl.py:1236-1238) runs the drop-in loop: the fused CFG + Euler
        kernel has no per-sample reduction."""
        unet, sched = self.unet, self.scheduler
        device = unet.device
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
        original_size = original_size or (height, width)
        target_size = target_size or (height, width)
        do_cfg = guidance_scale > 1 and unet.config.time_cond_proj_dim is None
        if do_cfg and (negative_prompt_embeds is None or negative_pooled_prompt_embeds is None):
            raise ValueError("classifier-free guidance needs negative_prompt_embeds and negative_pooled_prompt_embeds")
        batch = prompt_embeds.shape[0]
        dtype = prompt_embeds.dtype

        sched.set_timesteps(num_inference_steps, device=device)
        timesteps = sched.timesteps
        lat = self.prepare_latents(batch, unet.config.in_channels, height, width, dtype, device, generator, latents)

        add_text = pooled_prompt_embeds
        add_time_ids = torch.tensor([list(original_size) + list(crops_coords_top_left) + list(target_size)], dtype=dtype)
        passed = unet.config.addition_time_embed_dim * add_time_ids.shape[1] + int(pooled_prompt_embeds.shape[-1])
        if unet.add_embedding.linear_1.in_features != passed:
            raise ValueError(f"Model expects an added time embedding vector of length {unet.add_embedding.linear_1.in_features}, "
                             f"but a vector of {passed} was created.")
        neg_time_ids = add_time_ids
        if do_cfg:
            prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)
            add_text = torch.cat([negative_pooled_prompt_embeds, add_text], dim=0)
            add_time_ids = torch.cat([neg_time_ids, add_time_ids], dim=0)
        prompt_embeds = prompt_embeds.to(device)
        add_text = add_text.to(device)
        add_time_ids = add_time_ids.to(device).repeat(batch, 1)
        added = dict(text_embeds=add_text, time_ids=add_time_ids)
        sched.set_begin_index(0)

        from .schedulers import EulerDiscreteScheduler
        rescale = do_cfg and guidance_rescale > 0.0
        if fused and isinstance(sched, EulerDiscreteScheduler) and sched.config.prediction_type == "epsilon" and not rescale:
            lat = self._denoise_fused(lat, timesteps, prompt_embeds, added, guidance_scale, do_cfg)
        else:
            # drop-in loop (any scheduler with the reference's scale_model_input / step surface: DDIM, Euler-ancestral,
            # DPM-Solver++ here); `generator` reaches step() only if it takes one (pipeline_stable_diffusion_xl.py:599-607)
            import inspect
            extra = dict(generator=generator) if "generator" in inspect.signature(sched.step).parameters else {}
            for t in timesteps:
                inp = torch.cat([lat] * 2) if do_cfg else lat
                inp = sched.scale_model_input(inp, t)
                noise_pred = unet(inp, t, encoder_hidden_states=prompt_embeds, added_cond_kwargs=added, return_dict=False)[0]
                if do_cfg:
                    u, c = noise_pred.chunk(2)
                    noise_pred = u + guidance_scale * (c - u)
                    if rescale:
                        noise_pred = rescale_noise_cfg(noise_pred, c, guidance_rescale=guidance_rescale)
                lat = sched.step(noise_pred, t, lat, return_dict=False, **extra)[0]

        if output_type == "latent":
            image = lat
        else:
            if self.vae.dtype == torch.float16 and self.vae.config.force_upcast:
                # the reference up-casts such a VAE to fp32 for the decode (pipeline_stable_diffusion_xl.py:1262-1270 `needs_upcasting`:
                # the SDXL VAE overflows in fp16); there is no fp32 kernel path here, so refuse instead of returning NaN / black images
                raise NotImplementedError("an fp16 AutoencoderKL with force_upcast=True must be decoded in fp32 by the reference; "
                                          "build the AutoencoderKL in bfloat16 (same range as fp32) instead")
            lat = lat / self.vae.config.scaling_factor
            image = self.vae.decode(lat, return_dict=False)[0]
            image = postprocess_pt(image)
        if not return_dict:
            return (image,)
        return PipelineOutput(image)

    def _denoise_fused(self, lat, timesteps, prompt_embeds, added, guidance_scale, do_cfg):
        unet, sched = self.unet, self.scheduler
        B, C, H, W = lat.shape
        B2 = 2 * B if do_cfg else B
        lat = lat.to(unet.dtype).contiguous().clone()
        kv = unet._text_kv(prompt_embeds)
        st = self._get_graph(B2, H, W, kv, added)
        st["kv"].copy_(kv)
        st["added"]["text_embeds"].copy_(added["text_embeds"])
        st["added"]["time_ids"].copy_(added["time_ids"])
        sig = sched.sigmas  # host fp32 table
        # first model input: scale_model_input(cat([latents]*2), t0) in NHWC
        first = ops.scale_div(torch.cat([lat] * 2) if do_cfg else lat, float((sig[0] ** 2 + 1) ** 0.5))
        ops.nchw_to_nhwc(first, c_pad=unet.in_pad, out=st["x_in"])
        n = len(timesteps)
        for i in range(n):
            st["t"].copy_(timesteps[i], non_blocking=True)
            st["graph"].replay()
            ops._count(st["launches"])
            ops.cfg_euler_step(st["eps"], lat, st["x_in"], guidance_scale=guidance_scale, do_cfg=do_cfg,
                               sigma=float(sig[i]), sigma_next=float(sig[i + 1]))
        sched._step_index = n
        return lat


def rescale_noise_cfg(noise_cfg, noise_pred_text, guidance_rescale=0.0):
    """pipeline_stable_diffusion_xl.py:84-107 (Lin et al. 2305.08891, section 3.4): bring the per-sample standard deviation of the guided
    prediction back to the text-conditional one's, then blend by `guidance_rescale`.  Same op sequence as the reference (per-sample
    std over every non-batch axis, 16-bit tensor arithmetic), so the drop-in loop reproduces its rounding."""
    dims = list(range(1, noise_pred_text.ndim))
    factor = noise_pred_text.std(dim=dims, keepdim=True) / noise_cfg.std(dim=dims, keepdim=True)
    return guidance_rescale * (noise_cfg * factor) + (1 - guidance_rescale) * noise_cfg


def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15):
    """pipelines/flux/pipeline_flux.py:73-84"""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


class FluxPipeline:
    @classmethod
    def from_pretrained(cls, path, torch_dtype=torch.bfloat16, device="cuda", variant=None, text_encoders=False):
        """A reference `FluxPipeline.save_pretrained` directory -> this pipeline (transformer/, vae/ config for the latent
        geometry, scheduler/; with text_encoders=True also text_encoder/ (CLIP-L) and text_encoder_2/ (T5) for `encode_prompt` from
        token ids); output_type='latent' only."""
        from . import checkpoint
        from .config import FrozenConfig
        from .schedulers import FlowMatchEulerDiscreteScheduler
        from .transformer_flux import FluxTransformer2DModel
        c = checkpoint.load_pipeline_components(path, "FluxPipeline",
                                                dict(transformer=FluxTransformer2DModel, scheduler=FlowMatchEulerDiscreteScheduler),
                                                torch_dtype=torch_dtype, device=device, variant=variant)
        vae_cfg = checkpoint.public_config(checkpoint.load_config(os.path.join(path, "vae")))
        vae = type("VaeGeometry", (), dict(config=FrozenConfig(vae_cfg)))()
        te = {}
        if text_encoders:
            from .text_encoders import CLIPTextModel, T5EncoderModel
            te = dict(text_encoder=CLIPTextModel.from_pretrained(path, subfolder="text_encoder", variant=variant, torch_dtype=torch_dtype, device=device),
                      text_encoder_2=T5EncoderModel.from_pretrained(path, subfolder="text_encoder_2", variant=variant, torch_dtype=torch_dtype, device=device))
        return cls(c["scheduler"], vae, c["transformer"], **te)

    def __init__(self, scheduler, vae, transformer, text_encoder=None, text_encoder_2=None):
        self.scheduler, self.vae, self.transformer = scheduler, vae, transformer
        self.text_encoder, self.text_encoder_2 = text_encoder, text_encoder_2  # text_encoders.CLIPTextModel, T5EncoderModel (optional)
        self.vae_scale_factor = 2 ** (len(vae.config.block_out_channels) - 1) if vae is not None else 8
        self.default_sample_size = 128

    @torch.no_grad()
    def encode_prompt(self, clip_input_ids, t5_input_ids):
        """FluxPipeline.encode_prompt (pipeline_flux.py:217-387) from TOKEN IDS: pooled = CLIP `pooler_output`
        (_get_clip_prompt_embeds), prompt_embeds = T5 `[0]` (_get_t5_prompt_embeds), text_ids = zeros.
        -> (prompt_embeds, pooled_prompt_embeds, text_ids)"""
        if self.text_encoder is None or self.text_encoder_2 is None:
            raise ValueError("encode_prompt needs text_encoder (CLIP) and text_encoder_2 (T5) (diffusers_b200.text_encoders)")
        dev = self.transformer.device
        pooled = self.text_encoder(clip_input_ids.to(dev), output_hidden_states=False).pooler_output
        pe = self.text_encoder_2(t5_input_ids.to(dev), output_hidden_states=False)[0]
        dt = self.transformer.dtype
        return pe.to(dt), pooled.to(dt), torch.zeros(pe.shape[1], 3, device=dev, dtype=dt)

    @staticmethod
    def _prepare_latent_image_ids(height, width, device, dtype):
        ids = torch.zeros(height, width, 3)
        ids[..., 1] = ids[..., 1] + torch.arange(height)[:, None]
        ids[..., 2] = ids[..., 2] + torch.arange(width)[None, :]
        return ids.reshape(height * width, 3).to(device=device, dtype=dtype)

    @staticmethod
    def _pack_latents(latents, batch_size, num_channels_latents, height, width):
        latents = latents.view(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
        latents = latents.permute(0, 2, 4, 1, 3, 5)
        return latents.reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)

    @staticmethod
    def _unpack_latents(latents, height, width, vae_scale_factor):
        batch_size, num_patches, channels = latents.shape
        height = 2 * (int(height) // (vae_scale_factor * 2))
        width = 2 * (int(width) // (vae_scale_factor * 2))
        latents = latents.view(batch_size, height // 2, width // 2, channels // 4, 2, 2)
        latents = latents.permute(0, 3, 1, 4, 2, 5)
        return latents.reshape(batch_size, channels // 4, height, width)

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        height = 2 * (int(height) // (self.vae_scale_factor * 2))
        width = 2 * (int(width) // (self.vae_scale_factor * 2))
        ids = self._prepare_latent_image_ids(height // 2, width // 2, device, dtype)
        if latents is not None:
            return latents.to(device=device, dtype=dtype), ids
        latents = randn_tensor((batch_size, num_channels_latents, height, width), generator=generator, device=device, dtype=dtype)
        return self._pack_latents(latents, batch_size, num_channels_latents, height, width), ids

    @torch.no_grad()
    def __call__(self, prompt_embeds, pooled_prompt_embeds, height=None, width=None, num_inference_steps=28,
                 guidance_scale=3.5, generator=None, latents=None, output_type="latent", return_dict=True,
                 negative_prompt_embeds=None, negative_pooled_prompt_embeds=None, true_cfg_scale=1.0):
        """true_cfg_scale > 1 with negative embeddings = the reference's "true" classifier-free guidance for Flux
        (pipeline_flux.py:778-780, 911-927): a second transformer call per step under cache_context("uncond") and
        noise = neg + true_cfg_scale * (cond - neg)."""
        tr, sched = self.transformer, self.scheduler
        device = tr.device
        height = height or self.default_sample_size * self.vae_scale_factor
        width = width or self.default_sample_size * self.vae_scale_factor
        batch = prompt_embeds.shape[0]
        dtype = prompt_embeds.dtype
        prompt_embeds = prompt_embeds.to(device)
        pooled_prompt_embeds = pooled_prompt_embeds.to(device)
        text_ids = torch.zeros(prompt_embeds.shape[1], 3).to(device=device, dtype=dtype)
        do_true_cfg = true_cfg_scale > 1 and negative_prompt_embeds is not None
        if do_true_cfg:
            if negative_pooled_prompt_embeds is None:
                raise ValueError("true CFG needs negative_pooled_prompt_embeds next to negative_prompt_embeds")
            negative_prompt_embeds = negative_prompt_embeds.to(device)
            negative_pooled_prompt_embeds = negative_pooled_prompt_embeds.to(device)
            # all-zero ids of the negative prompt's length: the SAME tensor when the lengths agree, so the transformer's RoPE-table cache holds
            neg_text_ids = text_ids if negative_prompt_embeds.shape[1] == prompt_embeds.shape[1] else \
                torch.zeros(negative_prompt_embeds.shape[1], 3).to(device=device, dtype=dtype)
        lat, img_ids = self.prepare_latents(batch, tr.config.in_channels // 4, height, width, dtype, device, generator, latents)
        sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps)
        c = sched.config
        mu = calculate_shift(lat.shape[1], c.get("base_image_seq_len", 256), c.get("max_image_seq_len", 4096),
                             c.get("base_shift", 0.5), c.get("max_shift", 1.15))
        sched.set_timesteps(num_inference_steps, device=device, sigmas=sigmas, mu=mu)
        timesteps = sched.timesteps
        guidance = None
        if tr.config.guidance_embeds:
            guidance = torch.full([1], guidance_scale, device=device, dtype=torch.float32).expand(lat.shape[0])
        sched.set_begin_index(0)
        for t in timesteps:
            timestep = t.expand(lat.shape[0]).to(lat.dtype)
            with tr.cache_context("cond"):
                noise_pred = tr(hidden_states=lat, timestep=timestep / 1000, guidance=guidance,
                                pooled_projections=pooled_prompt_embeds, encoder_hidden_states=prompt_embeds,
                                txt_ids=text_ids, img_ids=img_ids, joint_attention_kwargs=None, return_dict=False)[0]
            if do_true_cfg:
                with tr.cache_context("uncond"):
                    neg_noise_pred = tr(hidden_states=lat, timestep=timestep / 1000, guidance=guidance,
                                        pooled_projections=negative_pooled_prompt_embeds, encoder_hidden_states=negative_prompt_embeds,
                                        txt_ids=neg_text_ids, img_ids=img_ids, joint_attention_kwargs=None, return_dict=False)[0]
                noise_pred = neg_noise_pred + true_cfg_scale * (noise_pred - neg_noise_pred)
            lat = sched.step(noise_pred, t, lat, return_dict=False)[0]
        if output_type == "latent":
            image = lat
        else:
            lat = self._unpack_latents(lat, height, width, self.vae_scale_factor)
            lat = (lat / self.vae.config.scaling_factor) + self.vae.config.shift_factor
            image = postprocess_pt(self.vae.decode(lat, return_dict=False)[0])
        if not return_dict:
            return (image,)
        return PipelineOutput(image)


class DDPMPipeline:
    """pipelines/ddpm/pipeline_ddpm.py:55-139 (BASELINE.json config 0): ancestral sampling with a fresh noise draw from
    the caller's generator every step."""

    @classmethod
    def from_pretrained(cls, path, torch_dtype=torch.bfloat16, device="cuda"):
        from . import checkpoint
        from .schedulers import DDPMScheduler
        from .unet_2d import UNet2DModel
        c = checkpoint.load_pipeline_components(path, "DDPMPipeline", dict(unet=UNet2DModel, scheduler=DDPMScheduler),
                                                torch_dtype=torch_dtype, device=device)
        return cls(c["unet"], c["scheduler"])

    def __init__(self, unet, scheduler):
        self.unet, self.scheduler = unet, scheduler

    @torch.no_grad()
    def __call__(self, batch_size=1, generator=None, num_inference_steps=1000, output_type="pt", return_dict=True):
        u = self.unet
        size = u.config.sample_size
        shape = (batch_size, u.config.in_channels, size, size) if isinstance(size, int) else (batch_size, u.config.in_channels, *size)
        # the reference draws the initial image in the model dtype (fp32 for config 0): same fp32 draw, then cast
        image = randn_tensor(shape, generator=generator, device=u.device, dtype=torch.float32).to(u.dtype)
        self.scheduler.set_timesteps(num_inference_steps)
        for t in self.scheduler._timesteps_cpu:
            eps = u(image, int(t)).sample
            image = self.scheduler.step(eps, int(t), image, generator=generator).prev_sample
        image = (image / 2 + 0.5).clamp(0, 1)
        if output_type == "np":
            image = image.float().cpu().permute(0, 2, 3, 1).numpy()
        if not return_dict:
            return (image,)
        return PipelineOutput(image)

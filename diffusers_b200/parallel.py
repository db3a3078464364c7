"""Batch data-parallel sampling over the GPUs of one box.

Samples are independent (GroupNorm / LayerNorm / attention are per sample), so rank r owns a contiguous slice of the
prompt batch and its cond+uncond pair stays on that rank: the data path has NO collective.  The only exchange is
one all-gather of the finished outputs over NCCL/NVLink when batch > 1 (decoded images, so VAE work stays sharded).
The seeded noise is drawn for the WHOLE batch from the single caller generator and sliced, exactly what a
single-process run would draw (reference utils/torch_utils.py:183-233 draws the full batch from one generator;
the docs recipe docs/source/en/training/distributed_inference.md:29-48 splits prompts per process).
"""
import torch
import torch.distributed as dist

from .pipelines import randn_tensor


def shard_bounds(n, rank, world):
    """Contiguous, balanced split of n samples: the first n % world ranks get one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t, rank, world):
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def seeded_latent_shard(shape, seed, rank, world, dtype, device):
    """Draw the full-batch N(0,1) latents from one CPU generator (device independent), return this rank's rows."""
    g = torch.Generator().manual_seed(seed)
    full = randn_tensor(shape, generator=g, device="cpu", dtype=dtype)
    lo, hi = shard_bounds(shape[0], rank, world)
    return full[lo:hi].to(device)


def all_gather_batch(local, n_total, group=None):
    """One all-gather of the per-rank outputs [n_local, ...] -> [n_total, ...] in global sample order.  Shards may
    differ by one sample, so they are padded to the largest shard (a single fixed-size collective)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    out = out.view((world, mx) + tuple(local.shape[1:]))
    assert sizes[rank][1] - sizes[rank][0] == local.shape[0]
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)


def sdxl_data_parallel(pipe, prompt_embeds, negative_prompt_embeds, pooled, negative_pooled, *, seed, height, width,
                       num_inference_steps, guidance_scale, output_type="pt", gather=True, group=None, **kw):
    """Every rank passes the FULL-batch embeddings; returns the full batch of images on every rank when gather."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = prompt_embeds.shape[0]
    dev = pipe.device
    lat_shape = (n, pipe.unet.config.in_channels, height // pipe.vae_scale_factor, width // pipe.vae_scale_factor)
    lat = seeded_latent_shard(lat_shape, seed, rank, world, prompt_embeds.dtype, dev)
    sh = lambda t: shard_batch(t, rank, world)  # noqa: E731
    if lat.shape[0] == 0:
        local = torch.empty((0, 3, height, width), dtype=prompt_embeds.dtype, device=dev)
    else:
        local = pipe(sh(prompt_embeds), sh(negative_prompt_embeds), sh(pooled), sh(negative_pooled), height=height,
                     width=width, num_inference_steps=num_inference_steps, guidance_scale=guidance_scale, latents=lat,
                     output_type=output_type, return_dict=False, **kw)[0]
    if gather and world > 1 and n > 1:
        return all_gather_batch(local, n, group)
    return local

"""BASELINE config 3 in miniature, one process per GPU (run by tests/test_parallel_gpu.py through torch.distributed.run):
the prompt batch is sharded over the ranks by parallel.sdxl_data_parallel - one seeded full-batch latent draw sliced per
rank, no collective on the data path, one NCCL all-gather of the decoded images - and must equal, bit for bit, the same
batch sampled by a single process."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusers_b200 import parallel, specs  # noqa: E402
from diffusers_b200.autoencoder_kl import AutoencoderKL  # noqa: E402
from diffusers_b200.pipelines import StableDiffusionXLPipeline  # noqa: E402
from diffusers_b200.schedulers import EulerDiscreteScheduler  # noqa: E402
from diffusers_b200.unet_2d_condition import UNet2DConditionModel  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ucfg = dict(specs.SDXL_UNET_CONFIG)
    ucfg.update(sample_size=16, block_out_channels=(64, 128), down_block_types=("DownBlock2D", "CrossAttnDownBlock2D"),
                up_block_types=("CrossAttnUpBlock2D", "UpBlock2D"), cross_attention_dim=64, transformer_layers_per_block=(1, 1),
                attention_head_dim=(1, 2), addition_time_embed_dim=8, projection_class_embeddings_input_dim=6 * 8 + 16, layers_per_block=1)
    vcfg = dict(specs.SDXL_VAE_CONFIG)
    vcfg.update(block_out_channels=(64, 64), layers_per_block=1, sample_size=32)
    usd = specs.random_state_dict(specs.unet2d_condition_params(ucfg), seed=1)
    vsd = specs.random_state_dict(specs.vae_decoder_params(vcfg), seed=2)
    skw = dict(beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", timestep_spacing="leading", steps_offset=1)
    pipe = StableDiffusionXLPipeline(AutoencoderKL(vcfg, vsd, device=dev), UNet2DConditionModel(ucfg, usd, device=dev), EulerDiscreteScheduler(**skw))
    n = 2 * world + 1  # uneven shards: the first ranks take one sample more
    g = torch.Generator().manual_seed(0)
    mk = lambda *s: torch.randn(*s, generator=g).bfloat16().to(dev)  # noqa: E731
    pe, npe, po, npo = mk(n, 7, 64), mk(n, 7, 64), mk(n, 16), mk(n, 16)
    kw = dict(seed=123, height=32, width=32, num_inference_steps=3, guidance_scale=7.5)
    imgs = parallel.sdxl_data_parallel(pipe, pe, npe, po, npo, **kw)
    assert tuple(imgs.shape) == (n, 3, 32, 32), imgs.shape
    # every rank holds the full batch after the all-gather, and all ranks hold the same bits
    ref0 = imgs.clone()
    dist.broadcast(ref0, src=0)
    assert torch.equal(ref0, imgs)
    if rank == 0:
        lat_hw = 32 // pipe.vae_scale_factor  # the shape parallel.sdxl_data_parallel draws for the whole batch
        lat = parallel.seeded_latent_shard((n, pipe.unet.config.in_channels, lat_hw, lat_hw), 123, 0, 1, torch.bfloat16, dev)
        single = pipe(pe, npe, po, npo, height=32, width=32, num_inference_steps=3, guidance_scale=7.5, latents=lat, output_type="pt").images
        per_sample = torch.cat([pipe(pe[i:i + 1], npe[i:i + 1], po[i:i + 1], npo[i:i + 1], height=32, width=32, num_inference_steps=3,
                                     guidance_scale=7.5, latents=lat[i:i + 1], output_type="pt").images for i in range(n)])
        d_batch = float((imgs.float() - single.float()).abs().max())
        d_single = float((imgs.float() - per_sample.float()).abs().max())
        print(f"config3 mini: {world} ranks, batch {n}: sharded vs one-process batch max |diff| {d_batch:.3g}; vs sample-by-sample {d_single:.3g}", flush=True)
        # GroupNorm reduces per sample but chunks its partial sums by batch size, so different batch shapes may differ in the last
        # bits; the same shard shape must be bit-exact, everything else within 16-bit rounding noise of the 3-step loop
        assert d_batch <= 2e-2 and d_single <= 2e-2
        print("CONFIG3_OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

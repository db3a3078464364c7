"""Context-parallel (Ulysses) Flux forward, one process per GPU (run by tests/test_parallel_gpu.py through
torch.distributed.run; `--full` = the 11.9 B Flux.1-dev shape, 4096 + 512 tokens).  Every rank builds the same model twice:
plain, and with `enable_parallelism(config=ContextParallelConfig(ulysses_degree=world))`; the sharded forward must return,
on EVERY rank, the output of the single-GPU forward (same kernels on the same numbers: row-sharded GEMMs and head-sharded
attention do not change any accumulation order)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffusers_b200 import ops, specs  # noqa: E402
from diffusers_b200.context_parallel import ContextParallelConfig, PeerGroup  # noqa: E402
from diffusers_b200.transformer_flux import FluxTransformer2DModel  # noqa: E402


def inputs(cfg, S, T, dev, B=1):
    g = torch.Generator().manual_seed(11)
    side = int(S ** 0.5)
    hs = torch.randn(B, S, cfg["in_channels"], generator=g).bfloat16().to(dev)
    ehs = torch.randn(B, T, cfg["joint_attention_dim"], generator=g).bfloat16().to(dev)
    pooled = torch.randn(B, cfg["pooled_projection_dim"], generator=g).bfloat16().to(dev)
    img_ids = torch.zeros(side, side, 3)
    img_ids[..., 1] += torch.arange(side)[:, None]
    img_ids[..., 2] += torch.arange(side)[None, :]
    return dict(hidden_states=hs, encoder_hidden_states=ehs, pooled_projections=pooled, timestep=torch.full((B,), 0.7).to(dev).bfloat16(),
                img_ids=img_ids.reshape(S, 3).to(dev).bfloat16(), txt_ids=torch.zeros(T, 3).to(dev).bfloat16(),
                guidance=torch.full((B,), 3.5).to(dev), return_dict=False)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    full = "--full" in sys.argv

    # ---- the primitives alone: peer stores through a mapping + the device-side barrier
    pg = PeerGroup(1 << 20)
    slots = pg.carve(torch.float32, (world, 1024))
    for it in range(3):
        for d in range(world):
            slots[d][rank].fill_(float(100 * it + rank))  # a torch kernel writing into rank d's memory over NVLink
        pg.barrier()
        got = slots[rank][:, 0].clone()
        want = torch.tensor([float(100 * it + r) for r in range(world)], device=dev)
        assert torch.equal(got, want), (rank, it, got, want)
        pg.barrier()  # nobody overwrites a slot before every rank has read it
    torch.cuda.synchronize()
    if rank == 0:
        print(f"peer memory + barrier: {world} ranks OK", flush=True)

    cases = [("flux_tiny_hd64", dict(patch_size=1, in_channels=16, num_layers=2, num_single_layers=3, attention_head_dim=64, num_attention_heads=2 * world,
                                      joint_attention_dim=96, pooled_projection_dim=48, guidance_embeds=True, axes_dims_rope=(8, 28, 28)), 256, 16 * world, 2),
             ("flux_tiny_hd128", dict(patch_size=1, in_channels=16, num_layers=1, num_single_layers=2, attention_head_dim=128, num_attention_heads=world,
                                       joint_attention_dim=64, pooled_projection_dim=32, guidance_embeds=True, axes_dims_rope=(16, 56, 56)), 1024, 24 * world, 1)]
    if full:
        cases = [("flux_dev_shape", dict(specs.FLUX_DEV_CONFIG), 4096, 512, 1)]
    for name, cfg, S, T, B in cases:
        if full:
            plain = FluxTransformer2DModel.random_init(cfg, seed=3, device=dev)
            sd = None
        else:
            sd = specs.random_state_dict(specs.flux_params(dict(specs.FLUX_DEV_CONFIG, **cfg)), seed=3)
            plain = FluxTransformer2DModel(cfg, sd, device=dev)
        kw = inputs(plain.config, S, T, dev, B)
        ref = plain(**kw)[0]
        torch.cuda.synchronize()
        if full:
            t0 = time.perf_counter()
            for _ in range(3):
                plain(**kw)
            torch.cuda.synchronize()
            t_plain = (time.perf_counter() - t0) / 3 * 1e3
            cp = plain  # 23.8 GB of weights: permute the QKV rows of the same model in place
        else:
            cp = FluxTransformer2DModel(cfg, sd, device=dev)
        cp.enable_parallelism(config=ContextParallelConfig(ulysses_degree=world))
        n0 = ops.launches()
        out = cp(**kw)[0]
        torch.cuda.synchronize()
        launches = ops.launches() - n0
        d = (out.float() - ref.float()).abs()
        allr = [None] * world
        dist.all_gather_object(allr, (float(d.max()), float(d.mean())))
        same = out.clone()
        dist.broadcast(same, src=0)
        assert torch.equal(same, out), "ranks disagree on the gathered output"
        if rank == 0:
            print(f"{name}: S={S} T={T} B={B} x {world} ranks: sharded vs single-GPU max |diff| per rank {[f'{a:.3g}' for a, _ in allr]} "
                  f"mean {[f'{b:.3g}' for _, b in allr]} (|ref| max {float(ref.float().abs().max()):.3g}); {launches} launches", flush=True)
        scale = float(ref.float().abs().max())
        assert max(a for a, _ in allr) <= 2e-2 * scale and max(b for _, b in allr) <= 2e-3 * scale, allr
        if full:
            dist.barrier()
            for _ in range(2):
                cp(**kw)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(5):
                cp(**kw)
            torch.cuda.synchronize()
            t_cp = (time.perf_counter() - t0) / 5 * 1e3
            if rank == 0:
                print(f"{name}: forward {t_plain:.2f} ms on one GPU, {t_cp:.2f} ms sharded over {world} (speed-up {t_plain / t_cp:.2f}x)", flush=True)
    dist.barrier()
    if rank == 0:
        print("FLUX_CP_OK", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

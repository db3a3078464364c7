"""world_size-2 `gloo` test of the batch data-parallel host logic (shard bounds, seeded-latent slicing, the single
all-gather with uneven shards) - the N>1 path of SURVEY.md §8e without GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusers_b200 import parallel


def test_shard_bounds_cover_batch():
    for n in (1, 2, 5, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape = (n, 4, 8, 8)
        mine = parallel.seeded_latent_shard(shape, 1234, rank, world, torch.float32, "cpu")
        full = torch.randn(shape, generator=torch.Generator().manual_seed(1234))
        lo, hi = parallel.shard_bounds(n, rank, world)
        assert torch.equal(mine, full[lo:hi])  # same numbers a single process would draw
        local = mine * 2 + 1                   # stand-in for the per-rank sampling result
        gathered = parallel.all_gather_batch(local, n)
        assert torch.equal(gathered, full * 2 + 1)

        class _Pipe:  # duck-typed pipeline: checks sdxl_data_parallel wiring without a GPU
            device = torch.device("cpu")
            vae_scale_factor = 8
            unet = type("U", (), dict(config=type("C", (), dict(in_channels=4))()))()

            def __call__(self, pe, npe, pool, npool, latents=None, **kw):
                return ((latents + pe.mean((1, 2))[:, None, None, None]).repeat(1, 1, 8, 8)[:, :3],)

        pe = torch.arange(n, dtype=torch.float32)[:, None, None].expand(n, 3, 5).contiguous()
        out = parallel.sdxl_data_parallel(_Pipe(), pe, pe, pe[:, 0], pe[:, 0], seed=7, height=64, width=64,
                                          num_inference_steps=1, guidance_scale=5.0)
        ref_lat = torch.randn((n, 4, 8, 8), generator=torch.Generator().manual_seed(7))
        ref = (ref_lat + torch.arange(n, dtype=torch.float32)[:, None, None, None]).repeat(1, 1, 8, 8)[:, :3]
        assert torch.equal(out, ref)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [4, 5])
def test_two_rank_gloo_shard_and_gather(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res

"""GPU diagnostic (test infrastructure - it imports the oracle, so it lives under tests/): product models (UNet / VAE)
vs the oracle on small configs + full-size timing.  Usage: python tests/diag_models.py [tiny_unet tiny_vae full_unet full_vae]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from diffusers_b200 import ops, specs
from diffusers_b200.autoencoder_kl import AutoencoderKL
from diffusers_b200.unet_2d_condition import UNet2DConditionModel
from oracle import unet as ounet
from oracle import vae as ovae

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def stats(name, out, ref, ref16=None):
    o = out.float().cpu()
    err = (o - ref).abs()
    res = dict(case=name, shape=list(o.shape), max_abs=float(err.max()), mean_abs=float(err.mean()),
               ref_absmax=float(ref.abs().max()), ref_absmean=float(ref.abs().mean()), nan=int(torch.isnan(o).sum()))
    if ref16 is not None:
        e16 = (ref16.float() - ref).abs()
        res["ref16_max_abs"] = float(e16.max())
        res["ref16_mean_abs"] = float(e16.mean())
    print("RESULT " + json.dumps(res), flush=True)


def tiny_unet():
    cfg = dict(specs.SDXL_UNET_CONFIG)
    cfg.update(sample_size=16, block_out_channels=(64, 128, 256), cross_attention_dim=128,
               transformer_layers_per_block=(1, 2, 3), attention_head_dim=(1, 2, 4), addition_time_embed_dim=32,
               projection_class_embeddings_input_dim=6 * 32 + 64)
    sd = specs.random_state_dict(specs.unet2d_condition_params(cfg), seed=1, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 16, 16, generator=g).bfloat16()
    ehs = torch.randn(2, 77, 128, generator=g).bfloat16()
    te = torch.randn(2, 64, generator=g).bfloat16()
    tid = torch.tensor([[128., 128, 0, 0, 128, 128]] * 2).bfloat16()
    t = torch.tensor(981.0)
    sd32 = {k: v.float() for k, v in sd.items()}
    ref = ounet.unet2d_condition_forward(sd32, cfg, x.float(), t, ehs.float(), dict(text_embeds=te.float(), time_ids=tid.float()))
    ref16 = ounet.unet2d_condition_forward(sd, cfg, x, t, ehs, dict(text_embeds=te, time_ids=tid))
    m = UNet2DConditionModel(cfg, sd, dtype=torch.bfloat16, device="cuda")
    out = m(x.cuda(), t.cuda(), ehs.cuda(), added_cond_kwargs=dict(text_embeds=te.cuda(), time_ids=tid.cuda()), return_dict=False)[0]
    torch.cuda.synchronize()
    stats("unet_tiny", out, ref, ref16)
    m.enable_cuda_graph(True)
    out2 = m(x.cuda(), t.cuda(), ehs.cuda(), added_cond_kwargs=dict(text_embeds=te.cuda(), time_ids=tid.cuda()), return_dict=False)[0]
    out3 = m(x.cuda(), t.cuda(), ehs.cuda(), added_cond_kwargs=dict(text_embeds=te.cuda(), time_ids=tid.cuda()), return_dict=False)[0]
    torch.cuda.synchronize()
    print("graph_equal", bool(torch.equal(out, out2)), bool(torch.equal(out2, out3)), flush=True)


def tiny_vae(boc=(64, 64, 128, 128), name="vae_tiny"):
    cfg = dict(specs.SDXL_VAE_CONFIG)
    cfg.update(block_out_channels=boc, sample_size=64)
    sd = specs.random_state_dict(specs.vae_decoder_params(cfg), seed=2, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(2, 4, 16, 16, generator=g).bfloat16()
    sd32 = {k: v.float() for k, v in sd.items()}
    ref = ovae.vae_decode(sd32, cfg, z.float())
    ref16 = ovae.vae_decode(sd, cfg, z)
    m = AutoencoderKL(cfg, sd, dtype=torch.bfloat16, device="cuda")
    out = m.decode(z.cuda(), return_dict=False)[0]
    torch.cuda.synchronize()
    stats(name, out, ref, ref16)


def time_it(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]


def full_unet():
    t0 = time.time()
    m = UNet2DConditionModel.random_init(seed=0, dtype=torch.bfloat16, device="cuda")
    print("sdxl unet init s", round(time.time() - t0, 1), flush=True)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(2, 4, 128, 128, generator=g, device="cuda").bfloat16()
    ehs = torch.randn(2, 77, 2048, generator=g, device="cuda").bfloat16()
    te = torch.randn(2, 1280, generator=g, device="cuda").bfloat16()
    tid = torch.tensor([[1024., 1024, 0, 0, 1024, 1024]] * 2, device="cuda").bfloat16()
    t = torch.tensor(981.0, device="cuda")
    kw = dict(added_cond_kwargs=dict(text_embeds=te, time_ids=tid), return_dict=False)
    n0 = ops.launches()
    out = m(x, t, ehs, **kw)[0]
    torch.cuda.synchronize()
    print("sdxl unet launches/forward", ops.launches() - n0, "out absmax", float(out.float().abs().max()), "nan", int(torch.isnan(out.float()).sum()), flush=True)
    ms = time_it(lambda: m(x, t, ehs, **kw))
    print("sdxl unet eager ms", [round(v, 2) for v in ms], flush=True)
    m.enable_cuda_graph(True)
    ms = time_it(lambda: m(x, t, ehs, **kw), n=10)
    print("sdxl unet graph ms", [round(v, 2) for v in ms], "TFLOP/s", round(13.52 / (min(ms) * 1e-3), 1), flush=True)
    out2 = m(x, t, ehs, **kw)[0]
    print("graph vs eager equal", bool(torch.equal(out, out2)), flush=True)
    del m
    torch.cuda.empty_cache()


def full_vae():
    m = AutoencoderKL.random_init(seed=0, dtype=torch.bfloat16, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    z = torch.randn(1, 4, 128, 128, generator=g, device="cuda").bfloat16()
    n0 = ops.launches()
    out = m.decode(z, return_dict=False)[0]
    torch.cuda.synchronize()
    print("vae launches", ops.launches() - n0, tuple(out.shape), "absmax", float(out.float().abs().max()), "nan", int(torch.isnan(out.float()).sum()), flush=True)
    ms = time_it(lambda: m.decode(z, return_dict=False), n=5)
    print("vae decode ms", [round(v, 2) for v in ms], "TFLOP/s", round(10.47 / (min(ms) * 1e-3), 1), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["tiny_unet", "tiny_vae", "vae512", "full_unet", "full_vae"]
    for w in which:
        try:
            if w == "vae512":
                tiny_vae(boc=(64, 64, 128, 512), name="vae_d512")
            else:
                globals()[w]()
        except Exception as e:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print("ERROR in", w, str(e)[:300], flush=True)
            if "CUDA" in str(e) or "launch" in str(e):
                break

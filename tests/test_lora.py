"""LoRA merged into the base weights before packing: W + scale * (alpha / r) * B A reproduces what the adapter layers
compute (Linear and Conv2d), for the key namings diffusers serialises; wrong / unknown targets are loud; a merged tiny UNet state dict still
packs and inverts exactly."""
import pytest
import torch
import torch.nn.functional as F

from diffusers_b200.lora import merge_lora


def test_merged_weights_equal_the_adapter_forward():
    g = torch.Generator().manual_seed(0)
    sd = {"blk.to_q.weight": torch.randn(48, 32, generator=g), "blk.to_q.bias": torch.randn(48, generator=g), "blk.conv.weight": torch.randn(16, 8, 3, 3, generator=g),
          "blk.other.weight": torch.randn(4, 4, generator=g)}
    A, B = torch.randn(4, 32, generator=g), torch.randn(48, 4, generator=g)
    Ac, Bc = torch.randn(2, 8, 3, 3, generator=g), torch.randn(16, 2, 1, 1, generator=g)
    lora = {"unet.blk.to_q.lora_A.weight": A, "unet.blk.to_q.lora_B.weight": B, "unet.blk.to_q.alpha": torch.tensor(8.0),
            "unet.blk.conv.lora.down.weight": Ac, "unet.blk.conv.lora.up.weight": Bc, "text_encoder.x.lora_A.weight": A, "text_encoder.x.lora_B.weight": B}
    out = merge_lora(sd, lora, prefix="unet", scale=0.7)
    x = torch.randn(5, 32, generator=g)
    want = F.linear(x, sd["blk.to_q.weight"]) + 0.7 * (8.0 / 4) * F.linear(F.linear(x, A), B)
    assert torch.allclose(F.linear(x, out["blk.to_q.weight"]), want, atol=1e-4)
    xi = torch.randn(2, 8, 9, 9, generator=g)
    wantc = F.conv2d(xi, sd["blk.conv.weight"], padding=1) + 0.7 * F.conv2d(F.conv2d(xi, Ac, padding=1), Bc)   # alpha defaults to the rank
    assert torch.allclose(F.conv2d(xi, out["blk.conv.weight"], padding=1), wantc, atol=1e-4)
    assert torch.equal(out["blk.other.weight"], sd["blk.other.weight"]) and torch.equal(out["blk.to_q.bias"], sd["blk.to_q.bias"]) and out is not sd
    # peft adapter-name infix and network_alphas
    out2 = merge_lora(sd, {"unet.blk.to_q.lora_A.default_0.weight": A, "unet.blk.to_q.lora_B.default_0.weight": B}, network_alphas={"unet.blk.to_q.alpha": 8.0}, scale=0.7)
    assert torch.allclose(out2["blk.to_q.weight"], out["blk.to_q.weight"])
    # 16-bit base weights: fp32 arithmetic, one rounding
    sd16 = {k: v.bfloat16() for k, v in sd.items()}
    o16 = merge_lora(sd16, lora, scale=0.7)["blk.to_q.weight"]
    assert o16.dtype == torch.bfloat16 and torch.equal(o16, (sd16["blk.to_q.weight"].float() + 0.7 * 2.0 * (B @ A)).bfloat16())


def test_merge_is_loud_about_what_it_cannot_place():
    sd = {"a.weight": torch.zeros(4, 4)}
    A, B = torch.zeros(2, 4), torch.zeros(4, 2)
    with pytest.raises(KeyError):
        merge_lora(sd, {"unet.b.lora_A.weight": A, "unet.b.lora_B.weight": B})           # module the model does not have
    with pytest.raises(KeyError):
        merge_lora(sd, {"unet.a.lora_A.weight": A})                                      # down without up
    with pytest.raises(ValueError):
        merge_lora(sd, {"unet.a.lora_A.weight": torch.zeros(2, 5), "unet.a.lora_B.weight": B})  # shapes
    with pytest.raises(KeyError):
        merge_lora(sd, {"transformer.a.lora_A.weight": A, "transformer.a.lora_B.weight": B})     # nothing under the prefix
    assert torch.equal(merge_lora(sd, {"unet.b.lora_A.weight": A, "unet.b.lora_B.weight": B, "unet.a.lora_A.weight": A, "unet.a.lora_B.weight": B}, strict=False)["a.weight"], sd["a.weight"])


def test_merged_unet_state_dict_packs_and_inverts():
    """A LoRA on the attention projections of a tiny SDXL-style UNet: the shell builds from the merged state dict (host side) and its packing
    inverse returns exactly the merged weights - the adapter costs nothing at run time."""
    from diffusers_b200 import specs
    from diffusers_b200.unet_2d_condition import UNet2DConditionModel
    cfg = dict(specs.SDXL_UNET_CONFIG)
    cfg.update(sample_size=16, block_out_channels=(64, 64), down_block_types=("DownBlock2D", "CrossAttnDownBlock2D"), up_block_types=("CrossAttnUpBlock2D", "UpBlock2D"),
               layers_per_block=1, cross_attention_dim=64, transformer_layers_per_block=1, attention_head_dim=(1, 1), addition_time_embed_dim=32,
               projection_class_embeddings_input_dim=256)
    sd = specs.random_state_dict(specs.unet2d_condition_params(cfg), seed=3, dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(1)
    lora = {}
    targets = [k[:-7] for k in sd if k.endswith((".to_q.weight", ".to_k.weight", ".to_v.weight", ".to_out.0.weight"))]
    for t in targets:
        n, k = sd[t + ".weight"].shape
        lora[f"unet.{t}.lora_A.weight"] = torch.randn(4, k, generator=g) * 0.1
        lora[f"unet.{t}.lora_B.weight"] = torch.randn(n, 4, generator=g) * 0.1
    merged = merge_lora(sd, lora, scale=0.8)
    assert len(targets) > 8 and all(not torch.equal(merged[t + ".weight"], sd[t + ".weight"]) for t in targets)
    m = UNet2DConditionModel(cfg, merged, device="cpu", fold_norms=False)
    back = m.reference_state_dict()
    assert all(torch.equal(back[k], merged[k]) for k in merged)

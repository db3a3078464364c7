"""GPU parity of the text encoders against the real transformers classes the reference's pipelines call (outputs
recorded in fp32 and bf16 CPU eager by oracle/make_golden.py text) and of b200_text_attention against an fp32 evaluation.
Criterion as in test_models_gpu.py: our error against the fp32 run is no larger than 1.5x the error of transformers' own bf16 run."""
import pytest
import torch

from diffusers_b200 import ops
from diffusers_b200 import text_encoders as T

pytestmark = pytest.mark.gpu


def _close(name, out, ref32, ref16):
    o = out.float().cpu()
    assert tuple(o.shape) == tuple(ref32.shape) and not torch.isnan(o).any()
    err, e16 = (o - ref32).abs(), (ref16.float() - ref32).abs()
    print(f"{name}: ours max {float(err.max()):.4g} mean {float(err.mean()):.4g} | transformers bf16 max {float(e16.max()):.4g} mean {float(e16.mean()):.4g}")
    assert float(err.mean()) <= 1.5 * float(e16.mean()) + 1e-3
    assert float(err.max()) <= 2.0 * float(e16.max()) + 1e-2


@pytest.mark.parametrize("name", ["clip_l_tiny", "clip_g_tiny", "t5_tiny", "t5_tiny_512"])
def test_text_encoder(golden, name):
    fx = golden("text")[name]
    kind = fx["kind"]
    spec = T.t5_encoder_params(fx["cfg"]) if kind == "t5" else T.clip_text_params(fx["cfg"], kind == "clip_proj")
    sd16 = T.random_state_dict(spec, fx["seed"])
    cls = dict(t5=T.T5EncoderModel, clip=T.CLIPTextModel, clip_proj=T.CLIPTextModelWithProjection)[kind]
    m = cls(fx["cfg"], sd16, dtype=torch.bfloat16, device="cuda")
    out = m(fx["ids"].cuda(), output_hidden_states=True)
    r32, r16 = fx["ref32"], fx["ref16"]
    _close(name + " last_hidden_state", out.last_hidden_state, r32["last_hidden_state"], r16["last_hidden_state"])
    assert len(out.hidden_states) == r32["n_hidden"]
    _close(name + " hidden_states[-2]", out.hidden_states[-2], r32["penultimate"], r16["penultimate"])
    if kind == "clip":
        _close(name + " pooler_output", out.pooler_output, r32["pooler_output"], r16["pooler_output"])
        assert out[0] is out.last_hidden_state and out[1] is out.pooler_output  # BaseModelOutputWithPooling order
    if kind == "clip_proj":
        _close(name + " text_embeds", out.text_embeds, r32["text_embeds"], r16["text_embeds"])
        assert out[0] is out.text_embeds and out[0].ndim == 2  # what encode_prompt takes as the pooled embedding
    if kind == "t5":
        plain = m(fx["ids"].cuda(), output_hidden_states=False)
        assert len(plain) == 1 and torch.equal(plain[0], out.last_hidden_state)
    with pytest.raises(NotImplementedError):
        m(fx["ids"].cuda(), attention_mask=torch.ones_like(fx["ids"]).cuda())


@pytest.mark.parametrize("B,H,Sq,Sk,causal,bias,dt", [(2, 12, 77, 77, True, False, torch.bfloat16), (1, 64, 512, 512, False, True, torch.bfloat16),
                                                       (3, 5, 96, 96, False, True, torch.float16), (1, 2, 33, 200, False, False, torch.bfloat16),
                                                       (2, 3, 130, 130, True, True, torch.bfloat16)])
def test_text_attention_kernel(B, H, Sq, Sk, causal, bias, dt):
    g = torch.Generator(device="cuda").manual_seed(3)
    D = H * 64
    q = torch.randn(B, Sq, D, generator=g, device="cuda").to(dt)
    kv = torch.randn(B, Sk, 2 * D, generator=g, device="cuda").to(dt)
    k, v = kv[:, :, :D], kv[:, :, D:]
    bt = (torch.randn(H, Sq, Sk, generator=g, device="cuda") * 2).contiguous() if bias else None
    scale = 0.125 if not bias else 0.3
    out = ops.text_attention(q, k, v, heads=H, scale=scale, causal=causal, bias=bt)
    torch.cuda.synchronize()
    qf, kf, vf = (t.float().reshape(B, -1, H, 64).transpose(1, 2) for t in (q, k, v))
    s = qf @ kf.transpose(-1, -2) * scale
    if bt is not None:
        s = s + bt[None]
    if causal:
        s = s + torch.full((Sq, Sk), float("-inf"), device="cuda").triu(1)
    ref = (torch.softmax(s, -1) @ vf).transpose(1, 2).reshape(B, Sq, D)
    err = (out.float() - ref).abs()
    tol = 2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11
    assert float((err - tol * ref.abs()).max()) <= 2e-3, float(err.max())  # fp32 arithmetic, one rounding of the result


def test_pipelines_encode_prompt_from_token_ids(golden):
    """encode_prompt of both pipelines composes the encoders exactly as the reference does (pipeline_stable_diffusion_xl.py:283-470,
    pipeline_flux.py:217-387): penultimate hidden states concatenated + projection pooled (SDXL); T5 states + CLIP pooler (Flux)."""
    from diffusers_b200.pipelines import FluxPipeline, StableDiffusionXLPipeline
    fx = golden("text")
    mk = lambda name, cls, spec_fn: cls(fx[name]["cfg"], T.random_state_dict(spec_fn(fx[name]["cfg"]), fx[name]["seed"]), device="cuda")  # noqa: E731
    te1 = mk("clip_l_tiny", T.CLIPTextModel, lambda c: T.clip_text_params(c, False))
    te2 = mk("clip_g_tiny", T.CLIPTextModelWithProjection, lambda c: T.clip_text_params(c, True))
    t5 = mk("t5_tiny", T.T5EncoderModel, T.t5_encoder_params)
    unet = type("U", (), dict(config=type("C", (), dict(sample_size=16))(), device=torch.device("cuda"), dtype=torch.bfloat16))()
    pipe = StableDiffusionXLPipeline(None, unet, None, text_encoder=te1, text_encoder_2=te2)
    ids1, ids2 = fx["clip_l_tiny"]["ids"][:2], fx["clip_g_tiny"]["ids"][:2]
    pe, npe, pooled, npooled = pipe.encode_prompt(ids1, ids2)
    assert tuple(pe.shape) == (2, 77, 128 + 192) and tuple(pooled.shape) == (2, 64)
    assert float(npe.abs().max()) == 0 and float(npooled.abs().max()) == 0
    _close("sdxl prompt_embeds[..., :128]", pe[..., :128], fx["clip_l_tiny"]["ref32"]["penultimate"], fx["clip_l_tiny"]["ref16"]["penultimate"])
    _close("sdxl prompt_embeds[..., 128:]", pe[..., 128:], fx["clip_g_tiny"]["ref32"]["penultimate"][:2], fx["clip_g_tiny"]["ref16"]["penultimate"][:2])
    _close("sdxl pooled", pooled, fx["clip_g_tiny"]["ref32"]["text_embeds"][:2], fx["clip_g_tiny"]["ref16"]["text_embeds"][:2])
    pe2, npe2, _, _ = pipe.encode_prompt(ids1, ids2, negative_input_ids=ids1.flip(0), negative_input_ids_2=ids2.flip(0))
    assert torch.equal(pe2, pe) and float((npe2.float() - pe.flip(0).float()).abs().max()) <= 1e-2  # rows are independent of their batch position
    tr = type("Tr", (), dict(device=torch.device("cuda"), dtype=torch.bfloat16))()
    vae = type("V", (), dict(config=type("C", (), dict(block_out_channels=(1, 1, 1, 1)))()))()
    fpipe = FluxPipeline(None, vae, tr, text_encoder=te1, text_encoder_2=t5)
    fpe, fpool, tids = fpipe.encode_prompt(ids1, fx["t5_tiny"]["ids"])
    _close("flux prompt_embeds", fpe, fx["t5_tiny"]["ref32"]["last_hidden_state"], fx["t5_tiny"]["ref16"]["last_hidden_state"])
    _close("flux pooled", fpool, fx["clip_l_tiny"]["ref32"]["pooler_output"], fx["clip_l_tiny"]["ref16"]["pooler_output"])
    assert tuple(tids.shape) == (96, 3) and float(tids.abs().max()) == 0

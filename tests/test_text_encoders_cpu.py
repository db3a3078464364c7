"""Host-side behaviour of the text-encoder shells without a GPU: construction from transformers-named state dicts, config defaults,
refusal of what the path does not implement, no CPU fallback, output objects that index like transformers' ModelOutput."""
import pytest
import torch

from diffusers_b200 import ops
from diffusers_b200 import text_encoders as T

CLIP = dict(vocab_size=200, hidden_size=128, intermediate_size=256, num_hidden_layers=1, num_attention_heads=2)
T5 = dict(vocab_size=200, d_model=128, d_kv=64, d_ff=256, num_layers=1, num_heads=2)


def test_shells_build_on_the_host_and_refuse_cpu_tensors():
    m = T.CLIPTextModelWithProjection(dict(CLIP, projection_dim=64), T.random_state_dict(T.clip_text_params(dict(T.CLIP_L_CONFIG, **CLIP, projection_dim=64), True)),
                                      device="cpu")
    assert m.config.hidden_act == "quick_gelu" and m.config.eos_token_id == 2 and m.dtype == torch.bfloat16 and m.device.type == "cpu"
    ids = torch.randint(0, 199, (1, 77))
    with pytest.raises(NotImplementedError):
        m(ids, attention_mask=torch.ones_like(ids))
    with pytest.raises(ValueError):
        m(torch.randint(0, 199, (1, 78)))  # longer than max_position_embeddings
    with pytest.raises(ops.B200Error):  # the numerics exist only as CUDA kernels
        m(ids)
    t5 = T.T5EncoderModel(T5, T.random_state_dict(T.t5_encoder_params(dict(T.T5_XXL_CONFIG, **T5))), device="cpu")
    assert t5.config.feed_forward_proj == "gated-gelu" and t5.inner == 128
    with pytest.raises(NotImplementedError):
        t5(torch.randint(0, 199, (1, 600)))  # more than 512 tokens
    with pytest.raises(ops.B200Error):
        t5(torch.randint(0, 199, (1, 64)))


def test_shells_refuse_configs_outside_the_kernels():
    sd = T.random_state_dict(T.clip_text_params(dict(T.CLIP_L_CONFIG, **CLIP), False))
    with pytest.raises(NotImplementedError):
        T.CLIPTextModel(dict(CLIP, num_attention_heads=4), sd, device="cpu")  # head_dim 32
    with pytest.raises(NotImplementedError):
        T.CLIPTextModel(dict(CLIP, hidden_act="relu"), sd, device="cpu")
    with pytest.raises(ValueError):
        T.CLIPTextModel(CLIP, {k: v for k, v in sd.items() if "final_layer_norm" not in k}, device="cpu")
    t5sd = T.random_state_dict(T.t5_encoder_params(dict(T.T5_XXL_CONFIG, **T5)))
    with pytest.raises(NotImplementedError):
        T.T5EncoderModel(dict(T5, feed_forward_proj="relu"), t5sd, device="cpu")
    with pytest.raises(NotImplementedError):
        T.T5EncoderModel(dict(T5, d_kv=32), t5sd, device="cpu")
    # tied embedding: a checkpoint that only carries encoder.embed_tokens.weight loads too
    alt = dict(t5sd)
    alt["encoder.embed_tokens.weight"] = alt.pop("shared.weight")
    assert torch.equal(T.T5EncoderModel(T5, alt, device="cpu").W("w0"), t5sd["shared.weight"])


def test_output_object_indexes_like_a_model_output():
    a, b = torch.zeros(2, 3), torch.ones(2, 4, 3)
    o = T._Output(text_embeds=a, last_hidden_state=b, hidden_states=None)
    assert o[0] is a and o[1] is b and len(o) == 2 and o.keys() == ["text_embeds", "last_hidden_state"] and o.hidden_states is None
    assert o["last_hidden_state"] is b
    with pytest.raises(IndexError):
        o[2]


def test_t5_bucket_properties():
    rel = torch.arange(-600, 601)
    b = T.t5_relative_position_bucket(rel, 32, 128)
    assert int(b.min()) == 0 and int(b.max()) == 31
    assert b[rel == 0].item() == 0 and b[rel == 3].item() == 16 + 3 and b[rel == -3].item() == 3       # exact buckets near zero, sign in the upper half
    assert b[rel == 500].item() == 31 and b[rel == -500].item() == 15                                    # saturate beyond max_distance
    assert bool((b[rel > 0][1:] >= b[rel > 0][:-1]).all())                                                # monotone in the distance


def test_pipelines_encode_prompt_composition_on_the_host(golden):
    """encode_prompt of both pipelines with stand-in encoders that answer from the ORACLE (same output objects as the CUDA shells):
    checks the composition the reference prescribes (pipeline_stable_diffusion_xl.py:283-470, pipeline_flux.py:217-387) against the
    recorded transformers outputs - penultimate hidden states concatenated, pooled = [0] of the projection encoder, zeros for an
    empty negative prompt, T5 states + CLIP pooler for Flux."""
    from diffusers_b200.pipelines import FluxPipeline, StableDiffusionXLPipeline
    from oracle import text as otext
    fx = golden("text")

    class Stand:
        def __init__(self, name):
            f = fx[name]
            self.kind, self.cfg = f["kind"], f["cfg"]
            spec = T.t5_encoder_params(self.cfg) if self.kind == "t5" else T.clip_text_params(self.cfg, self.kind == "clip_proj")
            self.sd = {k: v.float() for k, v in T.random_state_dict(spec, f["seed"]).items()}

        def __call__(self, ids, output_hidden_states=False):
            if self.kind == "t5":
                o = otext.t5_encoder_forward(self.sd, self.cfg, ids)
                return T._Output(last_hidden_state=o["last_hidden_state"], hidden_states=o["hidden_states"] if output_hidden_states else None)
            o = otext.clip_text_forward(self.sd, self.cfg, ids, with_projection=self.kind == "clip_proj")
            hs = o["hidden_states"] if output_hidden_states else None
            if self.kind == "clip_proj":
                return T._Output(text_embeds=o["text_embeds"], last_hidden_state=o["last_hidden_state"], hidden_states=hs)
            return T._Output(last_hidden_state=o["last_hidden_state"], pooler_output=o["pooler_output"], hidden_states=hs)

    te1, te2, t5 = Stand("clip_l_tiny"), Stand("clip_g_tiny"), Stand("t5_tiny")
    unet = type("U", (), dict(config=type("C", (), dict(sample_size=16))(), device=torch.device("cpu"), dtype=torch.float32))()
    pipe = StableDiffusionXLPipeline(None, unet, None, text_encoder=te1, text_encoder_2=te2)
    ids1, ids2 = fx["clip_l_tiny"]["ids"][:2], fx["clip_g_tiny"]["ids"][:2]
    pe, npe, pooled, npooled = pipe.encode_prompt(ids1, ids2)
    assert tuple(pe.shape) == (2, 77, 128 + 192) and tuple(pooled.shape) == (2, 64)
    assert float(npe.abs().max()) == 0 and float(npooled.abs().max()) == 0 and npe.shape == pe.shape
    close = lambda a, b: float((a - b).abs().max()) <= 2e-5 * max(1.0, float(b.abs().max()))  # noqa: E731
    assert close(pe[..., :128], fx["clip_l_tiny"]["ref32"]["penultimate"]) and close(pe[..., 128:], fx["clip_g_tiny"]["ref32"]["penultimate"][:2])
    assert close(pooled, fx["clip_g_tiny"]["ref32"]["text_embeds"][:2])
    pe2, npe2, _, npool2 = pipe.encode_prompt(ids1, ids2, negative_input_ids=ids1.flip(0), negative_input_ids_2=ids2.flip(0))
    assert torch.equal(pe2, pe) and close(npe2, pe.flip(0)) and close(npool2, pooled.flip(0))
    assert pipe.encode_prompt(ids1, ids2, do_classifier_free_guidance=False)[1] is None
    pipe.force_zeros_for_empty_prompt = False
    with pytest.raises(ValueError):
        pipe.encode_prompt(ids1, ids2)
    with pytest.raises(ValueError):
        StableDiffusionXLPipeline(None, unet, None).encode_prompt(ids1, ids2)
    tr = type("Tr", (), dict(device=torch.device("cpu"), dtype=torch.float32))()
    vae = type("V", (), dict(config=type("C", (), dict(block_out_channels=(1, 1, 1, 1)))()))()
    fpe, fpool, tids = FluxPipeline(None, vae, tr, text_encoder=te1, text_encoder_2=t5).encode_prompt(ids1, fx["t5_tiny"]["ids"])
    assert close(fpe, fx["t5_tiny"]["ref32"]["last_hidden_state"]) and close(fpool, fx["clip_l_tiny"]["ref32"]["pooler_output"])
    assert tuple(tids.shape) == (96, 3) and float(tids.abs().max()) == 0

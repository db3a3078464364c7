"""Pins the ORACLE (oracle/*.py, the CPU restatement the GPU parity tests are judged against) to the reference:
 * the reference's own golden vectors for this path (9 block slices of tests/models/unets/test_unet_2d_blocks.py,
   the Euler / DDPM scheduler known answers of tests/schedulers/),
 * outputs of the reference itself (models, schedulers, full pipelines) recorded by oracle/make_golden.py.
Runs on CPU; nothing here needs /root/reference."""
import numpy as np
import pytest
import torch

from conftest import state_dicts
from diffusers_b200 import specs
from oracle import blocks as Bk
from oracle import flux as oflux
from oracle import nn as N
from oracle import pipelines as opipe
from oracle import schedulers as osched
from oracle import unet as ounet
from oracle import vae as ovae
from oracle.make_golden import BLOCK_KATS, block_inputs

torch.set_grad_enabled(False)


def _run_block(name, sd, inp):
    x, temb = inp["hidden_states"], inp.get("temb")
    res = inp.get("res_hidden_states_tuple")
    p = "b"
    sd = {p + "." + k: v for k, v in sd.items()}
    if name == "DownBlock2D":
        return Bk.down_block_2d(sd, p, x, temb)[0]
    if name == "AttnDownBlock2D":
        return Bk.attn_down_block_2d(sd, p, x, temb, head_dim=1)[0]
    if name == "CrossAttnDownBlock2D":
        return Bk.cross_attn_down_block_2d(sd, p, x, temb, None, heads=1)[0]
    if name == "UNetMidBlock2D":
        return Bk.unet_mid_block_2d(sd, p, x, temb, head_dim=1)
    if name == "UNetMidBlock2DCrossAttn":
        return Bk.unet_mid_block_2d_cross_attn(sd, p, x, temb, None, heads=1)
    if name == "UpBlock2D":
        return Bk.up_block_2d(sd, p, x, res, temb)
    if name == "CrossAttnUpBlock2D":
        return Bk.cross_attn_up_block_2d(sd, p, x, res, temb, None, heads=1)
    if name == "AttnUpBlock2D":
        return Bk.attn_up_block_2d(sd, p, x, res, temb, head_dim=1)
    if name == "UpDecoderBlock2D":
        return Bk.up_decoder_block_2d(sd, p, x)
    raise KeyError(name)


@pytest.mark.parametrize("name", [k[0] for k in BLOCK_KATS])
def test_block_matches_reference_golden_slice(golden, name):
    fx = golden("blocks")[name]
    out = _run_block(name, fx["state_dict"], block_inputs(fx["flags"]))
    assert tuple(out.shape) == tuple(fx["output_shape"])
    sl = out[0, -1, -3:, -3:].flatten()
    # the reference's hard-coded slice and tolerance (test_unet_blocks_common.py:107-111)
    assert torch.allclose(sl, fx["expected_slice"], atol=5e-3), (sl, fx["expected_slice"])
    # and the reference's actual output on this machine class
    assert torch.allclose(sl, fx["output_slice"], atol=1e-5)
    assert abs(float(out.abs().mean()) - fx["output_abs_mean"]) < 1e-5


def _dummy_sample_deter():
    """tests/schedulers/test_schedulers.py:342-354"""
    n = 4 * 3 * 8 * 8
    s = torch.arange(n).reshape(3, 8, 8, 4) / n
    return s.permute(3, 0, 1, 2)


def _dummy_model(sample, t):
    if isinstance(t, torch.Tensor):
        t = t.reshape(-1, *(1,) * (sample.dim() - 1)).to(dtype=sample.dtype)
    return sample * t / (t + 1)


LAYER_KATS = ["upsample_default", "upsample_with_conv", "upsample_with_conv_out_dim", "downsample_with_conv",
              "downsample_with_conv_pad1", "downsample_with_conv_out_dim", "resnet_default", "resnet_use_in_shortcut",
              "transformer2d_default", "transformer2d_cross_attention_dim"]


@pytest.mark.parametrize("name", LAYER_KATS)
def test_layer_matches_reference_golden_slice(golden, name):
    """Layer-level known answers hard-coded in the reference's tests/models/test_layers_utils.py (:113-370): Upsample2D,
    Downsample2D, ResnetBlock2D, Transformer2DModel.  Weights = the reference's default init recorded in layers.pt, inputs
    re-drawn from the test's seed, expected values = the literals of the reference test (atol 1e-3 as there)."""
    import torch.nn.functional as F
    fx = golden("layers")[name]
    sd = fx["state_dict"]
    torch.manual_seed(0)
    x = torch.randn(*fx["shape"])
    temb = torch.randn(1, 128) if fx["extra"] == "temb" else None
    init = fx["init"]
    with torch.no_grad():
        if fx["cls"] == "Upsample2D":
            o = N.upsample2d({"u." + k: v for k, v in sd.items()}, "u", x) if init["use_conv"] else F.interpolate(x, scale_factor=2.0, mode="nearest")
        elif fx["cls"] == "Downsample2D":
            o = N.downsample2d({"d." + k: v for k, v in sd.items()}, "d", x, padding=init.get("padding", 1))
        elif fx["cls"] == "ResnetBlock2D":
            o = N.resnet_block({"r." + k: v for k, v in sd.items()}, "r", x, temb, groups=32, eps=1e-6)
        else:
            o = N.transformer_2d({"t." + k: v for k, v in sd.items()}, "t", x, fx["context"], heads=init["num_attention_heads"],
                                 num_layers=1, use_linear_projection=False)
    assert tuple(o.shape) == tuple(fx["output_shape"])
    sl = o[0, -1, -3:, -3:].flatten()
    assert torch.allclose(sl, fx["expected_slice"], atol=1e-3), (sl, fx["expected_slice"])
    assert abs(float(o.abs().mean()) - fx["output_abs_mean"]) < 1e-4


def test_timestep_embedding_known_answers():
    """tests/models/test_layers_utils.py:37-109 (EmbeddingsTests): the properties and the hard-coded sinusoid values for the
    three conventions (score-sde: shift 1 no flip; ldm / SDXL / Flux: shift 0 flipped; grad-tts: scale 1000)."""
    t = torch.arange(128)
    t1 = N.get_timestep_embedding(t, 64, downscale_freq_shift=1, flip_sin_to_cos=False)
    t2 = N.get_timestep_embedding(t, 64, downscale_freq_shift=0, flip_sin_to_cos=True)
    t3 = N.get_timestep_embedding(t, 64, scale=1000)
    assert torch.allclose(t1[23:26, 47:50].flatten(), torch.tensor([0.9646, 0.9804, 0.9892, 0.9615, 0.9787, 0.9882, 0.9582, 0.9769, 0.9872]), 1e-3)
    assert torch.allclose(t2[23:26, 47:50].flatten(), torch.tensor([0.3019, 0.2280, 0.1716, 0.3146, 0.2377, 0.1790, 0.3272, 0.2474, 0.1864]), 1e-3)
    assert torch.allclose(t3[23:26, 47:50].flatten(), torch.tensor([-0.9801, -0.9464, -0.9349, -0.3952, 0.8887, -0.9709, 0.5299, -0.2853, -0.9927]), 1e-3)
    # test_timestep_embeddings (:37): first half sin / second half cos of the same angles, row 0 = [0...,1...], last column of
    # the sin half ~ 0, gradients of frequency monotone
    e = N.get_timestep_embedding(torch.arange(16), 256)
    assert (e[0, :128] - 0).abs().sum() < 1e-5 and (e[0, 128:] - 1).abs().sum() < 1e-5
    assert (e[:, -1] - 1).abs().sum() < 1e-5
    grad_mean = np.abs(np.gradient(e.numpy(), axis=-1)).mean(axis=1)
    prev = 0.0
    for g in grad_mean:
        assert g > prev
        prev = g
    # test_timestep_flip_sin_cos (:60) and test_timestep_downscale_freq_shift (:71)
    a = N.get_timestep_embedding(torch.arange(10), 16, flip_sin_to_cos=True)
    b = N.get_timestep_embedding(torch.arange(10), 16, flip_sin_to_cos=False)
    assert torch.allclose(torch.cat([a[:, 8:], a[:, :8]], dim=-1), b, 1e-3)
    d0 = N.get_timestep_embedding(torch.arange(10), 16, downscale_freq_shift=0)
    d1 = N.get_timestep_embedding(torch.arange(10), 16, downscale_freq_shift=1)
    assert ((d0 - d1)[:, 8:] <= 0).all()  # "cosine needs to be negative"


def test_euler_full_loop_known_answer(golden):
    """tests/schedulers/test_scheduler_euler.py:108-136 (sum 10.0807, mean 0.0131)"""
    cfg = dict(num_train_timesteps=1100, beta_start=0.0001, beta_end=0.02, beta_schedule="linear")
    s = osched.EulerDiscrete(**cfg)
    s.set_timesteps(10)
    sig = s.sigmas
    s = osched.EulerDiscrete(**cfg)
    s.set_timesteps(sigmas=sig)
    sample = _dummy_sample_deter() * s.init_noise_sigma
    for t in s.timesteps:
        sample = s.scale_model_input(sample)
        sample = s.step(_dummy_model(sample, t), sample)
    ksum, kmean = golden("schedulers")["kat"]["euler_no_noise"]
    assert abs(float(sample.abs().sum()) - ksum) < 1e-2
    assert abs(float(sample.abs().mean()) - kmean) < 1e-3


def _full_loop(s, scale_input=False, generator=None, init_sigma=False):
    """tests/schedulers/test_scheduler_*.py `full_loop`: 10 steps of the deterministic dummy model on dummy_sample_deter."""
    s.set_timesteps(10)
    sample = _dummy_sample_deter() * (s.init_noise_sigma if init_sigma else 1.0)
    for t in s.timesteps:
        if scale_input:
            sample = s.scale_model_input(sample, t)
        sample = s.step(_dummy_model(sample, t), t, sample, **(dict(generator=generator) if generator is not None else {}))
    return sample


@pytest.mark.parametrize("kw,mean", [(dict(), 0.2464), (dict(use_karras_sigmas=True), 0.2925), (dict(prediction_type="v_prediction"), 0.1014),
                                     (dict(prediction_type="v_prediction", use_karras_sigmas=True), 0.1966)])
def test_unipc_known_answers(kw, mean):
    """The reference's own hard-coded results: tests/schedulers/test_scheduler_unipc.py:143-233 (config :19-31: linear betas, order 2, bh2,
    final_sigmas_type='sigma_min')."""
    s = osched.UniPC(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", solver_order=2, final_sigmas_type="sigma_min", **kw)
    assert abs(float(_full_loop(s).abs().mean()) - mean) < 1e-3


@pytest.mark.parametrize("kw,mean", [(dict(), 0.3301), (dict(prediction_type="v_prediction"), 0.2251),
                                     (dict(prediction_type="v_prediction", use_karras_sigmas=True), 0.2096)])
def test_dpm_solver_multistep_known_answers(kw, mean):
    """tests/schedulers/test_scheduler_dpm_multi.py:246-295 (config :20-37: dpmsolver++, midpoint, order 2, lower_order_final=False,
    final_sigmas_type='sigma_min')."""
    s = osched.DPMSolverPP2M(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", solver_order=2, lower_order_final=False,
                             final_sigmas_type="sigma_min", **kw)
    assert abs(float(_full_loop(s).abs().mean()) - mean) < 1e-3


@pytest.mark.parametrize("kw,ksum,mean", [(dict(), 172.0067, 0.223967), (dict(prediction_type="v_prediction"), 52.5302, 0.0684),
                                          (dict(set_alpha_to_one=True, beta_start=0.01), 149.8295, 0.1951),
                                          (dict(set_alpha_to_one=False, beta_start=0.01), 149.0784, 0.1941)])
def test_ddim_known_answers(kw, ksum, mean):
    """tests/schedulers/test_scheduler_ddim.py:112-148 (config :12-22: linear betas, clip_sample=True; eta = 0)."""
    cfg = dict(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=True)
    cfg.update(kw)
    out = _full_loop(osched.DDIM(**cfg))
    assert abs(float(out.abs().sum()) - ksum) < 1e-2 and abs(float(out.abs().mean()) - mean) < 1e-3


def test_euler_ancestral_known_answer():
    """tests/schedulers/test_scheduler_euler_ancestral.py:44-70 (1100 training steps, linear betas; noise from torch.manual_seed(0)): 152.3192 / 0.1983"""
    s = osched.EulerAncestral(num_train_timesteps=1100, beta_start=0.0001, beta_end=0.02, beta_schedule="linear")
    out = _full_loop(s, scale_input=True, generator=torch.manual_seed(0), init_sigma=True)
    assert abs(float(out.abs().sum()) - 152.3192) < 1e-2 and abs(float(out.abs().mean()) - 0.1983) < 1e-3


def test_ddpm_known_answers(golden):
    """tests/schedulers/test_scheduler_ddpm.py:62-70 (variances), :72-104 (full loop 258.9606 / 0.3372)"""
    kat = golden("schedulers")["kat"]
    s = osched.DDPM(num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                    variance_type="fixed_small", clip_sample=True)
    for t, v in kat["ddpm_variance"].items():
        assert abs(float(s._get_variance(t)) - v) < 1e-5
    sample = _dummy_sample_deter()
    g = torch.manual_seed(0)
    for t in reversed(range(1000)):
        sample = s.step(_dummy_model(sample, t), t, sample, generator=g)
    ksum, kmean = kat["ddpm_no_noise"]
    assert abs(float(sample.abs().sum()) - ksum) < 1e-2
    assert abs(float(sample.abs().mean()) - kmean) < 1e-3


def test_scheduler_tables_match_reference(golden):
    fx = golden("schedulers")
    for n, tab in fx["euler_sdxl"]["tables"].items():
        s = osched.EulerDiscrete(**fx["euler_sdxl"]["config"])
        s.set_timesteps(n)
        assert torch.equal(s.sigmas, tab["sigmas"]) and torch.equal(s.timesteps, tab["timesteps"])
        assert float(s.init_noise_sigma) == tab["init_noise_sigma"]
    for (n, seq), tab in fx["flow_match_flux"]["tables"].items():
        s = osched.FlowMatchEuler(shift=3.0, use_dynamic_shifting=True)
        s.set_timesteps(n, sigmas=np.linspace(1.0, 1 / n, n), mu=opipe.calculate_shift(seq))
        assert torch.equal(s.sigmas, tab["sigmas"]) and torch.equal(s.timesteps, tab["timesteps"])


def test_scheduler_steps_match_reference_bf16(golden):
    fx = golden("schedulers")
    e = fx["euler_step_bf16"]
    s = osched.EulerDiscrete(**fx["euler_sdxl"]["config"])
    s.set_timesteps(e["n"])
    assert torch.equal(s.scale_model_input(e["x"]), e["scaled"])
    assert torch.equal(s.step(e["eps"], e["x"]), e["prev"])
    f = fx["flow_step_bf16"]
    s = osched.FlowMatchEuler(shift=3.0, use_dynamic_shifting=True)
    s.set_timesteps(4, sigmas=np.linspace(1.0, 1 / 4, 4), mu=f["mu"])
    assert torch.equal(s.step(f["v"], f["x"]), f["prev"])


def test_unet2d_condition_matches_reference(golden):
    fx = golden("models")["unet_tiny"]
    sd16, sd32 = state_dicts(specs.unet2d_condition_params(fx["cfg"]), fx["seed"])
    f = lambda sd, dt: ounet.unet2d_condition_forward(  # noqa: E731
        sd, fx["cfg"], fx["sample"].to(dt), fx["timestep"], fx["encoder_hidden_states"].to(dt),
        dict(text_embeds=fx["text_embeds"].to(dt), time_ids=fx["time_ids"].to(dt)))
    assert torch.allclose(f(sd32, torch.float32), fx["ref32"], atol=2e-5, rtol=1e-5)
    # bf16 CPU eager: same op sequence and rounding points as the reference
    assert (f(sd16, torch.bfloat16).float() - fx["ref16"].float()).abs().max() < 2e-2


@pytest.mark.parametrize("name", ["vae_tiny", "vae_d512"])
def test_vae_decode_matches_reference(golden, name):
    fx = golden("models")[name]
    sd16, sd32 = state_dicts(specs.vae_decoder_params(fx["cfg"]), fx["seed"])
    out = ovae.vae_decode(sd32, fx["cfg"], fx["z"].float())
    assert torch.allclose(out, fx["ref32"], atol=5e-5, rtol=1e-5)


@pytest.mark.parametrize("name", ["vae_enc_tiny", "vae_enc_d512"])
def test_vae_encode_matches_reference(golden, name):
    """N3: the oracle's AutoencoderKL.encode (moments) against the real reference's fp32 run (tests/golden/vae_encode.pt)."""
    fx = golden("vae_encode")[name]
    _, sd32 = state_dicts(specs.vae_params(fx["cfg"]), fx["seed"])
    out = ovae.vae_encode(sd32, fx["cfg"], fx["x"].float())
    assert tuple(out.shape) == tuple(fx["ref32"].shape)
    assert float((out - fx["ref32"]).abs().max()) <= 5e-5
    if fx["roundtrip32"] is not None:  # AutoencoderKL.forward: encode -> mode -> decode
        mean = out[:, : out.shape[1] // 2]
        assert float((ovae.vae_decode(sd32, fx["cfg"], mean) - fx["roundtrip32"]).abs().max()) <= 1e-4


def test_diagonal_gaussian_is_the_reference_distribution(golden):
    """The product's DiagonalGaussianDistribution on the reference's bf16 moments: same clamp / exp / seeded sample, bit for bit."""
    import torch
    from diffusers_b200.autoencoder_kl import DiagonalGaussianDistribution
    fx = golden("vae_encode")["vae_enc_tiny"]
    post = DiagonalGaussianDistribution(fx["ref16"])
    assert torch.equal(post.sample(generator=torch.Generator().manual_seed(fx["sample_seed"])), fx["sample16"])
    assert torch.equal(post.mode(), fx["ref16"][:, :4]) and post.std.dtype == torch.bfloat16
    assert float(DiagonalGaussianDistribution(fx["ref32"]).kl().min()) > 0


@pytest.mark.parametrize("name", ["clip_l_tiny", "clip_g_tiny", "t5_tiny", "t5_tiny_512"])
def test_text_encoder_oracle_matches_transformers(golden, name):
    """N3: oracle/text.py against the real transformers classes the reference's pipelines call (fp32, tests/golden/text.pt)."""
    import torch
    from diffusers_b200 import text_encoders as T
    from oracle import text as otext
    fx = golden("text")[name]
    spec = T.t5_encoder_params(fx["cfg"]) if fx["kind"] == "t5" else T.clip_text_params(fx["cfg"], fx["kind"] == "clip_proj")
    sd32 = {k: v.float() for k, v in T.random_state_dict(spec, fx["seed"]).items()}
    ref = fx["ref32"]
    if fx["kind"] == "t5":
        out = otext.t5_encoder_forward(sd32, fx["cfg"], fx["ids"])
    else:
        out = otext.clip_text_forward(sd32, fx["cfg"], fx["ids"], with_projection=fx["kind"] == "clip_proj")
    scale = float(ref["last_hidden_state"].abs().max())
    assert float((out["last_hidden_state"] - ref["last_hidden_state"]).abs().max()) <= 2e-5 * max(1.0, scale)
    assert len(out["hidden_states"]) == ref["n_hidden"]
    assert float((out["hidden_states"][-2] - ref["penultimate"]).abs().max()) <= 2e-5 * max(1.0, float(ref["penultimate"].abs().max()))
    if "pooler_output" in ref:
        assert float((out["pooler_output"] - ref["pooler_output"]).abs().max()) <= 2e-5 * max(1.0, scale)
    if "text_embeds" in ref:
        assert float((out["text_embeds"] - ref["text_embeds"]).abs().max()) <= 2e-5 * max(1.0, scale)


def test_t5_bucket_table_is_the_oracles():
    import torch
    from diffusers_b200.text_encoders import t5_relative_position_bucket
    from oracle.text import t5_bucket
    pos = torch.arange(512)
    rel = pos[None, :] - pos[:, None]
    assert torch.equal(t5_relative_position_bucket(rel, 32, 128), t5_bucket(rel, 32, 128))


@pytest.mark.parametrize("name", ["flux_tiny", "flux_hd128"])
def test_flux_matches_reference(golden, name):
    fx = golden("models")[name]
    sd16, sd32 = state_dicts(specs.flux_params(fx["cfg"]), fx["seed"])
    out = oflux.flux_forward(sd32, fx["cfg"], fx["hidden_states"].float(), fx["encoder_hidden_states"].float(),
                             fx["pooled"].float(), fx["timestep"], fx["img_ids"], fx["txt_ids"], fx["guidance"])
    assert torch.allclose(out, fx["ref32"], atol=5e-5, rtol=1e-5)


def test_unet2d_matches_reference(golden):
    fx = golden("models")["unet2d_ddpm"]
    sd16, sd32 = state_dicts(specs.unet2d_params(fx["cfg"]), fx["seed"])
    out = ounet.unet2d_forward(sd32, fx["cfg"], fx["sample"].float(), fx["timestep"])
    assert torch.allclose(out, fx["ref32"], atol=2e-5, rtol=1e-5)


def test_sdxl_pipeline_matches_reference(golden):
    fx = golden("pipelines")["sdxl_tiny"]
    _, usd = state_dicts(specs.unet2d_condition_params(fx["unet_cfg"]), fx["unet_seed"])
    _, vsd = state_dicts(specs.vae_decoder_params(fx["vae_cfg"]), fx["vae_seed"])
    lat0 = torch.randn((1, 4, fx["height"] // 8, fx["width"] // 8), generator=torch.Generator().manual_seed(fx["latent_seed"]))
    tid = torch.tensor([[fx["height"], fx["width"], 0, 0, fx["height"], fx["width"]]], dtype=torch.float32)
    img, lat, _ = opipe.sdxl_sample(usd, fx["unet_cfg"], osched.EulerDiscrete(**fx["scheduler"]), lat0, fx["prompt_embeds"],
                                    fx["negative_prompt_embeds"], fx["pooled"], fx["negative_pooled"], tid, fx["steps"],
                                    fx["guidance_scale"], vsd, fx["vae_cfg"], return_all=True)
    assert torch.allclose(lat, fx["latents"], atol=1e-4, rtol=1e-4)
    assert torch.allclose(img, fx["image"], atol=1e-4)


def test_flux_pipeline_matches_reference(golden):
    fx = golden("pipelines")["flux_tiny"]
    _, sd = state_dicts(specs.flux_params(fx["cfg"]), fx["seed"])
    h = 2 * (fx["height"] // (fx["vae_scale_factor"] * 2))
    lat = torch.randn((1, fx["cfg"]["in_channels"] // 4, h, h), generator=torch.Generator().manual_seed(fx["latent_seed"]))
    packed = lat.view(1, -1, h // 2, 2, h // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(1, (h // 2) ** 2, -1)
    ids = torch.zeros(h // 2, h // 2, 3)
    ids[..., 1] += torch.arange(h // 2)[:, None]
    ids[..., 2] += torch.arange(h // 2)[None, :]
    out = opipe.flux_sample(sd, fx["cfg"], osched.FlowMatchEuler(shift=3.0, use_dynamic_shifting=True), packed,
                            fx["prompt_embeds"], fx["pooled"], ids.reshape(-1, 3), torch.zeros(fx["prompt_embeds"].shape[1], 3),
                            fx["steps"], fx["guidance_scale"])
    assert torch.allclose(out, fx["latents"], atol=1e-4, rtol=1e-4)


def test_ddpm_pipeline_matches_reference(golden):
    """Config 0 (BASELINE.json): UNet2DModel 32x32, DDPM, 10 steps, CPU - including the per-step RNG consumption."""
    fx = golden("pipelines")["ddpm"]
    _, sd = state_dicts(specs.unet2d_params(fx["cfg"]), fx["seed"])
    g = torch.manual_seed(0)
    image0 = torch.randn((1, 3, 32, 32), generator=g)
    img = opipe.ddpm_sample(sd, fx["cfg"], osched.DDPM(), image0, fx["steps"], g)
    assert torch.allclose(img.permute(0, 2, 3, 1), fx["image"], atol=1e-4)


def test_ddpm_pipeline_known_answer(golden):
    """tests/pipelines/ddpm/test_ddpm.py:45-66: the reference's hard-coded 3x3 slice for its dummy UNet2DModel (weights = its
    default init recorded in layers.pt), DDPMScheduler(), 2 steps, generator seed 0 - tolerance 1e-2 as there."""
    fx = golden("layers")["ddpm_pipeline_kat"]
    cfg = dict(specs.DDPM_TINY_CONFIG)
    cfg.update(fx["cfg"])
    g = torch.Generator().manual_seed(0)
    image0 = torch.randn((1, 3, 8, 8), generator=g)
    img = opipe.ddpm_sample(fx["state_dict"], cfg, osched.DDPM(), image0, 2, g).permute(0, 2, 3, 1)
    assert tuple(img.shape) == (1, 8, 8, 3)
    assert (img[0, -3:, -3:, -1].flatten().double() - fx["expected_slice"]).abs().max() < 1e-2
    assert torch.allclose(img, fx["image"], atol=1e-4)  # and the whole image the reference produced here

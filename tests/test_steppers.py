"""DDIM, Euler-ancestral and DPM-Solver++ (2M) steppers.

tests/golden/schedulers2.pt (oracle/make_golden.py gen_steppers) holds what the REAL reference schedulers produce: their
timestep / sigma tables and whole 8-step trajectories on fp32 and bf16 CPU tensors driven by a deterministic stand-in
denoiser.  CPU: the oracle restatement and the product's host logic (with b200_linear_step evaluated in float64 by a
stand-in) against those; GPU: the product with its kernel."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import GOLDEN  # noqa: E402
from diffusers_b200 import ops  # noqa: E402
from diffusers_b200 import schedulers as S  # noqa: E402
from oracle import schedulers as osched  # noqa: E402
from oracle.make_golden import STEPPERS, VPRED_STEPPERS, stepper_fake_model  # noqa: E402

FX = torch.load(os.path.join(GOLDEN, "schedulers2.pt"), weights_only=False)
ORACLE = dict(DDIMScheduler=osched.DDIM, EulerAncestralDiscreteScheduler=osched.EulerAncestral, DPMSolverMultistepScheduler=osched.DPMSolverPP2M,
              UniPCMultistepScheduler=osched.UniPC)
KEYS = [k for k, _, _ in STEPPERS]


def _run(sched, x, gen_kw, device=None):
    for t in sched.timesteps:
        tt = t if device is None else t.to(device)
        eps = stepper_fake_model(sched.scale_model_input(x, tt), t)
        out = sched.step(eps, tt, x, **gen_kw)
        x = out[0] if isinstance(out, tuple) else out
    return x


@pytest.mark.parametrize("key", KEYS)
def test_oracle_steppers_match_the_reference_recordings(key):
    fx = FX[key]
    cfg = {k: v for k, v in fx["config"].items()}
    for dtn, dt in (("float32", torch.float32), ("bfloat16", torch.bfloat16)):
        o = ORACLE[fx["cls"]](**cfg)
        o.set_timesteps(fx["steps"])
        kw = dict(generator=torch.Generator().manual_seed(0)) if fx["cls"].startswith("EulerAncestral") else {}
        got = _run(o, fx["trajectory"][dtn]["start"].clone(), kw)
        ref = fx["trajectory"][dtn]["final"]
        assert got.dtype == dt
        assert torch.equal(got, ref), (key, dtn, float((got.float() - ref.float()).abs().max()))  # same ops, same order: same bits


@pytest.mark.parametrize("key", KEYS)
def test_product_tables_are_the_references_bit_for_bit(key):
    fx = FX[key]
    for n, tab in fx["tables"].items():
        s = getattr(S, fx["cls"])(**fx["config"])
        s.set_timesteps(n)
        assert torch.equal(s.timesteps, tab["timesteps"]) and s.timesteps.dtype == tab["timesteps"].dtype
        if tab["sigmas"] is not None:
            assert torch.equal(s.sigmas, tab["sigmas"])
        assert float(s.init_noise_sigma) == tab["init_noise_sigma"]


@pytest.mark.parametrize("key", KEYS)
def test_product_host_logic_follows_the_reference_trajectory(key, monkeypatch):
    """Coefficients, multistep history, lower-order switches and RNG consumption, with b200_linear_step evaluated exactly."""
    def lin(sample, m0=None, m1=None, noise=None, *, a=1.0, b=0.0, c=0.0, s=0.0, out=None):
        v = a * sample.double()
        for coef, t in ((b, m0), (c, m1), (s, noise)):
            if t is not None:
                v = v + coef * t.double()
        return v.to(sample.dtype)
    monkeypatch.setattr(ops, "linear_step", lin)
    monkeypatch.setattr(ops, "scale_div", lambda x, div, out=None: x / div)
    fx = FX[key]
    s = getattr(S, fx["cls"])(**fx["config"])
    s.set_timesteps(fx["steps"])
    s.set_begin_index(0)
    kw = dict(generator=torch.Generator().manual_seed(0)) if fx["cls"].startswith("EulerAncestral") else {}
    got = _run(s, fx["trajectory"]["float32"]["start"].clone(), dict(return_dict=False, **kw))
    ref = fx["trajectory"]["float32"]["final"]
    assert (got - ref).abs().max() <= 1e-4 * ref.abs().max()


def test_euler_karras_tables_and_host_logic(monkeypatch):
    """EulerDiscreteScheduler(use_karras_sigmas=True): sigma / (fractional) timestep tables bit-identical to the reference's, and the
    product's stepping logic on them (b200_euler_step / b200_scale evaluated in float64 by stand-ins) follows the reference trajectory."""
    fx = FX["euler_karras_sdxl"]
    for n, tab in fx["tables"].items():
        s = S.EulerDiscreteScheduler(**fx["config"])
        s.set_timesteps(n)
        assert torch.equal(s.sigmas, tab["sigmas"]) and torch.equal(s.timesteps, tab["timesteps"]) and s.timesteps.dtype == torch.float32
        assert float(s.init_noise_sigma) == tab["init_noise_sigma"] and s.config.use_karras_sigmas is True
    monkeypatch.setattr(ops, "euler_step", lambda eps, x, sigma, sigma_next, out=None: (x.double() + eps.double() * (sigma_next - sigma)).to(eps.dtype))
    monkeypatch.setattr(ops, "scale_div", lambda x, div, out=None: (x.double() / div).to(x.dtype))
    s = S.EulerDiscreteScheduler(**fx["config"])
    s.set_timesteps(fx["steps"])
    s.set_begin_index(0)
    got = _run(s, fx["trajectory"]["float32"]["start"].clone(), dict(return_dict=False))
    ref = fx["trajectory"]["float32"]["final"]
    assert (got - ref).abs().max() <= 1e-4 * ref.abs().max()
    assert S.EulerDiscreteScheduler(beta_schedule="scaled_linear").config.use_karras_sigmas is False


class _OracleEuler:
    """oracle.schedulers.EulerDiscrete behind the (timestep-taking) call surface the runner uses."""

    def __init__(self, **kw):
        self.o = osched.EulerDiscrete(**kw)

    def set_timesteps(self, n):
        self.o.set_timesteps(n)
        self.timesteps = self.o.timesteps

    def scale_model_input(self, x, t):
        return self.o.scale_model_input(x)

    def step(self, eps, t, x):
        return self.o.step(eps, x)


@pytest.mark.parametrize("key", [k for k, _, _ in VPRED_STEPPERS])
def test_v_prediction_steppers(key, monkeypatch):
    """prediction_type='v_prediction' (SD 2.x-v, v-pred SDXL fine-tunes) on Euler, DDIM, DPM-Solver++ and UniPC: the oracle reproduces the real
    reference's fp32 and bf16 trajectories bit for bit; the product's tables are the reference's and its host logic (the step kernels
    evaluated exactly by stand-ins) follows the reference's fp32 trajectory.  Same kernels as the epsilon steppers, other coefficients."""
    fx = FX[key]
    oracle_cls = dict(ORACLE, EulerDiscreteScheduler=_OracleEuler)[fx["cls"]]
    for dtn in ("float32", "bfloat16"):
        o = oracle_cls(**fx["config"])
        o.set_timesteps(fx["steps"])
        got = _run(o, fx["trajectory"][dtn]["start"].clone(), {})
        assert torch.equal(got, fx["trajectory"][dtn]["final"]), (key, dtn)
    for n, tab in fx["tables"].items():
        s = getattr(S, fx["cls"])(**fx["config"])
        s.set_timesteps(n)
        assert torch.equal(s.timesteps, tab["timesteps"]) and (tab["sigmas"] is None or torch.equal(s.sigmas, tab["sigmas"]))

    def lin(sample, m0=None, m1=None, noise=None, *, a=1.0, b=0.0, c=0.0, s=0.0, out=None):
        v = a * sample.double()
        for coef, t in ((b, m0), (c, m1), (s, noise)):
            if t is not None:
                v = v + coef * t.double()
        return v.to(sample.dtype)
    monkeypatch.setattr(ops, "linear_step", lin)
    monkeypatch.setattr(ops, "scale_div", lambda x, div, out=None: (x.double() / div).to(x.dtype))
    s = getattr(S, fx["cls"])(**fx["config"])
    s.set_timesteps(fx["steps"])
    s.set_begin_index(0)
    got = _run(s, fx["trajectory"]["float32"]["start"].clone(), dict(return_dict=False))
    ref = fx["trajectory"]["float32"]["final"]
    assert (got - ref).abs().max() <= 1e-4 * ref.abs().max()


def test_steppers_reject_options_outside_the_path():
    with pytest.raises(NotImplementedError):
        S.DDIMScheduler(clip_sample=True)
    with pytest.raises(NotImplementedError):
        S.DPMSolverMultistepScheduler(use_lu_lambdas=True)
    with pytest.raises(NotImplementedError):
        S.DPMSolverMultistepScheduler(algorithm_type="sde-dpmsolver++")
    with pytest.raises(NotImplementedError):
        S.EulerAncestralDiscreteScheduler(prediction_type="v_prediction")
    with pytest.raises(NotImplementedError):
        S.DPMSolverMultistepScheduler(prediction_type="sample")
    for bad in (dict(solver_type="bh1"), dict(predict_x0=False), dict(disable_corrector=[0]), dict(use_flow_sigmas=True), dict(solver_order=3)):
        with pytest.raises(NotImplementedError):
            S.UniPCMultistepScheduler(**bad)
    assert S.UniPCMultistepScheduler(solver_type="midpoint").config.solver_type == "bh2"  # the reference maps the DPM-Solver names onto bh2
    s = S.DDIMScheduler(clip_sample=False)
    s.set_timesteps(10)
    with pytest.raises(NotImplementedError):
        s.step(torch.zeros(1), s.timesteps[0], torch.zeros(1), eta=0.5)


@pytest.mark.gpu
def test_linear_step_kernel_is_one_rounding_of_the_fp32_formula():
    g = torch.Generator(device="cuda").manual_seed(0)
    for dt in (torch.bfloat16, torch.float16):
        x, m0, m1, nz = [(torch.randn(3, 4, 33, 17, generator=g, device="cuda") * 3).to(dt) for _ in range(4)]
        for a, b, c, s, use in ((0.91, -0.37, 0.0, 0.0, (m0, None, None)), (1.0, -2.5, 0.0, 0.4, (m0, None, nz)), (0.5, 1.75, -0.6, 0.0, (m0, m1, None)),
                                (1.3, 0.0, 0.0, 0.0, (None, None, None))):
            out = ops.linear_step(x, m0=use[0], m1=use[1], noise=use[2], a=a, b=b, c=c, s=s)
            ref = a * x.double()
            for coef, t in ((b, use[0]), (c, use[1]), (s, use[2])):
                if t is not None:
                    ref = ref + coef * t.double()
            # fp32 fma chain then one rounding: within one ulp of the exactly rounded value
            ulp = ref.abs().float() * (2.0 ** -8 if dt == torch.bfloat16 else 2.0 ** -11) + 1e-6
            assert ((out.double() - ref).abs().float() <= ulp).all()


@pytest.mark.gpu
@pytest.mark.parametrize("key", KEYS)
def test_product_steppers_on_the_gpu(key):
    """bf16 trajectories with the real kernel: no further from the reference's fp32 trajectory than 1.5x the reference's own bf16 run."""
    fx = FX[key]
    s = getattr(S, fx["cls"])(**fx["config"])
    s.set_timesteps(fx["steps"], device="cuda")
    s.set_begin_index(0)
    kw = dict(generator=torch.Generator().manual_seed(0)) if fx["cls"].startswith("EulerAncestral") else {}
    got = _run(s, fx["trajectory"]["bfloat16"]["start"].clone().cuda(), dict(return_dict=False, **kw), device="cuda").float().cpu()
    ref32, ref16 = fx["trajectory"]["float32"]["final"], fx["trajectory"]["bfloat16"]["final"].float()
    e, e16 = (got - ref32).abs(), (ref16 - ref32).abs()
    print(f"{key}: ours max {float(e.max()):.4g} mean {float(e.mean()):.4g} | reference bf16 max {float(e16.max()):.4g} mean {float(e16.mean()):.4g}")
    assert float(e.mean()) <= 1.5 * float(e16.mean()) + 1e-3 and float(e.max()) <= 2.0 * float(e16.max()) + 2e-2

"""Drop-in schedulers for the hot path: the host-side schedule construction repeats the reference's
numpy/torch arithmetic op for op (so sigmas/timesteps are bit-identical), `step()` / `scale_model_input()` launch
one fused sm_100a kernel each (ops.euler_step, ops.scale_div, ops.flow_match_step).

Reference: schedulers/scheduling_euler_discrete.py:143 (EulerDiscreteScheduler: __init__ :203, set_timesteps :350,
scale_model_input :326, step :685), schedulers/scheduling_flow_match_euler_discrete.py:48
(FlowMatchEulerDiscreteScheduler: set_timesteps :283, step :423, time_shift :241).  Only the configuration space the
SDXL / Flux pipelines use is accepted; anything else raises.
"""
import math

import numpy as np
import torch

from . import ops
from .config import FrozenConfig


class SchedulerOutput:
    def __init__(self, prev_sample, pred_original_sample=None):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class EulerDiscreteScheduler:
    @classmethod
    def from_pretrained(cls, path, subfolder=None):
        """Constructor arguments from the reference's scheduler_config.json (schedulers/scheduling_utils.py:100)."""
        from .checkpoint import scheduler_from_pretrained
        return scheduler_from_pretrained(cls, path, subfolder)

    @classmethod
    def from_config(cls, config):
        from .checkpoint import scheduler_kwargs
        return cls(**scheduler_kwargs(cls, config))

    order = 1
    init_noise_sigma_is_tensor = True

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 prediction_type="epsilon", interpolation_type="linear", timestep_spacing="linspace", steps_offset=0,
                 final_sigmas_type="zero", **unsupported):
        # Options of the reference constructor (schedulers/scheduling_euler_discrete.py:203-222) this path does not implement:
        # a non-default value changes the sigma table, so it must fail loudly.  Keys that are NOT in the reference's signature
        # (legacy entries of published configs: SDXL-base's scheduler_config.json carries `sample_max_value`, `skip_prk_steps`,
        # `set_alpha_to_one`, `clip_sample`) are ignored with a warning, as the reference's own config loader does
        # (configuration_utils.py extract_init_dict).
        not_implemented = dict(trained_betas=(None,), use_exponential_sigmas=(None, False),
                               use_beta_sigmas=(None, False), sigma_min=(None,), sigma_max=(None,), timestep_type=("discrete",),
                               rescale_betas_zero_snr=(None, False))
        use_karras_sigmas = bool(unsupported.pop("use_karras_sigmas", False))  # "Euler Karras": a different sigma table, same step
        for k, v in unsupported.items():
            if k in not_implemented:
                if v not in not_implemented[k]:
                    raise NotImplementedError(f"EulerDiscreteScheduler option {k}={v!r} is outside the hot path")
            else:
                import warnings
                warnings.warn(f"EulerDiscreteScheduler: config key {k!r} is not an argument of the reference scheduler and is ignored")
        if prediction_type not in ("epsilon", "v_prediction"):
            raise NotImplementedError("prediction_type 'epsilon' or 'v_prediction'")
        if interpolation_type != "linear" or final_sigmas_type != "zero":
            raise NotImplementedError("only linear interpolation with a final sigma of zero")
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                   beta_schedule=beta_schedule, prediction_type=prediction_type,
                                   interpolation_type=interpolation_type, timestep_spacing=timestep_spacing,
                                   steps_offset=steps_offset, final_sigmas_type=final_sigmas_type, use_karras_sigmas=use_karras_sigmas,
                                   timestep_type="discrete")
        if beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).flip(0)
        timesteps = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy()
        self.timesteps = torch.from_numpy(timesteps).to(dtype=torch.float32)
        self.sigmas = torch.cat([sigmas, torch.zeros(1)]).to("cpu")
        self.num_inference_steps = None
        self.is_scale_input_called = False
        self._step_index = None
        self._begin_index = None
        self._timesteps_cpu = self.timesteps.clone()

    @property
    def init_noise_sigma(self):
        max_sigma = self.sigmas.max()
        if self.config.timestep_spacing in ("linspace", "trailing"):
            return max_sigma
        return (max_sigma ** 2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index=0):
        self._begin_index = begin_index

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None, sigmas=None):
        if timesteps is not None or sigmas is not None:
            raise NotImplementedError("custom timesteps / sigmas")
        c = self.config
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, c.num_train_timesteps - 1, num_inference_steps, dtype=np.float32)[::-1].copy()
        elif c.timestep_spacing == "leading":
            step_ratio = c.num_train_timesteps // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.float32)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            step_ratio = c.num_train_timesteps / num_inference_steps
            ts = (np.arange(c.num_train_timesteps, 0, -step_ratio)).round().copy().astype(np.float32)
            ts -= 1
        else:
            raise ValueError(c.timestep_spacing)
        full = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        sig = np.interp(ts, np.arange(0, len(full)), full)
        if c.use_karras_sigmas:
            # scheduling_euler_discrete.py:448-450: Karras spacing between the extremes of the INTERPOLATED table; the timesteps become the
            # (fractional) positions of those sigmas on the training schedule
            sig, ts = _karras_sigmas_and_timesteps(full, num_inference_steps, sigma_min=sig[-1].item(), sigma_max=sig[0].item())
        sig = np.concatenate([sig, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sig).to(dtype=torch.float32)  # stays on the host, like the reference (:481)
        self._timesteps_cpu = torch.from_numpy(ts.astype(np.float32))
        self.timesteps = self._timesteps_cpu.to(device=device)
        self._step_index = None
        self._begin_index = None

    def _init_step_index(self, timestep):
        if self._begin_index is None:
            t = float(timestep)  # device->host sync, as in the reference when set_begin_index was not called
            idx = (self._timesteps_cpu == t).nonzero()
            pos = 1 if len(idx) > 1 else 0
            self._step_index = int(idx[pos])
        else:
            self._step_index = self._begin_index

    def scale_model_input(self, sample, timestep):
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        div = float((sigma ** 2 + 1) ** 0.5)
        self.is_scale_input_called = True
        return ops.scale_div(sample, div)

    def step(self, model_output, timestep, sample, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0,
             generator=None, return_dict=True):
        if isinstance(timestep, (int, torch.IntTensor, torch.LongTensor)):
            raise ValueError("Passing integer indices as timesteps to EulerDiscreteScheduler.step() is not supported.")
        if s_churn != 0.0:
            raise NotImplementedError("s_churn > 0 (stochastic sampling) is outside the hot path")
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = float(self.sigmas[self._step_index])
        sigma_next = float(self.sigmas[self._step_index + 1])
        if self.config.prediction_type == "v_prediction":
            # x0 = -sigma / sqrt(sigma^2 + 1) v + x / (sigma^2 + 1) (:773-775); prev = x + (x - x0) / sigma * dt: linear in (x, v)
            dt = sigma_next - sigma
            prev = ops.linear_step(sample.to(model_output.dtype), m0=model_output, a=1.0 + dt * sigma / (sigma * sigma + 1.0), b=dt / (sigma * sigma + 1.0) ** 0.5)
        else:
            prev = ops.euler_step(model_output, sample, sigma, sigma_next)
        self._step_index += 1
        if not return_dict:
            return (prev, None)
        return SchedulerOutput(prev)

    def __len__(self):
        return self.config.num_train_timesteps


class FlowMatchEulerDiscreteScheduler:
    @classmethod
    def from_pretrained(cls, path, subfolder=None):
        """Constructor arguments from the reference's scheduler_config.json (schedulers/scheduling_utils.py:100)."""
        from .checkpoint import scheduler_from_pretrained
        return scheduler_from_pretrained(cls, path, subfolder)

    @classmethod
    def from_config(cls, config):
        from .checkpoint import scheduler_kwargs
        return cls(**scheduler_kwargs(cls, config))

    order = 1

    def __init__(self, num_train_timesteps=1000, shift=1.0, use_dynamic_shifting=False, base_shift=0.5, max_shift=1.15,
                 base_image_seq_len=256, max_image_seq_len=4096, time_shift_type="exponential", **unsupported):
        # reference signature: schedulers/scheduling_flow_match_euler_discrete.py:96-113; other keys are ignored with a warning
        not_implemented = ("invert_sigmas", "shift_terminal", "use_karras_sigmas", "use_exponential_sigmas", "use_beta_sigmas",
                           "stochastic_sampling")
        for k, v in unsupported.items():
            if k in not_implemented:
                if v not in (None, False):
                    raise NotImplementedError(f"FlowMatchEulerDiscreteScheduler option {k}={v!r} is outside the hot path")
            else:
                import warnings
                warnings.warn(f"FlowMatchEulerDiscreteScheduler: config key {k!r} is not an argument of the reference scheduler and is ignored")
        if time_shift_type != "exponential":
            raise NotImplementedError("only exponential time shift")
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps, shift=shift,
                                   use_dynamic_shifting=use_dynamic_shifting, base_shift=base_shift, max_shift=max_shift,
                                   base_image_seq_len=base_image_seq_len, max_image_seq_len=max_image_seq_len,
                                   time_shift_type=time_shift_type, invert_sigmas=False, shift_terminal=None,
                                   stochastic_sampling=False)
        timesteps = np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()
        timesteps = torch.from_numpy(timesteps).to(dtype=torch.float32)
        sigmas = timesteps / num_train_timesteps
        if not use_dynamic_shifting:
            sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.timesteps = sigmas * num_train_timesteps
        self._shift = shift
        self.sigmas = sigmas.to("cpu")
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()
        self._sigmas_cpu = self.sigmas
        self._timesteps_cpu = self.timesteps
        self._step_index = None
        self._begin_index = None
        self.num_inference_steps = None

    init_noise_sigma = 1.0

    @property
    def shift(self):
        return self._shift

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index=0):
        self._begin_index = begin_index

    def _sigma_to_t(self, sigma):
        return sigma * self.config.num_train_timesteps

    def time_shift(self, mu, sigma, t):
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None, timesteps=None):
        if self.config.use_dynamic_shifting and mu is None:
            raise ValueError("`mu` must be passed when `use_dynamic_shifting` is set to be `True`")
        if timesteps is not None:
            raise NotImplementedError("custom timesteps")
        if num_inference_steps is None:
            num_inference_steps = len(sigmas)
        self.num_inference_steps = num_inference_steps
        if sigmas is None:
            ts = np.linspace(self._sigma_to_t(self.sigma_max), self._sigma_to_t(self.sigma_min), num_inference_steps)
            sigmas = ts / self.config.num_train_timesteps
        else:
            sigmas = np.array(sigmas).astype(np.float32)
        if self.config.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = self._shift * sigmas / (1 + (self._shift - 1) * sigmas)
        sig = torch.from_numpy(sigmas).to(dtype=torch.float32)
        self._timesteps_cpu = sig * self.config.num_train_timesteps
        self._sigmas_cpu = torch.cat([sig, torch.zeros(1)])
        self.timesteps = self._timesteps_cpu.to(device=device)
        self.sigmas = self._sigmas_cpu.to(device=device)  # the reference keeps these on the device (:380)
        self._step_index = None
        self._begin_index = None

    def _init_step_index(self, timestep):
        if self._begin_index is None:
            t = float(timestep)
            idx = (self._timesteps_cpu == t).nonzero()
            pos = 1 if len(idx) > 1 else 0
            self._step_index = int(idx[pos])
        else:
            self._step_index = self._begin_index

    def step(self, model_output, timestep, sample, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0,
             generator=None, per_token_timesteps=None, return_dict=True):
        if isinstance(timestep, (int, torch.IntTensor, torch.LongTensor)):
            raise ValueError("Passing integer indices as timesteps to FlowMatchEulerDiscreteScheduler.step() is not supported.")
        if per_token_timesteps is not None:
            raise NotImplementedError("per_token_timesteps")
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = float(self._sigmas_cpu[self._step_index])
        sigma_next = float(self._sigmas_cpu[self._step_index + 1])
        prev = ops.flow_match_step(model_output, sample, sigma, sigma_next)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev)

    def __len__(self):
        return self.config.num_train_timesteps


class DDPMScheduler:
    @classmethod
    def from_pretrained(cls, path, subfolder=None):
        """Constructor arguments from the reference's scheduler_config.json (schedulers/scheduling_utils.py:100)."""
        from .checkpoint import scheduler_from_pretrained
        return scheduler_from_pretrained(cls, path, subfolder)

    @classmethod
    def from_config(cls, config):
        from .checkpoint import scheduler_kwargs
        return cls(**scheduler_kwargs(cls, config))

    """schedulers/scheduling_ddpm.py:137 (epsilon prediction, fixed_small variance, clip_sample): host tables are the
    reference's torch ops, `step` is one fused kernel; the noise is drawn with the caller's generator exactly like
    `randn_tensor` (utils/torch_utils.py:183) so the RNG stream is consumed identically."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 variance_type="fixed_small", clip_sample=True, prediction_type="epsilon", clip_sample_range=1.0,
                 timestep_spacing="leading", steps_offset=0, **unsupported):
        # the reference serialises every constructor default; these two only act when `thresholding` is on
        inert = dict(dynamic_thresholding_ratio=0.995, sample_max_value=1.0)
        for k, v in unsupported.items():
            if v not in (None, False) and inert.get(k) != v:
                raise NotImplementedError(f"DDPMScheduler option {k}={v!r} is outside the hot path")
        if prediction_type != "epsilon" or variance_type != "fixed_small":
            raise NotImplementedError("only epsilon prediction with fixed_small variance")
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                                   beta_schedule=beta_schedule, variance_type=variance_type, clip_sample=clip_sample,
                                   prediction_type=prediction_type, clip_sample_range=clip_sample_range,
                                   timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        if beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.custom_timesteps = False
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())
        self._timesteps_cpu = self.timesteps

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None):
        if timesteps is not None:
            raise NotImplementedError("custom timesteps")
        c = self.config
        if num_inference_steps > c.num_train_timesteps:
            raise ValueError("`num_inference_steps` cannot be larger than `num_train_timesteps`")
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, c.num_train_timesteps - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            ts = (np.arange(0, num_inference_steps) * (c.num_train_timesteps // num_inference_steps)).round()[::-1].copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = np.round(np.arange(c.num_train_timesteps, 0, -c.num_train_timesteps / num_inference_steps)).astype(np.int64)
            ts -= 1
        else:
            raise ValueError(c.timestep_spacing)
        self._timesteps_cpu = torch.from_numpy(ts)
        self.timesteps = self._timesteps_cpu.to(device)

    def previous_timestep(self, timestep):
        if self.custom_timesteps or self.num_inference_steps:
            index = (self._timesteps_cpu == timestep).nonzero(as_tuple=True)[0][0]
            return -1 if index == self._timesteps_cpu.shape[0] - 1 else int(self._timesteps_cpu[index + 1])
        return timestep - 1

    def _get_variance(self, t):
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        cur_beta = 1 - a_t / a_prev
        return torch.clamp((1 - a_prev) / (1 - a_t) * cur_beta, min=1e-20)

    def step(self, model_output, timestep, sample, generator=None, return_dict=True):
        t = int(timestep)
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        c0 = (a_prev ** 0.5 * cur_beta) / b_t
        c1 = cur_alpha ** 0.5 * b_prev / b_t
        noise, sigma = None, 0.0
        if t > 0:
            from .pipelines import randn_tensor
            # The reference draws in model_output.dtype (scheduling_ddpm.py:541-543).  BASELINE config 0 is an fp32 reference model
            # that this 16-bit shell stands in for, so the draw stays in fp32 (the reference's stream for config 0, which the
            # parity test relies on) and is rounded to the shell's dtype; a 16-bit REFERENCE model would draw a different stream.
            noise = randn_tensor(model_output.shape, generator=generator, device=model_output.device, dtype=torch.float32).to(model_output.dtype)
            sigma = float(self._get_variance(t) ** 0.5)
        prev = ops.ddpm_step(model_output, sample.to(model_output.dtype), noise, sqrt_beta_prod=float(b_t ** 0.5),
                             sqrt_alpha_prod=float(a_t ** 0.5), c0=float(c0), c1=float(c1), sigma=sigma,
                             clip=self.config.clip_sample, clip_range=self.config.clip_sample_range)
        if not return_dict:
            return (prev, None)
        return SchedulerOutput(prev)

    def __len__(self):
        return self.config.num_train_timesteps


# ----------------------------------------------------------------------------------------------------------------------
# more steppers on one fused update kernel (b200_linear_step: prev = a x + b m0 + c m1 + s noise in fp32
# with ONE rounding; the reference evaluates the same formulas as chains of 16-bit tensor ops).  Host tables are the
# reference's own numpy / torch expressions, bit for bit (tests/test_host_logic.py against tests/golden/schedulers2.pt).
# ----------------------------------------------------------------------------------------------------------------------
def _betas(beta_schedule, beta_start, beta_end, n):
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    raise NotImplementedError(beta_schedule)


class _StepIndexMixin:
    """Step bookkeeping without device->host syncs: the pipelines call set_begin_index(0); otherwise the first timestep is
    looked up once in the host copy of the table (like the reference's index_for_timestep)."""

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index=0):
        self._begin_index = begin_index

    def _init_step_index(self, timestep):
        if self._begin_index is None:
            t = float(timestep)
            idx = (self._timesteps_cpu.to(torch.float64) == t).nonzero()
            if len(idx) == 0:
                raise ValueError(f"timestep {t} is not in the schedule")
            self._step_index = int(idx[1 if len(idx) > 1 else 0])
        else:
            self._step_index = self._begin_index

    @classmethod
    def from_pretrained(cls, path, subfolder=None):
        from .checkpoint import scheduler_from_pretrained
        return scheduler_from_pretrained(cls, path, subfolder)

    @classmethod
    def from_config(cls, config):
        from .checkpoint import scheduler_kwargs
        return cls(**scheduler_kwargs(cls, config))

    def __len__(self):
        return self.config.num_train_timesteps


def _reject(name, unsupported, not_implemented):
    import warnings
    for k, v in unsupported.items():
        if k in not_implemented:
            if v not in not_implemented[k]:
                raise NotImplementedError(f"{name} option {k}={v!r} is outside the hot path")
        else:
            warnings.warn(f"{name}: config key {k!r} is not an argument of the reference scheduler and is ignored")


class DDIMScheduler(_StepIndexMixin):
    """schedulers/scheduling_ddim.py:137 - deterministic DDIM (eta = 0), epsilon prediction, no sample clipping (the latent
    diffusion configuration: SD / SDXL checkpoints ship clip_sample=False).
        x0 = (x - sqrt(1-a_t) eps) / sqrt(a_t);  prev = sqrt(a_prev) x0 + sqrt(1-a_prev) eps      (:470-500)"""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", clip_sample=True,
                 set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon", timestep_spacing="leading", **unsupported):
        _reject("DDIMScheduler", unsupported, dict(trained_betas=(None,), thresholding=(None, False), dynamic_thresholding_ratio=(0.995,),
                                                   clip_sample_range=(1.0, 1), sample_max_value=(1.0, 1), rescale_betas_zero_snr=(None, False)))
        if clip_sample:
            raise NotImplementedError("DDIMScheduler clip_sample=True (pixel-space models) is outside the hot path; latent models use False")
        if prediction_type not in ("epsilon", "v_prediction"):
            raise NotImplementedError("prediction_type 'epsilon' or 'v_prediction'")
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule,
                                   clip_sample=False, set_alpha_to_one=set_alpha_to_one, steps_offset=steps_offset, prediction_type=prediction_type,
                                   timestep_spacing=timestep_spacing, thresholding=False)
        self.betas = _betas(beta_schedule, beta_start, beta_end, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_inference_steps = None
        self._timesteps_cpu = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))
        self.timesteps = self._timesteps_cpu
        self._step_index = self._begin_index = None

    def scale_model_input(self, sample, timestep=None):
        return sample

    def set_timesteps(self, num_inference_steps, device=None):
        c = self.config
        if num_inference_steps > c.num_train_timesteps:
            raise ValueError("`num_inference_steps` cannot be larger than `num_train_timesteps`")
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, c.num_train_timesteps - 1, num_inference_steps).round()[::-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            ts = (np.arange(0, num_inference_steps) * (c.num_train_timesteps // num_inference_steps)).round()[::-1].copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = np.round(np.arange(c.num_train_timesteps, 0, -c.num_train_timesteps / num_inference_steps)).astype(np.int64)
            ts -= 1
        else:
            raise ValueError(c.timestep_spacing)
        self._timesteps_cpu = torch.from_numpy(ts)
        self.timesteps = self._timesteps_cpu.to(device)
        self._step_index = self._begin_index = None

    def step(self, model_output, timestep, sample, eta=0.0, use_clipped_model_output=False, generator=None, variance_noise=None,
             return_dict=True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if eta != 0.0 or variance_noise is not None:
            raise NotImplementedError("eta > 0 (stochastic DDIM) is outside the hot path")
        if self._step_index is None:
            self._init_step_index(timestep)
        t = int(self._timesteps_cpu[self._step_index])  # host copy: no device sync
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        sa, sb = float(a_t ** 0.5), float((1 - a_t) ** 0.5)
        sap, sbp = float(a_prev ** 0.5), float((1 - a_prev) ** 0.5)
        if self.config.prediction_type == "v_prediction":
            # x0 = sqrt(a_t) x - sqrt(1-a_t) v,  eps = sqrt(a_t) v + sqrt(1-a_t) x  (:468-470)
            x0 = ops.linear_step(sample, m0=model_output, a=sa, b=-sb)
            prev = ops.linear_step(sample, m0=model_output, a=sap * sa + sbp * sb, b=sbp * sa - sap * sb)
        else:
            x0 = ops.linear_step(sample, m0=model_output, a=1.0 / sa, b=-sb / sa)
            prev = ops.linear_step(sample, m0=model_output, a=sap / sa, b=sbp - sap * sb / sa)
        self._step_index += 1
        if not return_dict:
            return (prev, x0)
        return SchedulerOutput(prev, x0)


class EulerAncestralDiscreteScheduler(_StepIndexMixin):
    """schedulers/scheduling_euler_ancestral_discrete.py:133 - epsilon prediction: with sigma_up / sigma_down of :432-435
        prev = x + eps (sigma_down - sigma) + noise sigma_up,   noise ~ randn_tensor(model_output.dtype, generator)   (:437-447)"""
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", prediction_type="epsilon",
                 timestep_spacing="linspace", steps_offset=0, **unsupported):
        _reject("EulerAncestralDiscreteScheduler", unsupported, dict(trained_betas=(None,), rescale_betas_zero_snr=(None, False)))
        if prediction_type != "epsilon":
            raise NotImplementedError("only prediction_type='epsilon'")
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule,
                                   prediction_type=prediction_type, timestep_spacing=timestep_spacing, steps_offset=steps_offset)
        self.betas = _betas(beta_schedule, beta_start, beta_end, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        sigmas = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        sigmas = np.concatenate([sigmas[::-1], [0.0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)
        self.num_inference_steps = None
        self._timesteps_cpu = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy())
        self.timesteps = self._timesteps_cpu
        self.is_scale_input_called = False
        self._step_index = self._begin_index = None

    @property
    def init_noise_sigma(self):
        max_sigma = self.sigmas.max()
        if self.config.timestep_spacing in ("linspace", "trailing"):
            return max_sigma
        return (max_sigma ** 2 + 1) ** 0.5

    def set_timesteps(self, num_inference_steps, device=None):
        c = self.config
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, c.num_train_timesteps - 1, num_inference_steps, dtype=np.float32)[::-1].copy()
        elif c.timestep_spacing == "leading":
            ts = (np.arange(0, num_inference_steps) * (c.num_train_timesteps // num_inference_steps)).round()[::-1].copy().astype(np.float32)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = (np.arange(c.num_train_timesteps, 0, -c.num_train_timesteps / num_inference_steps)).round().copy().astype(np.float32)
            ts -= 1
        else:
            raise ValueError(c.timestep_spacing)
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))  # host table
        self._timesteps_cpu = torch.from_numpy(ts)
        self.timesteps = self._timesteps_cpu.to(device=device)
        self._step_index = self._begin_index = None

    def scale_model_input(self, sample, timestep):
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        self.is_scale_input_called = True
        return ops.scale_div(sample, float((sigma ** 2 + 1) ** 0.5))

    def step(self, model_output, timestep, sample, generator=None, return_dict=True):
        if isinstance(timestep, (int, torch.IntTensor, torch.LongTensor)):
            raise ValueError("Passing integer indices as timesteps to EulerAncestralDiscreteScheduler.step() is not supported.")
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma_from, sigma_to = self.sigmas[self._step_index], self.sigmas[self._step_index + 1]
        sigma_up = (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5
        sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
        from .pipelines import randn_tensor
        noise = randn_tensor(model_output.shape, dtype=model_output.dtype, device=model_output.device, generator=generator)
        prev = ops.linear_step(sample.to(model_output.dtype), m0=model_output, noise=noise, a=1.0, b=float(sigma_down - sigma_from), s=float(sigma_up))
        self._step_index += 1
        if not return_dict:
            return (prev, None)
        return SchedulerOutput(prev)


def _karras_sigmas_and_timesteps(train_sigmas, n, rho=7.0, sigma_min=None, sigma_max=None):
    """The reference's `_convert_to_karras` + `_sigma_to_t` (scheduling_dpmsolver_multistep.py:544-638; same numpy operations by necessity:
    the tables are compared bit for bit).  train_sigmas ascending (index = training timestep) -> (n descending sigmas, their float timesteps).
    sigma_min / sigma_max default to the extremes of the training schedule (DPM-Solver); Euler passes those of its interpolated table."""
    log_sigmas = np.log(train_sigmas)
    sigma_min = train_sigmas[0].item() if sigma_min is None else sigma_min
    sigma_max = train_sigmas[-1].item() if sigma_max is None else sigma_max
    ramp = np.linspace(0, 1, n)
    min_inv_rho, max_inv_rho = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sig = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
    ts = []
    for sigma in sig:
        log_sigma = np.log(np.maximum(sigma, 1e-10))
        dists = log_sigma - log_sigmas[:, np.newaxis]
        low_idx = np.cumsum((dists >= 0), axis=0).argmax(axis=0).clip(max=log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = log_sigmas[low_idx], log_sigmas[high_idx]
        w = np.clip((low - log_sigma) / (low - high), 0, 1)
        ts.append(((1 - w) * low_idx + w * high_idx).reshape(sigma.shape))
    return sig, np.array(ts)


class DPMSolverMultistepScheduler(_StepIndexMixin):
    """schedulers/scheduling_dpmsolver_multistep.py:132 - DPM-Solver++ (2M): algorithm_type 'dpmsolver++', solver_order 1 or 2,
    midpoint, epsilon prediction, sigmas interpolated from the training schedule or Karras-spaced (`use_karras_sigmas`: "DPM++ 2M
    Karras"; no Lu / exponential / beta / flow variants), final sigma 0.  Data prediction (:793-795) and the first / second order updates (:900-903, :980-992) are single
    b200_linear_step launches; the coefficients are computed from the fp32 sigma table exactly as the reference does."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", solver_order=2,
                 prediction_type="epsilon", algorithm_type="dpmsolver++", solver_type="midpoint", lower_order_final=True, euler_at_final=False,
                 final_sigmas_type="zero", lambda_min_clipped=-float("inf"), timestep_spacing="linspace", steps_offset=0, use_karras_sigmas=False,
                 **unsupported):
        _reject("DPMSolverMultistepScheduler", unsupported,
                dict(trained_betas=(None,), thresholding=(None, False), dynamic_thresholding_ratio=(0.995,), sample_max_value=(1.0, 1),
                     use_exponential_sigmas=(None, False), use_beta_sigmas=(None, False), use_lu_lambdas=(None, False),
                     use_flow_sigmas=(None, False), flow_shift=(1.0, 1), variance_type=(None,), rescale_betas_zero_snr=(None, False),
                     use_dynamic_shifting=(None, False), time_shift_type=("exponential",)))
        if (algorithm_type, solver_type, final_sigmas_type) != ("dpmsolver++", "midpoint", "zero") or solver_order not in (1, 2) or \
                prediction_type not in ("epsilon", "v_prediction"):
            raise NotImplementedError("DPMSolverMultistepScheduler: only epsilon / v_prediction, dpmsolver++ / midpoint / final sigma zero, order 1 or 2")
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule,
                                   solver_order=solver_order, prediction_type=prediction_type, algorithm_type=algorithm_type, solver_type=solver_type,
                                   lower_order_final=lower_order_final, euler_at_final=euler_at_final, final_sigmas_type=final_sigmas_type,
                                   lambda_min_clipped=lambda_min_clipped, timestep_spacing=timestep_spacing, steps_offset=steps_offset,
                                   use_karras_sigmas=bool(use_karras_sigmas), thresholding=False)
        self.betas = _betas(beta_schedule, beta_start, beta_end, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.alpha_t = torch.sqrt(self.alphas_cumprod)
        self.sigma_t = torch.sqrt(1 - self.alphas_cumprod)
        self.lambda_t = torch.log(self.alpha_t) - torch.log(self.sigma_t)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.num_inference_steps = None
        self._timesteps_cpu = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=np.float32)[::-1].copy())
        self.timesteps = self._timesteps_cpu
        self.model_outputs = [None] * solver_order
        self.lower_order_nums = 0
        self._step_index = self._begin_index = None

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def set_timesteps(self, num_inference_steps=None, device=None, timesteps=None):
        if timesteps is not None:
            raise NotImplementedError("custom timesteps")
        c = self.config
        clipped_idx = torch.searchsorted(torch.flip(self.lambda_t, [0]), c.lambda_min_clipped)
        last_timestep = ((c.num_train_timesteps - clipped_idx).numpy()).item()
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, last_timestep - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            step_ratio = last_timestep // (num_inference_steps + 1)
            ts = (np.arange(0, num_inference_steps + 1) * step_ratio).round()[::-1][:-1].copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = np.arange(last_timestep, 0, -c.num_train_timesteps / num_inference_steps).round().copy().astype(np.int64)
            ts -= 1
        else:
            raise ValueError(c.timestep_spacing)
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        if c.use_karras_sigmas:
            # "DPM++ 2M Karras" (scheduling_dpmsolver_multistep.py:444-449): Karras et al.'s rho = 7 spacing between the schedule's extreme
            # sigmas; a timestep is the rounded position of its sigma on the training schedule.  The step below reads only the sigma table.
            sig, ts = _karras_sigmas_and_timesteps(sig, num_inference_steps)
            ts = ts.round().astype(np.int64)
        else:
            sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0]]).astype(np.float32))  # host table
        self._timesteps_cpu = torch.from_numpy(ts)
        self.timesteps = self._timesteps_cpu.to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(ts)
        self.model_outputs = [None] * c.solver_order
        self.lower_order_nums = 0
        self._step_index = self._begin_index = None

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def step(self, model_output, timestep, sample, generator=None, variance_noise=None, return_dict=True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self._step_index is None:
            self._init_step_index(timestep)
        i, n, c = self._step_index, len(self._timesteps_cpu), self.config
        lower_order_final = (i == n - 1) and (c.euler_at_final or (c.lower_order_final and n < 15) or c.final_sigmas_type == "zero")
        # data prediction x0 = (x - sigma_t eps) / alpha_t at the current sigma (:793-795)
        a_cur, s_cur = self._alpha_sigma(self.sigmas[i])
        if c.prediction_type == "v_prediction":  # x0 = alpha_t x - sigma_t v (:797-799)
            x0 = ops.linear_step(sample.to(model_output.dtype), m0=model_output, a=float(a_cur), b=float(-s_cur))
        else:
            x0 = ops.linear_step(sample.to(model_output.dtype), m0=model_output, a=float(1.0 / a_cur), b=float(-s_cur / a_cur))
        for k in range(c.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = x0
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i + 1])
        alpha_s0, sigma_s0 = self._alpha_sigma(self.sigmas[i])
        lambda_t = torch.log(alpha_t) - torch.log(sigma_t)
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = lambda_t - lambda_s0
        A = float(sigma_t / sigma_s0)
        B = float(-(alpha_t * (torch.exp(-h) - 1.0)))
        if c.solver_order == 1 or self.lower_order_nums < 1 or lower_order_final:
            prev = ops.linear_step(sample.to(model_output.dtype), m0=x0, a=A, b=B)
        else:
            alpha_s1, sigma_s1 = self._alpha_sigma(self.sigmas[i - 1])
            lambda_s1 = torch.log(alpha_s1) - torch.log(sigma_s1)
            r0 = float((lambda_s0 - lambda_s1) / h)
            # x_t = A x + B D0 + 0.5 B D1,  D0 = m0,  D1 = (m0 - m1) / r0
            prev = ops.linear_step(sample.to(model_output.dtype), m0=x0, m1=self.model_outputs[-2], a=A, b=B * (1.0 + 0.5 / r0), c=-0.5 * B / r0)
        if self.lower_order_nums < c.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev)


class UniPCMultistepScheduler(_StepIndexMixin):
    """schedulers/scheduling_unipc_multistep.py:123 - UniPC (Zhao et al. 2302.04867) as the reference configures it by default: solver_order
    1 or 2, 'bh2', data prediction (predict_x0), epsilon prediction, corrector after every step, lower_order_final, sigmas interpolated from
    the training schedule (or Karras-spaced), final sigma 0.

    Every update of the reference is a linear combination of at most four tensors with scalar coefficients that depend only on the sigma
    table, so each is ONE b200_linear_step launch (fp32 arithmetic, one rounding) with the coefficients computed on the host in fp32 exactly
    as the reference computes its 0-dim tensors:
        x0        = x / alpha - (sigma / alpha) eps                                                        convert_model_output :796-799
        corrector = (s_t/s_s0) x_last - a_t h_phi_1 m0 - a_t B_h (rho_0 (m1 - m0)/r + rho_c (x0 - m0))     multistep_uni_c_bh_update :1080-1087
        predictor = (s_t/s_s0) x      - a_t h_phi_1 m0 - a_t B_h (1/2) (m1 - m0)/r                         multistep_uni_p_bh_update :940-947
    (rho = 1/2 at order 1; at order 2 the corrector's weights solve the reference's 2x2 system R rho = b).  The reference rounds those
    weights and every intermediate to the sample's 16-bit dtype; here they stay fp32 until the single final rounding."""
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", solver_order=2, prediction_type="epsilon",
                 predict_x0=True, solver_type="bh2", lower_order_final=True, disable_corrector=(), timestep_spacing="linspace", steps_offset=0,
                 final_sigmas_type="zero", use_karras_sigmas=False, **unsupported):
        _reject("UniPCMultistepScheduler", unsupported,
                dict(trained_betas=(None,), thresholding=(None, False), dynamic_thresholding_ratio=(0.995,), sample_max_value=(1.0, 1), solver_p=(None,),
                     use_exponential_sigmas=(None, False), use_beta_sigmas=(None, False), use_flow_sigmas=(None, False), flow_shift=(1.0, 1),
                     rescale_betas_zero_snr=(None, False), use_dynamic_shifting=(None, False), time_shift_type=("exponential",), sigma_min=(None,),
                     sigma_max=(None,), shift_terminal=(None,)))
        if solver_type in ("midpoint", "heun", "logrho"):
            solver_type = "bh2"  # the reference maps the DPM-Solver names onto bh2 (:274-276)
        if (predict_x0, solver_type, final_sigmas_type) != (True, "bh2", "zero") or solver_order not in (1, 2) or list(disable_corrector) or \
                prediction_type not in ("epsilon", "v_prediction"):
            raise NotImplementedError("UniPCMultistepScheduler: only epsilon / v_prediction, predict_x0 / bh2 / final sigma zero / corrector on, order 1 or 2")
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end, beta_schedule=beta_schedule,
                                   solver_order=solver_order, prediction_type=prediction_type, predict_x0=predict_x0, solver_type=solver_type,
                                   lower_order_final=lower_order_final, disable_corrector=[], timestep_spacing=timestep_spacing, steps_offset=steps_offset,
                                   final_sigmas_type=final_sigmas_type, use_karras_sigmas=bool(use_karras_sigmas), thresholding=False)
        self.betas = _betas(beta_schedule, beta_start, beta_end, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5
        self.num_inference_steps = None
        self._timesteps_cpu = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=np.float32)[::-1].copy())
        self.timesteps = self._timesteps_cpu
        self.model_outputs = [None] * solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = None
        self._step_index = self._begin_index = None

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None, mu=None):
        if sigmas is not None or mu is not None:
            raise NotImplementedError("custom sigmas / dynamic shifting (flow-matching variants)")
        c = self.config
        N = c.num_train_timesteps
        if c.timestep_spacing == "linspace":
            ts = np.linspace(0, N - 1, num_inference_steps + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif c.timestep_spacing == "leading":
            ts = (np.arange(0, num_inference_steps + 1) * (N // (num_inference_steps + 1))).round()[::-1][:-1].copy().astype(np.int64)
            ts += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ts = np.arange(N, 0, -N / num_inference_steps).round().copy().astype(np.int64)
            ts -= 1
        else:
            raise ValueError(c.timestep_spacing)
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        if c.use_karras_sigmas:
            sig, ts = _karras_sigmas_and_timesteps(sig, num_inference_steps)  # :375-383
            ts = ts.round().astype(np.int64)
        else:
            sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0]]).astype(np.float32))  # host table
        self._timesteps_cpu = torch.from_numpy(ts)
        self.timesteps = self._timesteps_cpu.to(device=device, dtype=torch.int64)
        self.num_inference_steps = len(ts)
        self.model_outputs = [None] * c.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.this_order = None
        self._step_index = self._begin_index = None

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1 / ((sigma ** 2 + 1) ** 0.5)
        return alpha_t, sigma * alpha_t

    def _lambda(self, idx):
        a, s = self._alpha_sigma(self.sigmas[idx])
        return torch.log(a) - torch.log(s)

    def _bh2(self, idx_t, idx_s0, idx_s1):
        """The scalars of one update from sigma index idx_s0 to idx_t (idx_s1: the older point of an order-2 update, or None):
        s_t/s_s0, a_t, h_phi_1 = B_h = expm1(-h), r = (lambda_s1 - lambda_s0) / h, and the reference's b vector (:905-916)."""
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[idx_t])
        alpha_s0, sigma_s0 = self._alpha_sigma(self.sigmas[idx_s0])
        lambda_s0 = torch.log(alpha_s0) - torch.log(sigma_s0)
        h = (torch.log(alpha_t) - torch.log(sigma_t)) - lambda_s0
        hh = -h
        h_phi_1 = torch.expm1(hh)
        B_h = torch.expm1(hh)
        r = (self._lambda(idx_s1) - lambda_s0) / h if idx_s1 is not None else None
        h_phi_k = h_phi_1 / hh - 1
        b1 = h_phi_k / B_h
        b2 = (h_phi_k / hh - 0.5) * 2 / B_h
        return sigma_t / sigma_s0, alpha_t, h_phi_1, B_h, r, b1, b2

    def step(self, model_output, timestep, sample, return_dict=True):
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self._step_index is None:
            self._init_step_index(timestep)
        i, n, c = self._step_index, len(self._timesteps_cpu), self.config
        dt = model_output.dtype
        a_cur, s_cur = self._alpha_sigma(self.sigmas[i])
        if c.prediction_type == "v_prediction":  # x0 = alpha_t x - sigma_t v (:800-801)
            x0 = ops.linear_step(sample.to(dt), m0=model_output, a=float(a_cur), b=float(-s_cur))
        else:
            x0 = ops.linear_step(sample.to(dt), m0=model_output, a=float(1.0 / a_cur), b=float(-s_cur / a_cur))
        if i > 0 and self.last_sample is not None:
            # ---- corrector: re-does the step that led here (from sigma[i-1] to sigma[i]) now that x0 at its end point is known
            m0 = self.model_outputs[-1]
            if self.this_order == 1:
                A, a_t, hp1, Bh, _, _, _ = self._bh2(i, i - 1, None)
                sample = ops.linear_step(self.last_sample, m0=m0, m1=x0, a=float(A), b=float(-a_t * hp1 + a_t * Bh * 0.5), c=float(-a_t * Bh * 0.5))
            else:
                A, a_t, hp1, Bh, r, b1, b2 = self._bh2(i, i - 1, i - 2)
                # R rho = b with R = [[1, 1], [r, 1]] (:1062-1076): rho_0 weighs D1 = (m1 - m0) / r, rho_c weighs x0 - m0
                rho_0 = (b1 - b2) / (1.0 - r)
                rho_c = b1 - rho_0
                m1 = self.model_outputs[-2]
                sample = ops.linear_step(self.last_sample, m0=m0, m1=m1, noise=x0, a=float(A), b=float(-a_t * hp1 + a_t * Bh * (rho_0 / r + rho_c)),
                                         c=float(-a_t * Bh * rho_0 / r), s=float(-a_t * Bh * rho_c))
        for k in range(c.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
        self.model_outputs[-1] = x0
        this_order = min(c.solver_order, n - i) if c.lower_order_final else c.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)  # warm-up of the multistep history
        self.last_sample = sample
        # ---- predictor: from sigma[i] to sigma[i+1]
        if self.this_order == 1:
            A, a_t, hp1, Bh, _, _, _ = self._bh2(i + 1, i, None)
            prev = ops.linear_step(sample.to(dt), m0=x0, a=float(A), b=float(-a_t * hp1))
        else:
            A, a_t, hp1, Bh, r, _, _ = self._bh2(i + 1, i, i - 1)
            prev = ops.linear_step(sample.to(dt), m0=x0, m1=self.model_outputs[-2], a=float(A), b=float(-a_t * hp1 + a_t * Bh * 0.5 / r), c=float(-a_t * Bh * 0.5 / r))
        if self.lower_order_nums < c.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev)

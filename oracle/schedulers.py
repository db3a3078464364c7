"""ORACLE (test infrastructure): CPU restatement of the three schedulers on the path, including their tensor
math, op for op (so dtype promotion / rounding points are the reference's).

Reference: schedulers/scheduling_euler_discrete.py (__init__ :203-276, set_timesteps :350-482, scale_model_input
:326-348, step :685-797), schedulers/scheduling_flow_match_euler_discrete.py (__init__ :77-140, time_shift :241,
set_timesteps :283-386, step :423-530), schedulers/scheduling_ddpm.py (__init__ :166-240, set_timesteps :274-346,
_get_variance :348-400, step :461-570, previous_timestep :648-670).
"""
import math

import numpy as np
import torch


class EulerDiscrete:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 timestep_spacing="linspace", steps_offset=0, prediction_type="epsilon"):
        self.N = num_train_timesteps
        self.spacing, self.offset, self.prediction_type = timestep_spacing, steps_offset, prediction_type
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).flip(0)
        self.sigmas = torch.cat([sig, torch.zeros(1)])
        self.timesteps = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy()).to(torch.float32)
        self.step_index = None

    @property
    def init_noise_sigma(self):
        m = self.sigmas.max()
        return m if self.spacing in ("linspace", "trailing") else (m ** 2 + 1) ** 0.5

    def _sigma_to_t(self, sigma, log_sigmas):
        """:484-516"""
        log_sigma = np.log(np.maximum(sigma, 1e-10))
        dists = log_sigma - log_sigmas[:, np.newaxis]
        low_idx = np.cumsum((dists >= 0), axis=0).argmax(axis=0).clip(max=log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = log_sigmas[low_idx], log_sigmas[high_idx]
        w = np.clip((low - log_sigma) / (low - high), 0, 1)
        t = (1 - w) * low_idx + w * high_idx
        return t.reshape(sigma.shape)

    def set_timesteps(self, n=None, sigmas=None):
        if sigmas is not None:  # custom sigma schedule (:424-427)
            log_sigmas = np.log(np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5))
            sigmas = np.array(sigmas).astype(np.float32)
            ts = np.array([self._sigma_to_t(sigma, log_sigmas) for sigma in sigmas[:-1]])
            self.sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32)
            self.timesteps = torch.from_numpy(ts.astype(np.float32))
            self.step_index = 0
            return
        if self.spacing == "linspace":
            ts = np.linspace(0, self.N - 1, n, dtype=np.float32)[::-1].copy()
        elif self.spacing == "leading":
            ts = (np.arange(0, n) * (self.N // n)).round()[::-1].copy().astype(np.float32)
            ts += self.offset
        elif self.spacing == "trailing":
            ts = (np.arange(self.N, 0, -self.N / n)).round().copy().astype(np.float32)
            ts -= 1
        else:
            raise ValueError(self.spacing)
        sig = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts.astype(np.float32))
        self.step_index = 0

    def scale_model_input(self, sample):
        sigma = self.sigmas[self.step_index]
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, sample):
        sample = sample.to(torch.float32)
        sigma = self.sigmas[self.step_index]
        if self.prediction_type == "epsilon":
            pred = sample - sigma * model_output
        elif self.prediction_type == "v_prediction":
            pred = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))
        else:
            raise ValueError(self.prediction_type)
        derivative = (sample - pred) / sigma
        dt = self.sigmas[self.step_index + 1] - sigma
        prev = (sample + derivative * dt).to(model_output.dtype)
        self.step_index += 1
        return prev


class FlowMatchEuler:
    def __init__(self, num_train_timesteps=1000, shift=1.0, use_dynamic_shifting=False):
        self.N, self.shift, self.dynamic = num_train_timesteps, shift, use_dynamic_shifting
        ts = torch.from_numpy(np.linspace(1, num_train_timesteps, num_train_timesteps, dtype=np.float32)[::-1].copy()).to(torch.float32)
        sig = ts / num_train_timesteps
        if not use_dynamic_shifting:
            sig = shift * sig / (1 + (shift - 1) * sig)
        self.sigmas = sig
        self.timesteps = sig * num_train_timesteps
        self.sigma_min, self.sigma_max = sig[-1].item(), sig[0].item()
        self.step_index = None

    def set_timesteps(self, n=None, sigmas=None, mu=None):
        if sigmas is None:
            ts = np.linspace(self.sigma_max * self.N, self.sigma_min * self.N, n)
            sigmas = ts / self.N
        else:
            sigmas = np.array(sigmas).astype(np.float32)
        if self.dynamic:
            sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1) ** 1.0)
        else:
            sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        sig = torch.from_numpy(sigmas).to(dtype=torch.float32)
        self.timesteps = sig * self.N
        self.sigmas = torch.cat([sig, torch.zeros(1)])
        self.step_index = 0

    def step(self, model_output, sample):
        sample = sample.to(torch.float32)
        dt = self.sigmas[self.step_index + 1] - self.sigmas[self.step_index]
        prev = (sample + dt * model_output).to(model_output.dtype)
        self.step_index += 1
        return prev


class DDPM:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 variance_type="fixed_small", clip_sample=True, clip_sample_range=1.0, timestep_spacing="leading",
                 steps_offset=0):
        self.N = num_train_timesteps
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.one = torch.tensor(1.0)
        self.variance_type, self.clip, self.clip_range = variance_type, clip_sample, clip_sample_range
        self.spacing, self.offset = timestep_spacing, steps_offset
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())
        self.num_inference_steps = None

    def set_timesteps(self, n):
        self.num_inference_steps = n
        if self.spacing == "linspace":
            ts = np.linspace(0, self.N - 1, n).round()[::-1].copy().astype(np.int64)
        elif self.spacing == "leading":
            ts = (np.arange(0, n) * (self.N // n)).round()[::-1].copy().astype(np.int64)
            ts += self.offset
        elif self.spacing == "trailing":
            ts = np.round(np.arange(self.N, 0, -self.N / n)).astype(np.int64)
            ts -= 1
        else:
            raise ValueError(self.spacing)
        self.timesteps = torch.from_numpy(ts)

    def previous_timestep(self, t):
        if self.num_inference_steps:
            index = (self.timesteps == t).nonzero(as_tuple=True)[0][0]
            return torch.tensor(-1) if index == self.timesteps.shape[0] - 1 else self.timesteps[index + 1]
        return t - 1

    def _get_variance(self, t):
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        cur_beta = 1 - a_t / a_prev
        var = (1 - a_prev) / (1 - a_t) * cur_beta
        return torch.clamp(var, min=1e-20)

    def step(self, model_output, t, sample, generator=None):
        prev_t = self.previous_timestep(t)
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        b_t = 1 - a_t
        b_prev = 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
        if self.clip:
            x0 = x0.clamp(-self.clip_range, self.clip_range)
        c0 = (a_prev ** 0.5 * cur_beta) / b_t
        c1 = cur_alpha ** 0.5 * b_prev / b_t
        prev = c0 * x0 + c1 * sample
        variance = 0
        if t > 0:
            noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
            variance = (self._get_variance(t) ** 0.5) * noise
        return prev + variance


# ----------------------------------------------------------------------------------------------------------------------
# SURVEY.md 8f N4 steppers, restated op for op (tensor ops on the caller's dtype, fp32 0-dim table entries as scalars, the
# upcasts where the reference has them) - pinned by tests/golden/schedulers2.pt (recorded from the real reference).
# ----------------------------------------------------------------------------------------------------------------------
def _betas(schedule, b0, b1, n):
    if schedule == "linear":
        return torch.linspace(b0, b1, n, dtype=torch.float32)
    if schedule == "scaled_linear":
        return torch.linspace(b0 ** 0.5, b1 ** 0.5, n, dtype=torch.float32) ** 2
    raise NotImplementedError(schedule)


class DDIM:
    """schedulers/scheduling_ddim.py: __init__ :193-245, set_timesteps :328-381, step :384-520 (eta = 0, epsilon, no clipping)"""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", set_alpha_to_one=True,
                 steps_offset=0, timestep_spacing="leading", clip_sample=False, prediction_type="epsilon", clip_sample_range=1.0):
        self.clip = clip_sample_range if clip_sample else None  # :476-479 (pixel-space models; the reference's own test configuration)
        self.N, self.offset, self.spacing, self.prediction_type = num_train_timesteps, steps_offset, timestep_spacing, prediction_type
        self.alphas_cumprod = torch.cumprod(1.0 - _betas(beta_schedule, beta_start, beta_end, num_train_timesteps), dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n):
        self.n = n
        if self.spacing == "linspace":
            ts = np.linspace(0, self.N - 1, n).round()[::-1].copy().astype(np.int64)
        elif self.spacing == "leading":
            ts = (np.arange(0, n) * (self.N // n)).round()[::-1].copy().astype(np.int64) + self.offset
        else:
            ts = np.round(np.arange(self.N, 0, -self.N / n)).astype(np.int64) - 1
        self.timesteps = torch.from_numpy(ts)

    def scale_model_input(self, x, t):
        return x

    def step(self, model_output, timestep, sample, generator=None):
        t = int(timestep)
        prev_t = t - self.N // self.n
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        if self.prediction_type == "v_prediction":  # :468-470
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        else:
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        if self.clip is not None:
            x0 = x0.clamp(-self.clip, self.clip)
        direction = (1 - a_prev) ** 0.5 * eps
        return a_prev ** 0.5 * x0 + direction


class EulerAncestral:
    """schedulers/scheduling_euler_ancestral_discrete.py: __init__ :164-210, set_timesteps :285-328, scale_model_input :258-283,
    step :376-455 (epsilon)"""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", timestep_spacing="linspace",
                 steps_offset=0):
        self.N, self.offset, self.spacing = num_train_timesteps, steps_offset, timestep_spacing
        self.alphas_cumprod = torch.cumprod(1.0 - _betas(beta_schedule, beta_start, beta_end, num_train_timesteps), dim=0)
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        self.sigmas = torch.from_numpy(np.concatenate([sig[::-1], [0.0]]).astype(np.float32))
        self.i = None

    @property
    def init_noise_sigma(self):
        m = self.sigmas.max()
        return m if self.spacing in ("linspace", "trailing") else (m ** 2 + 1) ** 0.5

    def set_timesteps(self, n):
        if self.spacing == "linspace":
            ts = np.linspace(0, self.N - 1, n, dtype=np.float32)[::-1].copy()
        elif self.spacing == "leading":
            ts = (np.arange(0, n) * (self.N // n)).round()[::-1].copy().astype(np.float32) + self.offset
        else:
            ts = (np.arange(self.N, 0, -self.N / n)).round().copy().astype(np.float32) - 1
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        sig = np.interp(ts, np.arange(0, len(sig)), sig)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.i = 0

    def scale_model_input(self, x, t):
        sigma = self.sigmas[self.i]
        return x / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, timestep, sample, generator=None):
        sigma = self.sigmas[self.i]
        sample = sample.to(torch.float32)
        x0 = sample - sigma * model_output
        s_from, s_to = self.sigmas[self.i], self.sigmas[self.i + 1]
        s_up = (s_to ** 2 * (s_from ** 2 - s_to ** 2) / s_from ** 2) ** 0.5
        s_down = (s_to ** 2 - s_up ** 2) ** 0.5
        prev = sample + (sample - x0) / sigma * (s_down - sigma)
        noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype)
        prev = (prev + noise * s_up).to(model_output.dtype)
        self.i += 1
        return prev


def karras_sigmas_and_timesteps(train_sigmas, n, rho=7.0):
    """_convert_to_karras (scheduling_dpmsolver_multistep.py:603-638, scheduling_euler_discrete.py:520-556) on the flipped training sigmas,
    then _sigma_to_t (:544-578) for each: log-linear interpolation of the sigma's position on the training schedule."""
    log_sigmas = np.log(train_sigmas)
    desc = np.flip(train_sigmas).copy()
    smin, smax = desc[-1].item(), desc[0].item()
    ramp = np.linspace(0, 1, n)
    lo, hi = smin ** (1 / rho), smax ** (1 / rho)
    sig = (hi + ramp * (lo - hi)) ** rho

    def sigma_to_t(sigma):
        log_sigma = np.log(np.maximum(sigma, 1e-10))
        dists = log_sigma - log_sigmas[:, np.newaxis]
        low_idx = np.cumsum((dists >= 0), axis=0).argmax(axis=0).clip(max=log_sigmas.shape[0] - 2)
        high_idx = low_idx + 1
        low, high = log_sigmas[low_idx], log_sigmas[high_idx]
        w = np.clip((low - log_sigma) / (low - high), 0, 1)
        return ((1 - w) * low_idx + w * high_idx).reshape(sigma.shape)

    return sig, np.array([sigma_to_t(x) for x in sig])


class DPMSolverPP2M:
    """schedulers/scheduling_dpmsolver_multistep.py: __init__ :215-330, set_timesteps :366-497, _sigma_to_alpha_sigma_t :577-600,
    convert_model_output :745-815 (dpmsolver++, epsilon), dpm_solver_first_order_update :855-915,
    multistep_dpm_solver_second_order_update :925-1010 (midpoint), step :1196-1282"""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", solver_order=2,
                 timestep_spacing="linspace", steps_offset=0, lower_order_final=True, use_karras_sigmas=False, prediction_type="epsilon",
                 final_sigmas_type="zero", euler_at_final=False):
        self.N, self.offset, self.spacing, self.order, self.lof = num_train_timesteps, steps_offset, timestep_spacing, solver_order, lower_order_final
        self.karras, self.prediction_type, self.final, self.eaf = use_karras_sigmas, prediction_type, final_sigmas_type, euler_at_final
        self.alphas_cumprod = torch.cumprod(1.0 - _betas(beta_schedule, beta_start, beta_end, num_train_timesteps), dim=0)
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n):
        last = self.N  # lambda_min_clipped = -inf
        if self.spacing == "linspace":
            ts = np.linspace(0, last - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.spacing == "leading":
            ts = (np.arange(0, n + 1) * (last // (n + 1))).round()[::-1][:-1].copy().astype(np.int64) + self.offset
        else:
            ts = np.arange(last, 0, -self.N / n).round().copy().astype(np.int64) - 1
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        if self.karras:
            # :444-449 "DPM++ 2M Karras": the n sigmas of Karras et al. (2206.00364, rho = 7) between the schedule's extremes, timesteps = the
            # (rounded) positions of those sigmas on the training schedule
            sig, ts = karras_sigmas_and_timesteps(sig, n)
            ts = ts.round()
        else:
            sig = np.interp(ts, np.arange(0, len(sig)), sig)
        last = 0 if self.final == "zero" else ((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5  # :473-481
        self.sigmas = torch.from_numpy(np.concatenate([sig, [last]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts).to(torch.int64)
        self.outs, self.lower, self.i = [None] * self.order, 0, 0

    def scale_model_input(self, x, t):
        return x

    @staticmethod
    def _as(sigma):
        a = 1 / ((sigma ** 2 + 1) ** 0.5)
        return a, sigma * a

    def step(self, model_output, timestep, sample, generator=None):
        n, i = len(self.timesteps), self.i
        final = (i == n - 1) and (self.eaf or (self.lof and n < 15) or self.final == "zero")  # :1235-1240
        a_c, s_c = self._as(self.sigmas[i])
        x0 = a_c * sample - s_c * model_output if self.prediction_type == "v_prediction" else (sample - s_c * model_output) / a_c  # :793-799
        self.outs = self.outs[1:] + [x0]
        sample = sample.to(torch.float32)
        a_t, s_t = self._as(self.sigmas[i + 1])
        a_s0, s_s0 = self._as(self.sigmas[i])
        lam_t, lam_s0 = torch.log(a_t) - torch.log(s_t), torch.log(a_s0) - torch.log(s_s0)
        h = lam_t - lam_s0
        if self.order == 1 or self.lower < 1 or final:
            prev = (s_t / s_s0) * sample - (a_t * (torch.exp(-h) - 1.0)) * x0
        else:
            a_s1, s_s1 = self._as(self.sigmas[i - 1])
            lam_s1 = torch.log(a_s1) - torch.log(s_s1)
            r0 = (lam_s0 - lam_s1) / h
            m0, m1 = self.outs[-1], self.outs[-2]
            d1 = (1.0 / r0) * (m0 - m1)
            prev = (s_t / s_s0) * sample - (a_t * (torch.exp(-h) - 1.0)) * m0 - 0.5 * (a_t * (torch.exp(-h) - 1.0)) * d1
        if self.lower < self.order:
            self.lower += 1
        self.i += 1
        return prev.to(model_output.dtype)


class UniPC:
    """schedulers/scheduling_unipc_multistep.py - UniPCMultistepScheduler with its defaults: solver_order 2, epsilon prediction,
    predict_x0, solver_type 'bh2', lower_order_final, corrector on, final sigma 0.  set_timesteps :318-482 (default branch),
    convert_model_output :760-832, multistep_uni_p_bh_update :833-960 (predictor), multistep_uni_c_bh_update :962-1098 (corrector),
    step :1153-1232.  Same torch operations in the same order (0-dim fp32 coefficient tensors against tensors of the sample's dtype),
    so the recorded reference trajectories are reproduced bit for bit."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear", solver_order=2,
                 timestep_spacing="linspace", steps_offset=0, prediction_type="epsilon", final_sigmas_type="zero", use_karras_sigmas=False):
        self.N, self.offset, self.spacing, self.order = num_train_timesteps, steps_offset, timestep_spacing, solver_order
        self.prediction_type, self.final, self.karras = prediction_type, final_sigmas_type, use_karras_sigmas
        self.alphas_cumprod = torch.cumprod(1.0 - _betas(beta_schedule, beta_start, beta_end, num_train_timesteps), dim=0)
        self.init_noise_sigma = 1.0

    def set_timesteps(self, n):
        if self.spacing == "linspace":
            ts = np.linspace(0, self.N - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.spacing == "leading":
            ts = (np.arange(0, n + 1) * (self.N // (n + 1))).round()[::-1][:-1].copy().astype(np.int64) + self.offset
        else:
            ts = np.arange(self.N, 0, -self.N / n).round().copy().astype(np.int64) - 1
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        if self.karras:  # :374-391: the last sigma of the Karras table itself is "sigma_min"
            sig, ts = karras_sigmas_and_timesteps(sig, n)
            ts = ts.round().astype(np.int64)
            last = 0 if self.final == "zero" else sig[-1]
        else:
            sig = np.interp(ts, np.arange(0, len(sig)), sig)
            last = 0 if self.final == "zero" else ((1 - self.alphas_cumprod[0]) / self.alphas_cumprod[0]) ** 0.5  # :455-462
        self.sigmas = torch.from_numpy(np.concatenate([sig, [last]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.outs, self.lower, self.i, self.last_sample, self.this_order = [None] * self.order, 0, 0, None, None

    def scale_model_input(self, x, t):
        return x

    @staticmethod
    def _as(sigma):
        a = 1 / ((sigma ** 2 + 1) ** 0.5)
        return a, sigma * a

    def _lam(self, idx):
        a, s = self._as(self.sigmas[idx])
        return torch.log(a) - torch.log(s)

    def _bh(self, h, rks, order):
        hh = -h
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        fact = 1
        B_h = torch.expm1(hh)  # bh2
        R, b = [], []
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return h_phi_1, B_h, torch.stack(R), torch.stack(b)

    def _predict(self, x, order):
        i, m0 = self.i, self.outs[-1]
        a_t, s_t = self._as(self.sigmas[i + 1])
        a_s0, s_s0 = self._as(self.sigmas[i])
        lam_s0 = torch.log(a_s0) - torch.log(s_s0)
        h = (torch.log(a_t) - torch.log(s_t)) - lam_s0
        rks, D1s = [], []
        for k in range(1, order):
            rk = (self._lam(i - k) - lam_s0) / h
            rks.append(rk)
            D1s.append((self.outs[-(k + 1)] - m0) / rk)
        rks.append(torch.ones(()))
        h_phi_1, B_h, R, b = self._bh(h, torch.stack(rks), order)
        x_t_ = s_t / s_s0 * x - a_t * h_phi_1 * m0
        if D1s:
            rhos_p = torch.ones(1, dtype=x.dtype) * 0.5  # order 2: the simplified weights
            pred_res = torch.einsum("k,bkc...->bc...", rhos_p, torch.stack(D1s, dim=1))
        else:
            pred_res = 0
        return (x_t_ - a_t * B_h * pred_res).to(x.dtype)

    def _correct(self, model_t, last_sample, order):
        i, m0, x = self.i, self.outs[-1], last_sample
        a_t, s_t = self._as(self.sigmas[i])
        a_s0, s_s0 = self._as(self.sigmas[i - 1])
        lam_s0 = torch.log(a_s0) - torch.log(s_s0)
        h = (torch.log(a_t) - torch.log(s_t)) - lam_s0
        rks, D1s = [], []
        for k in range(1, order):
            rk = (self._lam(i - (k + 1)) - lam_s0) / h
            rks.append(rk)
            D1s.append((self.outs[-(k + 1)] - m0) / rk)
        rks.append(torch.ones(()))
        h_phi_1, B_h, R, b = self._bh(h, torch.stack(rks), order)
        rhos_c = torch.ones(1, dtype=x.dtype) * 0.5 if order == 1 else torch.linalg.solve(R, b).to(x.dtype)
        x_t_ = s_t / s_s0 * x - a_t * h_phi_1 * m0
        corr_res = torch.einsum("k,bkc...->bc...", rhos_c[:-1], torch.stack(D1s, dim=1)) if D1s else 0
        D1_t = model_t - m0
        return (x_t_ - a_t * B_h * (corr_res + rhos_c[-1] * D1_t)).to(x.dtype)

    def step(self, model_output, timestep, sample, generator=None):
        n, i = len(self.timesteps), self.i
        a_c, s_c = self._as(self.sigmas[i])
        x0 = a_c * sample - s_c * model_output if self.prediction_type == "v_prediction" else (sample - s_c * model_output) / a_c  # :796-801
        if i > 0 and self.last_sample is not None:
            sample = self._correct(x0, self.last_sample, self.this_order)
        self.outs = self.outs[1:] + [x0]
        self.this_order = min(min(self.order, n - i), self.lower + 1)  # lower_order_final, multistep warm-up
        self.last_sample = sample
        prev = self._predict(sample, self.this_order)
        if self.lower < self.order:
            self.lower += 1
        self.i += 1
        return prev

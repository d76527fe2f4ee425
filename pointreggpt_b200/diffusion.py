"""`GaussianDiffusion` under the reference's name and constructor (SDD:1015-1151) with
`sample()` (SDD:1394-1409) running the whole T-step loop on the device through
`prg_sampler_run`: no per-step host sync, no per-step tensor ops in Python.

Only the sampling half is implemented (the north-star path has no backward pass).  Schedule
buffers are built exactly like the reference (float64 -> float32, same names and registration
order, so the 293-entry state-dict layout is unchanged); per-step coefficients are derived
from them with the same fp32 tensor arithmetic the reference performs per step.
"""
import ctypes
import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _ffi
from .geometry import get_mask_from_img_cond  # noqa: F401  (re-exported like the reference)


def sigmoid_beta_schedule(timesteps, start=-3, end=3, tau=1, clamp_min=1e-5):
    """SDD:997-1012."""
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    v_start = torch.tensor(start / tau).sigmoid()
    v_end = torch.tensor(end / tau).sigmoid()
    ac = (-((t * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


def linear_beta_schedule(timesteps):
    """SDD:976-980."""
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):
    """SDD:983-994."""
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)


class GaussianDiffusion(nn.Module):
    def __init__(self, model, *, image_size, timesteps=1000, sampling_timesteps=None,
                 loss_type='l1', objective='pred_noise', beta_schedule='cosine',
                 ddim_sampling_eta=1., min_snr_loss_weight=False, min_snr_gamma=5,
                 is_ddnm_sampling=True, ddnm_sampling_dropout=0., ddnm_dropout_schedule='none'):
        super().__init__()
        assert not (type(self) == GaussianDiffusion and model.channels != model.out_dim)
        assert not model.random_or_learned_sinusoidal_cond
        self.model = model
        self.channels = self.model.channels
        self.image_size = image_size
        self.objective = objective
        assert objective in {'pred_noise', 'pred_x0', 'pred_v'}
        if objective != 'pred_x0':
            raise NotImplementedError(
                "the native sampler implements the shipped objective 'pred_x0' (GD:41)")
        if beta_schedule == 'linear':
            betas = linear_beta_schedule(timesteps)
        elif beta_schedule == 'cosine':
            betas = cosine_beta_schedule(timesteps)
        elif beta_schedule == 'sigmoid':
            betas = sigmoid_beta_schedule(timesteps)
        else:
            raise ValueError(f'unknown beta schedule {beta_schedule}')
        alphas = 1. - betas
        alphas_cumprod = torch.cumprod(alphas, dim=0)
        alphas_cumprod_prev = F.pad(alphas_cumprod[:-1], (1, 0), value=1.)
        timesteps, = betas.shape
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else timesteps
        assert self.sampling_timesteps <= timesteps
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        self.ddim_sampling_eta = ddim_sampling_eta
        self.is_ddnm_sampling = is_ddnm_sampling
        self.ddnm_sampling_dropout = ddnm_sampling_dropout
        if ddnm_dropout_schedule not in ('none', 'linear'):
            raise ValueError(f'unknown ddnm dropout schedule {ddnm_dropout_schedule}')
        if ddnm_sampling_dropout != 0.:
            raise NotImplementedError("DDNM keep-mask dropout (SDD:1213-1216) is off on the "
                                      "data-generation path and not implemented natively")

        def reg(name, val):
            self.register_buffer(name, val.to(torch.float32))

        reg('betas', betas)
        reg('alphas_cumprod', alphas_cumprod)
        reg('alphas_cumprod_prev', alphas_cumprod_prev)
        reg('sqrt_alphas_cumprod', torch.sqrt(alphas_cumprod))
        reg('sqrt_one_minus_alphas_cumprod', torch.sqrt(1. - alphas_cumprod))
        reg('log_one_minus_alphas_cumprod', torch.log(1. - alphas_cumprod))
        reg('sqrt_recip_alphas_cumprod', torch.sqrt(1. / alphas_cumprod))
        reg('sqrt_recipm1_alphas_cumprod', torch.sqrt(1. / alphas_cumprod - 1))
        posterior_variance = betas * (1. - alphas_cumprod_prev) / (1. - alphas_cumprod)
        reg('posterior_variance', posterior_variance)
        reg('posterior_log_variance_clipped', torch.log(posterior_variance.clamp(min=1e-20)))
        reg('posterior_mean_coef1', betas * torch.sqrt(alphas_cumprod_prev) / (1. - alphas_cumprod))
        reg('posterior_mean_coef2',
            (1. - alphas_cumprod_prev) * torch.sqrt(alphas) / (1. - alphas_cumprod))
        snr = alphas_cumprod / (1 - alphas_cumprod)
        clipped = snr.clone()
        if min_snr_loss_weight:
            clipped.clamp_(max=min_snr_gamma)
        reg('loss_weight', clipped)   # objective == 'pred_x0' (SDD:1146-1147)

    # ------------------------------------------------------------------ step tables
    def sampling_steps(self, has_refine_step=False):
        """The list of `prg_step` the device loop executes, derived like SDD:1283-1392."""
        S = _ffi.Step
        steps = []
        if not self.is_ddim_sampling:
            c1 = self.posterior_mean_coef1.detach().cpu()
            c2 = self.posterior_mean_coef2.detach().cpu()
            lv = self.posterior_log_variance_clipped.detach().cpu()
            sig = (0.5 * lv).exp()                                   # SDD:1280
            for t in reversed(range(self.num_timesteps)):
                steps.append(S(t, _ffi.STEP_P_SAMPLE, int(t > 0), 0, float(c1[t]), float(c2[t]),
                               float(sig[t]), 0.0, 0.0))
            if has_refine_step:
                steps.append(S(0, _ffi.STEP_REFINE_P, 0, 0, float(c1[0]), float(c2[0]),
                               float(sig[0]), 0.0, 0.0))
        else:
            ac = self.alphas_cumprod.detach().cpu()
            r = self.sqrt_recip_alphas_cumprod.detach().cpu()
            rm1 = self.sqrt_recipm1_alphas_cumprod.detach().cpu()
            times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
            times = list(reversed(times.int().tolist()))
            eta = self.ddim_sampling_eta
            for t, t_next in zip(times[:-1], times[1:]):
                if t_next < 0:
                    steps.append(S(t, _ffi.STEP_DDIM_LAST, 0, 0, float(r[t]), float(rm1[t]),
                                   0.0, 0.0, 0.0))
                    continue
                alpha, alpha_next = ac[t], ac[t_next]
                sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
                c = (1 - alpha_next - sigma ** 2).sqrt()
                steps.append(S(t, _ffi.STEP_DDIM, 1, 0, float(r[t]), float(rm1[t]),
                               float(alpha_next.sqrt()), float(c), float(sigma)))
            if has_refine_step:
                steps.append(S(0, _ffi.STEP_REFINE_DDIM, 0, 0, float(r[0]), float(rm1[0]),
                               0.0, 0.0, 0.0))
        steps[-1].unnormalize = 1
        return steps

    def num_noise_draws(self, has_refine_step=False):
        """1 (x_T) + the number of `randn_like` draws of one `sample()` call."""
        return 1 + sum(s.add_noise for s in self.sampling_steps(has_refine_step))

    # ------------------------------------------------------------------ sampling
    @torch.no_grad()
    def sample(self, *, param_cond, img_cond=None, disable_tqdm=False, has_refine_step=False,
               noise=None, seed=None):
        """(b,4) intrinsics vector [+ (b,2,s,s) DDNM image condition] -> (b,1,s,s) in [0,1].

        `noise` (optional, (num_noise_draws, b,1,s,s)) injects the Gaussian draws (parity tests);
        otherwise the device Philox generator is seeded from `seed` or torch's global RNG.
        """
        _ffi.require_cuda(param_cond, img_cond, noise)
        b, s = param_cond.shape[0], self.image_size
        dev = param_cond.device
        steps = self.sampling_steps(has_refine_step)
        arr = (_ffi.Step * len(steps))(*steps)
        p = param_cond.float().contiguous()
        ic = None
        if img_cond is not None and self.is_ddnm_sampling:
            ic = img_cond.float().contiguous()
            assert ic.shape == (b, 2, s, s)
        if has_refine_step and ic is None:
            raise NotImplementedError("has_refine_step needs DDNM sampling with an image "
                                      "condition (SDD:1313)")
        if noise is not None:
            noise = noise.float().contiguous()
            need = 1 + sum(st.add_noise for st in steps)
            assert noise.shape[0] >= need and tuple(noise.shape[1:]) == (b, 1, s, s), \
                "noise must be (num_noise_draws, b, 1, s, s)"
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        out = torch.empty((b, self.channels, s, s), dtype=torch.float32, device=dev)
        h, cap = self.model.native_handle(b, s, dev)
        for i in range(0, b, cap):
            j = min(b, i + cap)
            nz = None if noise is None else noise[:, i:j].contiguous()
            _ffi.check(_ffi.lib().prg_sampler_run(
                h, arr, len(steps), _ffi.ptr(p[i:j]),
                _ffi.ptr(ic[i:j]) if ic is not None else None,
                _ffi.ptr(nz), ctypes.c_uint64(seed + i), _ffi.ptr(out[i:j]), j - i, _ffi.stream()))
        return out

    def forward(self, *args, **kwargs):
        raise NotImplementedError("training (p_losses, SDD:1464-1510) is outside the native "
                                  "data-generation path")

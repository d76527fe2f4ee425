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

from . import _ffi, rng
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
        # SDD:1076-1094: plain float64 tensors, not buffers (they are not part of the state dict)
        self.ddnm_dropouts = torch.linspace(
            ddnm_sampling_dropout, ddnm_sampling_dropout if ddnm_dropout_schedule == 'none' else 0.,
            timesteps, dtype=torch.float64)
        self.denoise_dropouts = torch.linspace(1., 0., timesteps, dtype=torch.float64) ** 100

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

    def keep_probabilities(self, steps, is_denoise=False):
        """Per step, the dropout probability of the Bernoulli keep-mask in model_predictions, or None
        where the condition mask is used as it is: DDNM sampling draws `uniform > ddnm_dropouts[t]`
        only where that entry is positive (SDD:1213-1216); `denoise()` on a model built with
        is_ddnm_sampling=False draws `uniform > denoise_dropouts[t]` at every step (SDD:1220-1225).
        Refine steps never draw (they ban the replacement, SDD:1308-1312)."""
        main = (_ffi.STEP_P_SAMPLE, _ffi.STEP_DDIM, _ffi.STEP_DDIM_LAST)
        out = []
        for st in steps:
            p = None
            if st.kind in main:
                if self.is_ddnm_sampling:
                    if self.ddnm_dropouts[st.t] > 0:
                        p = self.ddnm_dropouts[st.t]
                elif is_denoise:
                    p = self.denoise_dropouts[st.t]
            out.append(p)
        return out

    # ------------------------------------------------------------------ sampling
    @torch.no_grad()
    def sample(self, *, param_cond, img_cond=None, disable_tqdm=False, has_refine_step=False,
               noise=None, seed=None, keep_uniform=None, _is_denoise=False):
        """(b,4) intrinsics vector [+ (b,2,s,s) DDNM image condition] -> (b,1,s,s) in [0,1].

        `noise` (optional, (num_noise_draws, b,1,s,s)) injects the Gaussian draws (parity tests);
        otherwise every image draws from its own device Philox stream: `seed` is either a sequence of
        b integers (one key per image -- `rng.scene_seed` in the generator, so an image's draws do
        not depend on its batch, rank or world size) or one integer (image i gets mix(seed, i)), or
        None (a fresh key from torch's global RNG, like the unseeded reference).
        `keep_uniform` (optional, (number of keep-mask draws, b,1,s,s)) injects the uniform draws
        of the keep-mask dropout in step order.
        """
        _ffi.require_cuda(param_cond, img_cond, noise, keep_uniform)
        b, s = param_cond.shape[0], self.image_size
        dev = param_cond.device
        steps = self.sampling_steps(has_refine_step)
        arr = (_ffi.Step * len(steps))(*steps)
        p = param_cond.float().contiguous()
        ic = None
        replace_in_loop = self.is_ddnm_sampling or _is_denoise
        if img_cond is not None and (replace_in_loop or has_refine_step):
            ic = img_cond.float().contiguous()
            assert ic.shape == (b, 2, s, s)
        probs = self.keep_probabilities(steps, _is_denoise) if ic is not None else [None] * len(steps)
        if any(q is not None for q in probs) or (ic is not None and not replace_in_loop):
            return self._sample_stepwise(steps, probs, p, ic, replace_in_loop, noise, seed, keep_uniform)
        if has_refine_step and ic is None:
            raise NotImplementedError("has_refine_step needs DDNM sampling with an image "
                                      "condition (SDD:1313)")
        if noise is not None:
            noise = noise.float().contiguous()
            need = 1 + sum(st.add_noise for st in steps)
            assert noise.shape[0] >= need and tuple(noise.shape[1:]) == (b, 1, s, s), \
                "noise must be (num_noise_draws, b, 1, s, s)"
        seeds = self._image_seeds(seed, b)
        out = torch.empty((b, self.channels, s, s), dtype=torch.float32, device=dev)
        h, cap = self.model.native_handle(b, s, dev)
        for i in range(0, b, cap):
            j = min(b, i + cap)
            nz = None if noise is None else noise[:, i:j].contiguous()
            _ffi.check(_ffi.lib().prg_sampler_run(
                h, arr, len(steps), _ffi.ptr(p[i:j]),
                _ffi.ptr(ic[i:j]) if ic is not None else None,
                _ffi.ptr(nz), _ffi.seed_array(seeds[i:j]), _ffi.ptr(out[i:j]), j - i, _ffi.stream(p)))
        return out

    @staticmethod
    def _image_seeds(seed, b):
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if isinstance(seed, int):
            return rng.image_seeds(seed, b)
        seeds = [int(v) for v in (seed.tolist() if torch.is_tensor(seed) else seed)]
        assert len(seeds) == b, "one Philox key per image"
        return seeds

    @torch.no_grad()
    def denoise(self, *, param_cond, img_cond=None, disable_tqdm=False, has_refine_step=False,
                noise=None, seed=None, keep_uniform=None):
        """SDD:1411-1427: `sample` with is_denoise=True -- on a model built with
        is_ddnm_sampling=False the condition is imposed through a keep-mask that lets almost
        nothing through at large t and everything at t = 0 (denoise_dropouts = linspace(1,0,T)^100)."""
        return self.sample(param_cond=param_cond, img_cond=img_cond, disable_tqdm=disable_tqdm,
                           has_refine_step=has_refine_step, noise=noise, seed=seed,
                           keep_uniform=keep_uniform, _is_denoise=True)

    def _sample_stepwise(self, steps, probs, p, ic, replace_in_loop, noise, seed, keep_uniform):
        """The sampling loop driven step by step from the host, for the configurations whose image
        condition changes per step (keep-mask dropout) or is only used by the refine step: every step
        is one `prg_sampler_run` call of a single step whose `noise` argument carries the current
        sample (slab 0) and that step's Gaussian draw (slab 1).  Same kernels and arithmetic as the
        fused loop; the Gaussian / uniform draws come from torch's device generator."""
        b, s = p.shape[0], self.image_size
        dev = p.device
        gen = torch.Generator(device=dev)
        gen.manual_seed(rng.mix(*self._image_seeds(seed, b)) >> 1)
        shape = (b, 1, s, s)
        main = (_ffi.STEP_P_SAMPLE, _ffi.STEP_DDIM, _ffi.STEP_DDIM_LAST)
        mask = None if ic is None else (((ic[:, 1:2] + 1) * 0.5) > 0.5)          # SDD:507-508
        k = 1
        ku = 0
        x = noise[0].float().contiguous() if noise is not None else \
            torch.randn(shape, device=dev, generator=gen)
        h, cap = self.model.native_handle(b, s, dev)
        for st, q in zip(steps, probs):
            ic_step = ic
            if st.kind in main:
                if not replace_in_loop:
                    ic_step = None                     # is_ddnm_sampling=False: no replacement here
                elif q is not None:
                    if keep_uniform is not None:
                        u = keep_uniform[ku].float()
                    else:
                        u = torch.rand(shape, device=dev, generator=gen)
                    ku += 1
                    keep = (u > q.to(torch.float32).to(dev)) & mask
                    ic_step = ic.clone()
                    ic_step[:, 1:2] = torch.where(keep, 1.0, -1.0)
            slabs = [x]
            if st.add_noise:
                slabs.append(noise[k].float() if noise is not None else
                             torch.randn(shape, device=dev, generator=gen))
                k += 1
            x = self._run_single_step(h, cap, st, p, ic_step, torch.stack(slabs).contiguous())
        return x

    def _run_single_step(self, h, cap, st, p, ic_step, buf):
        """One sampler step on the device: buf[0] = x_t, buf[1] = the step's Gaussian draw (if any)."""
        return self._run_steps(h, cap, [st], p, ic_step, buf)

    def _run_steps(self, h, cap, steps, p, ic, buf):
        """A contiguous slice of the step list, started from a given state: buf[0] = x at the entry of
        steps[0], buf[1 + i] = the i-th Gaussian draw the slice consumes."""
        b = p.shape[0]
        out = torch.empty_like(buf[0])
        arr = (_ffi.Step * len(steps))(*steps)
        for i in range(0, b, cap):
            j = min(b, i + cap)
            _ffi.check(_ffi.lib().prg_sampler_run(
                h, arr, len(steps), _ffi.ptr(p[i:j]),
                _ffi.ptr(ic[i:j].contiguous()) if ic is not None else None,
                _ffi.ptr(buf[:, i:j].contiguous()), None, _ffi.ptr(out[i:j]), j - i,
                _ffi.stream(p)))
        return out

    @torch.no_grad()
    def run_steps(self, x, first, count, *, param_cond, img_cond=None, noise, has_refine_step=False):
        """Steps [first, first + count) of `sample()`'s step list applied to the state `x` (b,1,s,s)
        with injected draws `noise` (one (b,1,s,s) slab per noisy step of the slice, in order).
        `run_steps(x_T, 0, len(steps))` equals `sample(noise=...)`: the same device loop, entered in the
        middle -- used to compare trajectories with the reference state by state."""
        _ffi.require_cuda(x, param_cond, img_cond, noise)
        steps = self.sampling_steps(has_refine_step)[first:first + count]
        need = sum(st.add_noise for st in steps)
        b, s = param_cond.shape[0], self.image_size
        assert tuple(x.shape) == (b, 1, s, s)
        assert (noise.shape[0] if noise is not None else 0) >= need, "one draw per noisy step"
        slabs = [x.float()] + ([noise[i].float() for i in range(need)] if need else [])
        ic = None
        if img_cond is not None and (self.is_ddnm_sampling or has_refine_step):
            ic = img_cond.float().contiguous()
        h, cap = self.model.native_handle(b, s, x.device)
        return self._run_steps(h, cap, steps, param_cond.float().contiguous(), ic,
                               torch.stack(slabs).contiguous())

    def forward(self, *args, **kwargs):
        raise NotImplementedError("training (p_losses, SDD:1464-1510) is outside the native "
                                  "data-generation path")

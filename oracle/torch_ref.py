"""TEST INFRASTRUCTURE ONLY -- fp32 CPU restatement of the reference networks
and sampler, written against plain state-dicts (no nn.Module tree).

It is the checker for the CUDA path (tests/, __graft_entry__.smoke(), the
cpu_baseline leg of bench.py); the product package never imports it.

Pinned against the unmodified reference modules in this container by
tests/test_oracle_vs_reference.py and against tests/golden/*.npz (minted from
the reference by oracle/make_golden.py).

Reference citations (SDD = denoising_diffusion_pytorch/successive_ddnm_diffusion.py,
DC = depth_correction_pytorch/depth_correction.py):
  ws_conv            SDD:601-616   layer_norm        SDD:619-628
  block / resnet     SDD:681-734   linear_attention  SDD:737-769
  attention          SDD:772-796   unet_forward      SDD:920-964
  schedule           SDD:997-1012, 1056-1151
  model_predictions  SDD:1182-1232 p_sample          SDD:1234-1281
  p_sample_loop      SDD:1283-1317 ddim_sample       SDD:1319-1392
  depth_augment      DC:577-604    maskunet_forward  DC:871-906

`emulate` reproduces the operand roundings of the CUDA path (fp16 tensor-core
operands / fp16 stored activations, fp32 accumulation) so that design choices
can be evaluated on the CPU before spending GPU time; it is never the parity
target (the target is emulate=None, i.e. the fp32 reference semantics).
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# numerics emulation helpers
# --------------------------------------------------------------------------- #
class Emu:
    """Which tensors get rounded to fp16 in the CUDA path."""

    def __init__(self, operands=True, raw=True):
        self.operands = operands   # conv / gemm A and B operands
        self.raw = raw             # raw conv output stored before GroupNorm


def _h(x, on):
    return x.half().float() if on else x


def _conv(x, w, b, emu, stride=1, padding=0):
    if emu is not None and emu.operands:
        x = _h(x, True)
        w = _h(w, True)
    return F.conv2d(x, w, b, stride=stride, padding=padding)


# --------------------------------------------------------------------------- #
# building blocks
# --------------------------------------------------------------------------- #
def standardize_weight(w, eps=1e-5):
    """SDD:606-613 (fp32 branch: eps = 1e-5)."""
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return (w - mean) * (var + eps).rsqrt()


def layer_norm(x, g, eps=1e-5):
    """SDD:624-628: channel LayerNorm with gain only."""
    var = torch.var(x, dim=1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) * (var + eps).rsqrt() * g


def block(sd, pfx, x, scale_shift, emu, groups=8):
    """SDD:688-697."""
    w = standardize_weight(sd[pfx + ".proj.weight"])
    x = _conv(x, w, sd[pfx + ".proj.bias"], emu, padding=1)
    if emu is not None and emu.raw:
        # the CUDA path computes GroupNorm statistics from the fp32
        # accumulators but stores the raw tensor in fp16
        xs = x
        n, c = x.shape[:2]
        xg = xs.reshape(n, groups, -1)
        mean = xg.mean(-1, keepdim=True)
        var = xg.var(-1, unbiased=False, keepdim=True)
        xq = _h(x, True).reshape(n, groups, -1)
        x = ((xq - mean) * (var + 1e-5).rsqrt()).reshape(x.shape)
        x = x * sd[pfx + ".norm.weight"][None, :, None, None] + \
            sd[pfx + ".norm.bias"][None, :, None, None]
    else:
        x = F.group_norm(x, groups, sd[pfx + ".norm.weight"],
                         sd[pfx + ".norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(sd, pfx, x, cond, emu, groups=8):
    """SDD:720-734 (cond = cat(time_emb, param_emb) or None for DC:734-740)."""
    scale_shift = None
    if cond is not None and (pfx + ".mlp.1.weight") in sd:
        e = F.linear(F.silu(cond), sd[pfx + ".mlp.1.weight"],
                     sd[pfx + ".mlp.1.bias"])
        e = e[:, :, None, None]
        scale_shift = e.chunk(2, dim=1)
    h = block(sd, pfx + ".block1", x, scale_shift, emu, groups)
    h = block(sd, pfx + ".block2", h, None, emu, groups)
    if (pfx + ".res_conv.weight") in sd:
        res = _conv(x, sd[pfx + ".res_conv.weight"],
                    sd[pfx + ".res_conv.bias"], emu)
    else:
        res = x
    return h + res


def linear_attention(sd, pfx, x, emu, heads=4, dim_head=32):
    """Residual(PreNorm(LinearAttention)) -- SDD:583-589, 631-639, 748-769."""
    b, c, h, w = x.shape
    n = h * w
    xn = layer_norm(x, sd[pfx + ".fn.norm.g"])
    qkv = _conv(xn, sd[pfx + ".fn.fn.to_qkv.weight"], None, emu)
    q, k, v = [t.reshape(b, heads, dim_head, n) for t in qkv.chunk(3, dim=1)]
    q = q.softmax(dim=-2) * dim_head ** -0.5
    k = k.softmax(dim=-1)
    v = v / n
    context = torch.einsum('bhdn,bhen->bhde', k, v)
    out = torch.einsum('bhde,bhdn->bhen', context, q)
    out = out.reshape(b, heads * dim_head, h, w)
    out = _conv(out, sd[pfx + ".fn.fn.to_out.0.weight"],
                sd[pfx + ".fn.fn.to_out.0.bias"], emu)
    out = layer_norm(out, sd[pfx + ".fn.fn.to_out.1.g"])
    return out + x


def attention(sd, pfx, x, emu, heads=4, dim_head=32):
    """Residual(PreNorm(Attention)) -- SDD:782-796."""
    b, c, h, w = x.shape
    n = h * w
    xn = layer_norm(x, sd[pfx + ".fn.norm.g"])
    qkv = _conv(xn, sd[pfx + ".fn.fn.to_qkv.weight"], None, emu)
    q, k, v = [t.reshape(b, heads, dim_head, n) for t in qkv.chunk(3, dim=1)]
    q = q * dim_head ** -0.5
    sim = torch.einsum('bhdi,bhdj->bhij', q, k)
    attn = sim.softmax(dim=-1)
    out = torch.einsum('bhij,bhdj->bhid', attn, v)
    out = out.permute(0, 1, 3, 2).reshape(b, heads * dim_head, h, w)
    out = _conv(out, sd[pfx + ".fn.fn.to_out.weight"],
                sd[pfx + ".fn.fn.to_out.bias"], emu)
    return out + x


def sinusoidal_pos_emb(t, dim):
    """SDD:650-657."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, device=t.device) * -e)
    e = t[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def _num_levels(sd):
    n = 0
    while ("downs.%d.0.block1.proj.weight" % n) in sd:
        n += 1
    return n


def _trunk(sd, x, cond, emu, groups):
    """Shared encoder/decoder skeleton of Unet (SDD:929-964) and MaskUnet
    (DC:874-906): x is the stem output."""
    r = x.clone()
    hs = []
    levels = _num_levels(sd)
    for i in range(levels):
        p = "downs.%d" % i
        x = resnet_block(sd, p + ".0", x, cond, emu, groups)
        hs.append(x)
        x = resnet_block(sd, p + ".1", x, cond, emu, groups)
        x = linear_attention(sd, p + ".2", x, emu)
        hs.append(x)
        w = sd[p + ".3.weight"]
        if w.shape[-1] == 4:      # Downsample: conv 4x4 s2 p1 (SDD:597-598)
            x = _conv(x, w, sd[p + ".3.bias"], emu, stride=2, padding=1)
        else:                     # last level: conv 3x3 p1 (SDD:878-879)
            x = _conv(x, w, sd[p + ".3.bias"], emu, padding=1)
    x = resnet_block(sd, "mid_block1", x, cond, emu, groups)
    x = attention(sd, "mid_attn", x, emu)
    x = resnet_block(sd, "mid_block2", x, cond, emu, groups)
    for i in range(levels):
        p = "ups.%d" % i
        x = torch.cat((x, hs.pop()), dim=1)
        x = resnet_block(sd, p + ".0", x, cond, emu, groups)
        x = torch.cat((x, hs.pop()), dim=1)
        x = resnet_block(sd, p + ".1", x, cond, emu, groups)
        x = linear_attention(sd, p + ".2", x, emu)
        if (p + ".3.1.weight") in sd:   # Upsample: nearest x2 + conv3x3 (SDD:592-594)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = _conv(x, sd[p + ".3.1.weight"], sd[p + ".3.1.bias"], emu, padding=1)
        else:
            x = _conv(x, sd[p + ".3.weight"], sd[p + ".3.bias"], emu, padding=1)
    x = torch.cat((x, r), dim=1)
    x = resnet_block(sd, "final_res_block", x, cond, emu, groups)
    return x


@torch.no_grad()
def unet_forward(sd, x, time, param_cond, emu=None, groups=8):
    """Unet.forward -- SDD:920-964.  sd: state-dict of the Unet (no prefix)."""
    x = x.float()
    p = F.linear(param_cond.float(), sd["param_mlp.0.weight"], sd["param_mlp.0.bias"])
    p = F.linear(F.gelu(p), sd["param_mlp.2.weight"], sd["param_mlp.2.bias"])
    x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
    dim = sd["time_mlp.1.weight"].shape[1]
    t = sinusoidal_pos_emb(time.float(), dim)
    t = F.linear(t, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"])
    t = F.linear(F.gelu(t), sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])
    cond = torch.cat((t, p), dim=-1)
    x = _trunk(sd, x, cond, emu, groups)
    return F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"])


def depth_augment(depth):
    """DepthAugment.forward -- DC:582-604 (invalid_number = 0)."""
    cln = depth.clone()
    cln[cln == 0] = float("inf")
    mn = -F.max_pool2d(-cln, kernel_size=3, stride=1, padding=1)
    mn0 = -F.max_pool2d(-depth, kernel_size=3, stride=1, padding=1)
    mn = torch.where(mn.isinf(), mn0, mn)
    return torch.cat([depth, mn, mn - depth], dim=-3)


@torch.no_grad()
def maskunet_forward(sd, x, emu=None, groups=8):
    """MaskUnet.forward -- DC:871-906.  Returns the sigmoid keep-probability."""
    x = depth_augment(x.float())
    x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
    x = _trunk(sd, x, None, emu, groups)
    x = F.conv2d(x, sd["final_conv.0.weight"], sd["final_conv.0.bias"])
    return torch.sigmoid(x)


# --------------------------------------------------------------------------- #
# schedule + sampler
# --------------------------------------------------------------------------- #
def sigmoid_beta_schedule(timesteps, start=-3, end=3, tau=1):
    """SDD:997-1012 (float64)."""
    steps = timesteps + 1
    t = torch.linspace(0, timesteps, steps, dtype=torch.float64) / timesteps
    v_start = torch.tensor(start / tau).sigmoid()
    v_end = torch.tensor(end / tau).sigmoid()
    ac = (-((t * (end - start) + start) / tau).sigmoid() + v_end) / (v_end - v_start)
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clip(betas, 0, 0.999)


def linear_beta_schedule(timesteps):
    """SDD:976-980."""
    scale = 1000 / timesteps
    return torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)


def cosine_beta_schedule(timesteps, s=0.008):
    """SDD:983-994."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return torch.clip(betas, 0, 0.999)


def make_schedule(timesteps, beta_schedule="sigmoid"):
    """The fp32 buffers GaussianDiffusion registers -- SDD:1047-1134."""
    betas = {"sigmoid": sigmoid_beta_schedule, "linear": linear_beta_schedule,
             "cosine": cosine_beta_schedule}[beta_schedule](timesteps)
    alphas = 1. - betas
    ac = torch.cumprod(alphas, dim=0)
    ac_prev = F.pad(ac[:-1], (1, 0), value=1.)
    pv = betas * (1. - ac_prev) / (1. - ac)
    f = lambda v: v.to(torch.float32)
    return dict(
        betas=f(betas), alphas_cumprod=f(ac), alphas_cumprod_prev=f(ac_prev),
        sqrt_alphas_cumprod=f(torch.sqrt(ac)),
        sqrt_one_minus_alphas_cumprod=f(torch.sqrt(1. - ac)),
        log_one_minus_alphas_cumprod=f(torch.log(1. - ac)),
        sqrt_recip_alphas_cumprod=f(torch.sqrt(1. / ac)),
        sqrt_recipm1_alphas_cumprod=f(torch.sqrt(1. / ac - 1)),
        posterior_variance=f(pv),
        posterior_log_variance_clipped=f(torch.log(pv.clamp(min=1e-20))),
        posterior_mean_coef1=f(betas * torch.sqrt(ac_prev) / (1. - ac)),
        posterior_mean_coef2=f((1. - ac_prev) * torch.sqrt(alphas) / (1. - ac)),
    )


def dropout_tables(timesteps, ddnm_sampling_dropout=0., ddnm_dropout_schedule='none'):
    """(ddnm_dropouts, denoise_dropouts), float64 -- SDD:1076-1094."""
    end = ddnm_sampling_dropout if ddnm_dropout_schedule == 'none' else 0.
    ddnm = torch.linspace(ddnm_sampling_dropout, end, timesteps, dtype=torch.float64)
    denoise = torch.linspace(1., 0., timesteps, dtype=torch.float64) ** 100
    return ddnm, denoise


class KeepMask:
    """The Bernoulli keep-mask of model_predictions (SDD:1213-1216 DDNM dropout, SDD:1220-1225
    denoise): `uniform_(0, 1) > table[t]`, with the uniform draws injected (`draws` is consumed in
    step order).  mode 'ddnm' draws only where table[t] > 0; mode 'denoise' always."""

    def __init__(self, mode, table, draws):
        assert mode in ('ddnm', 'denoise')
        self.mode, self.table, self.draws = mode, table, iter(draws)

    def __call__(self, t, mask):
        if self.mode == 'ddnm' and not (self.table[t] > 0):
            return mask
        return (next(self.draws) > self.table[t]) & mask


def _predictions(sd, sch, x, t, param_cond, img_cond, clip_x_start, ban_ddnm, emu, keep=None):
    """model_predictions for objective='pred_x0' -- SDD:1182-1232.  `keep` (a KeepMask) adds the
    keep-mask dropout of the DDNM / denoise branches."""
    b = x.shape[0]
    tt = torch.full((b,), t, dtype=torch.long)
    x0 = unet_forward(sd, x, tt, param_cond, emu)
    if clip_x_start:
        x0 = x0.clamp(-1., 1.)
    pred_noise = (sch["sqrt_recip_alphas_cumprod"][t] * x - x0) / \
        sch["sqrt_recipm1_alphas_cumprod"][t]
    if img_cond is not None and not ban_ddnm:
        rpj = img_cond[:, 0:1]
        mask = ((img_cond[:, 1:2] + 1) * 0.5) > 0.5
        if keep is not None:
            mask = keep(t, mask)
        x0 = torch.where(mask, rpj, x0)
    return pred_noise, x0


@torch.no_grad()
def p_sample_step(sd, sch, img, t, param_cond, img_cond, noise, emu=None, keep=None):
    """One iteration of the p_sample_loop body (p_sample -> p_mean_variance -> q_posterior,
    SDD:1257-1281, 1234-1255, 1173-1180): x_t -> x_{t-1}; `noise` is the randn_like draw (unused at t = 0)."""
    _, x0 = _predictions(sd, sch, img, t, param_cond, img_cond, False, False, emu, keep)
    x0 = x0.clamp(-1., 1.)
    mean = sch["posterior_mean_coef1"][t] * x0 + sch["posterior_mean_coef2"][t] * img
    logvar = sch["posterior_log_variance_clipped"][t]
    if t > 0:
        return mean + (0.5 * logvar).exp() * noise
    return mean + (0.5 * logvar).exp() * 0.


@torch.no_grad()
def p_sample_loop(sd, sch, param_cond, img_cond, noises, has_refine_step=False,
                  emu=None, trajectory=None, keep=None):
    """SDD:1283-1317 with torch.randn replaced by the injected `noises`
    (noises[0] = x_T, noises[1 + i] = the i-th randn_like draw)."""
    T = sch["betas"].shape[0]
    img = noises[0].clone()
    k = 1
    for t in reversed(range(T)):
        noise = None
        if t > 0:
            noise = noises[k]
            k += 1
        img = p_sample_step(sd, sch, img, t, param_cond, img_cond, noise, emu, keep)
        if trajectory is not None:
            trajectory.append(img.clone())
    if has_refine_step:
        _, x0 = _predictions(sd, sch, img, 0, param_cond, img_cond, False, True, emu)
        x0 = x0.clamp(-1., 1.)
        mean = sch["posterior_mean_coef1"][0] * x0 + sch["posterior_mean_coef2"][0] * img
        mask = ((img_cond[:, 1:2] + 1) * 0.5) > 0.5
        img = torch.where(mask, mean, img)
    return (img + 1) * 0.5


def ddim_times(total_timesteps, sampling_timesteps):
    """SDD:1331-1337."""
    times = torch.linspace(-1, total_timesteps - 1, steps=sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


@torch.no_grad()
def ddim_step(sd, sch, img, t, t_next, param_cond, img_cond, noise, eta=1.0, emu=None, keep=None):
    """One iteration of the ddim_sample body (SDD:1343-1373): x_t -> x_{t_next}."""
    ac = sch["alphas_cumprod"]
    pred_noise, x0 = _predictions(sd, sch, img, t, param_cond, img_cond, True, False, emu, keep)
    if t_next < 0:
        return x0
    alpha, alpha_next = ac[t], ac[t_next]
    sigma = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
    c = (1 - alpha_next - sigma ** 2).sqrt()
    return x0 * alpha_next.sqrt() + c * pred_noise + sigma * noise


@torch.no_grad()
def ddim_sample(sd, sch, param_cond, img_cond, noises, sampling_timesteps, eta=1.0,
                has_refine_step=False, emu=None, trajectory=None, keep=None):
    """SDD:1319-1392 with injected noise (same convention as p_sample_loop)."""
    T = sch["betas"].shape[0]
    img = noises[0].clone()
    k = 1
    for t, t_next in ddim_times(T, sampling_timesteps):
        noise = None
        if t_next >= 0:
            noise = noises[k]
            k += 1
        img = ddim_step(sd, sch, img, t, t_next, param_cond, img_cond, noise, eta, emu, keep)
        if trajectory is not None:
            trajectory.append(img.clone())
    if has_refine_step:
        _, x0 = _predictions(sd, sch, img, 0, param_cond, img_cond, True, True, emu)
        mask = ((img_cond[:, 1:2] + 1) * 0.5) > 0.5
        img = torch.where(mask, x0, img)
    return (img + 1) * 0.5

"""`Unet` and `MaskUnet` under the reference's names, constructor arguments and state-dict
layout (SDD:802-918, DC:807-869), with `forward` running on libprg.so.

The module tree below exists only to own parameters with the reference's key names (so
reference checkpoints `load_state_dict` unchanged) and its default initialisation order (so
`torch.manual_seed(s)` + construction gives the same random weights as the reference).  No
torch op takes part in `forward`: parameters are packed once (re-packed when they change) and
handed to `prg_net_create`; `forward` forwards device pointers.
"""
import copy
import ctypes

import torch
from torch import nn

from . import _ffi, packing


class _Node(nn.Module):
    """Parameter container; children are attached by attribute assignment."""

    def forward(self, *a, **k):  # pragma: no cover - containers are never called
        raise RuntimeError("container module; call the owning network")


class _Gain(nn.Module):
    """Channel LayerNorm gain `g` (SDD:619-628)."""

    def __init__(self, dim):
        super().__init__()
        self.g = nn.Parameter(torch.ones(1, dim, 1, 1))


def _block(cin, cout, groups):
    b = _Node()
    b.proj = nn.Conv2d(cin, cout, 3, padding=1)       # weight-standardised at pack time
    b.norm = nn.GroupNorm(groups, cout)
    return b


def _resblock(cin, cout, cond_dim, groups):
    r = _Node()
    if cond_dim:
        r.mlp = nn.Sequential(nn.SiLU(), nn.Linear(cond_dim, cout * 2))
    r.block1 = _block(cin, cout, groups)
    r.block2 = _block(cout, cout, groups)
    r.res_conv = nn.Conv2d(cin, cout, 1) if cin != cout else nn.Identity()
    return r


def _attn(dim, linear, heads=4, dim_head=32):
    hidden = heads * dim_head
    inner = _Node()
    inner.to_qkv = nn.Conv2d(dim, hidden * 3, 1, bias=False)
    inner.to_out = nn.Sequential(nn.Conv2d(hidden, dim, 1), _Gain(dim)) if linear \
        else nn.Conv2d(hidden, dim, 1)
    pre = _Node()
    pre.fn = inner
    pre.norm = _Gain(dim)
    res = _Node()
    res.fn = pre
    return res


def _build_trunk(net, dim, init_dim, dim_mults, cond_dim, groups):
    dims = [init_dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    net.downs = nn.ModuleList([])
    net.ups = nn.ModuleList([])
    n = len(in_out)
    for i, (cin, cout) in enumerate(in_out):
        last = i >= n - 1
        net.downs.append(nn.ModuleList([
            _resblock(cin, cin, cond_dim, groups),
            _resblock(cin, cin, cond_dim, groups),
            _attn(cin, linear=True),
            nn.Conv2d(cin, cout, 4, 2, 1) if not last else nn.Conv2d(cin, cout, 3, padding=1)]))
    mid = dims[-1]
    net.mid_block1 = _resblock(mid, mid, cond_dim, groups)
    net.mid_attn = _attn(mid, linear=False)
    net.mid_block2 = _resblock(mid, mid, cond_dim, groups)
    for i, (cin, cout) in enumerate(reversed(in_out)):
        last = i == n - 1
        net.ups.append(nn.ModuleList([
            _resblock(cout + cin, cout, cond_dim, groups),
            _resblock(cout + cin, cout, cond_dim, groups),
            _attn(cout, linear=True),
            nn.Sequential(nn.Upsample(scale_factor=2, mode="nearest"),
                          nn.Conv2d(cout, cin, 3, padding=1)) if not last
            else nn.Conv2d(cout, cin, 3, padding=1)]))


class _NativeNet(nn.Module):
    """Shared handle management: pack -> prg_net_create, cached per (weights, batch, size)."""
    _kind = None

    def _init_native(self):
        self._handle = None
        self._handle_key = None
        self._blob = None
        self._blob_sig = None
        self.max_batch = None      # optional cap on the workspace batch (micro-batching)

    def _pack(self):
        raise NotImplementedError

    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _ensure(self, batch, size, device):
        sig = self._signature()
        if self._blob is None or sig != self._blob_sig:
            self._blob = self._pack()
            self._blob_sig = sig
            self._release()
        dev = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is not None:
            _, hb, hs, hd = self._handle_key
            if hs == size and hd == dev and hb >= batch:
                return self._handle
            self._release()
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(self._blob, len(self._blob))
        _ffi.check(_ffi.lib().prg_net_create(ctypes.byref(h), self._kind, buf, len(self._blob),
                                             int(batch), int(size), int(dev)))
        self._handle = h
        self._handle_key = (self._kind, int(batch), int(size), dev)
        return h

    _NATIVE_STATE = ("_handle", "_handle_key", "_blob", "_blob_sig")

    def invalidate(self):
        """Forget the packed blob and the device handle (they are rebuilt from the parameters on the
        next call).  Needed after parameters were written through `.data` (no version bump)."""
        self._release()
        self._blob = None
        self._blob_sig = None

    def __deepcopy__(self, memo):
        """A copy owns its parameters but not the native handle (a ctypes pointer cannot be copied
        or pickled): it packs and creates its own on first use (ema_pytorch-style `deepcopy(model)`)."""
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in self._NATIVE_STATE else copy.deepcopy(v, memo)
        return new

    def __getstate__(self):
        d = dict(self.__dict__)
        for k in self._NATIVE_STATE:
            d[k] = None
        return d

    def _release(self):
        if getattr(self, "_handle", None) is not None:
            _ffi.lib().prg_net_destroy(self._handle)
            self._handle = None
            self._handle_key = None

    def native_handle(self, batch, size, device):
        """(prg_net*, workspace batch) for up to `batch` images of `size` x `size`."""
        cap = self.max_batch or batch
        h = self._ensure(min(batch, cap), size, device)
        return h, self._handle_key[1]

    def device_bytes(self):
        return 0 if self._handle is None else int(_ffi.lib().prg_net_device_bytes(self._handle))

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


class Unet(_NativeNet):
    """Conditional U-Net denoiser (SDD:802-964)."""
    _kind = packing.KIND_UNET

    def __init__(self, dim, param_cond_dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8),
                 channels=1, resnet_block_groups=8, learned_variance=False,
                 learned_sinusoidal_cond=False, random_fourier_features=False,
                 learned_sinusoidal_dim=16):
        super().__init__()
        if learned_variance or learned_sinusoidal_cond or random_fourier_features:
            raise NotImplementedError("learned variance / learned sinusoidal embeddings are not on "
                                      "the data-generation path (GaussianDiffusion rejects them, "
                                      "SDD:1032-1034)")
        if channels != 1 or (init_dim not in (None, dim)) or (out_dim not in (None, 1)):
            raise NotImplementedError("the native kernels cover the shipped configuration: "
                                      "channels=1, init_dim=dim, out_dim=1")
        self.channels = channels
        self.out_dim = 1
        self.random_or_learned_sinusoidal_cond = False
        self.param_cond_dim = param_cond_dim
        self.resnet_block_groups = resnet_block_groups
        self.init_conv = nn.Conv2d(channels, dim, 7, padding=3)
        time_dim = dim * 4
        # index 0 of time_mlp is the parameter-free sinusoidal embedding (SDD:645-657)
        self.time_mlp = nn.Sequential(nn.Identity(), nn.Linear(dim, time_dim), nn.GELU(),
                                      nn.Linear(time_dim, time_dim))
        self.param_mlp = nn.Sequential(nn.Linear(param_cond_dim, time_dim), nn.GELU(),
                                       nn.Linear(time_dim, time_dim))
        _build_trunk(self, dim, dim, dim_mults, 2 * time_dim, resnet_block_groups)
        self.final_res_block = _resblock(dim * 2, dim, 2 * time_dim, resnet_block_groups)
        self.final_conv = nn.Conv2d(dim, 1, 1)
        self._init_native()

    def _pack(self):
        return packing.pack_unet(self.state_dict(), self.resnet_block_groups)

    @torch.no_grad()
    def forward(self, x, time, param_cond, img_cond=None):
        """x (b,1,s,s), time (b,) long, param_cond (b,4) -> (b,1,s,s); `img_cond` is accepted and
        ignored exactly like the reference network (SDD:920)."""
        _ffi.require_cuda(x, time, param_cond)
        b, c, s, s2 = x.shape
        assert c == 1 and s == s2
        x = x.float().contiguous()
        t = time.to(torch.int64).contiguous()
        p = param_cond.float().contiguous()
        out = torch.empty_like(x)
        h, cap = self.native_handle(b, s, x.device)
        for i in range(0, b, cap):
            j = min(b, i + cap)
            _ffi.check(_ffi.lib().prg_unet_forward(h, _ffi.ptr(x[i:j]), _ffi.ptr(t[i:j]),
                                                   _ffi.ptr(p[i:j]), _ffi.ptr(out[i:j]), j - i,
                                                   _ffi.stream(x)))
        return out


class MaskUnet(_NativeNet):
    """Depth-correction U-Net (DC:807-906): returns the sigmoid keep-probability."""
    _kind = packing.KIND_MASKUNET

    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8),
                 resnet_block_groups=8, learned_variance=False):
        super().__init__()
        if (init_dim not in (None, dim)) or (out_dim not in (None, 1)):
            raise NotImplementedError("the native kernels cover init_dim=dim, out_dim=1")
        self.out_dim = 1
        self.resnet_block_groups = resnet_block_groups
        self.init_aug = nn.Identity()        # DepthAugment has no parameters (fused into the stem)
        self.init_conv = nn.Conv2d(3, dim, 7, padding=3)
        _build_trunk(self, dim, dim, dim_mults, 0, resnet_block_groups)
        self.final_res_block = _resblock(dim * 2, dim, 0, resnet_block_groups)
        self.final_conv = nn.Sequential(nn.Conv2d(dim, 1, 1), nn.Sigmoid())
        self._init_native()

    def _pack(self):
        return packing.pack_maskunet(self.state_dict(), self.resnet_block_groups)

    def _run(self, x, want_prob, thresh):
        _ffi.require_cuda(x)
        b, c, s, s2 = x.shape
        assert c == 1 and s == s2
        x = x.float().contiguous()
        prob = torch.empty_like(x) if want_prob else None
        keep = None if thresh is None else torch.empty(x.shape, dtype=torch.uint8, device=x.device)
        h, cap = self.native_handle(b, s, x.device)
        for i in range(0, b, cap):
            j = min(b, i + cap)
            _ffi.check(_ffi.lib().prg_maskunet_forward(
                h, _ffi.ptr(x[i:j]), _ffi.ptr(prob[i:j]) if want_prob else None,
                _ffi.ptr(keep[i:j]) if keep is not None else None,
                float(thresh if thresh is not None else 0.0), j - i, _ffi.stream(x)))
        return prob, keep

    @torch.no_grad()
    def forward(self, x):
        return self._run(x, True, None)[0]

    @torch.no_grad()
    def keep_mask(self, x, thresh=0.99):
        """`MaskUnet(x) > thresh` (SDD:2564-2565) without materialising the probabilities."""
        return self._run(x, False, thresh)[1].view(torch.bool)

"""tcgen05 implicit-GEMM engine vs a plain fp32 torch convolution of the same fp16 operands."""
import pytest
import torch
import torch.nn.functional as F

from pointreggpt_b200 import _ffi, packing

pytestmark = pytest.mark.gpu


def _run(x_nchw, w, bias, mode):
    B, Cin, H, W = x_nchw.shape
    Cout = w.shape[0]
    x = x_nchw.permute(0, 2, 3, 1).contiguous().half().cuda()
    if mode == 3:
        wk = packing.upsample_fold_weight(w)
        Ho, Wo = H * 2, W * 2
    else:
        wk = packing.conv_weight_kmajor(w)
        Ho, Wo = (H // 2, W // 2) if mode == 2 else (H, W)
    wk = wk.half().cuda()
    y = torch.full((B, Ho, Wo, Cout), float("nan"), dtype=torch.float16, device="cuda")
    b = None if bias is None else bias.float().cuda()
    _ffi.check(_ffi.lib().prg_test_conv_f16(_ffi.ptr(x), _ffi.ptr(wk), _ffi.ptr(b), _ffi.ptr(y), B,
                                            H, W, Cin, Cout, mode, _ffi.stream()))
    torch.cuda.synchronize()
    return y.float().cpu().permute(0, 3, 1, 2)


def _ref(x, w, bias, mode):
    x = x.half().float()
    w = w.half().float()
    if mode == 0:
        return F.conv2d(x, w, bias)
    if mode == 1:
        return F.conv2d(x, w, bias, padding=1)
    if mode == 2:
        return F.conv2d(x, w, bias, stride=2, padding=1)
    return F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, bias, padding=1)


CASES = [
    # mode, B, H, W, Cin, Cout
    (0, 1, 16, 16, 64, 64),
    (1, 1, 16, 16, 64, 64),
    (1, 2, 32, 32, 64, 64),
    (1, 1, 64, 64, 128, 128),
    (1, 2, 32, 32, 192, 256),
    (1, 1, 32, 32, 256, 512),
    (0, 2, 32, 32, 128, 384),
    (0, 1, 128, 128, 64, 64),
    (1, 1, 256, 256, 64, 64),
    (2, 2, 32, 32, 64, 128),
    (2, 1, 64, 64, 128, 256),
    (2, 1, 256, 256, 64, 64),
    (3, 2, 16, 16, 128, 64),
    (3, 1, 32, 32, 256, 128),
]


@pytest.mark.parametrize("case", CASES)
def test_conv_matches_fp32_reference(case):
    mode, B, H, W, Cin, Cout = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    k = {0: 1, 1: 3, 2: 4, 3: 3}[mode]
    x = torch.randn(B, Cin, H, W, generator=g)
    # multiples of 1/16 so that pre-summed taps (mode 3) stay exact in fp16
    w = torch.randint(-8, 9, (Cout, Cin, k, k), generator=g).float() / 16.0
    bias = torch.randn(Cout, generator=g)
    got = _run(x, w, bias, mode)
    ref = _ref(x, w, bias, mode)
    assert got.shape == ref.shape
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    # fp32 accumulation of exact fp16 products; the only rounding is the fp16 store
    assert err <= scale * 1.5e-3, (err, scale)
    rel = ((got - ref).norm() / ref.norm()).item()
    assert rel < 5e-4, rel


def test_conv_no_bias():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 64, 16, 16, generator=g)
    w = torch.randn(128, 64, 3, 3, generator=g) * 0.05
    got = _run(x, w, None, 1)
    ref = _ref(x, w, None, 1)
    assert ((got - ref).norm() / ref.norm()).item() < 5e-4


def test_conv_halo_ring_many_rows():
    """3x3, Cin = Cout = 64 on >=128-wide maps takes the halo-ring path (each input row loaded
    once, nine shifted smem views); check several images / strips / segment boundaries."""
    g = torch.Generator().manual_seed(3)
    for (B, H, W) in [(3, 128, 128), (2, 256, 256), (1, 40, 128)]:
        x = torch.randn(B, 64, H, W, generator=g)
        w = torch.randn(64, 64, 3, 3, generator=g) * 0.06
        bias = torch.randn(64, generator=g)
        got = _run(x, w, bias, 1)
        ref = _ref(x, w, bias, 1)
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel < 5e-4, (B, H, W, rel)

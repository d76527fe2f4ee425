"""Micro-benchmark of one convolution through the tcgen05 engine (C-ABI test hook).

    python tools/bench_conv.py MODE B H W CIN COUT [iters]
MODE: 0 = 1x1, 1 = 3x3, 2 = 4x4 stride 2, 3 = nearest-x2 upsample + 3x3 (folded)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointreggpt_b200 import _ffi

mode, B, H, W, Cin, Cout = [int(v) for v in sys.argv[1:7]]
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 20
noflush = os.environ.get('NOFLUSH') is not None
taps = {0: 1, 1: 9, 2: 16, 3: 16}[mode]
Ho, Wo = {0: (H, W), 1: (H, W), 2: (H // 2, W // 2), 3: (2 * H, 2 * W)}[mode]
x = torch.randn(B, H, W, Cin, device="cuda").half()
w = (torch.randn(Cout, taps * Cin, device="cuda") * 0.05).half()
bias = torch.randn(Cout, device="cuda")
y = torch.empty(B, Ho, Wo, Cout, device="cuda", dtype=torch.float16)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def run():
    _ffi.check(_ffi.lib().prg_test_conv_f16(_ffi.ptr(x), _ffi.ptr(w), _ffi.ptr(bias), _ffi.ptr(y), B, H,
                                            W, Cin, Cout, mode, _ffi.stream()))


for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    if not noflush:
        flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
med = ts[len(ts) // 2]
eff_taps = {0: 1, 1: 9, 2: 16, 3: 9}[mode]   # algorithmic taps of the reference op
flops = 2.0 * B * Ho * Wo * Cout * eff_taps * Cin
print("mode %d B%d %dx%d %d->%d : %.1f us  %.0f TFLOP/s (algorithmic)  env=%s noflush=%s" %
      (mode, B, H, W, Cin, Cout, med * 1e3, flops / (med * 1e-3) / 1e12,
       os.environ.get("PRG_HALO_RSEG", ""), noflush))

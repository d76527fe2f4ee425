"""Tensor-pipe rate probe: SM cycles per tcgen05.mma (M=128, N, K=16) and the implied TFLOP/s.

    python tools/mma_rate.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointreggpt_b200 import _ffi

out = torch.zeros(148, dtype=torch.int64, device="cuda")
iters = 40000
for grid in (1, 148):
    for n in (64, 128, 192, 256):
        for shift in (0, 128):
            for same in (0, 1):
                for rep in range(2):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _ffi.check(_ffi.lib().prg_test_mma_rate(grid, n, iters, shift, same, _ffi.ptr(out), _ffi.stream()))
                    e1.record()
                    torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                cyc = out[:grid].double().mean().item() / iters
                flops = 2.0 * 128 * n * 16 * iters * grid
                print("grid %3d N %3d a_shift %3d same_ab %d: %.1f cyc/MMA (ideal %d)  %.3f ms  clock ~%.2f GHz  %.0f TFLOP/s"
                      % (grid, n, shift, same, cyc, n // 2, ms, cyc * iters / (ms * 1e6), flops / (ms * 1e-3) / 1e12))

"""Experiment: does running two half-batch samplers concurrently (two handles, two streams, two host
threads) beat one full-batch sampler?  Independent launch sequences let the tail of one kernel overlap
the head of the next and HBM-bound kernels overlap tensor-bound ones.

    python tools/dual_stream.py [--batch 32] [--steps 40]
"""
import argparse
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointreggpt_b200 import nets
from pointreggpt_b200.diffusion import GaussianDiffusion

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=40)
a = ap.parse_args()
dev = torch.device("cuda", 0)


def make():
    torch.manual_seed(0)
    u = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    return GaussianDiffusion(u, image_size=256, timesteps=a.steps, objective="pred_x0", beta_schedule="sigmoid").to(dev)


pc = torch.tensor([[303.9, 304.2, 128.5, 128.0]], device=dev)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


full = make()
t_full = timed(lambda: full.sample(param_cond=pc.repeat(a.batch, 1), seed=1))
print("one stream, B=%d: %.2f ms per step" % (a.batch, 1e3 * t_full / a.steps))
for parts in (2, 4):
    ds = [make() for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    hb = a.batch // parts

    def run_parts():
        def work(i):
            with torch.cuda.stream(streams[i]):
                ds[i].sample(param_cond=pc.repeat(hb, 1), seed=1 + i)
        ts = [threading.Thread(target=work, args=(i,)) for i in range(parts)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()

    t_p = timed(run_parts)
    print("%d streams, B=%d each: %.2f ms per step of the whole batch (%.3fx)" % (parts, hb, 1e3 * t_p / a.steps, t_full / t_p))
    del ds
    torch.cuda.empty_cache()

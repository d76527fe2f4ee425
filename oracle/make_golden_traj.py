"""Mints tests/golden/traj_*.npz: x_t snapshots of the UNMODIFIED reference's sampling loops at the
configurations that are benched and shipped (TEST INFRASTRUCTURE; build container only):

    python -m oracle.make_golden_traj p_sample     # 1000-step p_sample_loop (SDD:1283-1317), ~25 min on 8 cores
    python -m oracle.make_golden_traj ddim         # 250-step ddim_sample eta=1 (SDD:1319-1392, GD:34-39), ~6 min

B = 1, 256 x 256, random-init U-Net (`torch.manual_seed(0)`), DDNM image condition from the reference's
own `image_condition` of a synthetic frame.  The Gaussian draws are injected: draw k of a run is
`torch.randn(shape, generator=Generator().manual_seed(NOISE_SEED + k))` (k = 0 is x_T), so a test can
regenerate any single draw.  Stored: the state that ENTERS the listed steps (step index i counts U-Net
evaluations from 0), the final output, and the inputs' fingerprints.  The GPU tests
(tests/test_traj_gpu.py) (a) run the CUDA sampler free from x_T and compare at every snapshot and at
the end, (b) start one step from each reference snapshot and compare with the oracle's step.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.make_golden import fingerprint, sha  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402
from pointreggpt_b200 import synthetic as S  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
NOISE_SEED = 424200
SIZE = 256
# step indices (0 = the first U-Net evaluation, at t = T-1) whose INPUT state is stored
SNAP_P = [1, 100, 250, 500, 750, 900, 990, 999]      # of 1000 (t = 998, 899, ..., 9, 0)
SNAP_D = [1, 60, 125, 190, 240, 249]                 # of 250


_randn = torch.randn        # bound before main() patches torch.randn


def draw(k, shape=(1, 1, SIZE, SIZE)):
    return _randn(shape, generator=torch.Generator().manual_seed(NOISE_SEED + k))


def inputs(sdd):
    d01 = S.synthetic_depth_batch(50, 1, SIZE, SIZE)
    K = S.synthetic_intrinsics(1, SIZE, seed=7)
    P = S.synthetic_poses(1, seed=8)
    ic = sdd.image_condition(d01, torch.tensor(K), torch.tensor(P))
    pc = sdd.param_vector(torch.tensor(K))
    return d01, K, P, ic, pc


def main(which):
    sdd, _ = load_reference()
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    unet = sdd.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1).eval()
    d01, K, P, ic, pc = inputs(sdd)
    o = {"unet_fingerprint": fingerprint(unet.state_dict()), "noise_seed": np.int64(NOISE_SEED),
         "K": K, "P": P, "depth_sha": sha(d01.numpy()), "img_cond_sha": sha(ic.numpy()),
         "mask_fraction": np.float64(((ic[:, 1] + 1) * 0.5 > 0.5).float().mean())}
    if which == "p_sample":
        diff = sdd.GaussianDiffusion(unet, image_size=SIZE, timesteps=1000, sampling_timesteps=1000,
                                     objective="pred_x0", beta_schedule="sigmoid", is_ddnm_sampling=True)
        snaps, hook_name = SNAP_P, "p_sample"
    else:
        diff = sdd.GaussianDiffusion(unet, image_size=SIZE, timesteps=1000, sampling_timesteps=250,
                                     objective="pred_x0", beta_schedule="sigmoid", ddim_sampling_eta=1.0,
                                     is_ddnm_sampling=True)
        snaps, hook_name = SNAP_D, "model_predictions"
    state = {"k": 0, "step": 0, "t0": time.time()}
    rec = {}

    def next_noise(*a, **kw):
        z = draw(state["k"])
        state["k"] += 1
        return z

    inner = getattr(diff, hook_name)

    def hooked(x, *a, **kw):
        i = state["step"]
        if i in snaps:
            rec[i] = x.detach().clone().numpy()
        if i % 25 == 0:
            print("step %d  %.0f s" % (i, time.time() - state["t0"]), flush=True)
        state["step"] += 1
        return inner(x, *a, **kw)

    setattr(diff, hook_name, hooked)           # instance attribute: the reference file is untouched
    orig = (torch.randn, torch.randn_like)
    torch.randn, torch.randn_like = next_noise, next_noise
    try:
        out = diff.sample(param_cond=pc, img_cond=ic, disable_tqdm=True, has_refine_step=False)
    finally:
        torch.randn, torch.randn_like = orig
    o["num_draws"] = np.int64(state["k"])
    o["num_steps"] = np.int64(state["step"])
    o["snap_steps"] = np.array(snaps)
    for i in snaps:
        o["x_in_%d" % i] = rec[i]
    o["out"] = out.numpy()
    path = os.path.join(OUT, "traj_%s.npz" % which)
    np.savez_compressed(path, **o)
    print(path, os.path.getsize(path) // 1024, "KiB", "draws", state["k"], "steps", state["step"])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "ddim")

"""Mints tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded
synthetic inputs.  Run in the build container only:

    python -m oracle.make_golden

The fixtures pin (a) the oracle restatements in oracle/ (CPU tests) and (b) the CUDA path
(GPU tests) to the reference's own outputs.  Inputs are regenerated at test time from the same
seeds (pointreggpt_b200.synthetic / torch generators); the weight fingerprint guards the
assumption that `torch.manual_seed(0)` + construction reproduces the same random init.
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import load_reference  # noqa: E402
from pointreggpt_b200 import synthetic as S  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def fingerprint(sd):
    tot = 0.0
    for k in sorted(sd):
        tot += float(sd[k].double().abs().sum())
    return np.float64(tot)


def inject_noise(noises):
    it = iter(noises)
    orig = (torch.randn, torch.randn_like)
    torch.randn = lambda *a, **k: next(it).clone()
    torch.randn_like = lambda *a, **k: next(it).clone()
    return orig


def occlusion(sdd):
    """occlusion_filter (SDD:446-463) on the reprojection of the 256x256 geometry case and on a
    translated view (the Tester.sample use, SDD:2021-2037: 0.5 m forward, then the filter)."""
    o = {}
    B, H, W = 2, 256, 256
    d01 = S.synthetic_depth_batch(40, B, H, W)
    K = S.synthetic_intrinsics(B, 256, seed=5)
    fwd = np.stack([np.eye(4, dtype=np.float32)] * B)
    fwd[:, :3, 3] = [0, 0, 0.5]
    for tag, P in (("pose", S.synthetic_poses(B, seed=6)), ("fwd", fwd)):
        rd, rm = sdd.reproject_tensor(d01 * 10, torch.tensor(K), torch.tensor(P))
        fd, fm = sdd.occlusion_filter(rd.clone(), rm.clone())
        o["in_depth_sha_" + tag] = sha(rd.numpy())
        o["out_depth_sha_" + tag] = sha(fd.numpy())
        o["out_mask_sha_" + tag] = sha(fm.numpy())
        o["changed_" + tag] = np.int64((fd != rd).sum())
        o["out_crop_" + tag] = fd.numpy()[:, :, 96:160, 96:160]
    o["P_fwd"] = fwd
    np.savez_compressed(os.path.join(OUT, "occlusion.npz"), **o)
    print("occlusion.npz", os.path.getsize(os.path.join(OUT, "occlusion.npz")) // 1024, "KiB")


def main():
    os.makedirs(OUT, exist_ok=True)
    sdd, dc = load_reference()
    torch.set_num_threads(8)
    if sys.argv[1:] == ["occlusion"]:       # mint only the newer fixture, leave the others alone
        occlusion(sdd)
        return
    occlusion(sdd)

    # ------------------------------------------------------------------ geometry
    g = {}
    for tag, (B, H, W) in {"256": (2, 256, 256), "640": (1, 480, 640)}.items():
        d01 = S.synthetic_depth_batch(40, B, H, W)
        K = S.synthetic_intrinsics(B, 256 if H == 256 else None, seed=5)
        P = S.synthetic_poses(B, seed=6)
        dm = d01 * 10
        rd, rm = sdd.reproject_tensor(dm, torch.tensor(K), torch.tensor(P))
        pc, valid = sdd.depth2pc_tensor(dm, torch.tensor(K))
        g["K_" + tag], g["P_" + tag] = K, P
        g["depth_sha_" + tag] = sha(d01.numpy())
        if tag == "256":
            g["reproject_depth_256"] = rd.numpy()
            g["reproject_mask_256"] = np.packbits(rm.numpy())
        g["reproject_depth_sha_" + tag] = sha(rd.numpy())
        g["reproject_mask_sha_" + tag] = sha(rm.numpy())
        g["depth2pc_pc_sha_" + tag] = sha(pc.numpy())
        g["depth2pc_valid_sha_" + tag] = sha(valid.numpy())
        # Generator.generate path: numpy point_cloud -> numpy transform -> pc2depth_tensor per scene
        ds, ms, pcs_sha, back_sha = [], [], [], []
        for b in range(B):
            pcb = sdd.point_cloud(d01[b, 0].numpy() * 10, K[b], clip=[0.5, 10]).astype(np.float32)
            moved = pcb @ P[b, :3, :3].T + P[b, :3, 3]
            dd, mm = sdd.pc2depth_tensor(torch.tensor(moved[None]),
                                         torch.ones((1, moved.shape[0]), dtype=torch.bool),
                                         torch.tensor(K[b][None]), image_size=[H, W])
            ds.append(dd.numpy())
            ms.append(mm.numpy())
            pc64 = sdd.point_cloud(d01[b, 0].numpy() * 10, K[b], clip=[0.5, 10])
            pcs_sha.append(sha(pc64))
            back_sha.append(sha((pc64 - P[b, :3, 3]) @ P[b, :3, :3]))
        g["generate_pc2depth_depth_sha_" + tag] = sha(np.concatenate(ds))
        g["generate_pc2depth_mask_sha_" + tag] = sha(np.concatenate(ms))
        g["point_cloud_sha_" + tag] = np.array(pcs_sha)
        g["point_cloud_back_sha_" + tag] = np.array(back_sha)
    # host helpers
    cand = S.synthetic_intrinsics(6, None, seed=11)
    g["intrinsic_in"] = cand
    g["intrinsic_out"] = sdd.intrinsic_transform(cand, resize=256, centercrop=256)
    np.random.seed(21)
    g["pose_seed21"] = sdd.random_sample_pose(3)
    np.random.seed(22)
    g["intrinsic_seed22"] = sdd.random_sample_intrinsic(5)
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), **g)

    # ------------------------------------------------------------------ networks + sampler
    n = {}
    torch.manual_seed(0)
    unet = sdd.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1).eval()
    torch.manual_seed(0)
    mnet = dc.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8)).eval()
    n["unet_fingerprint"] = fingerprint(unet.state_dict())
    n["mask_fingerprint"] = fingerprint(mnet.state_dict())
    n["unet_probe"] = unet.state_dict()["downs.2.1.block2.proj.weight"][3, 5].numpy()
    gen = torch.Generator().manual_seed(77)
    pc = torch.tensor([[303.88547, 304.18253, 128.5, 128.0]])
    with torch.no_grad():
        x = torch.randn(1, 1, 128, 128, generator=gen)
        n["unet_128_t"] = np.array([417])
        n["unet_128_out"] = unet(x, torch.tensor([417]), pc).numpy()
        x2 = torch.randn(1, 1, 256, 256, generator=gen)
        n["unet_256_t"] = np.array([999])
        n["unet_256_out"] = unet(x2, torch.tensor([999]), pc).numpy()
        d = S.synthetic_depth_batch(3, 1, 128, 128)
        n["mask_128_out"] = mnet(d).numpy()
    # sampler with injected noise (p_sample T=3 + refine; ddim 12/3 + refine)
    draws = [torch.randn(1, 1, 128, 128, generator=gen) for _ in range(6)]
    dcond = S.synthetic_depth_batch(9, 1, 128, 128)
    ic = torch.cat([dcond, (dcond > 0).float()], 1) * 2 - 1
    diff = sdd.GaussianDiffusion(unet, image_size=128, timesteps=3, objective="pred_x0",
                                 beta_schedule="sigmoid", is_ddnm_sampling=True)
    orig = inject_noise(draws)
    try:
        n["p_sample_out"] = diff.sample(param_cond=pc, img_cond=ic, disable_tqdm=True,
                                        has_refine_step=True).numpy()
    finally:
        torch.randn, torch.randn_like = orig
    diff2 = sdd.GaussianDiffusion(unet, image_size=128, timesteps=12, sampling_timesteps=3,
                                  objective="pred_x0", beta_schedule="sigmoid",
                                  ddim_sampling_eta=1.0, is_ddnm_sampling=True)
    orig = inject_noise(draws)
    try:
        n["ddim_out"] = diff2.sample(param_cond=pc, img_cond=ic, disable_tqdm=True,
                                     has_refine_step=True).numpy()
    finally:
        torch.randn, torch.randn_like = orig
    # schedule buffers of the shipped configuration (T = 1000, sigmoid)
    diff3 = sdd.GaussianDiffusion(unet, image_size=256, timesteps=1000, sampling_timesteps=250,
                                  objective="pred_x0", beta_schedule="sigmoid")
    for k, v in diff3.state_dict().items():
        if not k.startswith("model."):
            n["sched_" + k] = v.numpy()
    n["ddim_times_1000_250"] = np.array(list(reversed(torch.linspace(-1, 999, steps=251).int().tolist())))
    np.savez_compressed(os.path.join(OUT, "networks.npz"), **n)
    for f in ("geometry.npz", "networks.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()

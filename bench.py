#!/usr/bin/env python
"""Benchmark of the per-pair data-generation hot path (BASELINE.json metric: generated
point-cloud pairs/sec at 1/2/4/8 B200; U-Net step ms; HBM GB/s of the geometry kernels).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pairs|geometry|dataset]
                    [--impl native|reference|reference-gpu]

Workloads (one "step" = one pass of the path over one batch of synthetic input; rank 0 prints ONE
JSON line; every rank works on its own shard, no data-path collective, `"scaling": "weak"`):

  pairs     (default) BASELINE configs[1]/[2]: `--batch` 256x256 pairs per GPU through z-buffer
            reprojection -> depth correction -> T-step DDNM p_sample loop -> depth correction ->
            depth->point cloud.  The line also carries the tensor roofline of the conv engine, the
            whole-step tensor fraction, and the HBM rooflines of the two geometry kernels measured
            on 640x480 maps right after the timed region (`roofline_geometry`).
  geometry  BASELINE configs[3]: z-buffer reprojection + dense unprojection over `--maps` 640x480
            maps per GPU (default 12 500 = 100 k over 8 GPUs), the first 256 checked bit for bit
            against the C oracle.
  dataset   BASELINE configs[4]: the `generate_dataset.py` driver (`Generator.generate`, files on:
            PLY / PNG / pose / intrinsics, resume logic) on synthetic source frames with the shipped
            sampler (250-step DDIM eta=1; `--batch` 4 = GD:47, or 32).

--impl reference      the reference algorithm's CPU port (oracle/, bit-equal to the reference's own
                      modules: tests/test_oracle_vs_reference.py) on the host cores, bounded sample,
                      extrapolated (a literal run takes ~20 min per pair).
--impl reference-gpu  secondary baseline: the same ATen call sequence as the reference's
                      modules, eager fp32 on cuda:0 through torch's cuDNN / cuBLAS -- "the existing
                      Blackwell path".
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNET_FLOP_PER_IMAGE = 236_282_087_424          # SURVEY.md section 8(d), per U-Net evaluation
MASK_FLOP_PER_IMAGE = 237_095_616_512
REPROJECT_BYTES_PER_PX = 9                     # 4 read + 4 write depth + 1 write mask
DEPTH2PC_BYTES_PER_PX = 17                     # 4 read + 12 write xyz + 1 write valid


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "reference-gpu"])
    ap.add_argument("--workload", default="pairs", choices=["pairs", "geometry", "dataset"])
    ap.add_argument("--batch", type=int, default=None, help="pairs per GPU per step (pairs: 32; dataset: 4)")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--timesteps", type=int, default=1000)
    ap.add_argument("--sampling-timesteps", type=int, default=None,
                    help="pairs: = --timesteps (p_sample_loop); dataset: 250 (DDIM, GD:38)")
    ap.add_argument("--micro-batch", type=int, default=0,
                    help="workspace batch of the network handles (0 = --batch)")
    ap.add_argument("--maps", type=int, default=12500, help="geometry: 640x480 maps per GPU per step")
    ap.add_argument("--pairs", type=int, default=64, help="dataset: scenes per GPU per step")
    ap.add_argument("--device-batch", type=int, default=None,
                    help="dataset: scenes per pass through the GPU (Generator's default: 32; = --batch for the "
                         "reference's literal batching)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    if a.batch is None:
        a.batch = 4 if a.workload == "dataset" else 32
    if a.sampling_timesteps is None:
        a.sampling_timesteps = 250 if a.workload == "dataset" else a.timesteps
    return a


def workload_name(a):
    if a.workload == "geometry":
        return ("configs[3]: z-buffer reprojection + depth->point-cloud unprojection over %d synthetic "
                "640x480 maps per GPU (first 256 bit-exact vs the C oracle)" % a.maps)
    if a.workload == "dataset":
        return ("configs[4]: generate_dataset.py driver (Generator.generate, files written), %d synthetic "
                "scenes per GPU per step, batch_size=%d, %d-step DDIM of T=%d, %dx%d"
                % (a.pairs, a.batch, a.sampling_timesteps, a.timesteps, a.size, a.size))
    return ("configs[1]: batch=%d %dx%d depth pairs, %d-step DDNM p_sample + z-buffer reprojection "
            "+ 2x depth-correction + depth->cloud, per GPU" % (a.batch, a.size, a.size, a.timesteps))


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU port timing
def cpu_reference_sample(size, timesteps, n_unet=3, unet_batch=4, n_mask=2, nmap=32):
    """Times the reference algorithm's CPU port (oracle/torch_ref + oracle/geometry_ref.c; bit-equal
    to the unmodified reference modules, tests/test_oracle_vs_reference.py) on a bounded sample of
    the pairs workload, as BASELINE.md section 4 prescribes -- scaled down to ~30 s of CPU work --
    and extrapolates pairs/s = 1 / (T * t_unet + 2 * t_mask + t_geom), all per image."""
    import torch
    from oracle import geometry_ref as G
    from oracle import torch_ref as R
    from pointreggpt_b200 import nets, synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd_u = {k: v.detach() for k, v in
            nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1).state_dict().items()}
    sd_m = {k: v.detach() for k, v in nets.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8)).state_dict().items()}
    x = torch.randn(unet_batch, 1, size, size)
    t = torch.full((unet_batch,), timesteps // 2)
    pc = torch.tensor([[303.9, 304.2, 128.5, 128.0]]).repeat(unet_batch, 1)
    R.unet_forward(sd_u, x[:1], t[:1], pc[:1])          # warm-up (thread pools, allocator)
    t0 = time.perf_counter()
    for _ in range(n_unet):
        R.unet_forward(sd_u, x, t, pc)
    t_unet = (time.perf_counter() - t0) / n_unet / unet_batch
    d = synthetic.synthetic_depth_batch(0, 1, size, size)
    t0 = time.perf_counter()
    for _ in range(n_mask):
        R.maskunet_forward(sd_m, d)
    t_mask = (time.perf_counter() - t0) / n_mask
    dd = synthetic.synthetic_depth_batch(0, 8, size, size).numpy().repeat(nmap // 8, 0)
    K = synthetic.synthetic_intrinsics(nmap, size)
    P = synthetic.synthetic_poses(nmap)
    t0 = time.perf_counter()
    G.reproject(dd * 10, K, P)
    G.depth2pc_compact(dd, K, P)
    t_geom = (time.perf_counter() - t0) / nmap
    per_pair = timesteps * t_unet + 2 * t_mask + t_geom
    sample = ("%d Unet.forward at B=%d + %d MaskUnet.forward at B=1, %dx%d fp32, + reprojection / "
              "point_cloud of %d maps, on %d host threads; CPU port of the reference (bit-equal to its "
              "modules); pairs/s extrapolated as 1/(%d*t_unet + 2*t_mask + t_geom) per image, "
              "t_unet=%.3fs t_mask=%.3fs t_geom=%.4fs"
              % (n_unet, unet_batch, n_mask, size, size, nmap, cores, timesteps, t_unet, t_mask, t_geom))
    return {"value": 1.0 / per_pair, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": sample, "t_unet_s": t_unet, "t_mask_s": t_mask, "t_geom_s": t_geom}


def cpu_geometry_sample(nmap=16, H=480, W=640):
    """CPU baseline of the geometry workload: the C oracle (scalar, one thread) on `nmap` maps."""
    import numpy as np
    from oracle import geometry_ref as G
    from pointreggpt_b200 import synthetic
    dd = synthetic.synthetic_depth_batch(0, 4, H, W).numpy().repeat(nmap // 4, 0) * np.float32(10)
    K = synthetic.synthetic_intrinsics(nmap, None)
    P = synthetic.synthetic_poses(nmap)
    G.reproject(dd[:1], K[:1], P[:1])
    t0 = time.perf_counter()
    G.reproject(dd, K, P)
    G.depth2pc(dd, K)
    dt = time.perf_counter() - t0
    return {"value": nmap / dt, "unit": "maps/s", "cores": 1, "kind": "port",
            "sample": "reproject + depth2pc of %d 640x480 maps, scalar C restatement of SDD:176-286 "
                      "(oracle/geometry_ref.c), one thread" % nmap}


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    vals, info = [], None
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        if a.workload == "geometry":
            info = cpu_geometry_sample()
        else:
            T = a.sampling_timesteps if a.workload == "dataset" else a.timesteps
            info = cpu_reference_sample(a.size, T, n_unet=1, unet_batch=min(4, a.batch), n_mask=1, nmap=8)
        if i >= a.warmup:
            vals.append((info["value"], time.perf_counter() - t0))
    v = sum(x for x, _ in vals) / len(vals)
    info["value"] = v
    unit = info["unit"]
    out = {
        "impl": "reference", "metric": "maps_per_sec" if a.workload == "geometry" else "pairs_per_sec",
        "value": v, "unit": unit,
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1000.0 * sum(t for _, t in vals) / len(vals),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "batch_per_gpu": a.batch, "image_size": a.size,
                   "timesteps": a.timesteps,
                   "sample": "bounded sample of the workload on the host cores, extrapolated (BASELINE.md section 4)"},
        "cpu_baseline": info,
        "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def run_reference_gpu_arm(a, rank, world):
    """The reference's ATen call sequence (oracle/torch_ref, bit-equal to its modules on CPU) run
    eagerly on cuda:0 in fp32: U-Net step ms at the bench batch, the number the native step is
    compared with.  TF32 off = the reference's default numerics; TF32 on is reported beside it."""
    if rank != 0:
        return
    import torch
    from oracle import torch_ref as R
    from pointreggpt_b200 import nets
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    sd = {k: v.detach().to(dev) for k, v in
          nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1).state_dict().items()}
    B = a.batch
    x = torch.randn(B, 1, a.size, a.size, device=dev)
    t = torch.full((B,), a.timesteps // 2, device=dev)
    pc = torch.tensor([[303.9, 304.2, 128.5, 128.0]], device=dev).repeat(B, 1)
    res = {}
    with torch.no_grad():
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cudnn.benchmark = True
            for _ in range(max(3, a.warmup)):
                R.unet_forward(sd, x, t, pc)
            torch.cuda.synchronize()
            n = max(5, a.steps)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                R.unet_forward(sd, x, t, pc)
            e1.record()
            torch.cuda.synchronize()
            res["tf32" if tf32 else "fp32"] = e0.elapsed_time(e1) / n
    T = a.timesteps
    pairs = B / (T * res["fp32"] * 1e-3)
    out = {"impl": "reference-gpu", "metric": "pairs_per_sec", "value": pairs, "unit": "pairs/s",
           "n_gpus": 1, "steps": a.steps, "warmup": a.warmup, "ms_per_step": T * res["fp32"],
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": workload_name(a), "batch_per_gpu": B, "image_size": a.size, "timesteps": T,
                      "sample": "Unet.forward only (>99.9 %% of the path), pairs/s = B / (T * step); eager "
                                "torch %s, cuDNN benchmark on" % torch.__version__},
           "unet_step_ms": res["fp32"], "unet_step_ms_tf32": res["tf32"],
           "unet_step_tflops": UNET_FLOP_PER_IMAGE * B / (res["fp32"] * 1e-3) / 1e12,
           "unet_step_tflops_tf32": UNET_FLOP_PER_IMAGE * B / (res["tf32"] * 1e-3) / 1e12,
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------- shared pieces
class Ctx:
    """torch / distributed state of one rank."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the native path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], device=self.dev, dtype=self.torch.int64)
        self.dist.all_reduce(t)
        return int(t.item())

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def geometry_roofline(ctx, maps=512, iters=10, H=480, W=640):
    """HBM rooflines of the two geometry kernels on `maps` 640x480 maps (inputs 630 MB > the 126 MB
    L2): CUDA events on the launching stream around each call, median of `iters`."""
    torch = ctx.torch
    from pointreggpt_b200 import geometry, synthetic
    peaks = load_peaks()
    peak = peaks.get("hbm_gbs")
    src = "MEASURED_PEAKS.json hbm_gbs (copy bandwidth; kernels timed alone)"
    if peak is None:
        peak, src = 6500.0, "fallback (B200_PROFILING.md measured copy bandwidth ~6.5 TB/s)"
    d = synthetic.synthetic_depth_batch(0, 8, H, W)
    d = (d * 10).repeat((maps + 7) // 8, 1, 1, 1)[:maps].contiguous().to(ctx.dev)       # metres
    K = torch.tensor(synthetic.synthetic_intrinsics(maps, None)).to(ctx.dev)
    P = torch.tensor(synthetic.synthetic_poses(maps)).to(ctx.dev)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return statistics.median(ts)

    px = maps * H * W
    out = {}
    for name, bpp, fn in (("reproject", REPROJECT_BYTES_PER_PX, lambda: geometry.reproject_tensor(d, K, P)),
                          ("depth2pc", DEPTH2PC_BYTES_PER_PX, lambda: geometry.depth2pc_tensor(d, K, clip=[0, 10]))):
        ms = timed(fn)
        gbs = px * bpp / ms / 1e6
        out[name] = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                     "traffic": None, "bytes_per_pixel": bpp, "maps": maps, "ms": ms,
                     "maps_per_sec": maps / ms * 1e3, "peak_source": src}
    return out


def build_models(ctx, a, sampling_timesteps):
    torch = ctx.torch
    from pointreggpt_b200 import dist as pdist
    from pointreggpt_b200 import nets
    from pointreggpt_b200.diffusion import GaussianDiffusion
    torch.manual_seed(0)        # rank 0 owns the seeded random init, everyone else receives it over NCCL
    unet = nets.Unet(dim=64, param_cond_dim=4, dim_mults=(1, 2, 4, 8), channels=1)
    mask = nets.MaskUnet(dim=64, dim_mults=(1, 2, 4, 8))
    with torch.no_grad():
        mask.final_conv[0].bias.fill_(8.0)     # random-init sigmoid(~0) would mask everything out
    diffusion = GaussianDiffusion(unet, image_size=a.size, timesteps=a.timesteps,
                                  sampling_timesteps=sampling_timesteps, objective="pred_x0",
                                  beta_schedule="sigmoid", ddim_sampling_eta=1.0,
                                  is_ddnm_sampling=True).to(ctx.dev)
    mask = mask.to(ctx.dev)
    if ctx.world > 1:
        pdist.broadcast_weights([diffusion, mask], src=0)
    if a.micro_batch:
        unet.max_batch = a.micro_batch
        mask.max_batch = a.micro_batch
    return unet, mask, diffusion


# --------------------------------------------------------------------------- workload: pairs
def run_pairs(a):
    ctx = Ctx(a)
    torch = ctx.torch
    from pointreggpt_b200 import _ffi, pipeline, rng, synthetic
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    unet, mask, diffusion = build_models(ctx, a, a.timesteps)

    # ---- synthetic inputs: distinct scenes per rank / step, resident in HBM
    B = a.batch
    nsets = 2
    host_inputs = []
    for s in range(nsets):
        base = (rank * nsets + s) * B
        d = synthetic.synthetic_depth_batch(base, B, a.size, a.size).pin_memory()
        K = torch.tensor(synthetic.synthetic_intrinsics(B, a.size, seed=base)).pin_memory()
        P = torch.tensor(synthetic.synthetic_poses(B, seed=base + 1)).pin_memory()
        host_inputs.append((d, K, P))
    dev_inputs = [tuple(t.to(dev) for t in hi) for hi in host_inputs]
    total_steps = a.warmup + 2 * a.steps

    def seeds(i):
        # one Philox key per pair, a function of the pair's absolute index in the whole job
        return [rng.scene_seed(1234, (rank * total_steps + i) * B + b, 0) for b in range(B)]

    def step(i, inputs):
        d, K, P = inputs[i % nsets]
        return pipeline.generate_batch(diffusion, mask, d, K, P, seed=seeds(i))

    for i in range(a.warmup):
        step(i, dev_inputs)
    ctx.barrier()

    # ---- timed region (device events on the launching stream; max over ranks)
    clocks = ClockSampler(ctx.local_rank)
    if rank == 0:
        clocks.start()
    every = max(1, (a.timesteps * a.steps) // 24)
    _ffi.profile_set(every)
    _ffi.profile_read(reset=True)
    _ffi.profile_ops(reset=True)
    l0 = _ffi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    e0.record()
    for i in range(a.steps):
        step(a.warmup + i, dev_inputs)
    e1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.sum_over_ranks(_ffi.launch_count() - l0)
    prof = _ffi.profile_read(reset=True)
    ops = _ffi.profile_ops(reset=True)
    _ffi.profile_set(0)
    clk = clocks.stop() if rank == 0 else None
    value = world * B * a.steps / (ms / 1000.0)

    # ---- end to end through the public API: pinned host inputs -> H2D -> path -> D2H result
    e2e = None
    if not a.no_e2e:
        host_out = torch.empty((B, a.size * a.size, 3), dtype=torch.float64).pin_memory()
        host_cnt = torch.empty((B,), dtype=torch.int64).pin_memory()
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(a.steps):
            hi = host_inputs[i % nsets]
            di = tuple(t.to(dev, non_blocking=True) for t in hi)
            pc, cnt, img = pipeline.generate_batch(diffusion, mask, di[0], di[1], di[2],
                                                   seed=seeds(a.warmup + a.steps + i))
            host_out.copy_(pc, non_blocking=True)
            host_cnt.copy_(cnt, non_blocking=True)
            torch.cuda.synchronize()
        te = ctx.max_over_ranks(time.perf_counter() - t0)
        h2d = sum(t.numel() * t.element_size() for t in host_inputs[0])
        d2h = host_out.numel() * 8 + host_cnt.numel() * 8
        e2e = {"value": world * B * a.steps / te, "unit": "pairs/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
        del host_out

    geo = None
    if rank == 0:
        del dev_inputs
        torch.cuda.empty_cache()
        try:
            geo = geometry_roofline(ctx)
        except Exception as ex:     # never take the headline down
            geo = {"error": repr(ex)}
    if rank != 0:
        ctx.close()
        return

    # ---- tensor roofline.  Numerator = the algorithmic FLOP of exactly the ops inside the timed family
    # (per-layer table of the plan, prg_profile_ops), not of the whole network.
    peaks = load_peaks()
    peak = peaks.get("bf16_tflops_sustained")
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernels timed inside a long step)"
    if peak is None:
        peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"
    roof, unet_ms = None, None
    eff_b = min(B, a.micro_batch or B)
    if "conv_tc" in prof and prof["conv_tc"]["forwards"] > 0:
        c = prof["conv_tc"]
        fw = c["forwards"]
        conv_ms = c["ms"] / fw
        u_ops = [o for o in ops if o[0].startswith("U")]       # (label, sampled launches, total ms, FLOP/image)
        n_u = max([n_l for _, n_l, _, _ in u_ops] or [1])      # sampled U-Net evaluations
        fams = {}
        for lab, n_l, ms_l, fl in u_ops:
            fam = lab.split(":", 1)[1].split(" ")[0] if ":" in lab else "tail"
            fam = "conv_tc" if fam.startswith("conv") else fam
            fams[fam] = fams.get(fam, 0.0) + ms_l / n_u
        conv_flop = sum(fl for lab, _, _, fl in u_ops if ":conv" in lab)
        conv_ms_u = fams.get("conv_tc", conv_ms)
        achieved = conv_flop * eff_b / (conv_ms_u * 1e-3) / 1e12
        unet_ms = sum(fams.values())
        pair_flop = a.timesteps * UNET_FLOP_PER_IMAGE + 2 * MASK_FLOP_PER_IMAGE
        whole = pair_flop * B * a.steps / (ms * 1e-3) / 1e12
        # DRAM bytes of all k_conv2 launches of one evaluation at batch 32, 256x256, from the ncu
        # capture summarised under profiles/ (see profiles/README.md for the capture it comes from)
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
                tj = json.load(f)
            if B == tj.get("batch") and a.size == tj.get("size") and not a.micro_batch:
                traffic = tj.get("dram_bytes_per_unet_eval")
        except Exception:
            pass
        roof = {"bound": "tensor",
                "kernel": "k_conv2 (persistent tcgen05 implicit-GEMM conv engine): the conv launches of one "
                          "U-Net evaluation, CUDA events on the launching stream, sampled inside the timed region",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "flop_per_image_in_family": conv_flop,
                "traffic": traffic, "traffic_unit": "bytes per U-Net evaluation (all conv launches)",
                "peak_source": peak_src,
                "launches_per_unet_eval": c["launches"] / fw, "ms_per_unet_eval": conv_ms_u,
                "whole_step_achieved": whole, "whole_step_frac": whole / peak,
                "whole_step_note": "(T x 236.282 + 2 x 237.096) GFLOP x pairs / un-instrumented wall time of the "
                                   "timed region (all kernels, launch gaps, geometry and depth correction included)",
                "families_ms_per_unet_eval": fams}

    out = {
        "metric": "pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": workload_name(a), "batch_per_gpu": B, "image_size": a.size,
                   "timesteps": a.timesteps, "micro_batch": a.micro_batch or B,
                   "weights": "random init, torch.manual_seed(0); MaskUnet final bias +8 so the "
                              "keep-mask is non-trivial",
                   "l2": "per-step activations/workspace (GBs) exceed the 126 MB L2; inputs "
                         "alternate between two resident sets",
                   "rng": "one device Philox stream per pair keyed by its absolute index (rank-independent)",
                   "parallelism": "independent pairs sharded by rank, weights NCCL-broadcast"},
        "unet_step_ms": unet_ms,
        "unet_step_ms_wall": (ms / a.steps) / a.timesteps,
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roof,
        "roofline_geometry": geo,
    }
    if world == 1 and not a.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_reference_sample(a.size, a.timesteps)
        except Exception as ex:  # the checker must never take the bench down
            out["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(out), flush=True)
    ctx.close()


# --------------------------------------------------------------------------- workload: geometry
def run_geometry(a):
    ctx = Ctx(a)
    torch = ctx.torch
    import numpy as np
    from pointreggpt_b200 import _ffi, geometry, synthetic
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    H, W = 480, 640
    POOL = 500                      # distinct maps resident in HBM: 614 MB of input, > the 126 MB L2
    chunk = POOL
    # 64 distinct synthetic maps per rank, the pool repeats them (distinct K / pose per pool entry)
    base = synthetic.synthetic_depth_batch(rank * 64, 64, H, W)
    depth_h = (base * 10).repeat((POOL + 63) // 64, 1, 1, 1)[:POOL].contiguous().pin_memory()
    K_h = torch.tensor(synthetic.synthetic_intrinsics(POOL, None, seed=rank)).pin_memory()
    P_h = torch.tensor(synthetic.synthetic_poses(POOL, seed=100 + rank)).pin_memory()
    depth, K, P = depth_h.to(dev), K_h.to(dev), P_h.to(dev)
    n_chunks = (a.maps + chunk - 1) // chunk
    maps = n_chunks * chunk if a.maps >= chunk else a.maps
    if a.maps < chunk:
        chunk, n_chunks = a.maps, 1
        depth, K, P = depth[:chunk], K[:chunk], P[:chunk]

    # ---- bit-exact subset check against the C oracle (TEST INFRASTRUCTURE; outside the timed region)
    checked = None
    if rank == 0:
        from oracle import geometry_ref as G
        ncheck = min(256, chunk)
        od, om = G.reproject(depth_h[:ncheck].numpy(), K_h[:ncheck].numpy(), P_h[:ncheck].numpy())
        opc, ov = G.depth2pc(depth_h[:ncheck].numpy(), K_h[:ncheck].numpy())
        gd, gm = geometry.reproject_tensor(depth[:ncheck], K[:ncheck], P[:ncheck])
        gpc, gv = geometry.depth2pc_tensor(depth[:ncheck], K[:ncheck], clip=[0, 10])
        ok = (np.array_equal(gd.cpu().numpy().view(np.uint32), od.view(np.uint32)) and
              np.array_equal(gm.cpu().numpy(), om) and
              np.array_equal(gpc.cpu().numpy().view(np.uint32), opc.view(np.uint32)) and
              np.array_equal(gv.cpu().numpy(), ov))
        checked = {"maps": ncheck, "bit_exact": bool(ok)}
        if not ok:
            raise SystemExit("bench.py geometry: CUDA results differ from the C oracle on the checked subset")
        del gd, gm, gpc, gv

    def one_pass(time_each=None):
        for c in range(n_chunks):
            if time_each is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
            geometry.reproject_tensor(depth, K, P)
            if time_each is not None:
                ev[1].record()
            geometry.depth2pc_tensor(depth, K, clip=[0, 10])
            if time_each is not None:
                ev[2].record()
                time_each.append(ev)

    for _ in range(a.warmup):
        one_pass()
    ctx.barrier()
    clocks = ClockSampler(ctx.local_rank)
    if rank == 0:
        clocks.start()
    l0 = _ffi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    evs = []
    ctx.barrier()
    e0.record()
    for _ in range(a.steps):
        one_pass(evs)
    e1.record()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.sum_over_ranks(_ffi.launch_count() - l0)
    clk = clocks.stop() if rank == 0 else None
    t_rp = sum(e[0].elapsed_time(e[1]) for e in evs) / len(evs)
    t_dp = sum(e[1].elapsed_time(e[2]) for e in evs) / len(evs)
    value = world * maps * a.steps / (ms / 1000.0)

    e2e = None
    if not a.no_e2e:
        # host buffers: H2D of the depth maps, both kernels, D2H of depth + mask + cloud + valid
        n_e = min(chunk, 250)
        ho = [torch.empty((n_e, 1, H, W), dtype=torch.float32).pin_memory(),
              torch.empty((n_e, 1, H, W), dtype=torch.bool).pin_memory(),
              torch.empty((n_e, H * W, 3), dtype=torch.float32).pin_memory(),
              torch.empty((n_e, H * W), dtype=torch.bool).pin_memory()]
        ctx.barrier()
        t0 = time.perf_counter()
        reps = max(1, a.steps)
        for _ in range(reps):
            d = depth_h[:n_e].to(dev, non_blocking=True)
            k = K_h[:n_e].to(dev, non_blocking=True)
            p = P_h[:n_e].to(dev, non_blocking=True)
            r = geometry.reproject_tensor(d, k, p) + geometry.depth2pc_tensor(d, k, clip=[0, 10])
            for dst, src in zip(ho, r):
                dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
        te = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": world * n_e * reps / te, "unit": "maps/s",
               "h2d_bytes_per_step": n_e * (H * W * 4 + 36 + 64),
               "d2h_bytes_per_step": sum(t.numel() * t.element_size() for t in ho),
               "maps_per_step": n_e}
    if rank != 0:
        ctx.close()
        return
    peaks = load_peaks()
    peak = peaks.get("hbm_gbs")
    src = "MEASURED_PEAKS.json hbm_gbs (sustained copy bandwidth)"
    if peak is None:
        peak, src = 6500.0, "fallback (B200_PROFILING.md measured copy bandwidth ~6.5 TB/s)"
    px = chunk * H * W
    rp = px * REPROJECT_BYTES_PER_PX / t_rp / 1e6
    dp = px * DEPTH2PC_BYTES_PER_PX / t_dp / 1e6
    both = px * (REPROJECT_BYTES_PER_PX + DEPTH2PC_BYTES_PER_PX) / (t_rp + t_dp) / 1e6
    out = {
        "metric": "maps_per_sec", "value": value, "unit": "maps/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "maps_per_gpu": maps, "height": H, "width": W,
                   "chunk": chunk, "l2": "each call reads %d MB of resident depth maps (> 126 MB L2) and writes "
                                         "%d MB" % (px * 4 >> 20, px * 22 >> 20),
                   "parallelism": "maps sharded by rank, no collective"},
        "hbm_gbs": both, "clocks": clk, "e2e": e2e, "gpu_launches": launches, "checked": checked,
        "roofline": {"bound": "hbm", "kernel": "k_reproject_fused (one persistent launch per call: splat + z-buffer ring in L2 + finalise) "
                     "and k_depth2pc; CUDA events around every call in the timed region",
                     "achieved": both, "peak": peak, "unit": "GB/s", "frac": both / peak, "traffic": None,
                     "peak_source": src,
                     "reproject": {"achieved": rp, "frac": rp / peak, "bytes_per_pixel": REPROJECT_BYTES_PER_PX, "ms_per_call": t_rp},
                     "depth2pc": {"achieved": dp, "frac": dp / peak, "bytes_per_pixel": DEPTH2PC_BYTES_PER_PX, "ms_per_call": t_dp}},
    }
    if world == 1 and not a.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_geometry_sample()
        except Exception as ex:
            out["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(out), flush=True)
    ctx.close()


# --------------------------------------------------------------------------- workload: dataset
def run_dataset(a):
    ctx = Ctx(a)
    torch = ctx.torch
    from pointreggpt_b200 import _ffi
    from pointreggpt_b200.generator import Generator
    rank, world = ctx.rank, ctx.world
    unet, mask, diffusion = build_models(ctx, a, a.sampling_timesteps)
    root = tempfile.mkdtemp(prefix="prg_dataset_r%d_" % rank)
    cwd = os.getcwd()
    os.chdir(root)                                  # the driver resolves checkpoints relative to the cwd
    try:
        gen = Generator(diffusion, "synthetic", batch_size=a.batch, results_folder=os.path.join(root, "res"),
                        samples_folder=os.path.join(root, "ds", "data"), device=ctx.dev,
                        device_batch=a.device_batch)
        gen.rank, gen.world_size = 0, 1             # this bench shards the scene ranges itself (weak scaling)
        n = a.pairs
        cursor = [rank * (a.warmup + a.steps) * n]

        def step():
            lo = cursor[0]
            cursor[0] += n
            done = gen.generate(lo, lo + n, num_samples=1, has_refine_step=False, depth_correction=mask,
                                base_seed=0)
            assert done == n * world          # generate() returns the all-reduced count of finished scenes
        for _ in range(a.warmup):
            step()
        ctx.barrier()
        clocks = ClockSampler(ctx.local_rank)
        if rank == 0:
            clocks.start()
        l0 = _ffi.launch_count()
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step()                                  # generate() returns after its writer thread has finished
        torch.cuda.synchronize()
        te = ctx.max_over_ranks(time.perf_counter() - t0)
        launches = ctx.sum_over_ranks(_ffi.launch_count() - l0)
        clk = clocks.stop() if rank == 0 else None
        files = sum(len(fs) for _, _, fs in os.walk(os.path.join(root, "ds")))
        nbytes = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(os.path.join(root, "ds")) for f in fs)
    finally:
        os.chdir(cwd)
        shutil.rmtree(root, ignore_errors=True)
    if rank != 0:
        ctx.close()
        return
    value = world * n * a.steps / te
    peaks = load_peaks()
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    pair_flop = a.sampling_timesteps * UNET_FLOP_PER_IMAGE + 2 * MASK_FLOP_PER_IMAGE
    whole = pair_flop * n * a.steps / te / 1e12
    out = {
        "metric": "pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1000.0 * te / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": workload_name(a), "pairs_per_gpu_per_step": n, "batch_size": a.batch,
                   "device_batch": gen.device_batch,
                   "batching": "batch_size is the reference's CLI value (skip-if-done bookkeeping); the scenes of "
                               "several such batches pass through the GPU together (device_batch) -- a scene's "
                               "files do not depend on the batch it travels in",
                   "image_size": a.size, "timesteps": a.timesteps, "sampling_timesteps": a.sampling_timesteps,
                   "files": "camera-intrinsics.txt, sample-*.{image.png,depth.png,pose.txt,cloud.ply}, "
                            "reprojected/corrected.image.png per scene (%d files, %.1f MB on this rank)" % (files, nbytes / 1e6),
                   "timing": "host wall clock around Generator.generate (includes source synthesis, H2D, all kernels, "
                             "D2H, voxel down-sampling, PNG/PLY/TXT writes), max over ranks",
                   "parallelism": "scene ranges sharded by rank, weights NCCL-broadcast"},
        "clocks": clk, "gpu_launches": launches,
        # this workload IS the end-to-end path: value == e2e
        "e2e": {"value": value, "unit": "pairs/s",
                "h2d_bytes_per_step": n * (a.size * a.size * 4 + 36 + 64),
                "d2h_bytes_per_step": int(nbytes / max(1, a.warmup + a.steps))},
        "roofline": {"bound": "tensor", "kernel": "whole driver step (every kernel + host tail)",
                     "achieved": whole, "peak": peak, "unit": "TFLOP/s", "frac": whole / peak, "traffic": None},
    }
    if world == 1 and not a.no_cpu_baseline:
        try:
            out["cpu_baseline"] = cpu_reference_sample(a.size, a.sampling_timesteps, n_unet=2)
        except Exception as ex:
            out["cpu_baseline"] = {"error": repr(ex)}
    print(json.dumps(out), flush=True)
    ctx.close()


def main():
    a = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if a.gpus > 1 and world == 1:
        # launched directly: re-exec under torchrun (one process per GPU)
        os.execv(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                  "--nproc-per-node", str(a.gpus), "--master-addr", "127.0.0.1",
                                  "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv)
    if a.impl == "reference":
        return run_reference_arm(a, rank, world)
    if a.impl == "reference-gpu":
        return run_reference_gpu_arm(a, rank, world)
    {"pairs": run_pairs, "geometry": run_geometry, "dataset": run_dataset}[a.workload](a)


if __name__ == "__main__":
    main()

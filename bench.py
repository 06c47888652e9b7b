#!/usr/bin/env python
"""bench.py — Mrays/s and frame time of the path-tracing hot path on B200 (see the contract in DESIGN.md).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (default N=1)
  python bench.py --impl reference --gpus N --steps K ...   the reference's algorithm on the host cores
                                                            (CPU oracle; the WGSL shader itself cannot run
                                                            in this image: no Rust toolchain, no Vulkan)
A step is one frame of the workload (all samples of every pixel)."""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mrays/s, RTIOW final scene 1920x1080 100 spp 10 bounces (ray = one raycast() call, raytrace.wgsl:190)"
METRIC_C3 = "Mrays/s, RTIOW final scene 3840x2160 1000 spp 10 bounces (ray = one raycast() call)"
METRIC_C4 = "Mrays/s, synthetic 2^20 random spheres 1920x1080 64 spp 10 bounces (ray = one raycast() call)"
UNIT = "Mrays/s"
SCENE_SEED = 1
BASE_SEED = 0.37

WORKLOADS = {
    # BASELINE.json configs[1]: RTIOW book-1 final scene, 1920x1080, 100 spp, 10 bounces, level Pure
    "c2": dict(name="C2 rtiow-final 1920x1080 100spp 10 bounces, book camera (13,2,3)->(0,0,0) vfov 20deg",
               width=1920, height=1080, spp=100, bounces=10, camera="book"),
    # BASELINE.json configs[3]: traversal / memory-bound stress, scene does not fit in shared memory
    "c4": dict(name="C4 synthetic 2^20 random spheres (cube side 200, r in [0.05,0.25], 80/15/5 % materials), "
                    "1920x1080 64spp 10 bounces, camera (0,0,130)->(0,0,0) fov pi/4",
               width=1920, height=1080, spp=64, bounces=10, camera="c4", scene=("random", 7, 1 << 20, 200.0, 0.05, 0.25)),
    # BASELINE.json configs[2]: 4K, 1000 spp, tile- or sample-sharded over the GPUs of the box (strong scaling: with
    # --shard samples every rank renders 1000/N samples of the whole frame with its own seed)
    "c3": dict(name="C3 rtiow-final 3840x2160 1000spp 10 bounces, book camera (13,2,3)->(0,0,0) vfov 20deg",
               width=3840, height=2160, spp=1000, bounces=10, camera="book", split_spp=True),
    # BASELINE.json configs[0] (plumbing / parity case)
    "c1": dict(name="C1 default scene 1280x720 1spp 4 bounces, repo camera (0,0,5)->(0,0,0) fov pi/4",
               width=1280, height=720, spp=1, bounces=4, camera="repo"),
}


def make_cam(bvr, wl, spp=None):
    aspect = wl["width"] / wl["height"]
    spp = wl["spp"] if spp is None else spp
    if wl["camera"] == "c4":
        return bvr.make_camera(position=(0.0, 0.0, 130.0), target=(0.0, 0.0, 0.0), aspect=aspect, sample_count=spp,
                               bounces=wl["bounces"])
    if wl["camera"] == "book":
        return bvr.make_camera(position=(13.0, 2.0, 3.0), target=(0.0, 0.0, 0.0), fov=float(np.deg2rad(20.0)),
                               aspect=aspect, sample_count=spp, bounces=wl["bounces"])
    return bvr.make_camera(aspect=aspect, sample_count=spp, bounces=wl["bounces"])


def make_scene(bvr, wl):
    sc = wl.get("scene")
    if sc and sc[0] == "random":
        return bvr.Scene.random(*sc[1:])
    return bvr.Scene.rtiow(SCENE_SEED)


def flops_per_ray(cnt):
    """Algorithmic FLOPs per ray from the oracle's work counters."""
    rays = max(cnt["rays"], 1)
    inner, tests, shaded, paths = cnt["inner_visits"], cnt["sphere_tests"], cnt["hits_shaded"], cnt["paths"]
    return 3.0 + (inner * 44.0 + tests * 27.0 + shaded * 83.0 + paths * 55.0) / rays


def bytes_per_ray(cnt):
    """Algorithmic bytes per ray in the reference layout, each datum fetched once."""
    rays = max(cnt["rays"], 1)
    return (cnt["inner_visits"] * 96.0 + cnt["sphere_tests"] * 32.0 + cnt["hits_shaded"] * 32.0) / rays


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region (nvidia-smi fields via NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(bvr, oracle, scene, wl, spp, threads=None):
    """Times the CPU oracle on a bounded sample of the workload: the same frame at `spp` samples."""
    threads = host_threads() if threads is None else threads
    cam = make_cam(bvr, wl, spp)
    win = bvr.make_window(BASE_SEED, wl["height"])
    t0 = time.perf_counter()
    _, cnt = oracle.render(scene.models, scene.materials, scene.nodes, cam, bvr.make_level(3), win, wl["width"],
                           threads=threads)
    dt = time.perf_counter() - t0
    return cnt, dt


def run_reference(args):
    """--impl reference: the reference's algorithm (CPU restatement of the WGSL shader, oracle/) on all host
    cores.  The shader itself cannot run here: no Rust toolchain, no Vulkan loader, no lavapipe."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import bevyray_b200 as bvr
    from oracle import oracle
    wl = WORKLOADS[args.workload]
    scene = make_scene(bvr, wl)
    sample_spp = max(1, min(wl["spp"], args.cpu_spp))
    cores = host_threads()
    for _ in range(args.warmup):
        cpu_sample(bvr, oracle, scene, dict(wl, width=wl["width"] // 4, height=wl["height"] // 4), 1)
    rays, total = 0, 0.0
    for _ in range(args.steps):
        cnt, dt = cpu_sample(bvr, oracle, scene, wl, sample_spp)
        rays += cnt["rays"]
        total += dt
    value = rays / total / 1e6
    sample = f"{wl['width']}x{wl['height']} x {sample_spp} spp of {wl['spp']} per step (same scene, camera, seed, bounces)"
    line = {"impl": "reference", "metric": {"c4": METRIC_C4, "c3": METRIC_C3}.get(args.workload, METRIC), "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3 * (wl["spp"] / sample_spp),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "l2": "not applicable (CPU)",
                       "note": "CPU restatement of the reference WGSL shader (oracle/), OpenMP over rows; "
                               "ms_per_step extrapolated linearly in spp from the sample"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import bevyray_b200 as bvr
    from bevyray_b200 import capi
    from bevyray_b200.distributed import ShardedRenderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — bevyray_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    wl = WORKLOADS[args.workload]
    W, H = wl["width"], wl["height"]
    scene = make_scene(bvr, wl)
    split_spp = bool(wl.get("split_spp")) and args.shard == "samples" and world > 1
    rank_spp = max(1, wl["spp"] // world) if split_spp else wl["spp"]
    cam = make_cam(bvr, wl, rank_spp)
    kernel = {"auto": capi.KERNEL_AUTO, "megakernel": capi.KERNEL_MEGAKERNEL, "wavefront": capi.KERNEL_WAVEFRONT,
              "cta-wavefront": capi.KERNEL_CTA_WAVEFRONT}[args.kernel]
    traversal = capi.TRAVERSAL_REFERENCE_ORDER if args.reference_order else capi.TRAVERSAL_AUTO

    r = ShardedRenderer(local_rank, rank, world, mode=args.shard, strip_rows=args.strip_rows)
    if args.gpu_bvh:
        r.ctx.upload_scene_gpu_bvh(scene.models, scene.materials)     # EXPERIMENT: LBVH built on the GPU
    else:
        r.upload_scene(scene.models, scene.materials, scene.nodes)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=r.device)   # > 126 MB L2

    fp32_peak = None
    if rank == 0:
        import ctypes as C
        pk = C.c_float()
        if capi.lib.bvr_bench_fp32_peak(local_rank, C.byref(pk)) == 0:
            fp32_peak = float(pk.value)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return r.render_frame(cam, 3, BASE_SEED, W, H, kernel, traversal)

    for _ in range(args.warmup):
        step()
        flush.zero_()
    barrier()
    launches0 = r.ctx.stats()["kernel_launches"]

    # ---- device-resident timing: K steps, CUDA events on the launching stream, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    rays = 0
    kernel_ms = []
    barrier()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        ev[i][0].record()
        step()
        ev[i][1].record()
        flush.zero_()
        st = r.ctx.stats()          # synchronises the stream; reads the device ray counter
        rays += st["rays"]
        kernel_ms.append(st["last_render_ms"])
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    launches = r.ctx.stats()["kernel_launches"] - launches0
    if world > 1:
        t = torch.tensor([total_ms, float(rays), float(launches)], dtype=torch.float64, device=r.device)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_ms, rays, launches = float(tmax[0]), int(t[1]), int(t[2])
    value = rays / (total_ms * 1e-3) / 1e6

    # ---- end to end through the public API: host buffers, upload + render + readback every step ----
    scene_bytes = scene.models.nbytes + scene.materials.nbytes + scene.nodes.nbytes
    host_rgba = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    host_np = host_rgba.numpy()
    e2e_rays = 0
    barrier()
    t0 = time.perf_counter()
    if world == 1:
        opts = bvr.make_options(W, kernel, traversal)
        win = bvr.make_window(BASE_SEED, H)
        for _ in range(args.steps):
            if args.gpu_bvh:
                r.ctx.upload_scene_gpu_bvh(scene.models, scene.materials)
            else:
                r.upload_scene(scene.models, scene.materials, scene.nodes)      # the reference re-uploads every frame
            r.ctx.render(cam, 3, win, opts, want=("rgba",), out={"rgba": host_np})   # bvr_render: H2D, kernels, D2H, sync
            e2e_rays += r.ctx.stats()["rays"]
        d2h = host_np.nbytes
    else:
        for _ in range(args.steps):
            r.upload_scene(scene.models, scene.materials, scene.nodes)
            frame = step()
            if rank == 0:
                host_rgba.copy_(frame, non_blocking=True)
            torch.cuda.synchronize()
            e2e_rays += r.ctx.stats()["rays"]
        d2h = host_np.nbytes if rank == 0 else 0
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=r.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
        t = torch.tensor([float(e2e_rays)], dtype=torch.float64, device=r.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        e2e_rays = int(t[0])
    e2e_value = e2e_rays / e2e_s / 1e6

    if rank == 0:
        line = {"metric": {"c4": METRIC_C4, "c3": METRIC_C3}.get(args.workload, METRIC), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak" if (args.shard == "samples" and not split_spp) else "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": wl["name"], "scene_seed": SCENE_SEED, "random_seed": BASE_SEED,
                           "spheres": int(len(scene.models)), "kernel": args.kernel,
                           "traversal": "reference-order" if args.reference_order else "near-first",
                           "bvh": "GPU LBVH" if args.gpu_bvh else "host PLOC (restated obvhs call, extract.rs:316-321)",
                           "sharding": "none" if world == 1 else
                           (f"samples: {rank_spp} spp per rank, distinct seed per rank, NCCL reduce to rank 0"
                            if args.shard == "samples" else
                            f"tiles: {args.strip_rows}-row strips interleaved over ranks, NCCL all_gather"),
                           "l2": "256 MiB buffer written between timed steps (L2 flush)"},
                "frame_ms": total_ms / args.steps, "mpaths_per_s": W * H * rank_spp * world * args.steps / (total_ms * 1e-3) / 1e6
                if args.shard == "samples" else W * H * wl["spp"] * args.steps / (total_ms * 1e-3) / 1e6,
                "rays_per_step": rays // args.steps, "wall_s_timed_region": t_wall,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(scene_bytes) * world,
                        "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_s / args.steps * 1e3},
                "gpu_launches": int(launches), "clocks": clocks}
        # ---- CPU baseline + roofline (rank 0, N=1 only) ----
        if world == 1 and not args.no_cpu:
            from oracle import oracle
            sample_spp = max(1, min(wl["spp"], args.cpu_spp))
            cnt, dt = cpu_sample(bvr, oracle, scene, wl, sample_spp)
            line["cpu_baseline"] = {"value": cnt["rays"] / dt / 1e6, "unit": UNIT, "cores": host_threads(),
                                    "kind": "port",
                                    "sample": f"{W}x{H} x {sample_spp} spp of {wl['spp']} (same scene, camera, seed, bounces); "
                                              "restated C++ CPU baseline, not lavapipe"}
            fpr, bpr = flops_per_ray(cnt), bytes_per_ray(cnt)
            avg_kernel_s = float(np.mean(kernel_ms)) * 1e-3
            rays_per_launch = rays / args.steps
            achieved = fpr * rays_per_launch / avg_kernel_s / 1e12
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tpath):
                try:
                    traffic = json.load(open(tpath)).get(args.workload)
                except Exception:
                    traffic = None
            fp32_roof = {"bound": "fp32", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                                "frac": achieved / fp32_peak if fp32_peak else None, "traffic": traffic,
                                "peak_source": "measured live: bvr_bench_fp32_peak FFMA probe (MEASURED_PEAKS.json has no FP32 figure)",
                                "flops_per_ray": fpr, "kernel_ms": avg_kernel_s * 1e3,
                                "counters_per_ray": {"inner_visits": cnt["inner_visits"] / cnt["rays"],
                                                     "sphere_tests": cnt["sphere_tests"] / cnt["rays"],
                                                     "hits_shaded": cnt["hits_shaded"] / cnt["rays"],
                                                     "rays_per_path": cnt["rays"] / cnt["paths"]}}
            hbm_peak = None
            ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(ppath):
                hbm_peak = json.load(open(ppath)).get("hbm_gbs")
            hbm_peak_src = "measured (MEASURED_PEAKS.json)" if hbm_peak else "fallback (B200_PROFILING.md)"
            hbm_peak = hbm_peak or 6650.0
            hbm_achieved = bpr * rays_per_launch / avg_kernel_s / 1e9
            hbm_roof = {"bound": "hbm", "achieved": hbm_achieved, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": hbm_achieved / hbm_peak, "traffic": traffic, "bytes_per_ray": bpr,
                                    "peak_source": hbm_peak_src,
                                    "note": "reference-layout bytes of the reference-order traversal, each fetched once "
                                            "(SURVEY 8d); on C2 they are served from shared memory (HBM is not the bound); on "
                                            "C4 the kernel's near-first 4-wide walk fetches ~3x fewer nodes than that count and "
                                            "97 % of its fetches hit L2 (ncu: 6.7 TB/s L2->SM, L1 wavefront pipe 83 % busy), "
                                            "which is why frac exceeds 1"}
            # the scene of C4 (160 MB in reference layout) exceeds shared memory and L2: node fetches bound it
            if args.workload == "c4":
                line["roofline"], line["roofline_fp32"] = hbm_roof, fp32_roof
            else:
                line["roofline"], line["roofline_hbm"] = fp32_roof, hbm_roof
        print(json.dumps(line), flush=True)
    r.close()
    if world > 1:
        dist.destroy_process_group()


def run_c5(args):
    """BASELINE.json configs[4]: animated 10k-sphere scene, per-frame BVH rebuild (host PLOC), dirty-range upload,
    render at the demo defaults (4 spp, 4 bounces, level FallbackRaytraced) with the fused depth composite
    against a synthetic raster colour/depth.  Reports the per-frame split."""
    import torch

    import bevyray_b200 as bvr
    from bevyray_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — bevyray_b200 has no CPU fallback")
    W, H, frames = 1280, 720, args.frames
    scene = bvr.Scene.random(11, 10000, 43.0, 0.05, 0.25)
    cam = bvr.make_camera(position=(0.0, 0.0, 40.0), target=(0.0, 0.0, 0.0), aspect=W / H, sample_count=4, bounces=4)
    ctx = bvr.Context(0)
    rs = np.random.RandomState(0)
    raster = torch.from_numpy(rs.rand(H, W, 4).astype(np.float32)).pin_memory().numpy()
    depth = torch.from_numpy((rs.rand(H, W) * 0.004).astype(np.float32)).pin_memory().numpy()
    out = {"rgba": torch.empty((H, W, 4), dtype=torch.float32).pin_memory().numpy()}
    opts = bvr.make_options(W)
    ctx.upload_scene(scene.models, scene.materials, scene.nodes)
    prev_models, prev_nodes = scene.models.copy(), scene.nodes.copy()
    t_build = t_upload = t_render = gpu_build_ms = 0.0
    rays = h2d = 0
    t_all0 = time.perf_counter()
    for f in range(frames):
        t0 = time.perf_counter()
        scene.animate(f + 1)                                  # closed-form motion + PLOC rebuild (host)
        t1 = time.perf_counter()
        m, n = scene.models, scene.nodes
        # dirty model ranges: runs of changed models, bridged over gaps < 16 (one copy per run)
        dm = np.nonzero((m.view(np.uint8).reshape(-1, 32) != prev_models.view(np.uint8).reshape(-1, 32)).any(axis=1))[0]
        ranges = []
        if len(dm):
            start = prev = int(dm[0])
            for i in dm[1:]:
                i = int(i)
                if i - prev > 16:
                    ranges.append((capi.ARRAY_MODELS, start, prev - start + 1))
                    start = i
                prev = i
            ranges.append((capi.ARRAY_MODELS, start, prev - start + 1))
        if args.gpu_bvh:
            # models only travel; the library rebuilds the BVH on the GPU (bvr_upload_scene_gpu_bvh).  The host
            # PLOC time inside scene.animate() is then not part of the frame: it is subtracted below.
            st0 = ctx.stats()["h2d_bytes"]
            ctx.upload_scene_gpu_bvh(m, scene.materials, ranges if f > 0 else None)
            st1 = ctx.stats()
            h2d += st1["h2d_bytes"] - st0
            gpu_build_ms += st1["last_upload_ms"]
            prev_models = m.copy()
            t2 = time.perf_counter()
            ctx.render(cam, 2, bvr.make_window((0.37 + 0.013 * f) % 1.0, H), opts, raster, depth, want=("rgba",), out=out)
            t3 = time.perf_counter()
            rays += ctx.stats()["rays"]
            t_build += t1 - t0
            t_upload += t2 - t1
            t_render += t3 - t2
            continue
        # ... plus the span of changed BVH nodes
        dn = np.nonzero((n.view(np.uint8).reshape(-1, 48) != prev_nodes.view(np.uint8).reshape(-1, 48)).any(axis=1))[0]
        if len(dn):
            ranges.append((capi.ARRAY_BVH_NODES, int(dn.min()), int(dn.max() - dn.min() + 1)))
        st0 = ctx.stats()["h2d_bytes"]
        ctx.upload_scene(m, scene.materials, n, ranges)
        h2d += ctx.stats()["h2d_bytes"] - st0
        prev_models, prev_nodes = m.copy(), n.copy()
        t2 = time.perf_counter()
        ctx.render(cam, 2, bvr.make_window((0.37 + 0.013 * f) % 1.0, H), opts, raster, depth, want=("rgba",), out=out)
        t3 = time.perf_counter()
        st = ctx.stats()
        rays += st["rays"]
        t_build += t1 - t0
        t_upload += t2 - t1
        t_render += t3 - t2
    total = time.perf_counter() - t_all0
    if args.gpu_bvh:
        total -= t_build      # host animate+PLOC is bench scaffolding in this mode (the tree comes from the GPU)
    line = {"metric": "frame ms, animated 10k spheres 1280x720 4spp 4 bounces level 2 (BVH rebuild + dirty upload + render + composite)",
            "value": total / frames * 1e3, "unit": "ms/frame", "n_gpus": 1, "steps": frames, "warmup": 0,
            "ms_per_step": total / frames * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 animated 10k random spheres, 300 frames, per-frame host PLOC rebuild, dirty-range upload, "
                                   "fused depth composite vs synthetic raster"},
            "split_ms": {"host_bvh_build": t_build / frames * 1e3, "dirty_detect_and_upload": t_upload / frames * 1e3,
                         "render_with_host_io": t_render / frames * 1e3},
            "bvh": "GPU LBVH (bvr_upload_scene_gpu_bvh)" if args.gpu_bvh else "host PLOC (csrc/host/ploc.cpp)",
            "gpu_upload_and_build_ms": gpu_build_ms / frames if args.gpu_bvh else None,
            "mrays_per_s": rays / total / 1e6, "scene_h2d_bytes_per_frame": h2d / frames,
            "full_scene_bytes": int(scene.models.nbytes + scene.materials.nbytes + scene.nodes.nbytes)}
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--gpu-bvh", action="store_true", help="c5: build the BVH on the GPU instead of the host PLOC")
    ap.add_argument("--frames", type=int, default=300, help="frames of the animated workload (c5)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "megakernel", "wavefront", "cta-wavefront"])
    ap.add_argument("--reference-order", action="store_true", help="reference traversal order (raytrace.wgsl:313-346)")
    ap.add_argument("--shard", default="samples", choices=["samples", "tiles"])
    ap.add_argument("--strip-rows", type=int, default=4)
    ap.add_argument("--cpu-spp", type=int, default=0,
                    help="samples per pixel of the bounded CPU-baseline sample (0 = per workload: about 10-30 s of CPU work)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--spp", type=int, default=0, help="PROFILING ONLY: override samples per pixel (the line is then not a bench value)")
    args = ap.parse_args()
    if args.cpu_spp <= 0:
        # ~17 Mrays/s (C2/C3) and ~1 Mrays/s (C4) on 16 host cores: 32 spp of C2 = 165 M rays ~ 10 s, 4 spp of C4 ~ 25 s
        args.cpu_spp = {"c1": 1, "c2": 32, "c3": 8, "c4": 4}.get(args.workload, 4)
    if args.spp:
        for wl in WORKLOADS.values():
            wl["spp"] = args.spp
            wl["name"] += f" [PROFILING OVERRIDE spp={args.spp}: not a bench value]"
    elif args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and args.impl == "ours":
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    if args.workload == "c5":
        if int(os.environ.get("RANK", "0")) == 0:
            run_c5(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

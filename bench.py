#!/usr/bin/env python
"""bench.py — Mrays/s and frame time of the path-tracing hot path on B200 (contract: DESIGN.md §6).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (default N=1)
  python bench.py --impl reference --gpus N --steps K ...   the reference's algorithm on the host cores
                                                            (CPU oracle; the WGSL shader itself cannot run
                                                            in this image: no Rust toolchain, no Vulkan)
A step is one frame of the workload (all samples of every pixel).  With N > 1 the SAME frame is shared out over the
ranks (strong scaling): by samples (default: rank g renders spp_g of the spp samples with its own seed, weighted NCCL
reduce) and, reported under extra.tiles, by interleaved row strips (bit-identical to one GPU, NCCL all-gather).  The
sharded frame is checked against single-GPU renders outside the timed region ("multi_gpu_parity")."""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRICS = {
    "c1": "Mrays/s, default scene 1280x720 1 spp 4 bounces (ray = one raycast() call, raytrace.wgsl:190)",
    "c2": "Mrays/s, RTIOW final scene 1920x1080 100 spp 10 bounces (ray = one raycast() call, raytrace.wgsl:190)",
    "c3": "Mrays/s, RTIOW final scene 3840x2160 1000 spp 10 bounces (ray = one raycast() call)",
    "c4": "Mrays/s, synthetic 2^20 random spheres 1920x1080 64 spp 10 bounces (ray = one raycast() call)",
}
UNIT = "Mrays/s"
SCENE_SEED = 1
BASE_SEED = 0.37
FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_rtiow.npz")

WORKLOADS = {
    # BASELINE.json configs[1]: RTIOW book-1 final scene, 1920x1080, 100 spp, 10 bounces, level Pure
    "c2": dict(name="C2 rtiow-final 1920x1080 100spp 10 bounces, book camera (13,2,3)->(0,0,0) vfov 20deg",
               width=1920, height=1080, spp=100, bounces=10, camera="book"),
    # BASELINE.json configs[3]: traversal / memory-bound stress, scene does not fit in shared memory
    "c4": dict(name="C4 synthetic 2^20 random spheres (cube side 200, r in [0.05,0.25], 80/15/5 % materials), "
                    "1920x1080 64spp 10 bounces, camera (0,0,130)->(0,0,0) fov pi/4",
               width=1920, height=1080, spp=64, bounces=10, camera="c4", scene=("random", 7, 1 << 20, 200.0, 0.05, 0.25)),
    # BASELINE.json configs[2]: 4K, 1000 spp, tile- or sample-sharded over the GPUs of the box
    "c3": dict(name="C3 rtiow-final 3840x2160 1000spp 10 bounces, book camera (13,2,3)->(0,0,0) vfov 20deg",
               width=3840, height=2160, spp=1000, bounces=10, camera="book"),
    # BASELINE.json configs[0] (plumbing / parity case)
    "c1": dict(name="C1 default scene 1280x720 1spp 4 bounces, repo camera (0,0,5)->(0,0,0) fov pi/4",
               width=1280, height=720, spp=1, bounces=4, camera="repo"),
}
BVH_NOTE = "host PLOC (restated obvhs call, extract.rs:316-321)"
L2_NOTE = "GPU arm: 256 MiB buffer written between timed steps (L2 flush); CPU arm: not applicable"


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


def sharding_note(world, shard, wl, strip_rows):
    if world == 1:
        return "none"
    if shard == "samples":
        base, rem = divmod(wl["spp"], world)
        per = (f"{base} per rank + one more on the 8x4 tiles of class (tx + ty + rank) % {world} < {rem} (every pixel gets {wl['spp']}, "
               "every rank the same work)") if rem and base else str([base + (1 if g < rem else 0) for g in range(world)])
        return (f"samples: the frame's {wl['spp']} spp shared out as {per}, distinct seed per rank, partial frames weighted "
                "by their share inside the render kernel, which stores them straight into the rank's slot on rank 0 (peer memory "
                "over NVLink); one small all-reduce, then rank 0 adds the slots in rank order")
    return f"tiles: {strip_rows}-row strips interleaved over ranks, NCCL all_gather + de-interleave kernel"


def config_for(key, wl, spheres, world, shard, strip_rows, gpu_bvh=False):
    """Identical in both arms (the driver compares them): what is computed, not how."""
    return {"workload": wl["name"], "scene_seed": SCENE_SEED, "random_seed": BASE_SEED, "spheres": int(spheres),
            "bvh": "GPU LBVH (bvr_upload_scene_gpu_bvh)" if gpu_bvh else BVH_NOTE,
            "sharding": sharding_note(world, shard, wl, strip_rows), "l2": L2_NOTE}


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
            self._stop_evt.wait(0.02)

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


# ------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (restated WGSL) on the host cores.  Needs nothing from the product library for C1-C3: scene and
# camera bytes come from tests/golden/bench_rtiow.npz (tests/golden/make_bench_fixture.py wrote them).
# ------------------------------------------------------------------------------------------------------------------
def fixture_scene(oracle, key, wl):
    if wl.get("scene") is None and os.path.exists(FIXTURE):
        z = np.load(FIXTURE)
        models = np.ascontiguousarray(z["models"]).view(oracle.MODEL_DTYPE).reshape(-1)
        materials = np.ascontiguousarray(z["materials"]).view(oracle.MATERIAL_DTYPE).reshape(-1)
        nodes = np.ascontiguousarray(z["nodes"]).view(oracle.BVH_NODE_DTYPE).reshape(-1)
        cam = oracle.Camera.from_buffer_copy(z["camera_" + key].tobytes())
        return models, materials, nodes, cam
    # C4 (160 MB of scene) is generated, not stored: this one needs the host layer of the product library
    import bevyray_b200 as bvr
    sc = make_scene(bvr, wl)
    cam = oracle.Camera.from_buffer_copy(bytes(make_cam(bvr, wl)))
    return sc.models.copy(), sc.materials.copy(), sc.nodes.copy(), cam


def cpu_sample(oracle, scene, cam, wl, spp, threads=None):
    """Times the CPU oracle on a bounded sample of the workload: the same frame at `spp` samples."""
    threads = host_threads() if threads is None else threads
    models, materials, nodes = scene
    c = oracle.Camera.from_buffer_copy(bytes(cam))
    c.sample_count = spp
    win = oracle.make_window(BASE_SEED, wl["height"])
    t0 = time.perf_counter()
    _, cnt = oracle.render(models, materials, nodes, c, oracle.make_level(3), win, wl["width"], threads=threads)
    dt = time.perf_counter() - t0
    return cnt, dt


def run_reference(args):
    """--impl reference: the reference's algorithm (CPU restatement of the WGSL shader, oracle/) on all host
    cores.  The shader itself cannot run here: no Rust toolchain, no Vulkan loader, no lavapipe."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    wl = WORKLOADS[args.workload]
    models, materials, nodes, cam = fixture_scene(oracle, args.workload, wl)
    scene = (models, materials, nodes)
    sample_spp = max(1, min(wl["spp"], args.cpu_spp))
    cores = host_threads()
    small = dict(wl, width=wl["width"] // 4, height=wl["height"] // 4)
    for _ in range(args.warmup):
        cpu_sample(oracle, scene, cam, small, 1)
    rays, total = 0, 0.0
    for _ in range(args.steps):
        cnt, dt = cpu_sample(oracle, scene, cam, wl, sample_spp)
        rays += cnt["rays"]
        total += dt
    value = rays / total / 1e6
    sample = f"{wl['width']}x{wl['height']} x {sample_spp} spp of {wl['spp']} per step (same scene, camera, seed, bounces)"
    line = {"impl": "reference", "metric": METRICS[args.workload], "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": config_for(args.workload, wl, len(models), args.gpus, args.shard, args.strip_rows),
            "impl_detail": {"what": "CPU restatement of the reference WGSL shader (oracle/), OpenMP over rows, reference-order "
                                    "traversal; a step is the bounded sample below, ms_per_step is MEASURED for that sample",
                            "ms_per_full_frame_extrapolated": total / args.steps * 1e3 * (wl["spp"] / sample_spp),
                            "rays_per_step": rays // args.steps},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------------
class Dist:
    """torch.distributed plumbing (or its single-process stand-in)."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — bevyray_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_sum(self, values):
        """(max over ranks, sum over ranks) of a list of floats."""
        if not self.dist:
            return list(values), list(values)
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        tmax = t.clone()
        self.dist.all_reduce(tmax, op=self.dist.ReduceOp.MAX)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(x) for x in tmax], [float(x) for x in t]

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def time_frames(d, r, step, flush, steps, warmup):
    """W untimed warm-up steps, then K steps timed on the device (CUDA events on the launching stream, L2 flushed
    between steps), bracketed by barrier + synchronize.  Returns per-rank-max total ms, summed rays / launches."""
    torch = d.torch
    for _ in range(warmup):
        step()
        flush.zero_()
    d.barrier()
    launches0 = r.ctx.stats()["kernel_launches"]
    sampler = ClockSampler(d.local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    rays, kernel_ms = 0, []
    d.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        ev[i][0].record()
        step()
        ev[i][1].record()
        flush.zero_()
        st = r.ctx.stats()          # synchronises the stream; reads the device ray counter
        rays += st["rays"]
        kernel_ms.append(st["last_render_ms"])
    d.barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    launches = r.ctx.stats()["kernel_launches"] - launches0
    mx, sm = d.max_sum([total_ms, float(rays), float(launches)])
    return {"total_ms": mx[0], "rays": int(sm[1]), "launches": int(sm[2]), "kernel_ms": kernel_ms, "wall_s": wall,
            "clocks": clocks, "my_rays": rays}


def time_e2e(d, r, scene, step, host_rgba, steps, gpu_bvh=False, single_call=None):
    """The same frames through the public API with HOST buffers: scene upload (the reference re-uploads every frame,
    pipeline.rs:136-138) + render + read-back of the frame into pinned host memory, wall clock, max over ranks."""
    torch = d.torch

    def one():
        if gpu_bvh:
            r.ctx.upload_scene_gpu_bvh(scene.models, scene.materials)
        else:
            r.upload_scene(scene.models, scene.materials, scene.nodes)
        if single_call is not None:
            single_call()
        else:
            step()
            torch.cuda.synchronize()
    for _ in range(2 if steps >= 3 else 1):   # untimed: the first host-buffer frame of a size allocates the library's pinned staging
        one()
    rays = 0
    d.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        if gpu_bvh:
            r.ctx.upload_scene_gpu_bvh(scene.models, scene.materials)
        else:
            r.upload_scene(scene.models, scene.materials, scene.nodes)
        if single_call is not None:
            single_call()                                   # bvr_render: H2D, kernels, D2H, sync in one call
        else:
            frame = step()
            if d.rank == 0:
                host_rgba.copy_(frame, non_blocking=True)
            torch.cuda.synchronize()
        rays += r.ctx.stats()["rays"]
    d.barrier()
    dt = time.perf_counter() - t0
    mx, sm = d.max_sum([dt, float(rays)])
    return mx[0], int(sm[1])


def check_sample_parity(d, bvr, r, scene, wl, W, H, kernel, traversal, frame):
    """Rank 0: the reduced frame must equal the sum of the ranks' weighted partial frames, each rendered here on one GPU
    with that rank's seed, sample count, flags and weight (ShardedRenderer.last_plan) and added up in rank order.  The fused exchange over peer memory adds
    the slots in that same order, so the two must agree BIT FOR BIT; only the NCCL fallback (no peer memory), whose
    summation order is its own, gets a few ulps (2e-6 absolute on values in [0,1])."""
    from bevyray_b200.distributed import seed_for_rank, split_samples
    torch = d.torch
    if d.rank != 0:
        return None
    ctx = bvr.Context(d.local_rank)
    ctx.upload_scene(scene.models, scene.materials, scene.nodes)
    acc = None
    part = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    for g, (count, flags, weight) in enumerate(r.last_plan):       # what rank g rendered: sample count, flags, weight
        if count == 0:
            continue
        cam = make_cam(bvr, wl, count)
        opts = bvr.make_options(W, kernel, traversal, output_weight=weight)
        opts.flags |= flags
        ctx.render_device(cam, 3, bvr.make_window(seed_for_rank(BASE_SEED, g, d.world, "samples"), H), opts, rgba=part.data_ptr())
        ctx.sync()
        acc = part.clone() if acc is None else acc + part
    torch.cuda.synchronize()
    diff = float((acc - frame).abs().max())
    ctx.close()
    tol = 0.0 if r._slots else 2e-6
    return "ok" if diff <= tol else f"FAILED: sample-sharded frame differs from the weighted per-seed frames by {diff:.3g}", diff


def check_tile_parity(d, bvr, scene, wl, W, H, kernel, traversal, frame):
    """Rank 0: the tile-gathered frame must be bit-identical to the frame one GPU renders alone."""
    torch = d.torch
    if d.rank != 0:
        return None
    ctx = bvr.Context(d.local_rank)
    ctx.upload_scene(scene.models, scene.materials, scene.nodes)
    one = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
    ctx.render_device(make_cam(bvr, wl), 3, bvr.make_window(BASE_SEED, H), bvr.make_options(W, kernel, traversal), rgba=one.data_ptr())
    ctx.sync()
    same = bool(torch.equal(one.view(torch.int32), frame.view(torch.int32)))
    ctx.close()
    return "ok" if same else "FAILED: tile-gathered frame is not bit-identical to the single-GPU frame"


def load_json(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return {}


def measured_peaks(capi, device):
    pk = {"fp32_tflops": None, "l2_gbs": None, "hbm_gbs": None, "hbm_source": None}
    v = C.c_float()
    if capi.lib.bvr_bench_fp32_peak(device, C.byref(v)) == 0:
        pk["fp32_tflops"] = float(v.value)
    v = C.c_float()
    if capi.lib.bvr_bench_l2_bandwidth(device, C.byref(v)) == 0:
        pk["l2_gbs"] = float(v.value)
    hbm = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json")).get("hbm_gbs")
    pk["hbm_gbs"] = hbm or 6650.0
    pk["hbm_source"] = "measured (MEASURED_PEAKS.json)" if hbm else "fallback (B200_PROFILING.md)"
    return pk


def run_leg(d, bvr, capi, key, args, peaks, steps, warmup, with_cpu, gpu_bvh=False):
    """One workload on the GPUs of this job: device-timed value, e2e, roofline.  Returns the line (rank 0) or None."""
    from bevyray_b200.distributed import ShardedRenderer
    torch = d.torch
    wl = WORKLOADS[key]
    W, H = wl["width"], wl["height"]
    scene = make_scene(bvr, wl)
    kernel = {"auto": capi.KERNEL_AUTO, "megakernel": capi.KERNEL_MEGAKERNEL, "wavefront": capi.KERNEL_WAVEFRONT,
              "cta-wavefront": capi.KERNEL_CTA_WAVEFRONT}[args.kernel]
    traversal = capi.TRAVERSAL_REFERENCE_ORDER if args.reference_order else capi.TRAVERSAL_AUTO
    cam = make_cam(bvr, wl)
    r = ShardedRenderer(d.local_rank, d.rank, d.world, mode=args.shard, strip_rows=args.strip_rows)
    if gpu_bvh:
        r.ctx.upload_scene_gpu_bvh(scene.models, scene.materials)
    else:
        r.upload_scene(scene.models, scene.materials, scene.nodes)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=r.device)   # > 126 MB L2

    def step():
        return r.render_frame(cam, 3, BASE_SEED, W, H, kernel, traversal,
                              split_samples_of=wl["spp"] if args.shard == "samples" else None)

    t = time_frames(d, r, step, flush, steps, warmup)
    value = t["rays"] / (t["total_ms"] * 1e-3) / 1e6
    frame = step()
    torch.cuda.synchronize()

    # ---- end to end: host buffers in, host frame out, every step ----
    host_rgba = torch.empty((H, W, 4), dtype=torch.float32).pin_memory()
    host_np = host_rgba.numpy()
    scene_bytes = scene.models.nbytes + scene.materials.nbytes + (0 if gpu_bvh else scene.nodes.nbytes)
    single = None
    if d.world == 1:
        opts, win = bvr.make_options(W, kernel, traversal), bvr.make_window(BASE_SEED, H)
        single = lambda: r.ctx.render(cam, 3, win, opts, want=("rgba",), out={"rgba": host_np})   # noqa: E731
    e2e_s, e2e_rays = time_e2e(d, r, scene, step, host_rgba, steps, gpu_bvh, single)

    # ---- multi-GPU: the sharded frame against single-GPU renders (outside every timed region) ----
    parity = None
    tiles = None
    if d.world > 1:
        if args.shard == "samples":
            res = check_sample_parity(d, bvr, r, scene, wl, W, H, kernel, traversal, frame)
            parity = {"samples": res[0], "samples_max_abs_diff": res[1]} if res else None
            if not args.no_extra:
                # the exact sharding, same frame: interleaved row strips, all-gather
                rt = ShardedRenderer(d.local_rank, d.rank, d.world, mode="tiles", strip_rows=args.strip_rows)
                rt.upload_scene(scene.models, scene.materials, scene.nodes)
                tstep = lambda: rt.render_frame(cam, 3, BASE_SEED, W, H, kernel, traversal)   # noqa: E731
                tt = time_frames(d, rt, tstep, flush, steps, warmup)
                tframe = tstep()
                torch.cuda.synchronize()
                tp = check_tile_parity(d, bvr, scene, wl, W, H, kernel, traversal, tframe)
                tiles = {"value": tt["rays"] / (tt["total_ms"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": tt["total_ms"] / steps,
                         "sharding": sharding_note(d.world, "tiles", wl, args.strip_rows), "parity": tp,
                         "limiter": "the per-pixel sequential RNG chain (raytrace.wgsl:89,161-167): one pixel's samples cannot "
                                    "be split over lanes, so the heaviest pixel bounds the frame once every GPU holds few pixels per lane"}
                if parity is not None:
                    parity["tiles"] = tp
                rt.close()
        else:
            tp = check_tile_parity(d, bvr, scene, wl, W, H, kernel, traversal, frame)
            parity = {"tiles": tp} if tp else None

    line = None
    if d.rank == 0:
        line = {"metric": METRICS[key], "value": value, "unit": UNIT, "n_gpus": d.world, "steps": steps, "warmup": warmup,
                "ms_per_step": t["total_ms"] / steps, "higher_is_better": True,
                "scaling": "strong" if d.world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_for(key, wl, len(scene.models), d.world, args.shard, args.strip_rows, gpu_bvh),
                "impl_detail": {"kernel": args.kernel, "traversal": "reference-order" if args.reference_order else "near-first"},
                "frame_ms": t["total_ms"] / steps, "mpaths_per_s": W * H * wl["spp"] * steps / (t["total_ms"] * 1e-3) / 1e6,
                "rays_per_step": t["rays"] // steps, "wall_s_timed_region": t["wall_s"],
                "e2e": {"value": e2e_rays / e2e_s / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(scene_bytes) * d.world,
                        "d2h_bytes_per_step": int(host_np.nbytes), "ms_per_step": e2e_s / steps * 1e3},
                "gpu_launches": t["launches"], "clocks": t["clocks"]}
        if d.world > 1:
            line["multi_gpu_parity"] = ("ok" if parity and all(v == "ok" for k, v in parity.items() if not k.endswith("diff"))
                                        else "FAILED")
            line["multi_gpu_parity_detail"] = parity
            line["limiter"] = ("samples: fixed per-frame cost that does not shrink with 1/N — kernel launch + scene staging into "
                               "shared memory per CTA, the pixel-queue tail (the heaviest pixel is a chain of spp_g samples), one small "
                               f"all-reduce and rank 0's sum over the {d.world} slots of {W * H * 16 / 1e6:.0f} MB"
                               if args.shard == "samples" else tiles and tiles["limiter"])
            if tiles:
                line.setdefault("extra", {})["tiles"] = tiles
        avg_kernel_s = float(np.mean(t["kernel_ms"])) * 1e-3
        rays_per_launch = t["my_rays"] / steps
        traffic = load_json(os.path.join(ROOT, "profiles", "traffic.json"))
        if d.world == 1 and key == "c4":
            # Scenes walked out of L2: the bound is L2 -> SM bandwidth.  Algorithmic bytes = the bytes of
            # node records and spheres the near-first walk needs per ray, counted by ncu as L2 sectors of this kernel
            # (profiles/traffic.json, per launch), against the L2 bandwidth probe measured live.
            l2_bytes = traffic.get("c4_l2_bytes")
            if l2_bytes and peaks["l2_gbs"]:
                ach = l2_bytes / avg_kernel_s / 1e9
                line["roofline"] = {"bound": "l2", "achieved": ach, "peak": peaks["l2_gbs"], "unit": "GB/s", "frac": ach / peaks["l2_gbs"],
                                    "traffic": traffic.get("c4"), "kernel_ms": avg_kernel_s * 1e3,
                                    "peak_source": "measured live: bvr_bench_l2_bandwidth (32 MiB L2-resident buffer streamed by every CTA)",
                                    "bytes_source": "lts__t_sectors x 32 B of one full-frame launch (ncu, profiles/traffic.json: c4_l2_bytes); "
                                                    "traffic = dram__bytes of the same launch"}
                if traffic.get("c4"):
                    hb = traffic["c4"] / avg_kernel_s / 1e9
                    line["roofline_hbm"] = {"bound": "hbm", "achieved": hb, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                            "frac": hb / peaks["hbm_gbs"], "traffic": traffic["c4"], "peak_source": peaks["hbm_source"]}
        if d.world == 1 and with_cpu:
            from oracle import oracle
            sample_spp = max(1, min(wl["spp"], args.cpu_spp or {"c1": 1, "c2": 32, "c3": 8, "c4": 4}[key]))
            cnt, dt = cpu_sample(oracle, (scene.models, scene.materials, scene.nodes), cam, wl, sample_spp)
            line["cpu_baseline"] = {"value": cnt["rays"] / dt / 1e6, "unit": UNIT, "cores": host_threads(), "kind": "port",
                                    "sample": f"{W}x{H} x {sample_spp} spp of {wl['spp']} (same scene, camera, seed, bounces); "
                                              "restated C++ CPU baseline, not lavapipe"}
            fpr, bpr = flops_per_ray(cnt), bytes_per_ray(cnt)
            achieved = fpr * rays_per_launch / avg_kernel_s / 1e12
            fp32_roof = {"bound": "fp32", "achieved": achieved, "peak": peaks["fp32_tflops"], "unit": "TFLOP/s",
                         "frac": achieved / peaks["fp32_tflops"] if peaks["fp32_tflops"] else None, "traffic": traffic.get(key),
                         "peak_source": "measured live: bvr_bench_fp32_peak FFMA probe (MEASURED_PEAKS.json has no FP32 figure)",
                         "flops_per_ray": fpr, "kernel_ms": avg_kernel_s * 1e3,
                         "counters_per_ray": {"inner_visits": cnt["inner_visits"] / cnt["rays"],
                                              "sphere_tests": cnt["sphere_tests"] / cnt["rays"],
                                              "hits_shaded": cnt["hits_shaded"] / cnt["rays"],
                                              "rays_per_path": cnt["rays"] / cnt["paths"]}}
            if key == "c4":
                line["roofline_fp32"] = fp32_roof
            else:
                # fp32 is the bound of scenes staged in shared memory (C1-C3); the reference-layout bytes the same launch
                # would move are served from shared memory, so an HBM figure for them would exceed the HBM peak: it is
                # reported as bytes per ray only
                line["roofline"] = fp32_roof
                line["reference_layout_bytes_per_ray"] = bpr
    r.close()
    del flush
    torch.cuda.empty_cache()
    return line


def dirty_model_ranges(capi, models, prev_models, bridge=16):
    """Runs of changed 32-byte model records as (array, first, count), bridged over gaps of up to `bridge` clean ones
    (one copy per run); numpy only — this is what a caller's change detection costs per frame."""
    changed = (models.view(np.uint8).reshape(-1, 32) != prev_models.view(np.uint8).reshape(-1, 32)).any(axis=1)
    dm = np.flatnonzero(changed)
    if not len(dm):
        return []
    cut = np.flatnonzero(np.diff(dm) > bridge)
    starts = np.concatenate(([dm[0]], dm[cut + 1]))
    ends = np.concatenate((dm[cut], [dm[-1]]))
    return [(capi.ARRAY_MODELS, int(a), int(b - a + 1)) for a, b in zip(starts, ends)]


def run_c5(args, frames, mode):
    """BASELINE.json configs[4]: animated 10k-sphere scene, per-frame BVH rebuild, dirty-range upload, render at the demo
    defaults (4 spp, 4 bounces, level FallbackRaytraced) with the fused depth composite against a synthetic raster
    colour/depth; host buffers in and out every frame.  mode:
      "host"      per-frame host PLOC build (the reference's flow, extract.rs:316-321) + dirty-range upload + bvr_render
      "gpu"       models only travel, bvr_upload_scene_gpu_bvh builds the tree on the GPU, bvr_render
      "pipelined" as "gpu" with two contexts: frame f renders asynchronously (bvr_render_async) in one while frame f+1's
                  models are uploaded and its tree is built in the other; a frame is complete when its pixels are in host memory"""
    import torch

    import bevyray_b200 as bvr
    from bevyray_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — bevyray_b200 has no CPU fallback")
    W, H = 1280, 720
    gpu_bvh = mode != "host"
    scene = bvr.Scene.random(11, 10000, 43.0, 0.05, 0.25)
    cam = bvr.make_camera(position=(0.0, 0.0, 40.0), target=(0.0, 0.0, 0.0), aspect=W / H, sample_count=4, bounces=4)
    refit_every = 16 if mode == "pipelined_refit" else 0     # rebuild every 16th upload of a context, refit in between
    if mode == "pipelined_refit":
        mode = "pipelined"
    n_ctx = 2 if mode == "pipelined" else 1
    ctxs = [bvr.Context(0) for _ in range(n_ctx)]
    rs = np.random.RandomState(0)
    raster = torch.from_numpy(rs.rand(H, W, 4).astype(np.float32)).pin_memory().numpy()
    depth = torch.from_numpy((rs.rand(H, W) * 0.004).astype(np.float32)).pin_memory().numpy()
    outs = [{"rgba": torch.empty((H, W, 4), dtype=torch.float32).pin_memory().numpy()} for _ in range(n_ctx)]
    opts = bvr.make_options(W)
    prevs = []
    for c in ctxs:
        if gpu_bvh:
            c.upload_scene_gpu_bvh(scene.models, scene.materials)
        else:
            c.upload_scene(scene.models, scene.materials, scene.nodes)
        prevs.append(scene.models.copy())
    prev_nodes = scene.nodes.copy()
    t_build = t_upload = t_render = gpu_build_ms = 0.0
    rays = h2d = 0
    sampler = ClockSampler(0)
    sampler.start()
    t_all0 = time.perf_counter()
    for f in range(frames):
        k = f % n_ctx
        ctx = ctxs[k]
        t0 = time.perf_counter()
        scene.animate(f + 1, rebuild_bvh=not gpu_bvh)        # closed-form motion (+ host PLOC rebuild in "host" mode)
        t1 = time.perf_counter()
        m = scene.models
        if mode == "pipelined":
            ctx.sync()                                        # frame f-2 of this context is complete: its buffers are free
            if f >= n_ctx:
                st = ctx.stats()
                rays += st["rays"]
                gpu_build_ms += st["last_upload_ms"]
        ranges = dirty_model_ranges(capi, m, prevs[k])
        win = bvr.make_window((0.37 + 0.013 * f) % 1.0, H)
        st0 = ctx.stats()["h2d_bytes"] if mode != "pipelined" else 0     # (bvr_get_stats waits for the stream)
        if gpu_bvh:
            ctx.upload_scene_gpu_bvh(m, scene.materials, ranges, refit=bool(refit_every) and (f // n_ctx) % refit_every != 0)
            if mode == "pipelined":
                h2d += sum(c for _, _, c in ranges) * 32                  # enqueued only: nothing here waits for the GPU
            else:
                st1 = ctx.stats()
                gpu_build_ms += st1["last_upload_ms"]
                h2d += st1["h2d_bytes"] - st0
        else:
            n = scene.nodes
            dn = np.flatnonzero((n.view(np.uint8).reshape(-1, 48) != prev_nodes.view(np.uint8).reshape(-1, 48)).any(axis=1))
            if len(dn):
                ranges.append((capi.ARRAY_BVH_NODES, int(dn[0]), int(dn[-1] - dn[0] + 1)))
            ctx.upload_scene(m, scene.materials, n, ranges)
            h2d += ctx.stats()["h2d_bytes"] - st0
            prev_nodes = n.copy()
        prevs[k] = m.copy()
        t2 = time.perf_counter()
        ctx.render(cam, 2, win, opts, raster, depth, want=("rgba",), out=outs[k], asynchronous=mode == "pipelined")
        t3 = time.perf_counter()
        if mode != "pipelined":
            rays += ctx.stats()["rays"]
        t_build += t1 - t0
        t_upload += t2 - t1
        t_render += t3 - t2
    for c in ctxs:
        c.sync()
        if mode == "pipelined":
            rays += c.stats()["rays"]
    total = time.perf_counter() - t_all0
    clocks = sampler.stop()
    line = {"metric": "frame ms, animated 10k spheres 1280x720 4spp 4 bounces level 2 (BVH rebuild + dirty upload + render + composite)",
            "value": total / frames * 1e3, "unit": "ms/frame", "n_gpus": 1, "steps": frames, "warmup": 0,
            "ms_per_step": total / frames * 1e3, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C5 animated 10k random spheres (every 4th moves), {frames} frames, per-frame BVH rebuild, dirty-range "
                                   "upload, fused depth composite vs synthetic raster, host buffers in and out"},
            "mode": {"host": "host PLOC (csrc/host/ploc.cpp) + bvr_upload_scene + bvr_render",
                     "gpu": "bvr_upload_scene_gpu_bvh (tree built on the GPU) + bvr_render",
                     "pipelined": "two contexts: bvr_upload_scene_gpu_bvh of frame f+1 overlaps bvr_render_async of frame f"
                                  + ("; the tree is rebuilt every 16th upload of a context and refitted in between "
                                     "(bvr_refit_scene_gpu_bvh)" if refit_every else "")}[mode],
            "split_ms": {"host_animate" + ("_and_bvh_build" if not gpu_bvh else ""): t_build / frames * 1e3,
                         "dirty_detect_and_upload": t_upload / frames * 1e3,
                         ("render_enqueue" if mode == "pipelined" else "render_with_host_io"): t_render / frames * 1e3},
            "gpu_upload_and_build_ms": gpu_build_ms / frames if gpu_bvh else None,
            "e2e": {"value": total / frames * 1e3, "unit": "ms/frame", "h2d_bytes_per_step": int(h2d / frames + W * H * 20),
                    "d2h_bytes_per_step": W * H * 16},
            "mrays_per_s": rays / total / 1e6, "scene_h2d_bytes_per_frame": h2d / frames,
            "full_scene_bytes": int(scene.models.nbytes + scene.materials.nbytes + scene.nodes.nbytes), "clocks": clocks}
    for c in ctxs:
        c.close()
    return line


def run_ours(args):
    import bevyray_b200 as bvr
    from bevyray_b200 import capi

    d = Dist()
    peaks = measured_peaks(capi, d.local_rank) if d.rank == 0 else {"fp32_tflops": None, "l2_gbs": None, "hbm_gbs": None, "hbm_source": None}
    d.barrier()
    line = run_leg(d, bvr, capi, args.workload, args, peaks, args.steps, args.warmup, with_cpu=not args.no_cpu, gpu_bvh=args.gpu_bvh)
    # ---- the other BASELINE configs, short legs in the same driver-run record (N=1 default run only) ----
    if d.world == 1 and args.workload == "c2" and not args.no_extra and not args.spp:
        extra = line.setdefault("extra", {})
        small = argparse.Namespace(**vars(args))
        small.no_extra = True
        for key, steps, warmup in (("c1", 20, 3), ("c3", 1, 1), ("c4", 2, 1)):
            try:
                leg = run_leg(d, bvr, capi, key, small, peaks, steps, warmup, with_cpu=False)
                extra[key] = {k: leg[k] for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "rays_per_step", "e2e",
                                                  "gpu_launches", "clocks", "config", "roofline", "roofline_hbm") if k in leg}
            except Exception as e:   # a leg must never take the headline line down with it
                extra[key] = {"error": repr(e)}
        for name, mode in (("c5_host_bvh", "host"), ("c5_gpu_bvh", "gpu"), ("c5_gpu_bvh_pipelined", "pipelined"),
                           ("c5_gpu_bvh_pipelined_refit", "pipelined_refit")):
            try:
                extra[name] = run_c5(args, 60, mode)
            except Exception as e:
                extra[name] = {"error": repr(e)}
    if d.rank == 0:
        print(json.dumps(line), flush=True)
    d.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--gpu-bvh", action="store_true", help="build the BVH on the GPU instead of the host PLOC")
    ap.add_argument("--frames", type=int, default=300, help="frames of the animated workload (c5)")
    ap.add_argument("--c5-mode", default=None, choices=["host", "gpu", "pipelined", "pipelined_refit"], help="c5: see run_c5")
    ap.add_argument("--kernel", default="auto", choices=["auto", "megakernel", "wavefront", "cta-wavefront"])
    ap.add_argument("--reference-order", action="store_true", help="reference traversal order (raytrace.wgsl:313-346)")
    ap.add_argument("--shard", default="samples", choices=["samples", "tiles"])
    ap.add_argument("--strip-rows", type=int, default=4)
    ap.add_argument("--cpu-spp", type=int, default=0,
                    help="samples per pixel of the bounded CPU-baseline sample (0 = per workload: about 10-30 s of CPU work)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short legs of the other BASELINE configs / shardings")
    ap.add_argument("--spp", type=int, default=0, help="PROFILING ONLY: override samples per pixel (the line is then not a bench value)")
    args = ap.parse_args()
    if args.cpu_spp <= 0 and args.impl == "reference":
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
            print(json.dumps(run_c5(args, args.frames, args.c5_mode or ("gpu" if args.gpu_bvh else "host"))), flush=True)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

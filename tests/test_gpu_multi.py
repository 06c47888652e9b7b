"""Two-GPU runs of ShardedRenderer (skipped on a single-GPU box): tile sharding reproduces the single-GPU frame bit for bit;
sample sharding equals the share-weighted sum of the per-seed frames — through the NCCL reduce (levels 1-2) and through the
exchange fused into the render kernel over peer memory (level Pure), which must also equal the NCCL form bit for bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 256, 144
SPP_TOTAL = 5          # samples mode: shared out as 3 + 2


def _raster():
    rs = np.random.RandomState(5)
    return rs.rand(H, W, 4).astype(np.float32), (rs.rand(H, W) * 0.06).astype(np.float32)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, mode, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import bevyray_b200 as bvr
    from bevyray_b200.distributed import ShardedRenderer
    scene = bvr.Scene.rtiow(1)
    cam = bvr.make_camera(sample_count=4, bounces=6, aspect=W / H)
    r = ShardedRenderer(rank, rank, world, mode="samples" if mode.startswith("samples") else mode, strip_rows=4)
    r.upload_scene(scene.models, scene.materials, scene.nodes)
    if mode == "tiles":
        frame = r.render_frame(cam, 3, 0.37, W, H)
    elif mode.startswith("samples-pure"):
        # level Pure: every rank renders into its slot on rank 0 (three frames: both halves of the double buffer and a
        # re-use), or — "-nccl" — scales and reduces
        r.peer_exchange = not mode.endswith("-nccl")
        for _ in range(3):
            frame = r.render_frame(cam, 3, 0.37, W, H, split_samples_of=SPP_TOTAL)
        assert r.peer_exchange is False or r._slots, "peer memory was not used"
        frame = frame.clone()
        # a second size: the slot buffer is re-made
        small = r.render_frame(bvr.make_camera(sample_count=4, bounces=6, aspect=2.0), 3, 0.37, 64, 32, split_samples_of=SPP_TOTAL)
        torch.cuda.synchronize()
        if rank == 0:
            np.save(os.path.join(out_dir, mode + "-small.npy"), small.cpu().numpy())
    else:
        # level 2: weighted partial frames, colour and depth reduced, composite once on rank 0
        rc, rd = _raster()
        d_rc, d_rd = torch.from_numpy(rc).cuda(), torch.from_numpy(rd).cuda()
        frame = r.render_frame(cam, 2, 0.37, W, H, d_raster_rgba=d_rc.data_ptr(), d_raster_depth=d_rd.data_ptr(),
                               split_samples_of=SPP_TOTAL)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(os.path.join(out_dir, mode + ".npy"), frame.cpu().numpy())
    r.close()
    dist.destroy_process_group()


def _weighted_sum(bvr, ctx, w, h, aspect):
    """The frame two ranks must produce: every rank's partial frame (its seed, its sample_plan entry) rendered here on one
    GPU, added in rank order."""
    from bevyray_b200.distributed import sample_plan, seed_for_rank
    acc = None
    for r, (count, flags, weight) in enumerate(sample_plan(2, 4, SPP_TOTAL)):
        c = bvr.make_camera(sample_count=count, bounces=6, aspect=aspect)
        o = bvr.make_options(w, output_weight=weight)
        o.flags |= flags
        part = ctx.render(c, 3, bvr.make_window(seed_for_rank(0.37, r, 2, "samples"), h), o, want=("rgba",))["rgba"]
        acc = part if acc is None else acc + part
    return acc


def test_two_gpu_sample_sharding_over_peer_memory(tmp_path, bvr):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    for mode in ("samples-pure", "samples-pure-nccl"):
        mp.spawn(_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    fused, nccl = np.load(tmp_path / "samples-pure.npy"), np.load(tmp_path / "samples-pure-nccl.npy")
    scene = bvr.Scene.rtiow(1)
    ctx = bvr.Context(0)
    ctx.upload_scene(scene.models, scene.materials, scene.nodes)
    want = _weighted_sum(bvr, ctx, W, H, W / H)
    assert np.array_equal(fused.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(nccl.view(np.uint32), want.view(np.uint32))      # two ranks: one addition, any order
    want_small = _weighted_sum(bvr, ctx, 64, 32, 2.0)
    assert np.array_equal(np.load(tmp_path / "samples-pure-small.npy").view(np.uint32), want_small.view(np.uint32))
    ctx.close()


@pytest.mark.parametrize("mode", ["tiles", "samples"])
def test_two_gpu_sharding(tmp_path, bvr, mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from bevyray_b200.distributed import seed_for_rank
    mp.spawn(_worker, args=(2, _free_port(), mode, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / (mode + ".npy"))
    scene = bvr.Scene.rtiow(1)
    cam = bvr.make_camera(sample_count=4, bounces=6, aspect=W / H)
    ctx = bvr.Context(0)
    ctx.upload_scene(scene.models, scene.materials, scene.nodes)
    if mode == "tiles":
        want = ctx.render(cam, 3, bvr.make_window(0.37, H), bvr.make_options(W), want=("rgba",))["rgba"]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    else:
        from bevyray_b200.distributed import split_samples
        acc = np.zeros((H, W, 4), np.float32)
        dep = np.zeros((H, W), np.float32)
        for r, share in enumerate(split_samples(SPP_TOTAL, 2)):
            c = bvr.make_camera(sample_count=share, bounces=6, aspect=W / H)
            p = ctx.render(c, 3, bvr.make_window(seed_for_rank(0.37, r, 2, "samples"), H), bvr.make_options(W), want=("rgba", "rt_depth"))
            wgt = np.float32(share / float(SPP_TOTAL))
            acc += p["rgba"] * wgt
            dep += p["rt_depth"] * wgt
        rc, rd = _raster()
        with np.errstate(divide="ignore"):
            d = np.where(dep > np.float32(cam.far_plane), np.float32(-1.0), np.float32(cam.near_plane) / dep)
        want = acc.copy()
        want[rd > d] = rc[rd > d]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        from_raster = (got == rc).all(axis=2)
        assert 0 < from_raster.sum() < from_raster.size
    ctx.close()

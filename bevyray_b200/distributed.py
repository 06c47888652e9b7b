"""Multi-GPU frame rendering: one process per GPU, torch.distributed (NCCL) for the exchange step.

Two shardings:
  * "tiles"   — interleaved row strips; every rank renders its strips for all samples with the
                reference's per-pixel RNG stream, so the gathered image is bit-identical to 1 GPU.
                Exchange: all_gather of equal-size shard planes + one de-interleave kernel.
  * "samples" — rank g renders its share of the samples of the whole frame with its own random_seed (what the
                reference does across frames, extract.rs:72-73): spp_g = split_samples(spp, world)[g].  Every rank
                scales its partial average by spp_g / spp, the weighted partials (colour and, for levels 1-2, the
                ray-traced depth) are summed with an NCCL reduce to rank 0, and the depth composite with the raster
                output runs once, on the sum (bvr_composite_device).
torch is plumbing only (device memory, streams, NCCL); every kernel on the render path is in
libbevyray_b200.so."""
import numpy as np
import torch
import torch.distributed as dist

from . import _capi as capi
from .api import Context, make_level, make_options, make_window


def shard_global_rows(height, shard_index, shard_count, strip_rows):
    """Global row of every shard-local row (>= height for padding rows): the host-side statement of
    shard_global_row() in csrc/kernels.cuh, used to check / undo the tile sharding."""
    opts = make_options(1, shard_index=shard_index, shard_count=shard_count, strip_rows=strip_rows)
    rows = capi.lib.bvr_shard_rows(height, opts)
    if shard_count <= 1:
        return np.arange(height)
    ly = np.arange(rows)
    return ((ly // strip_rows) * shard_count + shard_index) * strip_rows + (ly % strip_rows)


def split_samples(spp, world):
    """Samples per rank when `spp` samples are shared out: the first spp % world ranks get one more."""
    return [spp // world + (1 if g < spp % world else 0) for g in range(world)]


def seed_for_rank(base_seed, rank, world, mode):
    """Distinct, deterministic random_seed in [0,1) per rank for sample sharding."""
    if mode != "samples" or world == 1:
        return float(np.float32(base_seed))
    return float(np.float32((base_seed + 0.61803398875 * rank) % 1.0))


class ShardedRenderer:
    def __init__(self, device, rank=0, world=1, mode="samples", strip_rows=4):
        assert mode in ("tiles", "samples")
        self.rank, self.world, self.mode, self.strip_rows = rank, world, mode, strip_rows
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.ctx = Context(device)
        # a dedicated torch stream made current: the library, NCCL and torch.cuda.Event then share one
        # stream (torch's default stream has handle 0, which bvr_set_stream reads as "own stream")
        self.stream = torch.cuda.Stream(self.device)
        torch.cuda.set_stream(self.stream)
        self.ctx.set_stream(self.stream.cuda_stream)
        self._bufs = {}

    def upload_scene(self, models, materials, nodes, ranges=None):
        self.ctx.upload_scene(models, materials, nodes, ranges)

    def seed_for_rank(self, base_seed):
        return seed_for_rank(base_seed, self.rank, self.world, self.mode)

    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        if key not in self._bufs:
            self._bufs[key] = torch.empty(shape, dtype=dtype, device=self.device)
        return self._bufs[key]

    def options(self, width, kernel=capi.KERNEL_AUTO, traversal=capi.TRAVERSAL_AUTO):
        if self.mode == "tiles" and self.world > 1:
            return make_options(width, kernel, traversal, self.rank, self.world, self.strip_rows)
        return make_options(width, kernel, traversal)

    def render_frame(self, camera, level, base_seed, width, height, kernel=capi.KERNEL_AUTO,
                     traversal=capi.TRAVERSAL_AUTO, d_raster_rgba=0, d_raster_depth=0, split_samples_of=None):
        """Enqueues one frame on the current stream.  Returns the device tensor that holds the full
        fp32 RGBA frame on rank 0 (on every rank for "tiles").

        "samples" mode: `camera.sample_count` samples are rendered by THIS rank; with split_samples_of = S the
        frame is the S-sample frame shared out over the ranks (camera.sample_count is overridden by this rank's share,
        and the partial frames are weighted by their share).  Without it every rank contributes sample_count samples
        with equal weight."""
        opts = self.options(width, kernel, traversal)
        win = make_window(self.seed_for_rank(base_seed), height)
        lv = make_level(level)
        rows = self.ctx.shard_rows(height, opts)
        shard = self._buf("shard", (rows, width, 4), torch.float32)
        if self.world == 1 or self.mode == "tiles":
            self.ctx.render_device(camera, lv, win, opts, d_raster_rgba, d_raster_depth, rgba=shard.data_ptr())
            if self.world == 1:
                return shard
            gathered = self._buf("gathered", (self.world, rows, width, 4), torch.float32)
            dist.all_gather_into_tensor(gathered, shard)
            full = self._buf("full", (height, width, 4), torch.float32)
            self.ctx.unshard_device(gathered.data_ptr(), rows * width * 4, full.data_ptr(), width, height, 4,
                                    self.world, self.strip_rows)
            return full
        # ---- samples ----
        cam = camera
        weight = 1.0 / self.world
        if split_samples_of is not None:
            share = split_samples(int(split_samples_of), self.world)[self.rank]
            cam = type(camera).from_buffer_copy(camera)
            cam.sample_count = share
            weight = share / float(split_samples_of)
        composite = int(level) in (1, 2)
        depth = self._buf("depth", (rows, width), torch.float32) if composite else None
        if cam.sample_count == 0:
            shard.zero_()                    # more ranks than samples: this rank contributes nothing
            if composite:
                depth.zero_()
        else:
            opts.flags |= capi.RENDER_DEFER_COMPOSITE
            self.ctx.render_device(cam, lv, win, opts, 0, 0, rgba=shard.data_ptr(),
                                   **({"rt_depth": depth.data_ptr()} if composite else {}))
            self.ctx.axpby_device(shard.data_ptr(), weight, 0, 0.0, shard.numel())       # in-place scale
            if composite:
                self.ctx.axpby_device(depth.data_ptr(), weight, 0, 0.0, depth.numel())
        dist.reduce(shard, dst=0, op=dist.ReduceOp.SUM)
        if composite:
            dist.reduce(depth, dst=0, op=dist.ReduceOp.SUM)
            if self.rank == 0:
                self.ctx.composite_device(camera, lv, shard.data_ptr(), depth.data_ptr(), d_raster_rgba, d_raster_depth,
                                          rows * width)
        return shard

    def close(self):
        self.ctx.close()

"""Multi-GPU frame rendering: one process per GPU, torch.distributed (NCCL) for the exchange step.

Two shardings:
  * "tiles"   — interleaved row strips; every rank renders its strips for all samples with the
                reference's per-pixel RNG stream, so the gathered image is bit-identical to 1 GPU.
                Exchange: all_gather of equal-size shard planes + one de-interleave kernel.
  * "samples" — rank g renders its share of the samples of the whole frame with its own random_seed (what the
                reference does across frames, extract.rs:72-73): spp_g = split_samples(spp, world)[g].
                Level Pure (no composite): the exchange is FUSED into the render kernel over peer memory — every rank
                renders straight into its own slot of a buffer on rank 0 (the kernel weights each pixel by spp_g / spp
                and its stores travel over NVLink as the pixels finish), one tiny all-reduce tells rank 0 that all
                ranks are done, and rank 0 adds the slots up in rank order (bvr_sum_slots_device: fixed order, so the
                frame is a pure function of the partial frames).
                Levels 1-2: every rank scales its partial average by spp_g / spp, the weighted partials (colour and
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


def sample_plan(world, sample_count, split_samples_of=None, balanced=True):
    """What every rank renders of a sample-sharded frame: a list of (sample_count, flags, output_weight), one per rank.

    split_samples_of = S shares S samples out.  When `world` does not divide S and every rank gets at least one sample the
    split is BALANCED: every rank renders S // world samples, and on the 8x4 tiles of class (tx + ty + rank) % world < S %
    world one more (BVR_RENDER_EXTRA_SAMPLE) — every pixel gets exactly S samples over the ranks and every rank the same
    amount of work; a pixel with n samples is weighted n / S inside the kernel (output_weight = 1 / S per sample).
    Otherwise rank g renders split_samples(S, world)[g] samples of every pixel, weighted by its share.
    Without split_samples_of every rank contributes sample_count samples with weight 1 / world."""
    if split_samples_of is None:
        return [(int(sample_count), 0, 1.0 / world)] * world
    total = int(split_samples_of)
    base, rem = divmod(total, world)
    if balanced and rem and base >= 1 and world <= 255:
        return [(base, capi.render_extra_sample_bits(world, g, rem), 1.0 / total) for g in range(world)]
    return [(n, 0, n / float(total)) for n in split_samples(total, world)]


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
        self._slots = None        # (key, base pointer, floats per frame): peer buffer on rank 0, 2 x world frames
        self._slot_phase = 0
        self.peer_exchange = True   # False: always use the NCCL reduce
        self._unbalanced = False    # True once the library refused uneven sample counts for this scene
        self.last_plan = None       # sample_plan() of the last sample-sharded frame

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
        frame is the S-sample frame shared out over the ranks (sample_plan: camera.sample_count is overridden by this
        rank's share, the partial frames are weighted by their share).  Without it every rank contributes sample_count
        samples with equal weight."""
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
        composite = int(level) in (1, 2)
        # (levels 1-2 weight their partial frames after the render, by one number per rank: no uneven sample counts there)
        balanced = (not composite and not self._unbalanced and kernel in (capi.KERNEL_AUTO, capi.KERNEL_MEGAKERNEL)
                    and traversal == capi.TRAVERSAL_AUTO)          # uneven sample counts need the megakernel
        plan = sample_plan(self.world, camera.sample_count, split_samples_of, balanced=balanced)
        self.last_plan = plan
        count, flags, weight = plan[self.rank]
        cam = type(camera).from_buffer_copy(camera)
        cam.sample_count = count
        if not composite:
            # level Pure: the kernel weights every pixel as it stores it — straight into this rank's slot on rank 0 when
            # peer memory is available, into a local plane handed to an NCCL reduce otherwise
            opts.flags |= flags
            opts.output_weight = weight
            try:
                if self.peer_exchange and self.world <= 64 and self._peer_slots(rows, width):
                    mask = sum(1 << g for g, (n, _, _) in enumerate(plan) if n > 0)   # a rank without samples leaves its slot alone
                    return self._render_into_slots(cam, lv, win, opts, rows, width, shard, mask)
                if count == 0:
                    shard.zero_()
                else:
                    self.ctx.render_device(cam, lv, win, opts, 0, 0, rgba=shard.data_ptr())
            except RuntimeError:
                if not (flags & capi.RENDER_EXTRA_SAMPLE):
                    raise
                # the library renders this scene with a kernel that cannot take uneven sample counts (e.g. a tree of 32+
                # levels is walked in reference order): every rank gets the same refusal and re-plans the same way
                self._unbalanced = True
                return self.render_frame(camera, level, base_seed, width, height, kernel, traversal, d_raster_rgba,
                                         d_raster_depth, split_samples_of)
            dist.reduce(shard, dst=0, op=dist.ReduceOp.SUM)
            return shard
        depth = self._buf("depth", (rows, width), torch.float32)
        if count == 0:
            shard.zero_()                    # more ranks than samples: this rank contributes nothing
            depth.zero_()
        else:
            opts.flags |= capi.RENDER_DEFER_COMPOSITE
            self.ctx.render_device(cam, lv, win, opts, 0, 0, rgba=shard.data_ptr(), rt_depth=depth.data_ptr())
            self.ctx.axpby_device(shard.data_ptr(), weight, 0, 0.0, shard.numel())       # in-place scale
            self.ctx.axpby_device(depth.data_ptr(), weight, 0, 0.0, depth.numel())
        dist.reduce(shard, dst=0, op=dist.ReduceOp.SUM)
        dist.reduce(depth, dst=0, op=dist.ReduceOp.SUM)
        if self.rank == 0:
            self.ctx.composite_device(camera, lv, shard.data_ptr(), depth.data_ptr(), d_raster_rgba, d_raster_depth,
                                      rows * width)
        return shard

    # ---- sample sharding over peer memory ----
    def _peer_slots(self, rows, width):
        """Makes sure the slot buffer for frames of this size exists on rank 0 and is mapped on every rank.  Collective.
        Returns False (on every rank) when peer memory is not available: the caller then uses the NCCL reduce."""
        key = (rows, width)
        if self._slots is False:                # peer memory was refused once: stay with the NCCL reduce
            return False
        if self._slots is not None:
            if self._slots[0] == key:
                return True
            self._release_slots()
        floats = rows * width * 4
        ok, ptr = 1, 0
        handle = torch.zeros(capi.PEER_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
        if self.rank == 0:
            try:
                ptr, h = self.ctx.peer_alloc(2 * self.world * floats * 4)      # double-buffered: see _render_into_slots
                handle = torch.tensor(list(h), dtype=torch.uint8, device=self.device)
            except RuntimeError:
                ok = 0
        dist.broadcast(handle, src=0)
        if self.rank != 0:
            try:
                ptr = self.ctx.peer_open(bytes(handle.cpu().numpy().tobytes()))
            except RuntimeError:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if ptr:
                (self.ctx.peer_free if self.rank == 0 else self.ctx.peer_close)(ptr)
            self._slots = False
            return False
        self._slots = (key, ptr, floats)
        self._slot_phase = 0
        return True

    def _release_slots(self):
        if self._slots:
            _, ptr, _ = self._slots
            torch.cuda.synchronize(self.device)
            if self.rank != 0:
                self.ctx.peer_close(ptr)
            dist.barrier()                      # nobody maps the buffer any more
            if self.rank == 0:
                self.ctx.peer_free(ptr)
        self._slots = None

    def _render_into_slots(self, cam, lv, win, opts, rows, width, local, mask):
        """One frame of the fused exchange.  Slots are double-buffered: while rank 0 adds up the slots of frame f the other
        ranks may already be storing frame f+1 into the other half; they reach frame f+2 — the same half again — only
        after the all-reduce of frame f+1, which rank 0 joins after its sum of frame f (stream order)."""
        _, base, floats = self._slots
        phase = self._slot_phase
        half = base + phase * self.world * floats * 4
        if cam.sample_count > 0:
            self.ctx.render_device(cam, lv, win, opts, 0, 0, rgba=half + self.rank * floats * 4)   # (may refuse: see caller)
        self._slot_phase = phase ^ 1
        if ("done", (1,), torch.float32) not in self._bufs:
            self._buf("done", (1,), torch.float32).zero_()
        dist.all_reduce(self._buf("done", (1,), torch.float32))   # every rank's render kernel — and with it its pixel stores — is complete
        if self.rank != 0:
            return local
        full = self._buf("full_samples", (rows, width, 4), torch.float32)
        self.ctx.sum_slots_device(half, floats, self.world, mask, full.data_ptr(), floats)
        return full

    def close(self):
        if self._slots:
            self._release_slots()
        self.ctx.close()

"""Host logic of the sample sharding (CPU): sample_plan() must hand every pixel exactly S samples over the ranks, give every
rank the same work when the rank count does not divide S, and weight the partial frames so that the weights of a pixel add
up to one."""
import numpy as np
import pytest

from bevyray_b200 import _capi as capi
from bevyray_b200.distributed import sample_plan, split_samples


def extra_tiles(flags, tiles_x, tiles_y):
    """The tile classes BVR_RENDER_EXTRA_SAMPLE selects (include/bevyray_b200.h), as a boolean tile map."""
    if not flags & capi.RENDER_EXTRA_SAMPLE:
        return np.zeros((tiles_y, tiles_x), bool)
    modulus, phase, count = (flags >> 8) & 0xff, (flags >> 16) & 0xff, (flags >> 24) & 0xff
    ty, tx = np.mgrid[0:tiles_y, 0:tiles_x]
    return ((tx + ty + phase) % modulus) < count


@pytest.mark.parametrize("world", [1, 2, 3, 4, 7, 8, 16, 64])
@pytest.mark.parametrize("total", [1, 5, 64, 100, 1000])
def test_every_pixel_gets_all_samples_and_ranks_get_even_work(world, total):
    plan = sample_plan(world, 0, total)
    assert len(plan) == world
    tiles_x, tiles_y = 240, 270            # 1920 x 1080
    samples = np.zeros((tiles_y, tiles_x), np.int64)
    weights = np.zeros((tiles_y, tiles_x), np.float64)
    work = []
    for count, flags, weight in plan:
        n = count + extra_tiles(flags, tiles_x, tiles_y)
        samples += n
        # the kernel weights a pixel of an uneven-sample frame by n x output_weight, an even one by output_weight
        weights += n * weight if flags & capi.RENDER_EXTRA_SAMPLE else (weight if count else 0.0)
        work.append(int(n.sum()))
    assert (samples == total).all()
    assert np.allclose(weights, 1.0, atol=1e-12)
    base, rem = divmod(total, world)
    if rem and base:
        # balanced: every rank renders base samples everywhere and one more on rem / world of the tiles
        assert all(count == base and flags & capi.RENDER_EXTRA_SAMPLE for count, flags, _ in plan)
        assert max(work) - min(work) <= 0.01 * np.mean(work)       # tile classes are diagonals: even up to an edge effect
    else:
        assert [c for c, _, _ in plan] == split_samples(total, world)
        assert all(flags == 0 for _, flags, _ in plan)


def test_unbalanced_plan_and_equal_contributions():
    assert sample_plan(8, 0, 100, balanced=False) == [(n, 0, n / 100.0) for n in (13, 13, 13, 13, 12, 12, 12, 12)]
    assert sample_plan(4, 7) == [(7, 0, 0.25)] * 4                 # no split: every rank contributes sample_count samples
    # more ranks than samples: some ranks render nothing, nobody gets the extra-sample flag
    assert sample_plan(8, 0, 5) == [(1, 0, 0.2)] * 5 + [(0, 0, 0.0)] * 3


def test_flag_bits_match_the_header():
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "bevyray_b200.h")).read()
    assert re.search(r"BVR_RENDER_EXTRA_SAMPLE\s*=\s*2u", hdr)
    assert "<< 8) | ((uint32_t)(phase) << 16) | ((uint32_t)(count) << 24)" in hdr
    assert capi.render_extra_sample_bits(8, 3, 4) == 2 | (8 << 8) | (3 << 16) | (4 << 24)

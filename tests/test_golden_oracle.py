"""The oracle reproduces the committed golden fixtures bit for bit (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from golden_util import golden_names, load_golden


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(bvr, oracle, name):
    g = load_golden(bvr, name)
    planes, cnt = oracle.render(g["models"], g["materials"], g["nodes"], g["camera"], bvr.make_level(g["level"]),
                                bvr.make_window(g["seed"], g["height"]), g["width"], g["raster_rgba"], g["raster_depth"])
    for k, v in planes.items():
        assert np.array_equal(v.view(np.uint32), g["out_" + k].view(np.uint32)), k
    assert cnt["rays"] == int(g["rays"])
    assert bvr.validate_bvh(g["nodes"], g["models"]) is None


def test_host_scene_generator_matches_golden_bytes(bvr):
    """The seeded scene recipe + PLOC builder still produce the bytes frozen in the fixture."""
    g = load_golden(bvr, "rtiow_repo_cam")
    s = bvr.Scene.rtiow(1)
    assert s.models.tobytes() == g["models"].tobytes()
    assert s.materials.tobytes() == g["materials"].tobytes()
    assert s.nodes.tobytes() == g["nodes"].tobytes()

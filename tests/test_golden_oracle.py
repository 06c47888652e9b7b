"""The oracle reproduces the committed golden fixtures bit for bit (tests/golden/make_golden.py)."""
import os

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


def test_bench_fixture_is_what_the_host_layer_builds(bvr):
    """tests/golden/bench_rtiow.npz (scene + cameras of C1-C3 for `bench.py --impl reference`) must stay byte-identical
    to what the host layer hands the GPU arm: both arms then trace the same scene through the same camera."""
    import bench
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "bench_rtiow.npz"))
    sc = bvr.Scene.rtiow(int(z["scene_seed"]))
    assert np.array_equal(z["models"].reshape(-1), sc.models.view(np.uint8))
    assert np.array_equal(z["materials"].reshape(-1), sc.materials.view(np.uint8))
    assert np.array_equal(z["nodes"].reshape(-1), sc.nodes.view(np.uint8))
    for key in ("c1", "c2", "c3"):
        assert z["camera_" + key].tobytes() == bytes(bench.make_cam(bvr, bench.WORKLOADS[key])), key


def test_reference_arm_runs_without_the_product_library(tmp_path):
    """`bench.py --impl reference` times oracle/ only: it must work when the product library cannot even be loaded,
    and print the contract's line with the same config dict the GPU arm prints."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BEVYRAY_B200_LIB=str(tmp_path / "missing.so"))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mrays/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    import bench
    wl = bench.WORKLOADS["c1"]
    assert line["config"] == bench.config_for("c1", wl, 506, 1, "samples", 4)
    # measured, not extrapolated: steps x ms_per_step is the time the run really took
    assert line["ms_per_step"] * line["steps"] < 120e3

"""bench.py's device footprint for the driver's arguments (the round-1 bench held steps x 192 uploads)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_footprint_is_bounded_and_independent_of_steps():
    import bench
    for steps in (2, 20, 200):
        f = bench.device_footprint_gb(192, 64, 5_000_000, steps=steps)
        assert f["total_gb"] < 150.0, f
        assert f["depends_on_steps"] is False
    a = bench.device_footprint_gb(192, 64, 5_000_000, steps=2)
    b = bench.device_footprint_gb(192, 64, 5_000_000, steps=2000)
    assert a == b


def test_resident_inputs_are_created_once():
    """The resident arm builds its inputs from pairs_of_step(0) only, outside the step loop."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "resident = [abi.Index(*pairs[p]" in src and "for p in pairs_of_step(0)]" in src
    assert "for s in range(args.steps)]" not in src.split("def ours(")[1].split("resident = [")[1].split("\n")[0]

"""Host-side helpers of bench.py that decide where the launching threads run (no GPU)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def test_physical_cores_partition_the_allowed_cpus():
    cpus = sorted(os.sched_getaffinity(0))
    groups = bench.physical_cores(cpus)
    flat = sorted(c for g in groups for c in g)
    assert flat == cpus                                    # every allowed CPU in exactly one group
    assert all(g == sorted(g) and len(g) >= 1 for g in groups)


def test_rank_core_slices_are_disjoint_and_restore():
    before = os.sched_getaffinity(0)
    try:
        world = 2 if len(bench.physical_cores(before)) >= 2 else 1
        slices = []
        for r in range(world):
            os.sched_setaffinity(0, before)
            slices.append(set(bench.pin_rank_to_cores(r, world)))
        assert all(slices) and (world == 1 or not (slices[0] & slices[1]))
    finally:
        os.sched_setaffinity(0, before)

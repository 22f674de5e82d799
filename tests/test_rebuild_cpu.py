"""Row f4, host side: invariants of the accelerator-tree builder (pre-split references + SAH tree), checked by a small C++
program compiled from tests/rebuild_check.cpp against raytracinginaweekend_b200/csrc/rt_rebuild.hpp.  No GPU needed."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("rebuild") / "rebuild_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, os.path.join(HERE, "rebuild_check.cpp")])
    return exe


@pytest.mark.parametrize("n,seed,factor,extra,mode", [
    (1, 1, 8, 4, 0), (2, 2, 8, 4, 0), (3, 3, 1, 0, 0), (97, 4, 1, 0, 0), (500, 5, 8, 4, 0), (2000, 6, 8, 4, 1), (2000, 7, 2, 0, 1),
    (9000, 8, 4, 3, 1),        # above the threaded-build threshold, tight depth cap on a sliver nest
    (300, 9, 8, 4, 2),         # coincident centroids: median fallback
    (20000, 10, 1, 1, 0),      # no pre-splitting, threaded build, cap one level above the minimum
    (300000, 11, 2, 3, 0),     # 600 000 references: splits of >= 2^18 references scan, bin and partition on all cores (worker pool)
    (600000, 12, 2, 2, 1),     # sliver nests at that size: bucketed pre-split rounds with the budget ending inside a round
    (3000, 13, 1, 0, 3),       # the shared worker pool: concurrent callers with nested loops (no deadlock, no lost chunk)
])
def test_builder_invariants(checker, n, seed, factor, extra, mode):
    r = subprocess.run([checker, str(n), str(seed), str(factor), str(extra), str(mode)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK"), r.stdout

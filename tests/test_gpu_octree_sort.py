"""GPU: the CTA-parallel std::sort emulation of the octree's phase 2 (csrc/octree_core.h: oc_std_sort_cta -- one warp
per introsort partition, ballot-ranked Hoare swaps, rank pass instead of the final insertion sort) gives libstdc++'s
permutation on the device: tools/octree_sort_check.cu sorts ~1000 arrays (heavy ties, sorted / organ-pipe shapes, the
depth-limit input) with 128- and 256-thread CTAs and compares each with the host's std::sort
(the reference's `sort(vSizeAndPointerToNode...)`, ORBextractor.cc:985)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "octree_sort_check.cu")
EXE = os.path.join(ROOT, "tools", "octree_sort_check")
HDR = os.path.join(ROOT, "orb-slam3_byzyh_b200", "csrc", "octree_core.h")


@pytest.mark.gpu
def test_cta_sort_matches_std_sort():
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.check_call(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-o", EXE, SRC])
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr

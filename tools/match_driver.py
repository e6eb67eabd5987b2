#!/usr/bin/env python
"""Runs the matching kernels of the bench at their bench sizes a few times, nothing else -- the command
tools/profile_match.sh puts under ncu (kernel shares and per-launch counters; never a bench number).
C5: kNN-2 of 2000 descriptors against 1 M, SearchByProjection of 1 M map points against a 2000-keypoint frame;
C2 / C3: 64 stereo pairs through extraction + the batched stereo matcher / the batched kNN-2."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (sets up sys.path for the package and tests/)


def main():
    import torch
    import torch.distributed as dist
    import orbfe
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    orbfe.lib()

    def barrier():
        torch.cuda.synchronize()

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    m = bench.matching_leg(torch, dist, orbfe, dev, 0, 1, steps, barrier, lambda x: x)
    print("knn2 ms", m["ms_per_step"], "sbp", m["search_by_projection"].get("ms_per_step"), file=sys.stderr)
    for kind in ("c2", "c3"):
        r = bench.pairs_leg(torch, orbfe, dev, 0, kind, 64, steps, barrier, lambda x: x, 1, 0)
        print(kind, r["ms_per_step"], file=sys.stderr)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Generates tests/golden/matchers_ref.npz from the REFERENCE ITSELF: the matcher functions of
/root/reference/src/{ORBmatcher,Frame,KeyFrame,MapPoint}.cc and the vendored DBoW2, compiled verbatim into
oracle/_ref/libref_orbmatcher.so (oracle/ref_build.sh).  Inputs are the seeded synthetic cases of tests/golden_cases.py;
only the reference outputs and an 8-byte digest of every case's inputs are stored.  Run in the build container:
    python tests/golden/make_golden_matchers.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import golden_cases  # noqa: E402
from oracle import ref as R  # noqa: E402


def main():
    assert R.matcher_available(), "needs oracle/_ref/libref_orbmatcher.so (built from /root/reference)"
    out = {}
    for cls in golden_cases.CASES:
        case = cls()
        out[f"{cls.name}/inputs"] = case.inputs()
        for k, v in case.ref().items():
            out[f"{cls.name}/{k}"] = np.asarray(v)
        print(cls.name, {k: np.asarray(v).shape for k, v in case.ref().items()} if False else "ok")
    np.savez_compressed(os.path.join(HERE, "matchers_ref.npz"), **out)
    print("wrote matchers_ref.npz,", os.path.getsize(os.path.join(HERE, "matchers_ref.npz")), "bytes")


if __name__ == "__main__":
    main()

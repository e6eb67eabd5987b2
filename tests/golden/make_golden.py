#!/usr/bin/env python
"""Generates tests/golden/*.json|npz from the REFERENCE ITSELF: oracle/_ref/libref_orbextractor.so is
/root/reference/src/ORBextractor.cc compiled verbatim (oracle/ref_build.sh), and cv2 4.13's BFMatcher
for the brute-force matcher.  Run in the build container (needs /root/reference and cv2):
    python tests/golden/make_golden.py
Inputs are the seeded synthetic frames of tests/synth.py, so only digests and one small
raw vector set need to be committed."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import synth  # noqa: E402
from oracle import ref as R  # noqa: E402

CASES = [  # name, h, w, nfeatures, lapping, seeds  (BASELINE configs C1..C4 shapes + a small one)
    ("C1", 480, 752, 1000, (0, 1000), (0, 1, 2, 3)),
    ("C2", 480, 752, 1200, (0, 0), (0, 1)),
    ("C3a", 512, 512, 1500, (0, 511), (0, 1)),
    ("C3b", 512, 512, 1500, (100, 411), (0, 1)),
    ("C4", 720, 1280, 2000, (0, 1000), (0, 1)),
    ("small", 240, 320, 300, (0, 0), (0,)),
]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    assert R.available(), "needs oracle/_ref (built from /root/reference)"
    out = {}
    for name, h, w, nf, lap, seeds in CASES:
        for seed in seeds:
            img = synth.synth_frame(h, w, seed)
            ex = R.RefExtractor(nf)
            mono, kps, desc = ex(img, lap)
            out[f"{name}/seed{seed}"] = dict(
                h=h, w=w, nfeatures=nf, lapping=list(lap), seed=seed, image=sha(img), mono=int(mono), n=int(len(kps)),
                keypoints=sha(kps), descriptors=sha(desc), pyramid=[sha(ex.level_padded(l)) for l in range(8)])
            if name == "small":
                np.savez_compressed(os.path.join(HERE, "small_seed0.npz"), image=img, keypoints=kps, descriptors=desc,
                                    mono=np.int32(mono))
    json.dump(out, open(os.path.join(HERE, "extract_digests.json"), "w"), indent=1, sort_keys=True)
    # brute-force kNN-2 pinned to cv2's BFMatcher (Frame.cc:47,1553)
    import cv2
    rng = np.random.default_rng(42)
    q = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (700, 32), dtype=np.uint8)
    t[50] = t[10]; t[600] = q[3]; t[20] = q[3]
    for i in range(60):
        t[int(rng.integers(0, 700))] = synth.flip_bits(q[i], int(rng.integers(0, 70)), rng)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    idx = np.array([[p[0].trainIdx, p[1].trainIdx] for p in m], np.int32)
    dist = np.array([[int(p[0].distance), int(p[1].distance)] for p in m], np.int32)
    match = np.where(dist[:, 0].astype(np.float32) < dist[:, 1].astype(np.float32) * 0.7, idx[:, 0], -1).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "knn2_cv2.npz"), query=q, train=t, idx=idx, dist=dist, match=match)
    print("wrote", len(out), "extraction digests, small_seed0.npz, knn2_cv2.npz")


if __name__ == "__main__":
    main()

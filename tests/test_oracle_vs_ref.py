"""Pin the oracle restatement (oracle/orb_oracle.cpp) against the reference's own
src/ORBextractor.cc compiled verbatim (oracle/_ref/libref_orbextractor.so, built by
oracle/ref_build.sh from /root/reference; prebuilt .so travels to the GPU box).  CPU only."""
import numpy as np
import pytest

import synth
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (no /root/reference)")

CONFIGS = [  # (h, w, nfeatures, lapping)
    (480, 752, 1000, (0, 1000)),
    (480, 752, 1200, (0, 0)),
    (512, 512, 1500, (0, 511)),
    (512, 512, 1500, (100, 411)),
    (720, 1280, 2000, (0, 1000)),
    (240, 320, 500, (0, 0)),
]


@pytest.mark.parametrize("h,w,nf,lap", CONFIGS)
def test_extract_matches_reference(h, w, nf, lap):
    seeds = (0, 1) if h * w < 600000 else (0,)
    for seed in seeds:
        img = synth.synth_frame(h, w, seed)
        ex, rx = O.Extractor(nf), R.RefExtractor(nf)
        mo, ko, do = ex(img, lap)
        mr, kr, dr = rx(img, lap)
        assert mo == mr
        assert len(ko) == len(kr) and len(ko) >= nf * 0.9
        assert ko.tobytes() == kr.tobytes()
        assert np.array_equal(do, dr)
        for lvl in range(8):
            assert np.array_equal(ex.level(lvl)["padded"], rx.level_padded(lvl))


def test_noise_frame_and_small_targets():
    img = synth.noise_frame(200, 260, 3)
    for nf in (50, 300, 3000):
        mo, ko, do = O.Extractor(nf)(img, (0, 0))
        mr, kr, dr = R.RefExtractor(nf)(img, (0, 0))
        assert mo == mr and ko.tobytes() == kr.tobytes() and np.array_equal(do, dr)


def test_empty_image_returns_minus_one():
    empty = np.zeros((0, 0), np.uint8)
    assert O.Extractor(100)(empty)[0] == -1
    assert R.RefExtractor(100)(empty)[0] == -1


@pytest.mark.parametrize("seed", range(6))
def test_octree_alone(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 6000))
    W, H = int(rng.integers(100, 1300)), int(rng.integers(60, 700))
    # distinct integer positions, random scores with many ties
    pos = rng.choice(W * H, size=min(n, W * H), replace=False)
    xys = np.stack([pos % W, pos // W, rng.integers(7, 60, len(pos))], 1).astype(np.int32)
    order = np.lexsort((xys[:, 0], xys[:, 1]))
    xys = xys[order]
    N = int(rng.integers(5, 500))
    keep = O.octree(xys, 16, 16 + W, 16, 16 + H, N)
    refk = R.RefExtractor(1000).octree(xys, 16, 16 + W, 16, 16 + H, N)
    assert np.array_equal(xys[keep], refk)


def test_scale_tables_known_answers():
    t = O.Extractor(1000).tables()
    assert [hex(v) for v in t["scale"].view(np.uint32)] == [
        "0x3f800000", "0x3f99999a", "0x3fb851ec", "0x3fdd2f1c", "0x4004b5de", "0x401f40a4",
        "0x403f1a5f", "0x406552d9"]
    assert list(t["nfeatures"]) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(t["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert list(O.Extractor(2000).tables()["nfeatures"]) == [434, 362, 302, 251, 209, 175, 145, 122]

"""Golden vectors generated from the reference itself (tests/golden/make_golden.py: the reference's
ORBextractor.cc compiled verbatim, cv2 BFMatcher).  CPU: the oracle reproduces them; GPU: the CUDA
path reproduces them through the C ABI -- without touching /root/reference at run time."""
import hashlib
import json
import os

import numpy as np
import pytest

import synth
from oracle import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DIGESTS = json.load(open(os.path.join(G, "extract_digests.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("key", sorted(DIGESTS))
def test_oracle_reproduces_reference_digests(key):
    d = DIGESTS[key]
    img = synth.synth_frame(d["h"], d["w"], d["seed"])
    assert sha(img) == d["image"], "synthetic input generator changed"
    ex = O.Extractor(d["nfeatures"])
    mono, kps, desc = ex(img, tuple(d["lapping"]))
    assert (mono, len(kps)) == (d["mono"], d["n"])
    assert sha(kps) == d["keypoints"] and sha(desc) == d["descriptors"]
    assert [sha(ex.level(l)["padded"]) for l in range(8)] == d["pyramid"]


def test_oracle_small_raw_vectors_and_knn():
    z = np.load(os.path.join(G, "small_seed0.npz"))
    mono, kps, desc = O.Extractor(300)(z["image"], (0, 0))
    assert mono == int(z["mono"]) and kps.tobytes() == z["keypoints"].tobytes() and np.array_equal(desc, z["descriptors"])
    k = np.load(os.path.join(G, "knn2_cv2.npz"))
    match, idx, dist = O.fisheye_matches(k["query"], k["train"])
    assert np.array_equal(idx, k["idx"]) and np.array_equal(dist, k["dist"]) and np.array_equal(match, k["match"])


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(DIGESTS))
def test_cuda_reproduces_reference_digests(key):
    import orbfe
    d = DIGESTS[key]
    img = synth.synth_frame(d["h"], d["w"], d["seed"])
    ex = orbfe.ORBextractor(d["nfeatures"])
    mono, kps, desc = ex(img, None, tuple(d["lapping"]))
    assert (mono, len(kps)) == (d["mono"], d["n"])
    assert sha(kps) == d["keypoints"] and sha(desc) == d["descriptors"]
    assert [sha(ex.pyramid_level(l, with_border=True)) for l in range(8)] == d["pyramid"]


@pytest.mark.gpu
def test_cuda_small_raw_vectors_and_knn():
    import orbfe
    z = np.load(os.path.join(G, "small_seed0.npz"))
    mono, kps, desc = orbfe.ORBextractor(300)(z["image"], None, (0, 0))
    assert mono == int(z["mono"]) and kps.tobytes() == z["keypoints"].tobytes() and np.array_equal(desc, z["descriptors"])
    k = np.load(os.path.join(G, "knn2_cv2.npz"))
    idx, dist, match = orbfe.ORBmatcher().knn2(k["query"], k["train"])
    assert np.array_equal(idx, k["idx"]) and np.array_equal(dist, k["dist"]) and np.array_equal(match, k["match"])


# ---- matcher / BoW goldens: outputs of the reference's own functions (tests/golden/make_golden_matchers.py) ----------
import golden_cases  # noqa: E402

MATCHERS = np.load(os.path.join(G, "matchers_ref.npz"))


def _stored(name):
    pre = name + "/"
    return {k[len(pre):]: MATCHERS[k] for k in MATCHERS.files if k.startswith(pre)}


@pytest.mark.parametrize("cls", golden_cases.CASES, ids=lambda c: c.name)
def test_matcher_golden_inputs_are_reproducible(cls):
    """The seeded case builders give the inputs the goldens were generated from (numpy RNG streams are stable)."""
    assert np.array_equal(cls().inputs(), _stored(cls.name)["inputs"])


@pytest.mark.gpu
@pytest.mark.parametrize("cls", golden_cases.CASES, ids=lambda c: c.name)
def test_cuda_reproduces_reference_matcher_outputs(cls):
    """CUDA entry points against the stored outputs of the reference's own matcher functions (no oracle in between)."""
    import orbfe
    case = cls()
    stored = _stored(cls.name)
    assert np.array_equal(case.inputs(), stored.pop("inputs"))
    got = case.cuda(orbfe, stored)
    assert set(got) == set(stored)
    for k, v in stored.items():
        g = np.asarray(got[k])
        if v.dtype.kind == "f":
            assert np.array_equal(g.view(np.uint8), v.view(np.uint8)), (cls.name, k)      # float bit patterns
        else:
            assert np.array_equal(g, v), (cls.name, k)

"""CPU: the C-ABI library loads and exports every symbol include/orbfe.h declares, the ctypes
binding covers them all, and the product fails loudly (no CPU fallback) without a CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "orbfe.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(orbfe_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    import orbfe
    names = _declared()
    assert len(names) >= 30
    L = C.CDLL(orbfe.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/orbfe.h but not exported by libORBfe_b200.so"
    assert set(names) == set(orbfe.EXPORTS), set(names) ^ set(orbfe.EXPORTS)
    orbfe.lib()
    assert orbfe.lib().orbfe_version().startswith(b"orbfe-b200 sm_100a")


def test_keypoint_layout_is_cv_keypoint():
    import orbfe
    assert orbfe.KP_DTYPE.itemsize == 28
    assert [orbfe.KP_DTYPE.fields[f][1] for f in ("x", "y", "size", "angle", "response", "octave", "class_id")] == \
        [0, 4, 8, 12, 16, 20, 24]


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import numpy as np
    import orbfe
    with pytest.raises(orbfe.OrbfeError) as e:
        orbfe.ORBextractor(1000)
    assert e.value.code == orbfe.ERR_CUDA
    with pytest.raises(orbfe.OrbfeError):
        orbfe.ORBmatcher().knn2(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8))
    with pytest.raises(orbfe.OrbfeError):
        orbfe.ORBmatcher.DescriptorDistance(np.zeros(32, np.uint8), np.zeros(32, np.uint8))
    # the streaming entry points refuse a null handle instead of touching the buffers
    lib = orbfe.lib()
    img, out = np.zeros((1, 64, 64), np.uint8), np.zeros(64, np.uint8)
    p = lambda a: a.ctypes.data
    assert lib.orbfe_extract_batch_submit(None, p(img), 1, 64, 64, 64, 64 * 64, 0, 0, p(out), p(out), 1, p(out), p(out)) < 0
    assert lib.orbfe_extract_batch_wait(None) < 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "orb-slam3_byzyh_b200")
    pat = re.compile(r'^\s*(from\s+oracle|import\s+oracle|#\s*include\s+"[^"]*oracle/)', re.M)
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".inc")):
                txt = open(os.path.join(d, f), errors="replace").read()
                assert not pat.search(txt) and "liborb_oracle" not in txt, f

"""The header-only C++ adapter (orb-slam3_byzyh_b200/host/) keeps ORB-SLAM3's ORBextractor
signatures: it must compile and link against libORBfe_b200.so (CPU), and produce the oracle's
keypoints / descriptors / bordered pyramid when run (GPU)."""
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "orb-slam3_byzyh_b200")
EXE = os.path.join(HERE, "native", "host_adapter_check")


def _build():
    so = os.path.join(PKG, "libORBfe_b200.so")
    if not os.path.exists(so):
        pytest.skip("libORBfe_b200.so not built (run __graft_entry__.build())")
    src = os.path.join(HERE, "native", "host_adapter_check.cpp")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "oracle", "cvshim"),
                           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "host"), src,
                           os.path.join(ROOT, "oracle", "cvprims.cpp"), "-o", EXE, "-L", PKG, "-lORBfe_b200",
                           "-Wl,-rpath," + PKG, "-Wl,-rpath,/usr/local/cuda/lib64", "-L/usr/local/cuda/lib64", "-lcudart"])
    return EXE


def test_adapter_compiles_and_links():
    assert os.path.exists(_build())


@pytest.mark.gpu
def test_adapter_matches_oracle(tmp_path):
    import synth
    from oracle import oracle as O
    exe = _build()
    img = synth.synth_frame(480, 752, 11)
    raw = tmp_path / "in.raw"
    raw.write_bytes(img.tobytes())
    out = subprocess.check_output([exe, "480", "752", "1000", "0", "1000", str(raw), str(tmp_path / "o")], text=True)
    mono, n, levels, w1, self_d, sf = out.split()
    ex = O.Extractor(1000)
    omono, okps, odesc = ex(img, (0, 1000))
    assert int(mono) == omono and int(n) == len(okps) and int(levels) == 8 and int(self_d) == 0
    assert (tmp_path / "o.kps").read_bytes() == okps.tobytes()
    assert (tmp_path / "o.desc").read_bytes() == odesc.tobytes()
    assert (tmp_path / "o.pyr1").read_bytes() == ex.level(1)["padded"].tobytes()

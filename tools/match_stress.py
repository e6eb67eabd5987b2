"""Randomised parity sweep of the matcher entry points: the parity tests of tests/test_gpu_match.py (CUDA path through
the C ABI against the CPU oracle) re-run on random seeds and sizes.
usage: match_stress.py [rounds] [seed]   -- one line per failing case; exit 1 on any mismatch.
The tests also assert minimum match counts that hold for their own seeds; on a random seed such an assert can trip with
the two sides equal (seen once: SearchForTriangulation, 59 matches on both sides against a floor of 60) -- re-run the
case by hand before reading a line of this tool as a parity failure."""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "orb-slam3_byzyh_b200")]
import numpy as np
import orbfe
import test_gpu_match as T

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
orbfe.lib()
bad = done = 0
t0 = time.time()


def run(name, fn, *args):
    global bad, done
    done += 1
    try:
        fn(orbfe, *args)
    except Exception as e:
        bad += 1
        print("MISMATCH", name, args, "->", (str(e) or traceback.format_exc(limit=1))[:160], flush=True)


for r in range(rounds):
    s = int(rng.integers(100, 1 << 20))
    run("knn2", T.test_knn2_vs_oracle, int(rng.integers(1, 2500)), int(rng.integers(0, 30000)))
    for mode in (0, 1, 2):
        run("search_by_projection", T.test_search_by_projection, mode, int(rng.integers(50, 40000)), int(rng.integers(20, 2500)), s)
    run("stereo", T.test_stereo_matches, s)
    run("search_for_initialization", T.test_search_for_initialization, s)
    run("fuse", T.test_fuse_search, bool(r & 1), int(rng.integers(5, 30000)), int(rng.integers(1, 3000)), s)
    run("search_by_sim3", T.test_search_by_sim3, s)
    run("distinctive", T.test_distinctive_descriptors, s)

# ---- bag-of-words path, triangulation search, image intake ----
import synth
import test_gpu_bow as TB
import test_gpu_intake as TI
from oracle import oracle as O
voc = synth.make_vocabulary(10, 4, 3)
vocs = (voc, orbfe.ORBVocabulary(10, 4, voc["parent"], voc["desc"], voc["weight"]), O.Vocabulary(10, 4, voc["parent"], voc["desc"], voc["weight"]))


def runv(name, fn, *args):
    global bad, done
    done += 1
    try:
        fn(orbfe, vocs, *args)
    except Exception as e:
        bad += 1
        print("MISMATCH", name, args, "->", (str(e) or traceback.format_exc(limit=1))[:160], flush=True)


for r in range(rounds):
    s = int(rng.integers(100, 1 << 20))
    runv("search_by_bow", TB.test_search_by_bow_keyframe_frame, s, bool(r & 1), float(rng.choice([0.6, 0.7, 0.9])), int(rng.choice([-1, 700, 1000])))
    runv("search_by_bow_keyframes", TB.test_search_by_bow_keyframes, s, bool(r & 1))
    runv("search_for_triangulation", TB.test_search_for_triangulation, s, bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(r & 1), float(rng.uniform(0, 0.7)))
    runv("search_for_triangulation_rig", TB.test_search_for_triangulation_two_camera_keyframes, s, bool(rng.integers(0, 2)), bool(r & 1))
    run("resize", TI.test_resize, int(rng.integers(64, 1400)), int(rng.integers(48, 800)), int(rng.integers(64, 1400)), int(rng.integers(48, 800)))
    run("remap", TI.test_remap, int(rng.integers(60, 700)), int(rng.integers(80, 1200)), int(rng.integers(60, 700)), int(rng.integers(80, 1200)), s)
print(f"{done} cases compared, {bad} failing, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)

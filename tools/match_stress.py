"""Randomised parity sweep of the matcher entry points: the parity tests of tests/test_gpu_match.py (CUDA path through
the C ABI against the CPU oracle) re-run on random seeds and sizes.
usage: match_stress.py [rounds] [seed]   -- one line per failing case; exit 1 on any mismatch"""
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
print(f"{done} cases compared, {bad} failing, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)

"""Randomised pin of the matcher oracle (oracle/match_oracle.cpp) against the reference's own matcher functions compiled
verbatim (oracle/_ref): the seed-parametrised tests of tests/test_oracle_match_vs_ref.py on random seeds and sizes.  CPU only.
usage: oracle_match_stress.py [rounds] [seed]   (the tests' own floors on match counts can trip on a random seed with both
sides equal: re-run such a case by hand before reading it as a mismatch)"""
import sys, os, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from oracle import ref as R
if not R.available():
    print("oracle/_ref is not built here")
    sys.exit(0)
import test_oracle_match_vs_ref as T

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = done = 0
t0 = time.time()


def run(name, fn, *args):
    global bad, done
    done += 1
    try:
        fn(*args)
    except Exception as e:
        bad += 1
        print("MISMATCH", name, [a for a in args if not isinstance(a, tuple)], "->", (str(e) or traceback.format_exc(limit=1))[:160], flush=True)


for r in range(rounds):
    s = int(rng.integers(100, 1 << 20))
    run("features_in_area", T.test_features_in_area_vs_reference, s)
    run("search_mappoints", T.test_search_mappoints_vs_reference, int(rng.integers(200, 20000)), int(rng.integers(100, 2500)), s,
        float(rng.choice([1.0, 3.0, 5.0])), bool(r & 1))
    run("search_for_initialization", T.test_search_for_initialization_vs_reference, s, bool(r & 1))
    run("stereo", T.test_stereo_matches_vs_reference, s)
    run("fuse", T.test_fuse_vs_reference, s, bool(r & 1))
    run("search_by_sim3", T.test_search_by_sim3_vs_reference, s)
    run("distinctive", T.test_distinctive_descriptors_vs_reference, s)

# ---- BoW matchers and the triangulation search against DBoW2 / the reference bodies compiled verbatim ----
if R.matcher_available():
    import tempfile
    import synth
    from oracle import oracle as O
    import test_oracle_bow_vs_ref as TB
    voc = synth.make_vocabulary(10, 4, 3)
    path = os.path.join(tempfile.mkdtemp(), "voc.txt")
    synth.write_vocabulary_text(path, voc)
    vp = (voc, R.RefVocabulary(path), O.Vocabulary(10, 4, voc["parent"], voc["desc"], voc["weight"]))
    for r in range(rounds):
        s = int(rng.integers(100, 1 << 20))
        run("search_by_bow", TB.test_search_by_bow_kf_frame_vs_reference, vp, s, bool(r & 1), float(rng.choice([0.6, 0.7, 0.9])))
        run("search_by_bow_fisheye", TB.test_search_by_bow_kf_frame_fisheye_vs_reference, vp, s, bool(r & 1))
        run("search_by_bow_keyframes", TB.test_search_by_bow_kf_kf_vs_reference, vp, s, bool(r & 1))
        run("search_for_triangulation", TB.test_search_for_triangulation_vs_reference, vp, s, bool(rng.integers(0, 2)),
            bool(rng.integers(0, 2)), bool(r & 1), float(rng.uniform(0, 0.7)))
print(f"{done} cases compared, {bad} failing, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)

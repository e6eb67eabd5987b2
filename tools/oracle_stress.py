"""Randomised pin of the oracle restatement (oracle/orb_oracle.cpp) against the reference's own ORBextractor.cc compiled
verbatim (oracle/_ref, needs /root/reference at build time): the case generator of tools/parity_stress.py, CPU only.
usage: oracle_stress.py [cases] [seed]   -- one line per failing case; exit 1 on any mismatch"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
import synth
from oracle import oracle as O
from oracle import ref as R

if not R.available():
    print("oracle/_ref is not built here")
    sys.exit(0)
cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
MAXH, MAXW = int(os.environ.get("STRESS_MAXH", "800")), int(os.environ.get("STRESS_MAXW", "1400"))
src = open(os.path.join(ROOT, "tools", "parity_stress.py")).read()
ns = {}
exec(src[src.index("def image("):src.index("bad, t0, done")], {"np": np, "synth": synth}, ns)
image = ns["image"]

bad = done = 0
t0 = time.time()
for c in range(cases):
    h, w = int(rng.integers(160, MAXH)), int(rng.integers(200, MAXW))
    if w < h:
        h, w = w, h
    nf = int(rng.choice([5, 50, 200, 500, 1000, 1500, 2000, 4000, 12000]))
    sf = float(rng.choice([1.2, 1.2, 1.2, 1.1, 1.3, 1.5, 2.0]))
    nl = int(rng.integers(1, 9))
    ini = int(rng.choice([20, 20, 12, 30, 40, 7, 100, 0]))
    mn = int(rng.choice([7, 7, 5, 10, 2, 0, 25]))
    while nl > 1 and round(min(h, w) / sf ** (nl - 1)) < 48:
        nl -= 1
    lap = [(0, 1000), (0, 0), (0, w - 1), (w // 4, 3 * w // 4)][int(rng.integers(0, 4))]
    kind, seed = int(rng.integers(0, 5)), int(rng.integers(0, 1 << 30))
    tag = f"case {c}: {h}x{w} nf={nf} sf={sf} nl={nl} th={ini}/{mn} lap={lap} kind={kind} seed={seed}"
    img = image(h, w, kind, seed)
    ex, rx = O.Extractor(nf, sf, nl, ini, mn), R.RefExtractor(nf, sf, nl, ini, mn)
    mo, ko, do = ex(img, lap)
    mr, kr, dr = rx(img, lap)
    ok = mo == mr and ko.tobytes() == kr.tobytes() and np.array_equal(do, dr)
    if ok:
        for lvl in range(nl):
            if not np.array_equal(ex.level(lvl)["padded"], rx.level_padded(lvl)):
                ok = False
    if not ok:
        bad += 1
        print("MISMATCH", tag, f"mono {mo}/{mr} n {len(ko)}/{len(kr)}", flush=True)
    done += 1
print(f"{done} cases compared, {bad} failing, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)

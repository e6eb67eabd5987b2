"""Runs one frame repeatedly and compares the per-level FAST candidate lists with the oracle's (debug aid)."""
import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (R, os.path.join(R, "tests"), os.path.join(R, "orb-slam3_byzyh_b200")):
    sys.path.insert(0, p)
import numpy as np, synth, orbfe
from oracle import oracle as O
h, w, nf, seed = [int(a) for a in sys.argv[1:5]]
img = synth.synth_frame(h, w, seed)
oc = O.Extractor(nf)
oc(img, (0, 1000))
ex = orbfe.ORBextractor(nf)
for rep in range(int(sys.argv[5]) if len(sys.argv) > 5 else 10):
    ex(img, None, (0, 1000))
    for l in range(8):
        cg = ex.debug_candidates(l)
        co = oc.level(l)["cands"]
        if not np.array_equal(cg, co):
            sg = {tuple(r) for r in cg.tolist()}
            so = {tuple(r) for r in co.tolist()}
            L = oc.level(l)
            print(f"rep {rep} level {l}: gpu {len(cg)} oracle {len(co)}; only gpu {sorted(sg - so)[:8]}; only oracle {sorted(so - sg)[:8]}")
print("done")

"""Randomised parity sweep of the extraction path against the CPU oracle (same checks as tests/test_gpu_extract.py):
random frame sizes, feature counts, pyramid parameters, thresholds, lapping areas and image statistics.
usage: parity_stress.py [cases] [seed]   -- prints one line per failing case and a summary; exit 1 on any mismatch"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "orb-slam3_byzyh_b200")]
import numpy as np
import synth, orbfe
from oracle import oracle as O
from test_gpu_extract import _check_frame

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
orbfe.lib()
MAXH, MAXW = int(os.environ.get("STRESS_MAXH", "800")), int(os.environ.get("STRESS_MAXW", "1400"))


def image(h, w, kind, seed):
    r = np.random.default_rng(seed)
    if kind == 0:
        return synth.synth_frame(h, w, seed)
    if kind == 1:
        return synth.noise_frame(h, w, seed)
    if kind == 2:   # low contrast: most cells only pass at minThFAST, many stay empty
        f = synth.synth_frame(h, w, seed).astype(np.float32)
        return (100 + (f - 128) * r.uniform(0.05, 0.3)).clip(0, 255).astype(np.uint8)
    if kind == 3:   # saturated blocks: ties in the responses and in the octree keys
        f = synth.synth_frame(h, w, seed)
        return np.where(f > 128, 255, 0).astype(np.uint8)
    f = synth.synth_frame(h, w, seed)   # half flat, half textured
    f[:, : w // 2] = 97
    return f


bad, t0, done = 0, time.time(), 0
for c in range(cases):
    h, w = int(rng.integers(160, MAXH)), int(rng.integers(200, MAXW))
    if w < h:
        h, w = w, h                             # portrait levels can round width / height to 0 roots: the reference divides by zero there
    nf = int(rng.choice([5, 50, 200, 500, 1000, 1500, 2000, 4000, 12000]))
    sf = float(rng.choice([1.2, 1.2, 1.2, 1.1, 1.3, 1.5, 2.0]))
    nl = int(rng.integers(1, 9))
    ini = int(rng.choice([20, 20, 12, 30, 40, 7, 100, 0]))
    mn = int(rng.choice([7, 7, 5, 10, 2, 0, 25]))
    while nl > 1 and round(min(h, w) / sf ** (nl - 1)) < 48:
        nl -= 1                                  # the reference itself breaks on levels smaller than its 16-px border window
    lap = [(0, 1000), (0, 0), (0, w - 1), (w // 4, 3 * w // 4)][int(rng.integers(0, 4))]
    kind, seed = int(rng.integers(0, 5)), int(rng.integers(0, 1 << 30))
    tag = f"case {c}: {h}x{w} nf={nf} sf={sf} nl={nl} th={ini}/{mn} lap={lap} kind={kind} seed={seed}"
    try:
        exg, exc = orbfe.ORBextractor(nf, sf, nl, ini, mn), O.Extractor(nf, sf, nl, ini, mn)
    except Exception as e:
        continue
    img = image(h, w, kind, seed)
    try:
        exc(img, lap)
    except Exception:
        continue                                 # geometry the reference itself rejects
    try:
        nb = _check_frame(exg, exc, img, lap, stages=(c % 4 == 0))
        if nb:
            print("DESCRIPTOR DIFFS", nb, tag)
            bad += 1
    except AssertionError as e:
        print("MISMATCH", tag, "->", str(e)[:120])
        bad += 1
    except orbfe.OrbfeError as e:
        print("ERROR", tag, "->", str(e)[:120])
        bad += 1
    done += 1

# ---- handle reuse and batch paths: one extractor over changing frame sizes, single calls and host batches mixed,
# random chunk budgets (several chunks per batch, CUDA-graph replay for small batches, re-layout on size changes) ----
sessions = max(1, cases // 15)
for c in range(sessions):
    nf = int(rng.choice([300, 1000, 2000]))
    exg, exc = orbfe.ORBextractor(nf), O.Extractor(nf)
    sizes = [(int(rng.integers(200, 760)), int(rng.integers(760, 1300))) for _ in range(2)]
    for step in range(6):
        h, w = sizes[int(rng.integers(0, 2))]
        lap = [(0, 1000), (0, 0), (w // 4, 3 * w // 4)][int(rng.integers(0, 3))]
        B = int(rng.choice([1, 1, 2, 3, 4, 7, 12]))
        mb = int(rng.choice([48 << 20, 100 << 20, 6 << 30]))
        frames = np.stack([image(h, w, int(rng.integers(0, 5)), int(rng.integers(0, 1 << 30))) for _ in range(B)])
        tag = f"session {c} step {step}: {B} x {h}x{w} nf={nf} lap={lap} max_bytes={mb >> 20} MB"
        try:
            exg.set_max_bytes(mb)
            if B == 1 and step % 2:
                _check_frame(exg, exc, frames[0], lap, stages=False)
            else:
                n, mono, kps, desc = exg.extract_batch(frames, lap)
                for i in range(B):
                    mo, ko, do = exc(frames[i], lap)
                    assert mono[i] == mo and n[i] == len(ko), f"frame {i}: counts"
                    assert kps[i, :n[i]].tobytes() == ko.tobytes(), f"frame {i}: keypoints"
                    assert np.array_equal(desc[i, :n[i]], do), f"frame {i}: descriptors"
        except (AssertionError, orbfe.OrbfeError) as e:
            print("MISMATCH", tag, "->", str(e)[:120])
            bad += 1
        done += 1
print(f"{done} cases compared, {bad} failing, {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)

"""Seeded synthetic inputs for the five BASELINE configs.

numpy only (no cv2) so that the same bytes are produced here and on the GPU box.
"""
import numpy as np


def _gauss_kernel(sigma, radius):
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    k = np.exp(-0.5 * (x / sigma) ** 2)
    return k / k.sum()


def _blur(img, sigma, radius):
    k = _gauss_kernel(sigma, radius)
    p = np.pad(img, radius, mode="reflect")
    h, w = img.shape
    tmp = np.zeros((h + 2 * radius, w), np.float64)
    for i, kv in enumerate(k):
        tmp += kv * p[:, i:i + w]
    out = np.zeros((h, w), np.float64)
    for i, kv in enumerate(k):
        out += kv * tmp[i:i + h, :]
    return out


def synth_frame(h, w, seed):
    """Textured frame with rectangles, one low-contrast patch (minThFAST retry) and one flat
    patch (empty FAST cells)."""
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, (h, w)).astype(np.float64)
    img = _blur(img, 1.5, 4)
    img = (img - img.min()) / max(img.max() - img.min(), 1e-9) * 255.0
    nrect = h * w // 1500
    ys = rng.integers(0, h, nrect)
    xs = rng.integers(0, w, nrect)
    hs = rng.integers(6, 61, nrect)
    ws = rng.integers(6, 61, nrect)
    gs = rng.integers(0, 256, nrect)
    for y, x, hh, ww, g in zip(ys, xs, hs, ws, gs):
        img[y:y + hh, x:x + ww] = g
    img = _blur(img, 0.8, 1)
    # low-contrast region: amplitude ~12 gray levels around mid-gray
    ly, lx = int(h * 0.55), int(w * 0.6)
    lc = rng.integers(0, 256, (80, 80)).astype(np.float64)
    lc = _blur(lc, 1.0, 2)
    lc = (lc - lc.min()) / max(lc.max() - lc.min(), 1e-9) * 14.0 + 120.0
    img[ly:ly + 80, lx:lx + 80] = lc[:max(0, min(80, h - ly)), :max(0, min(80, w - lx))]
    # flat region
    fy, fx = int(h * 0.1), int(w * 0.15)
    img[fy:fy + 90, fx:fx + 90] = 97.0
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def stereo_pair(h, w, seed):
    """Rectified pair: right = left shifted by a smooth per-row-constant disparity in [2,60]
    px plus +-2 gray noise."""
    left = synth_frame(h, w, seed)
    rng = np.random.default_rng(seed + 1000)
    rows = np.arange(h)
    disp = 31.0 + 29.0 * np.sin(rows / h * 2.0 * np.pi * 1.5 + rng.uniform(0, 6.28))
    disp = np.clip(np.rint(disp), 2, 60).astype(np.int64)
    right = np.empty_like(left)
    cols = np.arange(w)
    for y in range(h):
        src = np.clip(cols + disp[y], 0, w - 1)  # x_right = x_left - d
        right[y] = left[y, src]
    noise = rng.integers(-2, 3, (h, w))
    right = np.clip(right.astype(np.int64) + noise, 0, 255).astype(np.uint8)
    return left, right


def shifted_pair(h, w, seed):
    """Fisheye-style pair (C3): second view = first shifted by (dx,dy), dx in [20,60]."""
    rng = np.random.default_rng(seed + 2000)
    big = synth_frame(h + 16, w + 64, seed)
    dx = int(rng.integers(20, 61))
    dy = int(rng.integers(0, 9))
    a = big[4:4 + h, 0:w]
    b = big[dy:dy + h, dx:dx + w]
    noise = rng.integers(-2, 3, (h, w))
    b = np.clip(b.astype(np.int64) + noise, 0, 255).astype(np.uint8)
    return np.ascontiguousarray(a), np.ascontiguousarray(b)


def noise_frame(h, w, seed):
    """Worst case for candidate counts: raw uniform noise."""
    return np.random.default_rng(seed).integers(0, 256, (h, w)).astype(np.uint8)


def random_descriptors(n, seed):
    return np.random.default_rng(seed).integers(0, 256, (n, 32)).astype(np.uint8)


def flip_bits(desc, k, rng):
    """Return a copy of 32-byte descriptor `desc` with k random distinct bits flipped."""
    d = desc.copy()
    bits = rng.choice(256, size=int(k), replace=False)
    for b in bits:
        d[b >> 3] ^= np.uint8(1 << (b & 7))
    return d


def map_vs_frame(n_map, n_frame, seed, w=1280, h=720, nlevels=8, scale=1.2):
    """C5-shaped matching workload: frame keypoints/descriptors plus a map whose first
    n_frame entries are noisy copies of the frame descriptors (true matches), 20% of those
    followed by a decoy; the rest uniform random.  Returns dict of arrays."""
    from oracle.oracle import KP_DTYPE
    rng = np.random.default_rng(seed)
    fdesc = rng.integers(0, 256, (n_frame, 32)).astype(np.uint8)
    share = np.array([scale ** -i for i in range(nlevels)])
    share /= share.sum()
    octave = rng.choice(nlevels, n_frame, p=share).astype(np.int32)
    keys = np.zeros(n_frame, KP_DTYPE)
    keys["x"] = rng.uniform(20, w - 20, n_frame).astype(np.float32)
    keys["y"] = rng.uniform(20, h - 20, n_frame).astype(np.float32)
    keys["octave"] = octave
    keys["angle"] = rng.uniform(0, 360, n_frame).astype(np.float32)
    keys["size"] = 31.0
    keys["class_id"] = -1
    mdesc = rng.integers(0, 256, (n_map, 32)).astype(np.uint8)
    src = np.full(n_map, -1, np.int64)
    order = rng.permutation(n_map)
    pos = 0
    for i in range(min(n_frame, n_map)):
        if pos >= n_map:
            break
        k = int(rng.integers(0, 61))
        mdesc[order[pos]] = flip_bits(fdesc[i], k, rng)
        src[order[pos]] = i
        pos += 1
        if rng.uniform() < 0.2 and pos < n_map:
            mdesc[order[pos]] = flip_bits(fdesc[i], min(k + int(rng.integers(0, 16)), 256), rng)
            src[order[pos]] = i
            pos += 1
    sf = np.float32(1.0) * np.cumprod(np.concatenate([[1.0], np.full(nlevels - 1, scale)])).astype(np.float32)
    u = rng.uniform(0, w, n_map).astype(np.float32)
    v = rng.uniform(0, h, n_map).astype(np.float32)
    lvl = rng.choice(nlevels, n_map, p=share).astype(np.int32)
    has = src >= 0
    u[has] = (keys["x"][src[has]] + rng.normal(0, 2, has.sum())).astype(np.float32)
    v[has] = (keys["y"][src[has]] + rng.normal(0, 2, has.sum())).astype(np.float32)
    lvl[has] = np.clip(octave[src[has]] + (rng.uniform(size=has.sum()) < 0.2) * rng.choice([-1, 1], has.sum()),
                       0, nlevels - 1)
    view_cos = rng.uniform(0.9, 1.0, n_map).astype(np.float32)
    return dict(keys=keys, fdesc=fdesc, mdesc=mdesc, src=src, u=u, v=v, level=lvl,
                view_cos=view_cos, scale_factors=sf, bounds=(0.0, 0.0, float(w), float(h)))


def make_vocabulary(k, L, seed, stop_frac=0.02, early_leaf_frac=0.0):
    """A DBoW2-shaped vocabulary tree (k children per node, L levels): node ids in the order DBoW2's
    HKmeansStep creates them (the k children of a node are consecutive, subtrees follow), child descriptors =
    parent descriptor with random bit flips, idf-like leaf weights with a few stopped (zero-weight) words.
    early_leaf_frac > 0 turns some nodes of the last inner level into leaves (unbalanced tree).
    Returns dict(k, L, parent, desc, weight) indexed by node id (0 = root)."""
    rng = np.random.default_rng(seed)
    parent, desc, level = [0], [np.zeros(32, np.uint8)], [0]

    def expand(pid, lvl):
        first = len(parent)
        for _ in range(k):
            if lvl == 1:
                d = rng.integers(0, 256, 32).astype(np.uint8)
            else:
                d = flip_bits(desc[pid], int(rng.integers(20, 60)) // lvl + 4, rng)
            parent.append(pid); desc.append(d); level.append(lvl)
        if lvl < L:
            for c in range(first, first + k):
                if lvl == L - 1 and rng.uniform() < early_leaf_frac:
                    continue
                expand(c, lvl + 1)
    expand(0, 1)
    n = len(parent)
    weight = rng.uniform(0.5, 9.0, n)
    weight[rng.uniform(size=n) < stop_frac] = 0.0
    return dict(k=k, L=L, parent=np.array(parent, np.int32), desc=np.stack(desc), weight=weight.astype(np.float64),
                level=np.array(level, np.int32))


def write_vocabulary_text(path, voc, scoring=0, weighting=0):
    """ORBvoc.txt format (TemplatedVocabulary::loadFromTextFile / saveToTextFile).  No trailing newline: DBoW2's
    `while(!f.eof())` loop would otherwise append a childless root child with an uninitialised descriptor."""
    parent, desc, weight = voc["parent"], voc["desc"], voc["weight"]
    has_child = np.zeros(len(parent), bool)
    has_child[parent[1:]] = True
    lines = ["%d %d %d %d" % (voc["k"], voc["L"], scoring, weighting)]
    for i in range(1, len(parent)):
        lines.append("%d %d %s %r" % (parent[i], 0 if has_child[i] else 1, " ".join(str(int(b)) for b in desc[i]),
                                      float(weight[i])))
    with open(path, "w") as f:
        f.write("\n".join(lines))


def descriptors_near_words(voc, n, seed, noise=12):
    """n descriptors = randomly chosen leaves of the vocabulary with `noise` bits flipped."""
    rng = np.random.default_rng(seed)
    has_child = np.zeros(len(voc["parent"]), bool)
    has_child[voc["parent"][1:]] = True
    leaves = np.flatnonzero(~has_child[1:]) + 1
    pick = rng.choice(leaves, n)
    return np.stack([flip_bits(voc["desc"][p], int(rng.integers(0, noise + 1)), rng) for p in pick])


def make_vocabulary_fast(k, L, seed):
    """Full-size vocabulary (k=10, L=6: 1.1 M nodes) built level by level with numpy: node ids are breadth-first,
    child descriptor = parent descriptor XOR a sparse random mask."""
    rng = np.random.default_rng(seed)
    parent = [np.zeros(1, np.int32)]
    desc = [np.zeros((1, 32), np.uint8)]
    first = 0
    for lvl in range(1, L + 1):
        n_prev = len(parent[-1])
        par = np.repeat(np.arange(first, first + n_prev, dtype=np.int32), k)
        if lvl == 1:
            d = rng.integers(0, 256, (len(par), 32)).astype(np.uint8)
        else:
            mask = (rng.uniform(size=(len(par), 256)) < 0.12 / lvl + 0.02)
            d = desc[-1][par - first] ^ np.packbits(mask, axis=1)
        first += n_prev
        parent.append(par); desc.append(d)
    parent = np.concatenate(parent); desc = np.concatenate(desc)
    weight = rng.uniform(0.5, 9.0, len(parent))
    weight[rng.uniform(size=len(parent)) < 0.01] = 0.0
    return dict(k=k, L=L, parent=parent, desc=desc, weight=weight)

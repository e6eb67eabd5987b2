"""ORBVocabulary (include/ORBVocabulary.h:30-31 = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) on the B200:
the tree lives in HBM, transform() descends it on the device (csrc/bow.cu) and folds the per-feature results into
BowVector / FeatureVector on the host exactly like DBoW2 (TemplatedVocabulary.h:1125-1197)."""
import ctypes as C

import numpy as np

from ._lib import check, lib, ptr

# DBoW2/BowVector.h:32-56
TF_IDF, TF, IDF, BINARY = 0, 1, 2, 3
L1_NORM, L2_NORM, CHI_SQUARE, KL, BHATTACHARYYA, DOT_PRODUCT = 0, 1, 2, 3, 4, 5


class ORBVocabulary:
    def __init__(self, k, L, parent, desc, weight, scoring=L1_NORM, weighting=TF_IDF, device=0):
        """parent / desc / weight indexed by node id, node 0 = root (its entries are ignored)."""
        self.k, self.L, self.scoring, self.weighting, self.device = int(k), int(L), scoring, weighting, device
        parent = np.ascontiguousarray(parent, np.int32)
        desc = np.ascontiguousarray(desc, np.uint8)
        weight = np.ascontiguousarray(weight, np.float64)
        h = C.c_void_p()
        check(lib().orbfe_vocabulary_create(self.k, self.L, len(parent), ptr(parent), ptr(desc), ptr(weight), device,
                                            C.byref(h)))
        self.h = h
        self.n_nodes = len(parent)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orbfe_vocabulary_destroy(self.h)
            self.h = None

    @classmethod
    def loadFromTextFile(cls, path, device=0):
        """ORBvoc.txt format (TemplatedVocabulary::loadFromTextFile, :1338-1424): header `k L scoring weighting`,
        then one node per line `parent isLeaf b0..b31 weight`.  A trailing empty line is ignored (DBoW2 itself turns
        it into a childless root child with an uninitialised descriptor)."""
        with open(path) as f:
            k, L, scoring, weighting = [int(x) for x in f.readline().split()[:4]]
            rows = np.loadtxt(f, dtype=np.float64, ndmin=2)
        n = len(rows) + 1
        parent = np.zeros(n, np.int32)
        desc = np.zeros((n, 32), np.uint8)
        weight = np.zeros(n, np.float64)
        parent[1:] = rows[:, 0].astype(np.int32)
        desc[1:] = rows[:, 2:34].astype(np.uint8)
        weight[1:] = rows[:, 34]
        return cls(k, L, parent, desc, weight, scoring, weighting, device)

    def transform_features(self, desc, levelsup=4):
        """Per feature: word id, word weight, node id `levelsup` levels above the leaves (:1226-1258)."""
        desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
        n = len(desc)
        word, node, w = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float64)
        check(lib().orbfe_bow_transform(self.h, ptr(desc), n, int(levelsup), ptr(word), ptr(w), ptr(node)))
        return word, w, node

    def transform(self, desc, levelsup=4):
        """transform(features, BowVector&, FeatureVector&, levelsup), :1125-1197.
        -> (word ids, values) in BowVector order, (node ids, start, feat) = FeatureVector in map order."""
        word, w, node = self.transform_features(desc, levelsup)
        keep = np.flatnonzero(w > 0)                                   # stopped words are skipped
        tf = self.weighting in (TF_IDF, TF)
        ids, inv = np.unique(word[keep], return_inverse=True)
        vals = np.zeros(len(ids), np.float64)
        if tf:
            np.add.at(vals, inv, w[keep])                              # addWeight: accumulates in feature order
        else:
            first = np.full(len(ids), -1)
            for pos, j in enumerate(inv):                              # addIfNotExist: the first weight stays
                if first[j] < 0:
                    first[j] = pos
            vals = w[keep][first] if len(ids) else vals
        must = self.scoring != DOT_PRODUCT
        if tf and len(ids) and not must:
            vals = vals / np.float64(len(ids))
        if must and len(ids):                                          # BowVector::normalize, BowVector.cpp:62-84
            if self.scoring == L2_NORM:
                norm = 0.0
                for v in vals:
                    norm += v * v
                norm = np.sqrt(norm)
            else:
                norm = 0.0
                for v in vals:
                    norm += abs(v)
            if norm > 0.0:
                vals = vals / np.float64(norm)
        nodeu = node[keep].astype(np.uint32)                           # NodeId is unsigned: -1 ("above the level") sorts last
        order = np.lexsort((keep, nodeu))                              # FeatureVector: node ascending, features in order
        nodes, counts = np.unique(nodeu, return_counts=True)
        start = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        return (ids.astype(np.uint32), vals), (nodes.astype(np.int32), start, keep[order].astype(np.int32))

    def transform_batch_device(self, d_desc, frame_start, capacity, levelsup=4, stream=None):
        """Frame::ComputeBoW for a batch of frames whose descriptors are resident in HBM: d_desc [N, 32] uint8 CUDA tensor,
        frame b = rows frame_start[b] .. frame_start[b+1].  Descent and fold both run on the device (TF_IDF / L1_NORM).
        -> dict of CUDA tensors: bow_word [B, cap], bow_value [B, cap], n_bow [B], fv_node [B, cap], fv_start [B, cap+1],
        fv_feat [B, cap], n_fv [B]."""
        import torch
        assert self.weighting == TF_IDF and self.scoring == L1_NORM
        dev = d_desc.device
        n = d_desc.shape[0]
        fs = torch.as_tensor(np.ascontiguousarray(frame_start, np.int32), device=dev)
        B = len(frame_start) - 1
        word = torch.empty(n, dtype=torch.int32, device=dev)
        node = torch.empty(n, dtype=torch.int32, device=dev)
        w = torch.empty(n, dtype=torch.float64, device=dev)
        st = C.c_void_p(stream.cuda_stream if stream is not None else torch.cuda.current_stream(dev).cuda_stream)
        check(lib().orbfe_bow_transform_device(self.h, ptr(d_desc), n, int(levelsup), ptr(word), ptr(w), ptr(node), st))
        out = dict(bow_word=torch.empty((B, capacity), dtype=torch.int32, device=dev),
                   bow_value=torch.empty((B, capacity), dtype=torch.float64, device=dev),
                   n_bow=torch.empty(B, dtype=torch.int32, device=dev),
                   fv_node=torch.empty((B, capacity), dtype=torch.int32, device=dev),
                   fv_start=torch.empty((B, capacity + 1), dtype=torch.int32, device=dev),
                   fv_feat=torch.empty((B, capacity), dtype=torch.int32, device=dev),
                   n_fv=torch.empty(B, dtype=torch.int32, device=dev))
        check(lib().orbfe_bow_fold_device(ptr(word), ptr(w), ptr(node), ptr(fs), B, int(capacity), ptr(out["bow_word"]),
                                          ptr(out["bow_value"]), ptr(out["n_bow"]), ptr(out["fv_node"]), ptr(out["fv_start"]),
                                          ptr(out["fv_feat"]), ptr(out["n_fv"]), st))
        return out

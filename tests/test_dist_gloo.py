"""N>1 host logic on CPU: world_size-2 gloo run of the map-shard gather (orbfe.dist.gather_tables,
shard_range).  The per-shard tables come from the CPU oracle here (the CUDA kernels are covered by
tests/test_gpu_match.py::test_knn2_sharded_merge_equals_whole); the merged result must equal the
whole-map answer, ties included."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth
from oracle import oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _merge_numpy(g_idx, g_dist):
    """(distance, index) lexicographic top-2 over shards -- the order orbfe_knn2_merge_device uses."""
    G, nq, _ = g_idx.shape
    idx, dst = np.full((nq, 2), -1, np.int32), np.full((nq, 2), -1, np.int32)
    for q in range(nq):
        c = sorted((int(g_dist[s, q, k]), int(g_idx[s, q, k])) for s in range(G) for k in range(2) if g_idx[s, q, k] >= 0)
        for k, (d, i) in enumerate(c[:2]):
            idx[q, k], dst[q, k] = i, d
    return idx, dst


def _worker(rank, world, port, q_desc, t_desc, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import orbfe.dist as D
    lo, hi = D.shard_range(len(t_desc), rank, world)
    idx, dst = O.knn2(q_desc, t_desc[lo:hi])
    idx = np.where(idx >= 0, idx + lo, -1).astype(np.int32)      # train_offset = shard begin
    g_idx, g_dist = D.gather_tables(torch.from_numpy(idx), torch.from_numpy(dst))
    assert g_idx.shape == (world, len(q_desc), 2)
    m_idx, m_dist = _merge_numpy(g_idx.numpy(), g_dist.numpy())
    ret[rank] = (m_idx, m_dist)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    import orbfe.dist as D
    for n in (0, 1, 7, 4096, 1000000):
        for world in (1, 2, 3, 8):
            r = [D.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_world2_gloo_gather_and_merge():
    q, t = synth.random_descriptors(120, 1), synth.random_descriptors(901, 2)
    t[5] = q[0]; t[800] = q[0]          # equal distances in different shards: lower global index first
    t[450] = t[449]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, q, t, ret), nprocs=2, join=True)
    eidx, edist = O.knn2(q, t)
    for rank in range(2):
        assert np.array_equal(ret[rank][0], eidx) and np.array_equal(ret[rank][1], edist)


# ---- sharded SearchByProjection: the claim fixpoint across shards (host logic; the passes themselves are CUDA) ------
def _sequential_claims(pref, blocks, static):
    """The reference's loop (ORBmatcher.cc:54-167) on a toy model: point j takes the first keypoint of its preference
    list that no EARLIER blocking point holds; returns (choice per point, claim table)."""
    n = len(static)
    claims = static.copy()
    choice = np.full(len(pref), -1, np.int64)
    for j, cand in enumerate(pref):
        for k in cand:
            if claims[k] < j:
                continue
            choice[j] = k
            if blocks[j]:
                claims[k] = min(claims[k], j)
            break
    return choice, claims


def _toy_pass(pref, blocks, lo, cin, cout):
    """What orbfe_map_shard_pass does for a shard [lo, lo + len(pref)): every point against the global table cin."""
    choice = np.full(len(pref), -1, np.int64)
    for jl, cand in enumerate(pref):
        j = lo + jl
        for k in cand:
            if cin[k] < j:
                continue
            choice[jl] = k
            if blocks[jl]:
                cout[k] = min(cout[k], j)
            break
    return choice


def _toy_case(seed, m=400, n=60):
    rng = np.random.default_rng(seed)
    pref = [list(rng.choice(n, size=int(rng.integers(0, 5)), replace=False)) for _ in range(m)]   # heavy contention
    blocks = rng.uniform(size=m) < 0.8
    static = np.where(rng.uniform(size=n) < 0.1, -1, 2 ** 31 - 1).astype(np.int32)
    return pref, blocks, static


def _fixpoint_worker(rank, world, port, seed, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import orbfe.dist as D
    pref, blocks, static = _toy_case(seed)
    lo, hi = D.shard_range(len(pref), rank, world)
    last = {}

    def run_pass(cin, cout):
        c = cout.numpy()
        last["choice"] = _toy_pass(pref[lo:hi], blocks[lo:hi], lo, cin.numpy(), c)

    def combine(cout):
        dist.all_reduce(cout, op=dist.ReduceOp.MIN)      # the one exchange of a pass: 4 bytes per keypoint
        return cout
    claims, passes = D.claim_fixpoint(run_pass, torch.from_numpy(static), combine)
    ret[rank] = (lo, hi, last["choice"], claims.numpy().copy(), passes)
    dist.barrier()
    dist.destroy_process_group()


def test_claim_fixpoint_single_shard_equals_sequential_loop():
    import orbfe.dist as D
    for seed in range(5):
        pref, blocks, static = _toy_case(seed)
        last = {}

        def run_pass(cin, cout):
            last["choice"] = _toy_pass(pref, blocks, 0, cin.numpy(), cout.numpy())
        claims, passes = D.claim_fixpoint(run_pass, torch.from_numpy(static), lambda c: c)
        choice, eclaims = _sequential_claims(pref, blocks, static)
        assert np.array_equal(last["choice"], choice) and np.array_equal(claims.numpy(), eclaims) and passes >= 2


def test_world2_gloo_claim_fixpoint_equals_sequential_loop():
    seed = 11
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_fixpoint_worker, args=(2, _free_port(), seed, ret), nprocs=2, join=True)
    pref, blocks, static = _toy_case(seed)
    choice, eclaims = _sequential_claims(pref, blocks, static)
    got = np.full(len(pref), -2, np.int64)
    for rank in range(2):
        lo, hi, ch, claims, passes = ret[rank]
        got[lo:hi] = ch
        assert np.array_equal(claims, eclaims)           # the same table on every rank
    assert np.array_equal(got, choice)

"""Real multi-GPU parity (skipped below 2 devices): one process per GPU over NCCL, the map sharded over the ranks.
Brute-force kNN-2 (peer-load merge and all-gather forms, a query count that changes between calls) and the sharded
SearchByProjection (peer-load and all-reduce exchanges) against the single-GPU answers."""
import os
import socket

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import sys
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import orbfe
    import orbfe.dist as D
    out = {}
    # kNN-2: two query counts through the same ShardedMap (the symmetric tables are re-viewed per call)
    t = synth.random_descriptors(100003, 6)
    t[123] = synth.random_descriptors(700, 5)[9]
    lo, hi = D.shard_range(len(t), rank, world)
    for exchange in ("p2p", "nccl"):
        smap = D.ShardedMap(t[lo:hi], lo, dev, exchange=exchange)
        for nq in (700, 333, 1500):
            q = synth.random_descriptors(nq, 5 if nq == 700 else nq)
            idx, dst, match = smap.knn2(torch.from_numpy(q).to(dev))
            out[("knn", exchange, nq)] = (idx.cpu().numpy(), dst.cpu().numpy(), match.cpu().numpy())
        out[("knn", exchange, "how")] = smap.exchange
    # SearchByProjection, map sharded
    F, pts = D.c5_projection_case(200000, 2000, 3)
    lo, hi = D.shard_range(len(pts["u"]), rank, world)
    for exchange in ("p2p", "nccl"):
        sp = D.ShardedProjection({k: v[lo:hi] for k, v in pts.items()}, lo, dev, exchange=exchange)
        sp.set_frame(F)
        claimed = torch.zeros(len(F.keys), dtype=torch.uint8, device=dev)
        claimed[::17] = 1
        nm, asg = sp.search(claimed)
        out[("sbp", exchange)] = (nm, asg.cpu().numpy(), sp.passes, sp.exchange)
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs at least 2 GPUs")
def test_map_sharded_over_real_gpus_equals_single_gpu():
    import torch.multiprocessing as mp
    import orbfe
    import orbfe.dist as D
    world = min(_ngpu(), 8)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    t = synth.random_descriptors(100003, 6)
    t[123] = synth.random_descriptors(700, 5)[9]
    m = orbfe.ORBmatcher(0.8, True)
    for nq in (700, 333, 1500):
        q = synth.random_descriptors(nq, 5 if nq == 700 else nq)
        eidx, edist, ematch = orbfe.ORBmatcher().knn2(q, t)
        for rank in range(world):
            for exchange in ("p2p", "nccl"):
                idx, dst, match = ret[rank][("knn", exchange, nq)]
                assert np.array_equal(idx, eidx) and np.array_equal(dst, edist) and np.array_equal(match, ematch)
    F, pts = D.c5_projection_case(200000, 2000, 3)
    claimed = np.zeros(len(F.keys), np.uint8)
    claimed[::17] = 1
    en, easg, _, _ = m.SearchByProjection(F, pts, claimed, np.full(len(F.keys), -1, np.int32))
    for rank in range(world):
        for exchange in ("p2p", "nccl"):
            nm, asg, passes, how = ret[rank][("sbp", exchange)]
            assert nm == en and np.array_equal(asg, easg), (rank, exchange, how)

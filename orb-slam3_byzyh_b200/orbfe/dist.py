"""Multi-GPU plumbing for the two shardable parts of the path, one process per GPU
over torch.distributed:

  * extraction: frames are independent units -> contiguous frame ranges per rank, NO collective;
  * matching against a large map: the map is sharded by contiguous index ranges; every rank
    computes the best two train rows per query over its shard (orbfe_knn2_device with
    train_offset = shard begin, so indices are global), ONE all_gather moves the nq x 2 tables
    (NCCL over NVLink on GPUs; gloo in the CPU tests) and every rank merges them with the
    associative (distance, index) order (orbfe_knn2_merge_device), which reproduces the
    single-GPU tie-breaks exactly.

On GPUs the exchange is fused into the merge: every rank writes its shard's table into a symmetric-memory buffer
(torch.distributed._symmetric_memory: the same allocation mapped into every process over NVLink), a device barrier
orders the writes, and the merge kernel of every rank reads all shards' tables with peer loads
(orbfe_knn2_merge_peers_device) -- no all-gather.  Tables are double-buffered, so one barrier per step also
guarantees that nobody still reads a buffer when it is written again.  exchange="nccl" keeps the all-gather form.

  * SearchByProjection against a sharded map (ShardedProjection): the reference's loop is sequential over the map
    points and earlier acceptances block later points, so every pass of the claim fixpoint is run by all shards
    against ONE global claim table and the shards' new tables are combined by an elementwise minimum -- 4 bytes per
    frame keypoint per pass, by peer loads over NVLink (symmetric memory) or all_reduce(MIN).

The compute calls are CUDA only; the host-side logic here (ranges, gather layout, the fixpoint loop) is what the
world_size-2 gloo tests exercise on CPU."""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def shard_range(n, rank, world):
    """Contiguous, balanced, order-preserving partition of range(n): rank r owns [lo, hi)."""
    return rank * n // world, (rank + 1) * n // world


def gather_tables(idx2, dist2, group=None):
    """all_gather of per-shard (idx2, dist2) tables [nq,2] int32 -> [world, nq, 2] on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    g_idx = torch.empty((world,) + tuple(idx2.shape), dtype=idx2.dtype, device=idx2.device)
    g_dist = torch.empty_like(g_idx)
    if world == 1:
        g_idx[0].copy_(idx2)
        g_dist[0].copy_(dist2)
    else:
        dist.all_gather(list(g_idx.unbind(0)), idx2.contiguous(), group=group)
        dist.all_gather(list(g_dist.unbind(0)), dist2.contiguous(), group=group)
    return g_idx, g_dist


class ShardedMap:
    """This rank's shard of a map-point descriptor table, resident in HBM."""

    def __init__(self, map_desc_shard, lo, device, exchange="p2p"):
        """exchange: "p2p" = peer loads inside the merge kernel (symmetric memory over NVLink), "nccl" = all-gather.
        "p2p" falls back to "nccl" when symmetric memory cannot be set up (reported in self.exchange)."""
        self.device = device
        self.lo = int(lo)
        self.train = map_desc_shard if isinstance(map_desc_shard, torch.Tensor) else torch.from_numpy(map_desc_shard)
        self.train = self.train.to(device).contiguous()
        self.exchange = exchange
        self._symm = None      # (capacity nq, [tensor, tensor], [handle, handle], [peer pointer array, ...])
        self._step = 0

    def _symmetric_tables(self, nq, group):
        """Two symmetric buffers of {idx2[nq][2], dist2[nq][2]} int32 and the peer pointers of each."""
        if self._symm is not None and self._symm[0] >= nq:
            return self._symm
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        bufs, hdls, ptrs = [], [], []
        nq = max(nq, 4096)      # flat buffers, re-viewed per call: the query count varies from frame to frame
        for _ in range(2):
            t = symm_mem.empty(4 * nq, dtype=torch.int32, device=self.device)
            h = symm_mem.rendezvous(t, grp)
            bufs.append(t)
            hdls.append(h)
            ptrs.append((C.c_void_p * h.world_size)(*[int(p) for p in h.buffer_ptrs]))
        self._symm = (nq, bufs, hdls, ptrs)
        return self._symm

    def knn2(self, d_query, group=None):
        """d_query: [nq,32] uint8 CUDA tensor (same on every rank).  Returns (idx2, dist2, match)
        CUDA tensors with GLOBAL map indices, identical on every rank."""
        L = _lib.lib()
        nq = d_query.shape[0]
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        st = torch.cuda.current_stream(self.device).cuda_stream
        f_idx = torch.empty((nq, 2), dtype=torch.int32, device=self.device)
        f_dist = torch.empty_like(f_idx)
        match = torch.empty(nq, dtype=torch.int32, device=self.device)
        if world > 1 and self.exchange == "p2p":
            try:
                cap, bufs, hdls, ptrs = self._symmetric_tables(nq, group)
            except Exception as e:      # noqa: BLE001 -- no symmetric memory on this system: keep the all-gather
                self.exchange = "nccl (symmetric memory unavailable: %s)" % str(e)[:80]
            else:
                k = self._step & 1
                self._step += 1
                tab = bufs[k][:4 * nq].view(2, nq, 2)    # the peers kernel addresses dist2 at tab + 2 * nq
                _lib.check(L.orbfe_knn2_device(_lib.ptr(d_query), nq, _lib.ptr(self.train), self.train.shape[0], self.lo,
                                               _lib.ptr(tab[0]), _lib.ptr(tab[1]), st))
                hdls[k].barrier(channel=0)          # every shard's table is written (device-side, on this stream)
                _lib.check(L.orbfe_knn2_merge_peers_device(ptrs[k], world, nq, _lib.ptr(f_idx), _lib.ptr(f_dist),
                                                           _lib.ptr(match), st))
                return f_idx, f_dist, match
        tab = torch.empty((2, nq, 2), dtype=torch.int32, device=self.device)   # {idx2, dist2} of this shard
        _lib.check(L.orbfe_knn2_device(_lib.ptr(d_query), nq, _lib.ptr(self.train), self.train.shape[0], self.lo,
                                       _lib.ptr(tab[0]), _lib.ptr(tab[1]), st))
        if world > 1:
            packed = torch.empty((world, 2, nq, 2), dtype=torch.int32, device=self.device)
            dist.all_gather_into_tensor(packed, tab, group=group)      # the one exchange step (NCCL / NVLink)
        else:
            packed = tab
        _lib.check(L.orbfe_knn2_merge_packed_device(_lib.ptr(packed), world, nq, _lib.ptr(f_idx), _lib.ptr(f_dist),
                                                    _lib.ptr(match), st))
        return f_idx, f_dist, match


def extract_sharded(extractor, frames, lapping, rank, world):
    """frames: [B,rows,cols] host array visible to every rank (or already this rank's slice when
    `world` is None).  Extracts this rank's contiguous frame range; returns (lo, hi, n, mono, kps, desc).
    No collective: ranks never exchange data for extraction."""
    lo, hi = shard_range(frames.shape[0], rank, world)
    n, mono, kps, desc = extractor.extract_batch(frames[lo:hi], lapping)
    return lo, hi, n, mono, kps, desc


INT_MAX = 2 ** 31 - 1


def claim_fixpoint(run_pass, static_claims, combine, max_passes=1 << 20):
    """The claim fixpoint of the sharded SearchByProjection, independent of where the passes run.
    run_pass(claims_in, claims_out): this shard's points against the global table claims_in; lowers claims_out
    (pre-set to the static claims) to its own first acceptors.  combine(claims_out) -> elementwise minimum over all
    shards (identical on every rank).  Returns (final claims, number of passes)."""
    cin = static_claims.clone()
    for p in range(1, max_passes + 1):
        cout = static_claims.clone()
        run_pass(cin, cout)
        cout = combine(cout)
        if torch.equal(cin, cout):        # same table on every rank: the same decision everywhere, no extra collective
            return cout, p
        cin = cout
    raise RuntimeError("claim fixpoint did not converge")


class ShardedProjection:
    """ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, ...) (ORBmatcher.cc:46-240) against this rank's
    contiguous range [lo, lo + m) of a map that is sharded over the ranks.  Results are those of the single-GPU call on
    the whole map (orbfe_search_by_projection), on every rank."""

    def __init__(self, pts_shard, lo, device, exchange="p2p", nnratio=0.8, th_high=100):
        self.device = device if isinstance(device, torch.device) else torch.device("cuda", int(device))
        self.lo = int(lo)
        self.exchange = exchange
        self.prm = _lib.SearchParams(0, int(th_high), float(nnratio), 0)
        keep = []
        pp = _lib.ProjPoints()
        pp.m = len(pts_shard["u"])
        for name, dt in [("u", np.float32), ("v", np.float32), ("ur", np.float32), ("radius", np.float32),
                         ("min_level", np.int32), ("max_level", np.int32), ("angle", np.float32),
                         ("valid", np.uint8), ("blocks", np.uint8), ("desc", np.uint8)]:
            if pts_shard.get(name) is None:
                continue
            a = np.ascontiguousarray(pts_shard[name], dt)
            keep.append(a)
            setattr(pp, name, a.ctypes.data)
        self.m = pp.m
        h = C.c_void_p()
        _lib.check(_lib.lib().orbfe_map_shard_create(C.byref(pp), self.lo, self.device.index or 0, C.byref(h)))
        self.h = h
        self._symm = None
        self._step = 0
        self.passes = 0

    def __del__(self):
        try:
            if getattr(self, "h", None):
                _lib.lib().orbfe_map_shard_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_frame(self, F):
        """F: orbfe.FrameData (keypoints, descriptors, grid bounds, stereo coordinates) -- the same on every rank."""
        keep = []
        fv = F.view(keep)
        self.n = fv.n
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().orbfe_map_shard_set_frame(self.h, C.byref(fv), st))
        torch.cuda.current_stream(self.device).synchronize()      # the host arrays in `keep` may go now

    def _symmetric_claims(self, n, group):
        if self._symm is not None and self._symm[0] >= n:
            return self._symm
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        cap = max(n, 8192)
        bufs, hdls, ptrs = [], [], []
        for _ in range(2):
            t = symm_mem.empty(cap, dtype=torch.int32, device=self.device)
            h = symm_mem.rendezvous(t, grp)
            bufs.append(t)
            hdls.append(h)
            ptrs.append((C.c_void_p * h.world_size)(*[int(p) for p in h.buffer_ptrs]))
        self._symm = (cap, bufs, hdls, ptrs)
        return self._symm

    def search(self, claimed=None, assigned=None, group=None):
        """claimed: uint8[n] CUDA tensor or None (F.mvpMapPoints[i] holds a point with observations); assigned: int32[n]
        CUDA tensor (in/out, -1 = none) or None.  Returns (nmatches, assigned) -- identical on every rank."""
        L = _lib.lib()
        n = self.n
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        st = torch.cuda.current_stream(self.device).cuda_stream
        static = torch.empty(n, dtype=torch.int32, device=self.device)
        _lib.check(L.orbfe_claims_init_device(_lib.ptr(claimed) if claimed is not None else None, n, _lib.ptr(static), st))
        use_p2p = world > 1 and self.exchange == "p2p"
        if use_p2p:
            try:
                cap, bufs, hdls, ptrs = self._symmetric_claims(n, group)
            except Exception as e:      # noqa: BLE001 -- no symmetric memory on this system: all_reduce instead
                self.exchange = "nccl (symmetric memory unavailable: %s)" % str(e)[:80]
                use_p2p = False

        run_pass = self.run_pass

        def combine(cout):
            if world == 1:
                return cout
            if use_p2p:
                k = self._step & 1
                self._step += 1
                bufs[k][:n].copy_(cout)
                hdls[k].barrier(channel=0)          # every shard's table is written (device-side, on this stream)
                merged = torch.empty_like(cout)
                _lib.check(L.orbfe_claims_min_peers_device(ptrs[k], world, n, _lib.ptr(merged), None, None, st))
                return merged
            dist.all_reduce(cout, op=dist.ReduceOp.MIN, group=group)
            return cout
        claims, self.passes = claim_fixpoint(run_pass, static, combine)
        out = assigned.clone() if assigned is not None else torch.full((n,), -1, dtype=torch.int32, device=self.device)
        # keypoints accepted by a point lose their previous content first (the reference overwrites, :156); an entry
        # some shard raised is >= 0, the others keep what they held
        nm = torch.zeros(1, dtype=torch.int32, device=self.device)
        mine = torch.full((n,), -1, dtype=torch.int32, device=self.device)
        self.finish(mine, nm)
        if world > 1:
            dist.all_reduce(mine, op=dist.ReduceOp.MAX, group=group)
            dist.all_reduce(nm, op=dist.ReduceOp.SUM, group=group)
        out = torch.where(mine >= 0, mine, out)
        return int(nm.item()), out

    def run_pass(self, claims_in, claims_out):
        """One pass of this shard's points (device int32[n] tables; see claim_fixpoint)."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().orbfe_map_shard_pass(self.h, C.byref(self.prm), _lib.ptr(claims_in), _lib.ptr(claims_out), st))

    def finish(self, assigned, nmatches):
        """Raises assigned[k] to the global index of this shard's last accepting point; adds its match count."""
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().orbfe_map_shard_finish(self.h, _lib.ptr(assigned), _lib.ptr(nmatches), st))

    def results(self):
        bi, bd = np.empty(self.m, np.int32), np.empty(self.m, np.int32)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().orbfe_map_shard_results(self.h, _lib.ptr(bi), _lib.ptr(bd), st))
        return bi, bd


def c5_projection_case(n_map=1000000, n_frame=2000, seed=5):
    """BASELINE config 5 for SearchByProjection: 1 M projected map points against a 2000-keypoint frame (tests/synth.py,
    SURVEY 8(d)): the first n_frame points are noisy copies of the frame descriptors projected near their keypoint."""
    import synth
    from .matcher import FrameData
    d = synth.map_vs_frame(n_map, n_frame, seed)
    m = len(d["u"])
    sf = d["scale_factors"]
    radius = (np.float32(4.0) * sf[d["level"]]).astype(np.float32)      # RadiusByViewingCos = 4.0 (ORBmatcher.cc:243-250)
    pts = dict(u=d["u"], v=d["v"], ur=d["u"], radius=radius, min_level=(d["level"] - 1).astype(np.int32),
               max_level=d["level"].astype(np.int32), valid=np.ones(m, np.uint8), blocks=np.ones(m, np.uint8),
               desc=d["mdesc"])
    F = FrameData(d["keys"], d["fdesc"], d["bounds"], None)
    return F, pts


def bench_sharded_projection(torch_, dist_, orbfe, dev, rank, world, steps, barrier, max_over_ranks):
    """map points/s of the sharded SearchByProjection on the C5 case; checks on every run that the sharded result equals
    the single-GPU call (rank 0 runs it on the whole map)."""
    F, pts = c5_projection_case()
    m = len(pts["u"])
    lo, hi = shard_range(m, rank, world)
    shard = {k: (v[lo:hi] if v is not None else None) for k, v in pts.items()}
    sp = ShardedProjection(shard, lo, dev)
    sp.set_frame(F)
    for _ in range(2):
        nm, asg = sp.search()
    barrier()
    e0, e1 = torch_.cuda.Event(enable_timing=True), torch_.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        nm, asg = sp.search()
    e1.record()
    torch_.cuda.synchronize()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    out = {"metric": "SearchByProjection map points/s (1 M projected map points x 2000-keypoint frame)",
           "value": m / (ms / 1e3), "unit": "map points/s", "ms_per_step": ms, "map_shards": world,
           "passes": sp.passes, "matches": nm,
           "exchange": "none" if world == 1 else sp.exchange + ": elementwise min of the claim tables, 4 B per frame keypoint per pass",
           "algorithmic_bytes_per_point": 60, "hbm_frac": m * 60 / (ms / 1e3) / 1e9 / (hbm_peak_gbs() * world)}
    if rank == 0:
        mt = orbfe.ORBmatcher(0.8, True, device=dev.index or 0)
        cl, a0 = np.zeros(len(F.keys), np.uint8), np.full(len(F.keys), -1, np.int32)
        n1, a1, _, _ = mt.SearchByProjection(F, pts, cl, a0)
        same = bool(n1 == nm and np.array_equal(a1, asg.cpu().numpy()))
        assert same, "sharded SearchByProjection differs from the single-GPU call"
        out["equals_single_gpu_call"] = same
    return out


def hbm_peak_gbs():
    import json
    import os
    f = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(f))["hbm_gbs"])
    except Exception:
        return 6650.0

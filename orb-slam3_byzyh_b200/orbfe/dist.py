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

The compute calls are CUDA only; the host-side logic here (ranges, gather layout) is what the
world_size-2 gloo tests exercise on CPU."""
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
        import ctypes as C
        import torch.distributed._symmetric_memory as symm_mem
        grp = group if group is not None else dist.group.WORLD
        bufs, hdls, ptrs = [], [], []
        for _ in range(2):
            t = symm_mem.empty((2, nq, 2), dtype=torch.int32, device=self.device)
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
                assert cap == nq, "the symmetric tables are laid out for a fixed query count"
                k = self._step & 1
                self._step += 1
                tab = bufs[k]
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

"""Multi-GPU plumbing for the two shardable parts of the path, one process per GPU
over torch.distributed:

  * extraction: frames are independent units -> contiguous frame ranges per rank, NO collective;
  * matching against a large map: the map is sharded by contiguous index ranges; every rank
    computes the best two train rows per query over its shard (orbfe_knn2_device with
    train_offset = shard begin, so indices are global), ONE all_gather moves the nq x 2 tables
    (NCCL over NVLink on GPUs; gloo in the CPU tests) and every rank merges them with the
    associative (distance, index) order (orbfe_knn2_merge_device), which reproduces the
    single-GPU tie-breaks exactly.

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

    def __init__(self, map_desc_shard, lo, device):
        self.device = device
        self.lo = int(lo)
        self.train = map_desc_shard if isinstance(map_desc_shard, torch.Tensor) else torch.from_numpy(map_desc_shard)
        self.train = self.train.to(device).contiguous()

    def knn2(self, d_query, group=None):
        """d_query: [nq,32] uint8 CUDA tensor (same on every rank).  Returns (idx2, dist2, match)
        CUDA tensors with GLOBAL map indices, identical on every rank."""
        L = _lib.lib()
        nq = d_query.shape[0]
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        tab = torch.empty((2, nq, 2), dtype=torch.int32, device=self.device)   # {idx2, dist2} of this shard
        st = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(L.orbfe_knn2_device(_lib.ptr(d_query), nq, _lib.ptr(self.train), self.train.shape[0], self.lo,
                                       _lib.ptr(tab[0]), _lib.ptr(tab[1]), st))
        if world > 1:
            packed = torch.empty((world, 2, nq, 2), dtype=torch.int32, device=self.device)
            dist.all_gather_into_tensor(packed, tab, group=group)      # the one exchange step (NCCL / NVLink)
        else:
            packed = tab
        f_idx = torch.empty((nq, 2), dtype=torch.int32, device=self.device)
        f_dist = torch.empty_like(f_idx)
        match = torch.empty(nq, dtype=torch.int32, device=self.device)
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

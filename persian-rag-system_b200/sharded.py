"""Row-sharded exact search across the GPUs of one box.

One process per GPU (torchrun).  Rank g holds the contiguous row block
[g*ceil(N/G), min(N, (g+1)*ceil(N/G))) in its own FlatIndex; queries are replicated.  A search is
  local fused scan + top-k (csrc kernels)  ->  ncclAllGather of the [nq, k] (score, global id)
  lists over NVLink  ->  on-device G-way merge (prs_merge_topk_device),
so only 12*nq*k bytes per rank cross NVLink.  Ties are broken on GLOBAL ids, so the sharded
result is bit-identical to the unsharded one.  With the `gloo` backend (CPU tests) the gather
goes through host tensors; the merge is still the device kernel when a GPU is present.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import METRIC_INNER_PRODUCT, METRIC_L2, check
from .flat import FlatIndex


def shard_bounds(n_total: int, world: int, rank: int):
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def merge_topk(D_parts, I_parts, largest: bool, tie_high_id: bool = False):
    """[G, nq, k] CUDA tensors -> merged ([nq, k], [nq, k]) via the device kernel."""
    import torch
    G, nq, k = (int(s) for s in D_parts.shape)
    D_parts = D_parts.contiguous()
    I_parts = I_parts.contiguous()
    D = torch.empty((nq, k), dtype=torch.float32, device=D_parts.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=D_parts.device)
    st = torch.cuda.current_stream(D_parts.device).cuda_stream
    check(_lib.lib().prs_merge_topk_device(ctypes.c_void_p(D_parts.data_ptr()), ctypes.c_void_p(I_parts.data_ptr()), G, nq, k,
                                           1 if largest else 0, 1 if tie_high_id else 0, ctypes.c_void_p(D.data_ptr()),
                                           ctypes.c_void_p(I.data_ptr()), int(D_parts.device.index or 0), ctypes.c_void_p(st)))
    return D, I


class ShardedFlatIndex:
    """FlatIndex whose rows are split across the ranks of a torch.distributed process group."""

    def __init__(self, d: int, metric: int = METRIC_L2, storage="fp16", group=None, device: int | None = None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local = FlatIndex(d, metric, storage, device)
        self.metric = metric
        self.d = d
        self.offset = 0
        self.ntotal_global = 0

    def add_local(self, x, global_offset: int, n_total_global: int) -> None:
        """Add this rank's row block; `global_offset` is the global id of its first row."""
        self.local.add(x)
        self.offset = int(global_offset)
        self.local.set_id_offset(self.offset)
        self.ntotal_global = int(n_total_global)

    @property
    def ntotal(self) -> int:
        return self.ntotal_global

    def search(self, q, k: int):
        """q: CUDA tensor [nq, d] replicated on every rank -> (D, I) CUDA tensors on every rank."""
        import torch
        D, I = self.local.search(q, k)
        if self.world == 1:
            return D, I
        nq = int(D.shape[0])
        Dg = torch.empty((self.world * nq, k), dtype=D.dtype, device=D.device)      # rank-major concatenation
        Ig = torch.empty((self.world * nq, k), dtype=I.dtype, device=I.device)
        self.dist.all_gather_into_tensor(Dg, D, group=self.group)
        self.dist.all_gather_into_tensor(Ig, I, group=self.group)
        return merge_topk(Dg.view(self.world, nq, k), Ig.view(self.world, nq, k), largest=self.metric == METRIC_INNER_PRODUCT)


def merge_topk_host_lists(D_parts: np.ndarray, I_parts: np.ndarray, largest: bool, tie_high_id: bool = False):
    """Host-side restatement of the merge RULE (not a compute fallback for search): used by the
    gloo CPU tests to check the sharding arithmetic (bounds, offsets, tie order on global ids)."""
    G, nq, k = D_parts.shape
    D = np.empty((nq, k), D_parts.dtype)
    I = np.empty((nq, k), np.int64)
    for q in range(nq):
        d = D_parts[:, q, :].reshape(-1)
        i = I_parts[:, q, :].reshape(-1)
        valid = i >= 0
        key = np.where(valid, -d if largest else d, np.inf)
        tie = -i if tie_high_id else i
        order = np.lexsort((tie, key))[:k]
        D[q], I[q] = d[order], i[order]
        bad = ~valid[order]
        I[q][bad] = -1
    return D, I

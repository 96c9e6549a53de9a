"""Row-sharded exact search across the GPUs of one box.

One process per GPU (torchrun).  Rank g holds the contiguous row block
[g*ceil(N/G), min(N, (g+1)*ceil(N/G))) in its own FlatIndex; queries are replicated.  A search is
the local fused scan + top-k, then ONE kernel (csrc/xchg.cuh) that merges the local per-CTA lists,
stores the local top-k straight into every peer's exchange buffer over NVLink (CUDA IPC mapped peer
memory), waits for the peers' lists and merges the G lists -- no collective launch, 12*nq*k bytes per
rank pair.  `exchange="nccl"` keeps the plain variant (ncclAllGather of the [nq, k] lists +
prs_merge_topk_device), which is also what searches larger than the exchange buffer use.  Ties are broken on GLOBAL ids, so the sharded
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


def merge_topk_f64(S_parts, I_parts, tie_high_id: bool = True):
    """[G, nq, k] float64 scores + int64 global ids (CUDA tensors) -> merged ([nq, k], [nq, k]), largest
    first; ties by id descending (the sparse path's order) unless `tie_high_id` is False."""
    import torch
    G, nq, k = (int(s) for s in S_parts.shape)
    S_parts = S_parts.to(torch.float64).contiguous()
    I_parts = I_parts.to(torch.int64).contiguous()
    S = torch.empty((nq, k), dtype=torch.float64, device=S_parts.device)
    I = torch.empty((nq, k), dtype=torch.int64, device=S_parts.device)
    st = torch.cuda.current_stream(S_parts.device).cuda_stream
    check(_lib.lib().prs_merge_topk_f64_device(ctypes.c_void_p(S_parts.data_ptr()), ctypes.c_void_p(I_parts.data_ptr()), G, nq, k,
                                               1 if tie_high_id else 0, ctypes.c_void_p(S.data_ptr()), ctypes.c_void_p(I.data_ptr()),
                                               int(S_parts.device.index or 0), ctypes.c_void_p(st)))
    return S, I


class ShardedSparseIndex:
    """Sparse (BM25 / TF-IDF) index split by DOC RANGE across the ranks of a process group:
    rank g holds the postings of docs [lo_g, hi_g) -- the rows of the global doc-by-term weight matrix that
    belong to its block; idf / avgdl / vocabulary are global quantities baked into those weights by whoever
    built the matrix (e.g. `sparse.build_bm25_csr` over the whole corpus, then sliced by rows).  A search
    scores the local block, all-gathers the [nq, k] (score, global id) lists over NCCL and merges them on
    the device with ties on GLOBAL ids, so the result equals the unsharded index's."""

    def __init__(self, indptr, indices, values, n_terms: int, doc_offset: int, n_docs_global: int, group=None,
                 device: int | None = None, mode: str = "exact"):
        import torch.distributed as dist
        from .sparse import SparseIndex
        self.dist, self.group = dist, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.local = SparseIndex(indptr, indices, values, n_terms, device, mode)
        self.local.set_id_offset(int(doc_offset))
        self.offset, self.ndocs_global = int(doc_offset), int(n_docs_global)
        if self.world > 1:
            dist.barrier(group=group)

    @classmethod
    def from_global_csr(cls, indptr, indices, values, n_terms: int, group=None, device: int | None = None, mode: str = "exact"):
        """Every rank holds (or can read) the global CSR: keep only this rank's contiguous row block."""
        import torch.distributed as dist
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        indptr = np.asarray(indptr, dtype=np.int64)
        n = int(indptr.shape[0] - 1)
        lo, hi = shard_bounds(n, world, rank)
        a, b = int(indptr[lo]), int(indptr[hi])
        return cls(indptr[lo:hi + 1] - indptr[lo], np.asarray(indices)[a:b], np.asarray(values)[a:b], n_terms, lo, n, group, device, mode)

    def search_device(self, q_indptr, q_terms, q_weights, k: int):
        """Query CSR as CUDA tensors (replicated on every rank) -> (S float64 [nq, k], I int64 [nq, k]) CUDA
        tensors with global doc ids, identical on every rank."""
        import torch
        S, I = self.local.search_device(q_indptr, q_terms, q_weights, k)
        if self.world == 1:
            return S, I
        nq = int(S.shape[0])
        Sg = torch.empty((self.world * nq, k), dtype=S.dtype, device=S.device)
        Ig = torch.empty((self.world * nq, k), dtype=I.dtype, device=I.device)
        self.dist.all_gather_into_tensor(Sg, S, group=self.group)
        self.dist.all_gather_into_tensor(Ig, I, group=self.group)
        return merge_topk_f64(Sg.view(self.world, nq, k), Ig.view(self.world, nq, k), tie_high_id=True)

    def search(self, q_indptr, q_terms, q_weights, k: int):
        """numpy in / numpy out convenience around `search_device`."""
        import torch
        dev = torch.device("cuda", self.local.device)
        S, I = self.search_device(torch.from_numpy(np.ascontiguousarray(q_indptr, dtype=np.int64)).to(dev),
                                  torch.from_numpy(np.ascontiguousarray(q_terms, dtype=np.int32)).to(dev),
                                  torch.from_numpy(np.ascontiguousarray(q_weights, dtype=np.float64)).to(dev), k)
        return S.cpu().numpy(), I.cpu().numpy()


class ShardedFlatIndex:
    """FlatIndex whose rows are split across the ranks of a torch.distributed process group."""

    def __init__(self, d: int, metric: int = METRIC_L2, storage="fp16", group=None, device: int | None = None,
                 exchange: str = "p2p", nq_cap: int = 1024, k_cap: int = 128, lanes: int = 2):
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
        self.exchange = exchange
        # `lanes` independent exchange contexts, used round-robin: consecutive searches issued on
        # different CUDA streams can then be in flight together (each context has its own slots and
        # generation counter; every rank must issue the same sequence of searches)
        self._xs = []
        self._calls = 0
        self._cap = (int(nq_cap), int(nq_cap) * int(k_cap))
        if self.world > 1 and exchange == "p2p":
            for _ in range(max(1, int(lanes))):
                self._xs.append(self._open_exchange(nq_cap, k_cap))

    def _open_exchange(self, nq_cap: int, k_cap: int):
        """Create one exchange buffer of this rank and map every peer's over CUDA IPC (NVLink peer memory)."""
        import torch
        L = _lib.lib()
        dev = self.local.device
        x = ctypes.c_void_p()
        check(L.prs_xchg_create(dev, self.world, self.rank, int(nq_cap), int(k_cap), ctypes.byref(x)))
        hb = int(L.prs_xchg_handle_bytes())
        mine = (ctypes.c_ubyte * hb)()
        check(L.prs_xchg_get_handle(x, mine))
        on_gpu = self.dist.get_backend(self.group) == "nccl"
        tdev = torch.device("cuda", dev) if on_gpu else torch.device("cpu")
        t = torch.tensor(list(mine), dtype=torch.uint8, device=tdev)
        allh = torch.empty(self.world * hb, dtype=torch.uint8, device=tdev)
        self.dist.all_gather_into_tensor(allh, t, group=self.group)
        raw = bytes(allh.cpu().tolist())
        check(L.prs_xchg_open_peers(x, raw))
        self.dist.barrier(group=self.group)
        return x

    def search_into(self, q_ptr: int, nq: int, k: int, D_ptr: int, I_ptr: int, stream: int = 0) -> None:
        """Raw-pointer variant for page-locked host buffers (or device buffers): float32 queries at
        `q_ptr`, results written to `D_ptr` / `I_ptr`.  Under unified addressing pinned host memory is
        directly readable and writable by the kernels, so a search needs no staging copies; the caller
        synchronises `stream` before reading the results.  Falls back to an error when the fused
        exchange is not available (use `search` with CUDA tensors then)."""
        if not (self.world > 1 and self._xs and int(nq) <= self._cap[0] and int(nq) * int(k) <= self._cap[1]):
            if self.world == 1:
                check(_lib.lib().prs_index_search_device(self.local._h, ctypes.c_void_p(q_ptr), _lib.F32, int(nq), int(k),
                                                         ctypes.c_void_p(D_ptr), ctypes.c_void_p(I_ptr), ctypes.c_void_p(stream)))
                return
            raise _lib.PrsError(_lib.EINVAL, "search_into needs the p2p exchange and nq, k within its capacity")
        x = self._xs[self._calls % len(self._xs)]
        self._calls += 1
        check(_lib.lib().prs_index_search_sharded_device(self.local._h, x, ctypes.c_void_p(q_ptr), _lib.F32, int(nq), int(k),
                                                         ctypes.c_void_p(D_ptr), ctypes.c_void_p(I_ptr), ctypes.c_void_p(stream)))

    def check_exchange(self) -> None:
        """Raises if any fused search timed out waiting for a peer (synchronises the device)."""
        for x in self._xs:
            check(_lib.lib().prs_xchg_status(x))

    def __del__(self):
        try:
            for x in getattr(self, "_xs", []):
                if x.value:
                    _lib.lib().prs_xchg_free(x)
            self._xs = []
        except Exception:
            pass

    def write(self, dirpath: str) -> None:
        """Persist the sharded corpus: every rank writes its block as one shard file (the HBM image + norms),
        rank 0 the manifest (container.py)."""
        from .container import write_sharded
        write_sharded(self, dirpath)

    def load(self, dirpath: str) -> None:
        """Load this rank's shard file(s) of a container written by `write` (any number of files >= ranks),
        in parallel with the other ranks: a straight disk -> pinned -> HBM copy."""
        from .container import read_manifest, read_sharded
        man = read_manifest(dirpath)
        idx = read_sharded(dirpath, self.local.device, self.rank, self.world)
        if idx.d != self.d or idx.metric_type != self.metric or idx.storage != self.local.storage:
            raise _lib.PrsError(_lib.EINVAL, "container does not match this index (d / metric / storage)")
        self.local = idx
        from .container import shards_for_rank
        mine = list(shards_for_rank(len(man["shards"]), self.world, self.rank))
        self.offset = int(man["shards"][mine[0]]["id_offset"])
        self.ntotal_global = int(man["ntotal"])
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def add_local(self, x, global_offset: int, n_total_global: int) -> None:
        """Add this rank's row block; `global_offset` is the global id of its first row."""
        self.local.add(x)
        self.offset = int(global_offset)
        self.local.set_id_offset(self.offset)
        self.ntotal_global = int(n_total_global)
        # every rank must hold its block before anyone searches: the fused exchange waits (bounded) for
        # the peers' lists, and a rank still building its shard would make the others time out
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def set_exchange_timeout(self, seconds: float) -> None:
        """How long a fused search waits for a peer's lists (default 2 s).  A search that gives up
        returns ids -1 for the affected queries and the NEXT search (or `check_exchange`) raises."""
        for x in self._xs:
            check(_lib.lib().prs_xchg_set_timeout_ms(x, max(1, int(seconds * 1000))))

    @property
    def ntotal(self) -> int:
        return self.ntotal_global

    def search(self, q, k: int):
        """q: CUDA tensor [nq, d] replicated on every rank -> (D, I) CUDA tensors on every rank."""
        import torch
        if self.world > 1 and self._xs and int(q.shape[0]) <= self._cap[0] and int(q.shape[0]) * int(k) <= self._cap[1]:
            x = self._xs[self._calls % len(self._xs)]
            self._calls += 1
            from .flat import _torch_dtype_code
            q = q.contiguous()
            nq = int(q.shape[0])
            D = torch.empty((nq, k), dtype=torch.float32, device=q.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=q.device)
            st = torch.cuda.current_stream(q.device).cuda_stream
            check(_lib.lib().prs_index_search_sharded_device(self.local._h, x, ctypes.c_void_p(q.data_ptr()), _torch_dtype_code(q),
                                                             nq, int(k), ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()),
                                                             ctypes.c_void_p(st)))
            return D, I
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

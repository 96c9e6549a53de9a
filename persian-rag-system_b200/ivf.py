"""IVF-Flat on the flat-scan kernels -- the reference's index for >= 1000 embeddings.

  index = faiss.IndexIVFFlat(faiss.IndexFlatL2(d), d, nlist)      scripts/phase3_pdf_chunking.py:49-50
  index.train(embeddings[:10000]); index.add(batch)               :53-54, :64
  index.search(q, k)  with faiss's default nprobe = 1             (queried through src/retrieval.py:102)

faiss 1.7.4 semantics restated: k-means with the Level1Quantizer defaults (10 iterations, seed 1234, at most 256
training points per centroid, initial centroids = the first nlist entries of faiss's `rand_perm(n, seed + 1)`,
centroid = float32 sum in row order x 1/count, empty clusters re-seeded by splitting a big one), vectors appended to
the inverted list of their nearest centroid, a search scans the `nprobe` nearest lists and keeps the k smallest
squared L2 distances (padding: id -1, distance FLT_MAX).  Every distance computation -- assignment, coarse probe,
list scan -- is an exact flat search on the GPU (FlatIndex: one small index per inverted list); the centroid
update is csrc/ivf.cu.  This is an APPROXIMATE index by construction (the reference silently switches to it); the
exact FlatIndex is a superset in recall.  `faiss.write_index` of an IVF index is not reproduced.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional

import numpy as np

from . import _lib
from ._lib import METRIC_L2, PrsError, check
from .flat import FlatIndex, _default_device

FLT_MAX = np.finfo(np.float32).max


class _MT19937:
    """std::mt19937 seeded like faiss's RandomGenerator (`mt((unsigned)seed)`)."""

    def __init__(self, seed: int):
        mt = [0] * 624
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt, self.idx = mt, 624

    def _twist(self):
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            v = mt[(i + 397) % 624] ^ (y >> 1)
            if y & 1:
                v ^= 0x9908B0DF
            mt[i] = v
        self.idx = 0

    def u32(self) -> int:
        if self.idx >= 624:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def rand_int(self, mx: int) -> int:          # faiss RandomGenerator::rand_int(max)
        return self.u32() % mx

    def rand_float(self) -> float:               # faiss RandomGenerator::rand_float()
        return float(np.float32(self.u32()) / np.float32(0xFFFFFFFF))


def faiss_rand_perm(n: int, seed: int) -> np.ndarray:
    """faiss::rand_perm: Fisher-Yates with `i + rng.rand_int(n - i)`."""
    rng = _MT19937(seed)
    perm = np.arange(n, dtype=np.int64)
    for i in range(n - 1):
        j = i + rng.rand_int(n - i)
        perm[i], perm[j] = perm[j], perm[i]
    return perm


def split_empty_clusters(centroids: np.ndarray, counts: np.ndarray, n: int) -> int:
    """faiss Clustering split_clusters: every empty cluster takes (a perturbed copy of) a cluster picked with
    probability proportional to its size; in place on float32 `centroids` [k, d] and `counts` [k]."""
    k, d = centroids.shape
    eps = np.float32(1.0 / 1024.0)
    rng = _MT19937(1234)
    nsplit = 0
    sign = np.where(np.arange(d) % 2 == 0, np.float32(1), np.float32(-1))
    for ci in range(k):
        if counts[ci] != 0:
            continue
        cj = 0
        while True:
            p = (float(counts[cj]) - 1.0) / float(n - k)
            if rng.rand_float() < p:
                break
            cj = (cj + 1) % k
        centroids[ci] = centroids[cj]
        centroids[ci] *= (np.float32(1) + sign * eps)
        centroids[cj] *= (np.float32(1) - sign * eps)
        counts[ci] = counts[cj] // 2
        counts[cj] -= counts[ci]
        nsplit += 1
    return nsplit


class IndexIVFFlat:
    """`faiss.IndexIVFFlat(faiss.IndexFlatL2(d), d, nlist)` with every scan on the GPU."""

    MAX_POINTS_PER_CENTROID = 256
    NITER = 10
    SEED = 1234

    def __init__(self, d: int, nlist: int, metric: int = METRIC_L2, storage: str = "fp32", device: Optional[int] = None):
        if metric != METRIC_L2:
            raise PrsError(_lib.EUNSUP, "IndexIVFFlat: only METRIC_L2 (what the reference builds)")
        self.d, self.nlist, self.nprobe = int(d), int(nlist), 1
        self.metric_type, self.storage = metric, storage
        self.device = _default_device() if device is None else int(device)
        self.quantizer = FlatIndex(self.d, METRIC_L2, "fp32", self.device)
        self.is_trained = False
        self.ntotal = 0
        self._lists: List[Optional[FlatIndex]] = [None] * self.nlist
        self._ids: List[np.ndarray] = [np.empty(0, np.int64) for _ in range(self.nlist)]
        self.centroids: Optional[np.ndarray] = None
        self.last_nsplit = 0

    # ------------------------------------------------------------------ training (faiss Clustering::train)
    def train(self, x) -> None:
        import torch
        x = self._as_f32(x)
        n = int(x.shape[0])
        k = self.nlist
        if n < k:
            raise PrsError(_lib.EINVAL, f"Number of training points ({n}) should be at least as large as number of clusters ({k})")
        if n > k * self.MAX_POINTS_PER_CENTROID:                       # subsample_training_set
            perm = faiss_rand_perm(n, self.SEED)
            x = np.ascontiguousarray(x[perm[: k * self.MAX_POINTS_PER_CENTROID]])
            n = int(x.shape[0])
        dev = torch.device("cuda", self.device)
        xt = torch.from_numpy(x).to(dev)
        perm = faiss_rand_perm(n, self.SEED + 1)
        cent = xt[torch.from_numpy(perm[:k]).to(dev)].contiguous()
        counts = torch.zeros(k, dtype=torch.int64, device=dev)
        L = _lib.lib()
        self.last_nsplit = 0
        for _ in range(self.NITER):
            q = FlatIndex(self.d, METRIC_L2, "fp32", self.device)
            q.add(cent)
            _D, I = q.search(xt, 1)                                     # nearest centroid of every training row
            assign = I[:, 0].contiguous()
            st = torch.cuda.current_stream(dev).cuda_stream
            check(L.prs_centroid_update_device(ctypes.c_void_p(xt.data_ptr()), n, self.d, ctypes.c_void_p(assign.data_ptr()), k,
                                               ctypes.c_void_p(cent.data_ptr()), ctypes.c_void_p(counts.data_ptr()), self.device,
                                               ctypes.c_void_p(st)))
            hc = counts.cpu().numpy()
            if (hc == 0).any():                                         # rare: re-seed empty clusters like faiss does
                ch = cent.cpu().numpy()
                self.last_nsplit += split_empty_clusters(ch, hc, n)
                cent = torch.from_numpy(ch).to(dev)
        self.centroids = cent.cpu().numpy()
        self.quantizer = FlatIndex(self.d, METRIC_L2, "fp32", self.device)
        self.quantizer.add(cent)
        self.is_trained = True

    def set_centroids(self, centroids) -> None:
        """A pre-trained coarse quantizer (faiss: an IndexFlatL2 that already holds nlist vectors)."""
        c = self._as_f32(centroids)
        if c.shape != (self.nlist, self.d):
            raise PrsError(_lib.EINVAL, f"set_centroids: expected [{self.nlist}, {self.d}], got {c.shape}")
        self.centroids = c.copy()
        self.quantizer = FlatIndex(self.d, METRIC_L2, "fp32", self.device)
        self.quantizer.add(c)
        self.is_trained = True

    # ------------------------------------------------------------------ add / search
    @staticmethod
    def _as_f32(x) -> np.ndarray:
        if hasattr(x, "detach"):
            x = x.detach().float().cpu().numpy()
        return np.ascontiguousarray(x, dtype=np.float32)

    def add(self, x) -> None:
        if not self.is_trained:
            raise PrsError(_lib.EINVAL, "IndexIVFFlat.add: the index is not trained")
        x = self._as_f32(x)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise PrsError(_lib.EINVAL, f"add: expected [n, {self.d}], got {x.shape}")
        if x.shape[0] == 0:
            return
        _D, I = self.quantizer.search(x, 1)
        assign = I[:, 0]
        ids = np.arange(self.ntotal, self.ntotal + x.shape[0], dtype=np.int64)
        for l in np.unique(assign):
            sel = np.nonzero(assign == l)[0]
            if self._lists[l] is None:
                self._lists[l] = FlatIndex(self.d, METRIC_L2, self.storage, self.device)
            self._lists[l].add(x[sel])
            self._ids[l] = np.concatenate([self._ids[l], ids[sel]])
        self.ntotal += int(x.shape[0])

    def search(self, x, k: int):
        """(D float32 [nq, k] squared L2 ascending, I int64 [nq, k]); fewer than k rows in the probed lists -> (FLT_MAX, -1)."""
        k = int(k)
        as_tensor = hasattr(x, "is_cuda") and x.is_cuda
        q = self._as_f32(x)
        if q.ndim != 2 or q.shape[1] != self.d:
            raise PrsError(_lib.EINVAL, f"search: expected [nq, {self.d}], got {q.shape}")
        nq = q.shape[0]
        D = np.full((nq, k), FLT_MAX, np.float32)
        I = np.full((nq, k), -1, np.int64)
        if nq == 0 or self.ntotal == 0:
            return self._out(D, I, x, as_tensor)
        nprobe = min(int(self.nprobe), self.nlist)
        _Dc, Ic = self.quantizer.search(q, nprobe)
        cand_d = [[] for _ in range(nq)]
        cand_i = [[] for _ in range(nq)]
        for l in np.unique(Ic[Ic >= 0]):
            lst = self._lists[l]
            if lst is None:
                continue
            rows = np.nonzero((Ic == l).any(axis=1))[0]
            Dl, Il = lst.search(q[rows], min(k, max(1, lst.ntotal)))
            for r, dr, ir in zip(rows, Dl, Il):
                ok = ir >= 0
                cand_d[r].append(dr[ok])
                cand_i[r].append(self._ids[l][ir[ok]])
        for r in range(nq):
            if not cand_d[r]:
                continue
            d = np.concatenate(cand_d[r])
            i = np.concatenate(cand_i[r])
            order = np.lexsort((i, d))[:k]
            D[r, : order.size], I[r, : order.size] = d[order], i[order]
        return self._out(D, I, x, as_tensor)

    @staticmethod
    def _out(D, I, x, as_tensor):
        if not as_tensor:
            return D, I
        import torch
        return torch.from_numpy(D).to(x.device), torch.from_numpy(I).to(x.device)

    def list_sizes(self) -> np.ndarray:
        return np.array([a.shape[0] for a in self._ids], dtype=np.int64)

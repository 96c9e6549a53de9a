"""Hybrid fusion (dense + BM25) on the device -- the fusion loop of `retrieve_hybrid`
(src/retrieval.py:181-216) as one kernel over a whole batch of queries (csrc/hybrid.cu).

Inputs are the two top-2k lists exactly as the flat scan and the sparse scoring kernels leave them
in HBM; only the fused [nq, k] lists are copied back."""
from __future__ import annotations

import ctypes

from . import _lib
from ._lib import check


def hybrid_fuse(D_dense, I_dense, S_sparse, I_sparse, n_chunks: int, top_k: int,
                dense_weight: float = 0.6, bm25_weight: float = 0.4):
    """D_dense float32 [nq, kd] squared L2 (ascending), I_dense int64 [nq, kd]; S_sparse float64 [nq, ks],
    I_sparse int64 [nq, ks] -- CUDA tensors.  Returns (S float64 [nq, top_k], I int64 [nq, top_k]) CUDA
    tensors: fused score descending, ties in the reference's insertion order, -1 padded.
    Asynchronous on the current stream."""
    import torch
    if not (D_dense.is_cuda and I_dense.is_cuda and S_sparse.is_cuda and I_sparse.is_cuda):
        raise _lib.PrsError(_lib.ECUDA, "hybrid_fuse needs CUDA tensors: there is no CPU fallback")
    nq, kd = (int(v) for v in D_dense.shape)
    ks = int(S_sparse.shape[1])
    if tuple(I_dense.shape) != (nq, kd) or tuple(S_sparse.shape) != (nq, ks) or tuple(I_sparse.shape) != (nq, ks):
        raise _lib.PrsError(_lib.EINVAL, "hybrid_fuse: shape mismatch")
    D_dense = D_dense.to(torch.float32).contiguous()
    I_dense = I_dense.to(torch.int64).contiguous()
    S_sparse = S_sparse.to(torch.float64).contiguous()
    I_sparse = I_sparse.to(torch.int64).contiguous()
    dev = D_dense.device
    S = torch.empty((nq, int(top_k)), dtype=torch.float64, device=dev)
    I = torch.empty((nq, int(top_k)), dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    check(_lib.lib().prs_hybrid_fuse_device(ctypes.c_void_p(D_dense.data_ptr()), ctypes.c_void_p(I_dense.data_ptr()), kd,
                                            ctypes.c_void_p(S_sparse.data_ptr()), ctypes.c_void_p(I_sparse.data_ptr()), ks,
                                            nq, int(n_chunks), float(dense_weight), float(bm25_weight), int(top_k),
                                            ctypes.c_void_p(S.data_ptr()), ctypes.c_void_p(I.data_ptr()),
                                            int(dev.index or 0), ctypes.c_void_p(st)))
    return S, I

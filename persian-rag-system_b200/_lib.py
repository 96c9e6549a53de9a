"""ctypes binding of libprs.so (C ABI in include/prs.h).

There is no CPU fallback: if the shared library is missing this module raises, and every compute
entry point of the library itself fails with PRS_ECUDA when no sm_100a device is present."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# PRS_LIB_PATH selects another build of the same library (A/B runs of experiment builds: `make EXPERIMENTS=1 OUT=...`)
LIB_PATH = os.environ.get("PRS_LIB_PATH") or os.path.join(_HERE, "libprs.so")

OK, EINVAL, ECUDA, EIO, ENOMEM, EUNSUP = 0, -1, -2, -3, -4, -5
METRIC_INNER_PRODUCT, METRIC_L2 = 0, 1
F32, F16, BF16, F64 = 0, 1, 2, 3
MAX_K = 1024

c_void_p, c_int, c_i64, c_char_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_char_p

# name -> (restype, argtypes); every symbol declared in include/prs.h
SIGNATURES = {
    "prs_last_error": (c_char_p, []),
    "prs_launch_count": (c_i64, []),
    "prs_device_arch": (c_int, [c_int]),
    "prs_index_create": (c_int, [c_int, c_int, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "prs_index_free": (None, [c_void_p]),
    "prs_index_reserve": (c_int, [c_void_p, c_i64]),
    "prs_index_add_host": (c_int, [c_void_p, c_void_p, c_i64]),
    "prs_index_add_device": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_void_p]),
    "prs_index_ntotal": (c_i64, [c_void_p]),
    "prs_index_d": (c_int, [c_void_p]),
    "prs_index_metric": (c_int, [c_void_p]),
    "prs_index_storage": (c_int, [c_void_p]),
    "prs_index_search_host": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p]),
    "prs_index_search_device": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_int, c_void_p, c_void_p, c_void_p]),
    "prs_index_set_id_offset": (c_int, [c_void_p, c_i64]),
    "prs_index_set_path": (c_int, [c_void_p, c_int]),
    "prs_index_last_path": (c_int, [c_void_p]),
    "prs_index_set_fused": (c_int, [c_void_p, c_int]),
    "prs_index_last_fused": (c_int, [c_void_p]),
    "prs_index_set_timing": (c_int, [c_void_p, c_int]),
    "prs_index_scan_time": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_i64)]),
    "prs_index_phase_times": (c_int, [c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
    "prs_index_reconstruct_host": (c_int, [c_void_p, c_i64, c_i64, c_void_p]),
    "prs_index_write": (c_int, [c_void_p, c_char_p]),
    "prs_index_read": (c_int, [c_char_p, c_int, c_int, ctypes.POINTER(c_void_p)]),
    "prs_index_write_shard": (c_int, [c_void_p, c_char_p]),
    "prs_index_read_shard": (c_int, [c_char_p, c_int, ctypes.POINTER(c_void_p)]),
    "prs_merge_topk_device": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "prs_merge_topk_f64_device": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "prs_xchg_create": (c_int, [c_int, c_int, c_int, c_i64, c_int, ctypes.POINTER(c_void_p)]),
    "prs_xchg_handle_bytes": (c_int, []),
    "prs_xchg_get_handle": (c_int, [c_void_p, c_void_p]),
    "prs_xchg_open_peers": (c_int, [c_void_p, c_void_p]),
    "prs_xchg_status": (c_int, [c_void_p]),
    "prs_xchg_set_timeout_ms": (c_int, [c_void_p, c_i64]),
    "prs_xchg_free": (None, [c_void_p]),
    "prs_index_search_sharded_device": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_i64, c_int, c_void_p, c_void_p, c_void_p]),
    "prs_xchg_link_local": (c_int, [c_void_p, c_int]),
    "prs_group_create": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_i64, c_int, ctypes.POINTER(c_void_p)]),
    "prs_group_free": (None, [c_void_p]),
    "prs_group_ndev": (c_int, [c_void_p]),
    "prs_group_ntotal": (c_i64, [c_void_p]),
    "prs_group_d": (c_int, [c_void_p]),
    "prs_group_metric": (c_int, [c_void_p]),
    "prs_group_storage": (c_int, [c_void_p]),
    "prs_group_shard_rows": (c_i64, [c_void_p, c_int]),
    "prs_group_reserve": (c_int, [c_void_p, c_i64]),
    "prs_group_add_host": (c_int, [c_void_p, c_void_p, c_i64]),
    "prs_group_add_device": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_void_p]),
    "prs_group_search_host": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p]),
    "prs_group_search_device": (c_int, [c_void_p, c_void_p, c_int, c_i64, c_int, c_void_p, c_void_p, c_void_p]),
    "prs_group_reconstruct_host": (c_int, [c_void_p, c_i64, c_i64, c_void_p]),
    "prs_sparse_build": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_i64, ctypes.c_int32, c_int, ctypes.POINTER(c_void_p)]),
    "prs_sparse_free": (None, [c_void_p]),
    "prs_sparse_ndocs": (c_i64, [c_void_p]),
    "prs_sparse_nnz": (c_i64, [c_void_p]),
    "prs_sparse_search_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p]),
    "prs_sparse_search_device": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p, c_void_p, c_void_p]),
    "prs_sparse_set_mode": (c_int, [c_void_p, c_int]),
    "prs_sparse_mode": (c_int, [c_void_p]),
    "prs_sparse_last_postings": (c_i64, [c_void_p]),
    "prs_sparse_set_id_offset": (c_int, [c_void_p, c_i64]),
    "prs_hybrid_fuse_device": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_i64, c_i64, ctypes.c_double, ctypes.c_double,
                                       c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "prs_centroid_update_device": (c_int, [c_void_p, c_i64, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "prs_pool_norm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
}

_lib = None


class PrsError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libprs error {code}: {message}")
        self.code = code


def build(verbose: bool = False) -> str:
    """Compile libprs.so in-tree (nvcc, sm_100a).  Cross-compiles without a GPU."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4)]
    subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension is the only compute path of this package "
                f"(no CPU fallback). Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                f"or `make -C {os.path.join(_HERE, 'csrc')}`.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError here == header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().prs_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != 0:
        raise PrsError(rc, last_error())

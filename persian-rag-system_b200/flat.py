"""faiss-shaped flat index on top of libprs (C ABI: include/prs.h).

Mirrors exactly the faiss surface the reference touches:
  faiss.IndexFlatL2(d)            src/create_embeddings.py:130, scripts/phase3_pdf_chunking.py:47
  index.add(x)                    src/create_embeddings.py:133
  index.search(x, k) -> (D, I)    src/retrieval.py:102
  index.ntotal / index.d          src/retrieval.py:56, src/create_embeddings.py:286
  faiss.write_index / read_index  src/create_embeddings.py:136, src/retrieval.py:55
plus the knobs the B200 engine adds: `storage` (fp32 exact-parity / fp16 / bf16 throughput),
device-resident tensors in and out (no host hop between the encoder and the search), and a
global-id offset for row shards.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _lib
from ._lib import BF16, F16, F32, METRIC_INNER_PRODUCT, METRIC_L2, PrsError, check

_STORAGE = {"fp32": F32, "float32": F32, "f32": F32, F32: F32,
            "fp16": F16, "float16": F16, "f16": F16, "half": F16, F16: F16,
            "bf16": BF16, "bfloat16": BF16, BF16: BF16}
_STORAGE_NAME = {F32: "fp32", F16: "fp16", BF16: "bf16"}


def _is_tensor(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda")


def _torch_dtype_code(t):
    import torch
    code = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}.get(t.dtype)
    if code is None:
        raise PrsError(_lib.EINVAL, f"unsupported tensor dtype {t.dtype}")
    return code


def _default_device() -> int:
    import sys
    if "torch" in sys.modules:
        torch = sys.modules["torch"]
        try:
            if torch.cuda.is_available():
                return int(torch.cuda.current_device())
        except Exception:
            pass
    return int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("PRS_DEVICE_FROM_LOCAL_RANK") else 0


class FlatIndex:
    """Exact (brute-force) dense index resident in HBM."""

    def __init__(self, d: int, metric: int = METRIC_L2, storage="fp32", device: int | None = None, _handle=None,
                 devices=None, nq_cap: int = 1024, k_cap: int = 128):
        self._h = ctypes.c_void_p()
        self._g = None                      # prs_group handle when the rows are split over several devices
        self._L = _lib.lib()
        if devices is not None and len(devices) > 1:
            # one process, several GPUs (SURVEY 8b `devices=[...]`): contiguous row blocks per device, fused
            # merge + NVLink exchange kernel, results identical to the single index (csrc/group.cu)
            if storage not in _STORAGE:
                raise PrsError(_lib.EINVAL, f"unknown storage {storage!r}")
            devs = (ctypes.c_int * len(devices))(*[int(v) for v in devices])
            self._g = ctypes.c_void_p()
            check(self._L.prs_group_create(int(d), int(metric), _STORAGE[storage], devs, len(devices), int(nq_cap), int(k_cap),
                                           ctypes.byref(self._g)))
            self.devices = [int(v) for v in devices]
            self.device = self.devices[0]
            self.is_trained = True
            return
        if devices is not None and len(devices) == 1:
            device = int(devices[0])
        if _handle is not None:
            self._h = _handle
        else:
            if storage not in _STORAGE:
                raise PrsError(_lib.EINVAL, f"unknown storage {storage!r}")
            dev = _default_device() if device is None else int(device)
            check(self._L.prs_index_create(int(d), int(metric), _STORAGE[storage], dev, ctypes.byref(self._h)))
        self.device = _default_device() if device is None else int(device)
        self.is_trained = True

    # ---- faiss attributes ----
    @property
    def d(self) -> int:
        return int(self._L.prs_group_d(self._g) if self._g else self._L.prs_index_d(self._h))

    @property
    def ntotal(self) -> int:
        return int(self._L.prs_group_ntotal(self._g) if self._g else self._L.prs_index_ntotal(self._h))

    @property
    def metric_type(self) -> int:
        return int(self._L.prs_group_metric(self._g) if self._g else self._L.prs_index_metric(self._h))

    @property
    def storage(self) -> str:
        return _STORAGE_NAME[int(self._L.prs_group_storage(self._g) if self._g else self._L.prs_index_storage(self._h))]

    @property
    def shard_rows(self):
        """Rows held by each device (one entry for a single-device index)."""
        if not self._g:
            return [self.ntotal]
        return [int(self._L.prs_group_shard_rows(self._g, i)) for i in range(len(self.devices))]

    def _single_only(self, what: str) -> None:
        if self._g:
            raise PrsError(_lib.EUNSUP, f"{what} is not available on a multi-device index")

    @property
    def last_path(self) -> str:
        if self._g:
            return "group"
        return {0: "none", 1: "cuda-core", 2: "tcgen05"}[int(self._L.prs_index_last_path(self._h))]

    def set_path(self, path) -> None:
        """'auto' | 'cuda-core' | 'tcgen05' (tests and benchmarks)."""
        self._single_only("set_path")
        code = {"auto": 0, "cuda-core": 1, "tcgen05": 2, 0: 0, 1: 1, 2: 2}[path]
        check(self._L.prs_index_set_path(self._h, code))

    def set_fused(self, enable: bool) -> None:
        """One-launch search (prep + scan + merge in one cooperative kernel, default on) vs the three-kernel sequence.
        `enable=2` also fuses row-sharded searches (push in the scan kernel's tail + a small pull kernel); `enable=3`:
        two launches (the scan prepares its own queries, a separate merge kernel follows -- the row-sharded default)."""
        self._single_only("set_fused")
        check(self._L.prs_index_set_fused(self._h, int(enable)))

    @property
    def last_fused(self) -> bool:
        return (not self._g) and int(self._L.prs_index_last_fused(self._h)) == 1

    def set_timing(self, enable: bool) -> None:
        """Bracket every scan-kernel launch with CUDA events on its stream (bench instrumentation)."""
        self._single_only("set_timing")
        check(self._L.prs_index_set_timing(self._h, 1 if enable else 0))

    def scan_time(self):
        """(summed scan-kernel device time in ms, launches) since the previous call; synchronises."""
        ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
        check(self._L.prs_index_scan_time(self._h, ctypes.byref(ms), ctypes.byref(n)))
        return float(ms.value), int(n.value)

    def phase_times(self):
        """(prep kernel ms, merge kernel ms) summed since the previous call (timing must be enabled)."""
        a, b = ctypes.c_double(0.0), ctypes.c_double(0.0)
        check(self._L.prs_index_phase_times(self._h, ctypes.byref(a), ctypes.byref(b)))
        return float(a.value), float(b.value)

    def set_id_offset(self, offset: int) -> None:
        self._single_only("set_id_offset")
        check(self._L.prs_index_set_id_offset(self._h, int(offset)))

    def reserve(self, n_total: int) -> None:
        if self._g:
            return check(self._L.prs_group_reserve(self._g, int(n_total)))
        check(self._L.prs_index_reserve(self._h, int(n_total)))

    # ---- faiss methods ----
    def add(self, x) -> None:
        if _is_tensor(x):
            import torch
            if not x.is_cuda:
                return self.add(x.detach().float().numpy())
            if x.dim() != 2 or x.shape[1] != self.d:
                raise PrsError(_lib.EINVAL, f"add: expected [n, {self.d}], got {tuple(x.shape)}")
            x = x.contiguous()
            st = torch.cuda.current_stream(x.device).cuda_stream
            fn, h = (self._L.prs_group_add_device, self._g) if self._g else (self._L.prs_index_add_device, self._h)
            check(fn(h, ctypes.c_void_p(x.data_ptr()), _torch_dtype_code(x), int(x.shape[0]), ctypes.c_void_p(st)))
            return
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise PrsError(_lib.EINVAL, f"add: expected [n, {self.d}], got {x.shape}")
        fn, h = (self._L.prs_group_add_host, self._g) if self._g else (self._L.prs_index_add_host, self._h)
        check(fn(h, x.ctypes.data_as(ctypes.c_void_p), int(x.shape[0])))

    def search(self, x, k: int):
        """numpy in -> (D float32 [nq,k], I int64 [nq,k]) numpy out (faiss contract);
        CUDA torch tensor in -> torch tensors out, asynchronous on the current stream."""
        k = int(k)
        if _is_tensor(x) and x.is_cuda:
            import torch
            if x.dim() != 2 or x.shape[1] != self.d:
                raise PrsError(_lib.EINVAL, f"search: expected [nq, {self.d}], got {tuple(x.shape)}")
            x = x.contiguous()
            nq = int(x.shape[0])
            D = torch.empty((nq, k), dtype=torch.float32, device=x.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=x.device)
            st = torch.cuda.current_stream(x.device).cuda_stream
            if self._g:
                if int(x.device.index or 0) != self.device:
                    raise PrsError(_lib.EINVAL, f"search: queries must live on the group's first device (cuda:{self.device})")
                check(self._L.prs_group_search_device(self._g, ctypes.c_void_p(x.data_ptr()), _torch_dtype_code(x), nq, k,
                                                      ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()), ctypes.c_void_p(st)))
                return D, I
            check(self._L.prs_index_search_device(self._h, ctypes.c_void_p(x.data_ptr()), _torch_dtype_code(x), nq, k,
                                                  ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()),
                                                  ctypes.c_void_p(st)))
            return D, I
        if _is_tensor(x):
            x = x.detach().float().numpy()
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.d:
            raise PrsError(_lib.EINVAL, f"search: expected [nq, {self.d}], got {x.shape}")
        nq = int(x.shape[0])
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        fn, h = (self._L.prs_group_search_host, self._g) if self._g else (self._L.prs_index_search_host, self._h)
        # (ndarray.ctypes.data_as costs 2.4 us per call; three of them are a tenth of a small-index search)
        check(fn(h, ctypes.c_void_p(x.ctypes.data), nq, k, ctypes.c_void_p(D.ctypes.data), ctypes.c_void_p(I.ctypes.data)))
        return D, I

    def search_to_host(self, x, k: int):
        """CUDA tensor in -> numpy out: the kernels store the [nq, k] results straight into page-locked host memory
        (no device result tensors, no device->host copies); synchronises the current stream."""
        import torch
        k = int(k)
        if self._g or not (_is_tensor(x) and x.is_cuda):
            D, I = self.search(x, k)
            return (D.cpu().numpy(), I.cpu().numpy()) if _is_tensor(D) else (D, I)
        if x.dim() != 2 or x.shape[1] != self.d:
            raise PrsError(_lib.EINVAL, f"search: expected [nq, {self.d}], got {tuple(x.shape)}")
        x = x.contiguous()
        nq = int(x.shape[0])
        cache = self.__dict__.setdefault("_pinned_out", {})
        buf = cache.get((nq, k))
        if buf is None:
            if len(cache) >= 8:
                cache.clear()
            buf = cache[(nq, k)] = (torch.empty((nq, k), dtype=torch.float32).pin_memory(), torch.empty((nq, k), dtype=torch.int64).pin_memory())
        D, I = buf
        stream = torch.cuda.current_stream(x.device)
        check(self._L.prs_index_search_device(self._h, ctypes.c_void_p(x.data_ptr()), _torch_dtype_code(x), nq, k,
                                              ctypes.c_void_p(D.data_ptr()), ctypes.c_void_p(I.data_ptr()), ctypes.c_void_p(stream.cuda_stream)))
        stream.synchronize()
        return D.numpy().copy(), I.numpy().copy()

    def search_into(self, q_ptr: int, nq: int, k: int, D_ptr: int, I_ptr: int) -> None:
        """Raw host-pointer search (pinned buffers in benchmarks): no allocation per call."""
        self._single_only("search_into")
        check(self._L.prs_index_search_host(self._h, ctypes.c_void_p(q_ptr), int(nq), int(k),
                                            ctypes.c_void_p(D_ptr), ctypes.c_void_p(I_ptr)))

    def reconstruct_n(self, i0: int = 0, n: int | None = None) -> np.ndarray:
        n = self.ntotal - i0 if n is None else n
        out = np.empty((n, self.d), dtype=np.float32)
        fn, h = (self._L.prs_group_reconstruct_host, self._g) if self._g else (self._L.prs_index_reconstruct_host, self._h)
        check(fn(h, int(i0), int(n), out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def __del__(self):
        try:
            if getattr(self, "_g", None) is not None and self._g.value:
                self._L.prs_group_free(self._g)
                self._g = None
            if getattr(self, "_h", None) is not None and self._h.value:
                self._L.prs_index_free(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


def IndexFlatL2(d: int, storage="fp32", device: int | None = None, devices=None) -> FlatIndex:
    """faiss.IndexFlatL2(d) -- src/create_embeddings.py:130"""
    return FlatIndex(d, METRIC_L2, storage, device, devices=devices)


def IndexFlatIP(d: int, storage="fp32", device: int | None = None, devices=None) -> FlatIndex:
    """faiss.IndexFlatIP(d)"""
    return FlatIndex(d, METRIC_INNER_PRODUCT, storage, device, devices=devices)


def write_index(index: FlatIndex, path: str) -> None:
    """faiss.write_index(index, path) -- src/create_embeddings.py:136.  fp32 storage writes faiss's
    byte-exact IndexFlat file."""
    if index._g:
        # gather the row blocks of all devices into a single-device index of the same storage, then write that
        tmp = FlatIndex(index.d, index.metric_type, index.storage, index.device)
        tmp.reserve(index.ntotal)
        step = max(1, (64 << 20) // (4 * index.d))
        for i0 in range(0, index.ntotal, step):
            tmp.add(index.reconstruct_n(i0, min(step, index.ntotal - i0)))
        index = tmp
    check(index._L.prs_index_write(index._h, os.fsencode(path)))


def read_index(path: str, storage="fp32", device: int | None = None, devices=None) -> FlatIndex:
    """faiss.read_index(path) -- src/retrieval.py:55.  `devices=[...]`: the rows are split over those GPUs."""
    if storage not in _STORAGE:
        raise PrsError(_lib.EINVAL, f"unknown storage {storage!r}")
    L = _lib.lib()
    h = ctypes.c_void_p()
    dev = (_default_device() if device is None else int(device)) if not devices else int(devices[0])
    check(L.prs_index_read(os.fsencode(path), _STORAGE[storage], dev, ctypes.byref(h)))
    one = FlatIndex(0, _handle=h, device=dev)
    if not devices or len(devices) < 2:
        return one
    grp = FlatIndex(one.d, one.metric_type, storage, devices=devices)
    grp.reserve(one.ntotal)
    step = max(1, (64 << 20) // (4 * one.d))
    for i0 in range(0, one.ntotal, step):
        grp.add(one.reconstruct_n(i0, min(step, one.ntotal - i0)))      # values are exactly representable in `storage`
    return grp

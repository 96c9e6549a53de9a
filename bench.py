#!/usr/bin/env python
"""bench.py -- QPS @ k=10 of the exact flat search (the hot path of src/retrieval.py:102) on B200.

Workload (BASELINE.json configs[1]): multilingual-e5-base shape, d = 768, 1 M synthetic unit-norm
chunks, cosine == inner product, k = 10, fp16 storage / fp32 accumulate.  One STEP = one batch of
`--batch` queries (default 64, the top of the north-star's bandwidth-bound range) searched against
the whole corpus: scan + fused top-k + merge.  The corpus (1.536 GB) is 12x the 126 MB L2, so every
step streams it from HBM ("inputs larger than L2"); when a shard is smaller than 2x L2 (N = 8)
the L2 is flushed between steps and steps are timed one by one.

  python bench.py [--gpus N --steps K --warmup W]          this engine (libprs.so), one rank per GPU
  python bench.py --impl reference [...]                   the CPU implementation on the host cores

N > 1 (torchrun): STRONG scaling -- the same 1 M rows are row-sharded over the ranks, every rank
scans its block, the [B, k] lists are all-gathered over NVLink and merged on the device
(SURVEY.md 8e).  value = queries answered per second by the whole job.

JSON keys beyond the base contract: `roofline` (scan kernel, HBM bound, timed with CUDA events on
the launching stream inside the timed region), `cpu_baseline` (oracle port timed on the host cores,
rank 0, N = 1), `e2e` (same metric through the C-ABI host entry point, pinned host buffers, copies
inside the timed region), `sweep` (batch 1..1024, device-resident, outside the K timed steps).
"""
from __future__ import annotations

import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC_NAME = "QPS @k=10, 1M x 768 corpus (exact flat search)"
L2_BYTES = 126 * 1024 * 1024


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--d", type=int, default=768)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--storage", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--metric", default="ip", choices=["ip", "l2"])
    ap.add_argument("--path", default="auto", choices=["auto", "cuda-core", "tcgen05"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N>1: fused peer-memory exchange, or ncclAllGather + merge")
    ap.add_argument("--inflight", type=int, default=2,
                    help="searches in flight in the extra `pipelined` measurement (alternating CUDA streams / index workspaces); 1 = skip it")
    ap.add_argument("--capacity-rows", type=int, default=50_000_000,
                    help="rows per GPU of the secondary weak-scaling measurement (configs[4] share: 400M x 384 over 8 GPUs); 0 = skip")
    ap.add_argument("--extras", default="c0,c2,c3,c4",
                    help="secondary north-star configs measured after the headline (extra JSON keys): c0 = the reference's own call shape "
                         "(125 x 384 fp32, nq=1, k=5, latency), c2 = 50M x 512 fp16 k=100, c3 = BM25 on 10M docs x 4096 queries, "
                         "c4 = batch x k sweep on the 50M x 384 per-GPU share; 'none' skips them")
    ap.add_argument("--no-fuse", action="store_true", help="three-kernel search (prep, scan, merge) instead of the one-launch search (A/B)")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--secondary-budget-s", type=float, default=540.0,
                    help="seconds the measurements AFTER the headline (sweep, CPU baseline, extra configs) may take before the line is printed without them")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep-out", default="")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed regions, read through NVML (the library behind
    nvidia-smi; nvidia_ml_py) from a thread every ~2 ms -- the timed region of a default run is only
    tens of milliseconds long, too short for `nvidia-smi -lms`."""

    def __init__(self, cuda_index: int):
        self.rows, self.on, self.stop_flag, self.h, self.mx = [], False, False, None, None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(cuda_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if bytes is not str else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:                      # no NVML: report that, do not invent numbers
            self.err = repr(e)
            self.h = None

    def start(self):
        if self.h is None:
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            if self.on:
                try:
                    mhz = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                    try:
                        rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    self.rows.append((float(mhz), int(rs)))
                except Exception:
                    pass
            time.sleep(0.002)

    def stop(self):
        self.stop_flag = True

    def summary(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "error": getattr(self, "err", "no NVML")}
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r in self.rows))
        sm = [m for m, _ in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(sm), "source": "NVML, sampled every ~2 ms during the timed regions"}


def host_threads():
    """Cores this process may really use: the affinity mask, capped by the cgroup CPU quota (a container can see every
    core of the box in its mask while its quota is a fraction of them; BLAS threads beyond the quota only spin and
    get throttled -- a 12 s CPU leg then takes minutes)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        quota = None
        if os.path.exists("/sys/fs/cgroup/cpu.max"):                      # cgroup v2: "<quota|max> <period>"
            q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
            if q != "max":
                quota = float(q) / float(per)
        elif os.path.exists("/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):       # cgroup v1
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0 and per > 0:
                quota = q / per
        if quota is not None:
            n = max(1, min(n, int(quota + 0.999)))
    except Exception:
        pass
    return n


def use_all_host_threads():
    """The CPU legs use every host core this process may run on, whatever OMP_NUM_THREADS says
    (torchrun exports OMP_NUM_THREADS=1 to its workers).  Returns the BLAS thread count in effect."""
    n = host_threads()
    try:
        import torch
        torch.set_num_threads(n)
    except Exception:
        pass
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=n)
        got = [int(i.get("num_threads", 1)) for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(got) if got else n
    except Exception:
        return n


# ------------------------------------------------------------------------------------------------
# CPU leg: the oracle port (faiss IndexFlat restated: sgemm blocks + threshold/heap), all host cores
# ------------------------------------------------------------------------------------------------
def cpu_flat_search():
    """(callable(x32, q, k, metric) -> (D, I), kind, note): real faiss-cpu when it is importable (site-packages or
    baseline/_ref), else the numpy/OpenBLAS restatement of faiss IndexFlat.search (the oracle "port")."""
    from oracle import oracle as O
    faiss = O.reference_library("faiss")
    if faiss is not None:
        cache = {}

        def run(x32, q, k, metric):
            key = (id(x32), metric)
            if key not in cache:                             # index.add is not part of a search step
                cache.clear()
                idx = faiss.IndexFlatL2(x32.shape[1]) if metric == O.METRIC_L2 else faiss.IndexFlatIP(x32.shape[1])
                idx.add(np.ascontiguousarray(x32, dtype=np.float32))
                cache[key] = idx
            return cache[key].search(np.ascontiguousarray(q, dtype=np.float32), k)
        return run, "reference", f"faiss {getattr(faiss, '__version__', '?')} IndexFlat.search (the reference's own library)"
    return (lambda x32, q, k, metric: O.flat_search_np_threshold(x32, q, k, metric)), "port", \
        "faiss-cpu/rank_bm25 wheels absent; numpy+OpenBLAS restatement of faiss IndexFlat.search"


def cpu_search_timed(x32, q, k, metric_code, budget_s, min_reps=1):
    """Times the CPU search of `q` over all of `x32` for about `budget_s` seconds.  Returns (times, result, kind, note,
    rows_used): if a probe over 1/8 of the rows says that ONE full search would already exceed the budget (a box whose
    host cores are oversubscribed or throttled), only a prefix of the rows is searched -- rows_used < len(x32), the
    caller scales the rate and skips the result comparison -- so that this leg stays bounded."""
    search, kind, note = cpu_flat_search()
    probe_rows = min(len(x32), max(131072, len(x32) // 8))
    search(x32[: min(len(x32), 131072)], q, k, metric_code)       # warm BLAS threads
    t0 = time.perf_counter()
    search(x32[:probe_rows], q, k, metric_code)
    est_full = (time.perf_counter() - t0) * len(x32) / probe_rows
    rows_used = len(x32)
    if est_full > budget_s:
        rows_used = max(probe_rows, int(len(x32) * budget_s / est_full))
    xs = x32 if rows_used == len(x32) else x32[:rows_used]
    times, res = [], None
    t_all = time.perf_counter()
    while len(times) < min_reps or (time.perf_counter() - t_all) < budget_s:
        t0 = time.perf_counter()
        res = search(xs, q, k, metric_code)
        times.append(time.perf_counter() - t0)
        if len(times) >= 50:
            break
    return times, res, kind, note, rows_used


def run_reference(a):
    """--impl reference: faiss-cpu is not installable here (no wheel, no network), so the reference
    arm is the oracle port of faiss IndexFlat.search on the host cores, same config and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import oracle as O
    cores = use_all_host_threads()
    metric_code = O.METRIC_IP if a.metric == "ip" else O.METRIC_L2
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(a.rows, a.d, generator=g)
    x /= x.norm(dim=1, keepdim=True)
    if a.storage != "fp32":
        x = x.to(torch.float16 if a.storage == "fp16" else torch.bfloat16).float()
    x32 = x.numpy()
    gq = torch.Generator().manual_seed(4321)
    Q = torch.randn(a.warmup + a.steps, a.batch, a.d, generator=gq)
    Q /= Q.norm(dim=2, keepdim=True)
    Q = Q.numpy()
    search, kind, note = cpu_flat_search()
    for w in range(a.warmup):
        search(x32, Q[w], a.k, metric_code)
    t0 = time.perf_counter()
    for s in range(a.steps):
        search(x32, Q[a.warmup + s], a.k, metric_code)
    dt = time.perf_counter() - t0
    qps = a.steps * a.batch / dt
    sample = f"full workload per step: {a.batch} queries x {a.rows} x {a.d} fp32 rows, k={a.k}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a, a.gpus),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": kind, "sample": sample, "note": note},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(a, world):
    return {"workload": f"configs[1]: multilingual-e5-base shape d={a.d}, {a.rows} synthetic unit-norm chunks, "
                        f"exact {'cosine/IP' if a.metric == 'ip' else 'L2'} search, k={a.k}, 1xB200",
            "rows": a.rows, "d": a.d, "k": a.k, "batch": a.batch, "storage": a.storage, "metric": a.metric,
            "sharding": f"rows/{world}, exchange={a.exchange}" if world > 1 else "none",
            "cache": "inputs larger than L2 (corpus shard vs 126 MB L2)"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    import persian_rag_system_b200 as P
    from persian_rag_system_b200.sharded import ShardedFlatIndex, shard_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: libprs has no CPU path"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = P.lib()
    metric_code = P.METRIC_INNER_PRODUCT if a.metric == "ip" else P.METRIC_L2
    tdt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[a.storage]
    es = 4 if a.storage == "fp32" else 2

    # ---- corpus: on-device Philox, unit-norm rows in fp32, cast to storage ----
    lo, hi = shard_bounds(a.rows, world, rank)
    n_local = hi - lo
    sh = ShardedFlatIndex(a.d, metric_code, a.storage, device=local, exchange=a.exchange, nq_cap=max(a.batch, 64), k_cap=max(a.k, 16))
    idx = sh.local
    idx.reserve(n_local)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    slab = max(1, (256 << 20) // (a.d * 4))
    added = 0
    while added < n_local:
        c = min(slab, n_local - added)
        xb = torch.randn(c, a.d, generator=gen, device=dev)
        xb /= xb.norm(dim=1, keepdim=True)
        idx.add(xb.to(tdt))
        added += c
    del xb
    sh.offset, sh.ntotal_global = lo, a.rows
    idx.set_id_offset(lo)
    idx.set_path(a.path)
    if a.no_fuse:
        idx.set_fused(False)
    torch.cuda.synchronize()

    gq = torch.Generator(device=dev).manual_seed(4321)              # same queries on every rank
    nbatches = a.warmup + a.steps
    Qd = torch.randn(nbatches, a.batch, a.d, generator=gq, device=dev)
    Qd /= Qd.norm(dim=2, keepdim=True)
    Qh = torch.empty((nbatches, a.batch, a.d), dtype=torch.float32).pin_memory()
    Qh.copy_(Qd)
    Dh = torch.empty((a.batch, a.k), dtype=torch.float32).pin_memory()
    Ih = torch.empty((a.batch, a.k), dtype=torch.int64).pin_memory()

    pitch = (a.d + 63) // 64 * 64
    shard_bytes = n_local * pitch * es
    # a shard streamed cyclically through the 126 MB L2 keeps nothing between steps once it is > 2x L2;
    # smaller shards (N = 8: 192 MB) get an explicit flush between steps: WRITE a 256 MB buffer, then
    # read a second one so that the dirty lines are written back before the timed step, not during it
    flush = shard_bytes < 2 * L2_BYTES
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if flush else None
    flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device=dev) if flush else None

    def l2_flush(i):
        flush_buf.fill_(i & 0xFF)
        flush_rd.sum()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The headline `value` is strictly serial (one search after the other on one stream), so that the
    # scan kernel's CUDA-event time inside the timed region is its real duration.  The extra `pipelined`
    # measurement below keeps `--inflight` searches in flight on alternating streams: the merge /
    # exchange kernel of batch i overlaps the scan of batch i+1 (one index workspace per search).
    n_inflight = 1 if (a.inflight < 2 or flush) else 2
    pipe_streams = [torch.cuda.Stream(device=dev) for _ in range(n_inflight)] if n_inflight > 1 else []
    streams = []

    def step_device(i):
        if not streams:
            return sh.search(Qd[i], a.k)
        with torch.cuda.stream(streams[i % len(streams)]):
            return sh.search(Qd[i], a.k)

    def fork_streams():
        ev = torch.cuda.Event()
        ev.record()
        for s_ in streams:
            s_.wait_event(ev)

    def join_streams():
        for s_ in streams:
            ev = torch.cuda.Event()
            ev.record(s_)
            torch.cuda.current_stream().wait_event(ev)

    # ---- device-resident throughput: `value` ----
    fork_streams()
    for w in range(a.warmup):
        step_device(w)
    join_streams()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    idx.set_timing(True)
    idx.scan_time()
    launches0 = L.prs_launch_count()
    sampler.on = True
    barrier()
    if not flush:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fork_streams()
        for s in range(a.steps):
            step_device(a.warmup + s)
        join_streams()
        e1.record()
        barrier()
        dev_ms = e0.elapsed_time(e1)
    else:
        evs = []
        tok = torch.zeros(1, device=dev)
        for s in range(a.steps):
            l2_flush(s)
            # every rank starts the timed step together: a one-element NCCL all-reduce ENQUEUED on the stream (no
            # host synchronisation, so the launches of the step stay queued behind it) releases all GPUs within
            # microseconds of each other; without it the ranks drift out of phase through the flush kernels and the
            # fused exchange would measure that skew
            if world > 1:
                dist.all_reduce(tok)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_device(a.warmup + s)
            e1.record()
            evs.append((e0, e1))
        barrier()
        dev_ms = sum(x.elapsed_time(y) for x, y in evs)
    sampler.on = False
    launches = L.prs_launch_count() - launches0
    scan_ms, scan_launches = idx.scan_time()
    prep_ms, merge_ms = idx.phase_times()
    idx.set_timing(False)
    last_path = idx.last_path

    # ---- the same K steps with `--inflight` searches in flight (throughput of a serving loop) ----
    pipelined = None
    if pipe_streams:
        streams = pipe_streams
        fork_streams()
        for w in range(a.warmup):
            step_device(w)
        join_streams()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fork_streams()
        for s in range(a.steps):
            step_device(a.warmup + s)
        join_streams()
        e1.record()
        barrier()
        pipe_ms = e0.elapsed_time(e1)
        streams = []
        tp = torch.tensor([pipe_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        pipe_ms = float(tp.item())
        pipelined = {"searches_in_flight": len(pipe_streams), "value": a.steps * a.batch / (pipe_ms * 1e-3), "unit": "queries/s",
                     "ms_per_step": pipe_ms / a.steps,
                     "note": "same steps, two CUDA streams round-robin; kernel event times overlap here, so the roofline is taken from the serial region"}

    # ---- end to end through the host entry point: pinned host queries in, host (D, I) out ----
    def step_e2e(i):
        if world == 1:
            idx.search_into(Qh[i].data_ptr(), a.batch, a.k, Dh.data_ptr(), Ih.data_ptr())
        elif a.exchange == "p2p":
            # pinned host queries in, pinned host results out: the kernels address them directly (UVA)
            st = torch.cuda.current_stream().cuda_stream
            sh.search_into(Qh[i].data_ptr(), a.batch, a.k, Dh.data_ptr(), Ih.data_ptr(), st)
            torch.cuda.synchronize()
        else:
            qd = Qh[i].to(dev, non_blocking=True)
            D, I = sh.search(qd, a.k)
            Dh.copy_(D, non_blocking=True)
            Ih.copy_(I, non_blocking=True)
            torch.cuda.synchronize()

    for w in range(a.warmup):
        step_e2e(w)
    barrier()
    sampler.on = True
    if not flush:
        t0 = time.perf_counter()
        for s in range(a.steps):
            step_e2e(a.warmup + s)
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
    else:                                   # steps timed one by one, the L2 flush between them is not part of a step
        e2e_ms = 0.0
        for s in range(a.steps):
            l2_flush(s)
            torch.cuda.synchronize()
            barrier()                       # every rank starts the step together, like the device-timed steps do
            t0 = time.perf_counter()
            step_e2e(a.warmup + s)          # synchronous: returns with the results on the host
            e2e_ms += (time.perf_counter() - t0) * 1e3
        barrier()
    sampler.on = False
    if rank == 0:
        sampler.stop()
    # the reference's literal call -- index.search(numpy array) with ordinary (pageable) host memory, fresh output arrays
    # per call (src/retrieval.py:102) -- next to the pinned-buffer number above (N = 1 only; not the headline)
    e2e_pageable = None
    if world == 1:
        qn = [Qh[a.warmup + s].numpy().copy() for s in range(a.steps)]
        for w in range(3):
            idx.search(qn[w % len(qn)], a.k)
        t0 = time.perf_counter()
        for s in range(a.steps):
            idx.search(qn[s], a.k)
        tp = (time.perf_counter() - t0) / a.steps
        e2e_pageable = {"value": a.batch / tp, "unit": "queries/s", "ms_per_step": tp * 1e3,
                        "api": "FlatIndex.search(numpy float32 [B, d]) -> numpy (D, I): pageable host memory (one host memcpy into the index's page-locked buffer, zero-copy from there), outputs allocated per call"}
    last_I = Ih.numpy().copy()
    last_D = Dh.numpy().copy()
    # the host-buffer result of the last step must equal the device-resident search of the same batch
    Dd, Id = sh.search(Qd[nbatches - 1], a.k)
    torch.cuda.synchronize()
    e2e_same = bool(np.array_equal(Id.cpu().numpy(), last_I) and np.array_equal(Dd.cpu().numpy(), last_D))

    # max over ranks
    t = torch.tensor([dev_ms, e2e_ms, scan_ms / max(scan_launches, 1)], dtype=torch.float64, device=dev)
    cnt = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, scan_avg_ms = (float(v) for v in t.tolist())
    launches = int(cnt.item())

    qps = a.steps * a.batch / (dev_ms * 1e-3)
    qps_e2e = a.steps * a.batch / (e2e_ms * 1e-3)
    peak, peak_src = peaks()
    alg_bytes = n_local * a.d * es + (4 * n_local if a.metric == "l2" else 0)
    achieved = alg_bytes / (scan_avg_ms * 1e-3) / 1e9 if scan_avg_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")          # per-launch dram bytes from the committed ncu capture
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"{a.storage}_{a.rows // world}x{a.d}_b{a.batch}_{last_path}")
        except Exception:
            traffic = None

    out = {
        "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": {"fp16": "f16 storage, f32 accumulate", "bf16": "bf16 storage, f32 accumulate", "fp32": "f32"}[a.storage],
        "data": "synthetic", "config": workload_config(a, world), "kernel_path": last_path,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "flat_scan_umma_kernel" if last_path == "tcgen05" else "flat_scan_simt_kernel",
                     "bytes_per_launch": alg_bytes, "avg_launch_ms": scan_avg_ms, "launches_timed": scan_launches,
                     "frac_of_nominal_8TBs": achieved / 8000.0,
                     "share_of_step": scan_avg_ms * (scan_launches / max(a.steps, 1)) / (dev_ms / a.steps)},
        "e2e": {"value": qps_e2e, "unit": "queries/s", "h2d_bytes_per_step": a.batch * a.d * 4,
                "d2h_bytes_per_step": a.batch * a.k * 12, "ms_per_step": e2e_ms / a.steps, "equals_device_resident_result": e2e_same,
                "api": "prs_index_search_host (pinned host q, D, I)" if world == 1 else
                       ("ShardedFlatIndex.search_into (pinned host q, D, I addressed by the kernels)" if a.exchange == "p2p" else "ShardedFlatIndex.search with pinned H2D/D2H")},
        "step_breakdown_ms": {"prep_kernel": prep_ms / a.steps, "scan_kernel": scan_ms / a.steps, "merge_kernel": merge_ms / a.steps,
                              "rest (launch gaps, event records)": dev_ms / a.steps - (prep_ms + scan_ms + merge_ms) / a.steps},
        "gpu_launches": launches,
        "l2_flush_between_steps": bool(flush),
        "one_launch_search": bool(idx.last_fused),
    }
    if e2e_pageable:
        out["e2e_pageable_numpy"] = e2e_pageable

    if pipelined:
        out["pipelined"] = pipelined
    if rank == 0:
        out["clocks"] = sampler.summary()
    # From here on only SECONDARY measurements follow (sweep, CPU baseline, extra configs).  If one of them ever stalls,
    # the headline line must still come out: a watchdog prints what has been measured so far and ends the process.
    watchdog = _arm_watchdog(out, rank, a.secondary_budget_s)

    # ---- batch sweep (device resident, 1 GPU): the B = 1..1024 picture of configs[1] ----
    if world == 1 and not a.no_sweep:
        sweep = []
        sw_clk = ClockSampler(local)            # tensor-bound batches draw the most power: record clocks / cap reasons per batch size
        sw_clk.start()
        for B in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024):
            q = torch.randn(B, a.d, generator=gq, device=dev)
            q /= q.norm(dim=1, keepdim=True)
            for _ in range(3):
                idx.search(q, a.k)
            iters = 20 if B <= 256 else 8
            torch.cuda.synchronize()
            idx.set_timing(True); idx.scan_time()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            r0 = len(sw_clk.rows)
            sw_clk.on = True
            e0.record()
            for _ in range(iters):
                idx.search(q, a.k)
            e1.record()
            torch.cuda.synchronize()
            sw_clk.on = False
            ms = e0.elapsed_time(e1) / iters
            sms, sn = idx.scan_time()
            idx.set_timing(False)
            gbs = alg_bytes / (ms * 1e-3) / 1e9
            row = {"batch": B, "ms": round(ms, 4), "qps": round(B / (ms * 1e-3), 1), "path": idx.last_path,
                   "corpus_gbs": round(gbs, 1), "frac_hbm": round(gbs / peak, 4),
                   "scan_ms": round(sms / iters, 4), "tflops": round(2.0 * n_local * a.d * B / (ms * 1e-3) / 1e12, 2)}
            rows = sw_clk.rows[r0:]
            if rows:
                row["sm_mhz"] = float(np.median([m for m, _ in rows]))
                cap = getattr(sw_clk.nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)
                row["sw_power_cap"] = bool(any(r & cap for _, r in rows))
            sweep.append(row)
        sw_clk.stop()
        out["sweep"] = sweep
        if a.sweep_out:
            json.dump(sweep, open(a.sweep_out, "w"), indent=1)

    # ---- CPU baseline beside it (rank 0, N = 1): oracle port on the same corpus and queries ----
    if world == 1 and rank == 0 and not a.no_cpu_baseline:
        from oracle import oracle as O
        blas_threads = use_all_host_threads()
        x32 = idx.reconstruct_n(0, n_local)                       # the stored (rounded) rows, as fp32
        qlast = Qh[nbatches - 1].numpy()
        times, (Dc, Ic), cpu_kind, cpu_note, cpu_rows = cpu_search_timed(x32, qlast, a.k, O.METRIC_IP if a.metric == "ip" else O.METRIC_L2, budget_s=12.0)
        cpu_scale = cpu_rows / float(n_local)                     # < 1 only when one full search would not fit the time budget
        cpu_qps = a.batch / float(np.median(times)) * cpu_scale
        out["cpu_baseline"] = {"value": cpu_qps, "unit": "queries/s", "cores": blas_threads, "kind": cpu_kind,
                               "sample": (f"{len(times)} x (1 batch of {a.batch} queries over all {n_local} rows, fp32), median" if cpu_rows == n_local else
                                          f"{len(times)} x (1 batch of {a.batch} queries over the first {cpu_rows} of {n_local} rows, fp32), median, rate scaled by {cpu_scale:.4f}"),
                               "ms_per_batch": float(np.median(times)) * 1e3 / cpu_scale, "note": cpu_note}
        # parity of the timed GPU result (last e2e step) against the CPU port: ids identical except ties within 1e-3
        try:
            if cpu_rows != n_local:
                raise AssertionError("CPU leg searched a prefix of the corpus only (time budget): no result comparison in this run")
            qh = torch.from_numpy(qlast).to(tdt).float().numpy() if last_path == "tcgen05" else qlast
            if last_path == "tcgen05":
                Dc, Ic = O.flat_search_np_threshold(x32, qh, a.k, O.METRIC_IP if a.metric == "ip" else O.METRIC_L2)
            flips = O.check_topk_lists(last_I, last_D, Ic, Dc, rtol=1e-3, atol=2e-5, what="bench parity")
            out["parity"] = {"queries": a.batch, "k": a.k, "id_positions_differing_within_tie_tolerance": int(flips),
                             "identical_lists": bool(np.array_equal(last_I, Ic))}
        except AssertionError as e:
            out["parity"] = {"error": str(e)[:300]}

    # ---- secondary: capacity (weak) scaling -- every GPU holds the per-GPU share of configs[4]
    # (400M x 384 fp16 over 8 GPUs = 50M rows = 38.4 GB per GPU); the global corpus grows with N and the
    # time per batch should stay flat.  Not the headline (the headline keeps the 1M x 768 corpus).
    if world > 1:
        sh.check_exchange()
    extras = set() if a.extras == "none" else set(a.extras.split(","))
    del sh, idx, Qd
    torch.cuda.empty_cache()
    if "c0" in extras and world == 1:
        try:
            out["configs0_reference_call_shape"] = measure_configs0(local, dev)
        except Exception as e:                                    # never lose the headline line
            out["configs0_reference_call_shape"] = {"error": repr(e)[:300]}
    if a.capacity_rows > 0:
        try:
            out["capacity_scaling"] = measure_capacity(a, world, rank, local, dev, peak, sweep="c4" in extras)
        except Exception as e:
            out["capacity_scaling"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    if "c2" in extras:
        try:
            out["configs2_50M_x_512_k100"] = measure_configs2(a, world, rank, local, dev, peak)
        except Exception as e:
            out["configs2_50M_x_512_k100"] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    if "c3" in extras and world == 1:
        try:
            out["configs3_bm25_10M_docs"] = measure_configs3(dev, peak)
        except Exception as e:
            out["configs3_bm25_10M_docs"] = {"error": repr(e)[:300]}
    watchdog.cancel()
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()                       # nobody leaves while a peer still uses its exchange buffers


def _timed_search(sh, idx, q, k, iters, world):
    """(ms per batch, scan-kernel ms per batch) of `iters` back-to-back searches, max over ranks."""
    import torch
    import torch.distributed as dist
    for _ in range(2):                      # both search workspaces of the index (used round-robin) get their buffers
        sh.search(q, k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    idx.set_timing(True); idx.scan_time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        sh.search(q, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    scan_ms, _n = idx.scan_time()
    idx.set_timing(False)
    t = torch.tensor([ms, scan_ms / iters], dtype=torch.float64, device=q.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0].item()), float(t[1].item())


def measure_configs0(local, dev):
    """BASELINE configs[0] -- the only shape the reference itself ever runs: its shipped MiniLM index (125 x 384 fp32,
    IndexFlatL2), ONE query per call, k = 5, through `RetrievalSystem.retrieve` (src/retrieval.py:222 -> :92-115), 1 000
    queries (seeded perturbations of cyclic rows), latency per call.  Encoder = a table lookup (the
    transformer is out of scope): numpy output (the reference's call, pageable host buffers) and CUDA-tensor output (no
    host hop).  The CPU path beside it is the scalar C restatement of faiss's nq < 20 direct-form search on the same
    queries.  At 125 rows the search is pure launch latency: the GPU is EXPECTED to lose to a CPU here."""
    import torch
    import persian_rag_system_b200 as P
    from oracle import oracle as O
    path = os.path.join(ROOT, "tests", "golden", "indices", "paraphrase-multilingual-MiniLM-L12-v2_finetuned_drugs_word_chunks.index")
    x, _ = O.read_faiss_flat(path)
    n, d = x.shape
    rng = np.random.default_rng(100)
    rows = x[np.arange(1000) % n]
    q = (rows + 0.1 * np.linalg.norm(rows, axis=1, keepdims=True) / np.sqrt(d) * rng.standard_normal((1000, d))).astype(np.float32)
    texts = [f"q{i}" for i in range(1000)]
    table = {t: q[i:i + 1] for i, t in enumerate(texts)}
    qdev = torch.from_numpy(q).to(dev)
    dtable = {t: qdev[i:i + 1] for i, t in enumerate(texts)}

    class HostEnc:
        def encode(self, s, device=None):
            return table[s[0]]

    class DevEnc:
        def encode(self, s, device=None, convert_to_tensor=False):
            return dtable[s[0]]

    chunks = [{"id": f"word_chunk_{i}", "text": "x", "chunk_type": "word_based"} for i in range(n)]
    res = {"workload": f"configs[0]: reference MiniLM index {n} x {d} fp32 (IndexFlatL2), nq=1, k=5, 1000 queries, latency per RetrievalSystem.retrieve call"}
    got = {}
    for name, enc in (("numpy_embeddings_pageable_host_path", HostEnc()), ("cuda_tensor_embeddings_no_host_hop", DevEnc())):
        r = P.RetrievalSystem(method="dense", encoder=enc, device=f"cuda:{local}")
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            assert r.load_chunks(chunks, path)
        for t in texts[:50]:
            r.retrieve(t, 5)
        lat, ids = [], []
        for t in texts:
            t0 = time.perf_counter()
            hits = r.retrieve(t, 5)
            lat.append(time.perf_counter() - t0)
            ids.append([int(c["id"].rsplit("_", 1)[1]) for c, _ in hits])
        lat = np.array(lat) * 1e6
        got[name] = np.array(ids)
        res[name] = {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)), "mean_us": float(lat.mean()),
                     "qps_single_caller": float(1e6 / lat.mean())}
    # raw index call (no Python retriever around it): FlatIndex.search(numpy [1, d], 5)
    idx = P.read_index(path, device=local)
    lat = []
    for i in range(1000):
        t0 = time.perf_counter()
        idx.search(q[i:i + 1], 5)
        lat.append(time.perf_counter() - t0)
    lat = np.array(lat) * 1e6
    res["flat_index_search_numpy"] = {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99))}
    # CPU beside it: faiss when importable, else the scalar C restatement (nq = 1 -> direct form), one query per call
    faiss = O.reference_library("faiss")
    if faiss is not None:
        fi = faiss.IndexFlatL2(d)
        fi.add(x)
        cpu_call, kind = (lambda qq: fi.search(qq, 5)), "reference"
    else:
        cpu_call, kind = (lambda qq: O.flat_search_c(x, qq, 5, O.METRIC_L2, form=1)), "port"
    lat, Ic = [], []
    for i in range(1000):
        t0 = time.perf_counter()
        _D, I1 = cpu_call(q[i:i + 1])
        lat.append(time.perf_counter() - t0)
        Ic.append(I1[0])
    lat = np.array(lat) * 1e6
    res["cpu_baseline"] = {"p50_us": float(np.percentile(lat, 50)), "p99_us": float(np.percentile(lat, 99)), "cores": 1, "kind": kind,
                           "sample": "the same 1000 single-query searches (search call only, no retriever around it)"}
    Ic = np.array(Ic)
    res["parity"] = {name: {"identical_id_lists": int((g == Ic).all(axis=1).sum()), "of": 1000} for name, g in got.items()}
    return res


def measure_configs2(a, world, rank, local, dev, peak):
    """BASELINE configs[2]: distiluse shape d = 512, 50 M synthetic unit-norm chunks fp16 (51.2 GB), row-sharded over the N
    GPUs of the run, k = 100 (the wide-k path: sampled threshold -> collecting scan -> select, fully asynchronous) with
    the top-k merged across GPUs (fused exchange / NCCL), batches of 1, 64 and 1024 queries."""
    import torch
    import torch.distributed as dist
    import persian_rag_system_b200 as P
    from persian_rag_system_b200.sharded import ShardedFlatIndex, shard_bounds
    d, rows_total, k = 512, 50_000_000, 100
    lo, hi = shard_bounds(rows_total, world, rank)
    rows = hi - lo
    sh = ShardedFlatIndex(d, P.METRIC_INNER_PRODUCT, "fp16", device=local, exchange=a.exchange, nq_cap=1024, k_cap=128, lanes=1)
    idx = sh.local
    idx.reserve(rows)
    gen = torch.Generator(device=dev).manual_seed(512 + rank)
    slab = (512 << 20) // (d * 4)
    done = 0
    while done < rows:
        c = min(slab, rows - done)
        xb = torch.randn(c, d, generator=gen, device=dev)
        xb /= xb.norm(dim=1, keepdim=True)
        idx.add(xb.half())
        done += c
    del xb
    sh.offset, sh.ntotal_global = lo, rows_total
    idx.set_id_offset(lo)
    if world > 1:
        dist.barrier()
    gq = torch.Generator(device=dev).manual_seed(78)
    cases = []
    shard_bytes = rows * d * 2
    for B in (1, 64, 1024):
        q = torch.randn(B, d, generator=gq, device=dev)
        q /= q.norm(dim=1, keepdim=True)
        ms, scan = _timed_search(sh, idx, q, k, 8 if B <= 64 else 3, world)
        passes = (B + 127) // 128
        cases.append({"batch": B, "ms_per_batch": ms, "qps": B / (ms * 1e-3), "collect_scan_ms": scan,
                      "collect_scan_gbs_per_gpu": passes * shard_bytes / (scan * 1e-3) / 1e9,
                      "frac_hbm_collect_scan": passes * shard_bytes / (scan * 1e-3) / 1e9 / peak,
                      "whole_search_gbs_per_gpu": passes * shard_bytes / (ms * 1e-3) / 1e9})
    D, I = sh.search(q[:4], k)
    ok = bool((I >= 0).all().item() and (I < rows_total).all().item() and (D[:, :-1] >= D[:, 1:]).all().item())
    res = {"workload": f"configs[2]: {rows_total} x {d} fp16 rows over {world} GPU(s) ({rows} per GPU), inner product, k={k}",
           "cases": cases, "results_valid": ok, "exchange": a.exchange if world > 1 else "none"}
    if world > 1:
        sh.check_exchange()
    # CPU beside it (rank 0, N = 1): the corpus (102 GB as fp32) exceeds what a host arm should hold, so the CPU path is
    # timed on the first 1 M rows and extrapolated linearly in N (a flat scan is linear), as SURVEY 8d prescribes
    if world == 1 and rank == 0 and not a.no_cpu_baseline:
        from oracle import oracle as O
        threads = use_all_host_threads()
        x32 = idx.reconstruct_n(0, 1_000_000)
        qh = q[:64].cpu().numpy()
        search, kind, note = cpu_flat_search()
        search(x32[:65536], qh, k, O.METRIC_IP)
        t0 = time.perf_counter()
        search(x32, qh, k, O.METRIC_IP)
        dt = time.perf_counter() - t0
        res["cpu_baseline"] = {"value": 64 / (dt * rows_total / 1_000_000), "unit": "queries/s", "cores": threads, "kind": kind,
                               "sample": "64 queries over the first 1M rows (fp32), time scaled x50 to the 50M-row corpus", "note": note}
    return res


def measure_configs3(dev, peak):
    """BASELINE configs[3]: BM25 over a synthetic Persian-vocabulary CSR matrix -- 10 M docs, 200 k terms, ~100 draws per
    doc (Zipf 1.07; ~73 distinct terms), 4 096 queries of 1 + Poisson(6) Zipf tokens, k = 10.  Both scoring
    kernels through prs_sparse_search_device (query CSR and results on the device), CUDA events around the call.
    Roofline: HBM, algorithmic bytes = 8 x sum of df over the query tokens (one doc id + one fp32 weight per posting)."""
    import torch
    import scipy.sparse as sp
    import persian_rag_system_b200 as P
    from oracle import oracle as O
    from tools.bench_aux import gen_sparse
    docs, terms, nq, k = 10_000_000, 200_000, 4096, 10
    t0 = time.time()
    indptr, indices, values, cdf, _df = gen_sparse(docs, terms, 7, dev)
    t_gen = time.time() - t0
    t0 = time.time()
    idx = P.SparseIndex(indptr, indices, values, terms, device=int(dev.index or 0))
    t_build = time.time() - t0
    rng = np.random.default_rng(11)
    qlen = 1 + rng.poisson(6, size=nq)
    q_indptr = np.zeros(nq + 1, np.int64)
    q_indptr[1:] = np.cumsum(qlen)
    u = torch.from_numpy(rng.random(int(q_indptr[-1]))).to(dev)
    q_terms = torch.searchsorted(cdf, u).clamp_(max=terms - 1).to(torch.int32)
    d_ip = torch.from_numpy(q_indptr).to(dev)
    d_qw = torch.ones(q_terms.shape[0], dtype=torch.float64, device=dev)
    res = {"workload": f"configs[3]: BM25, {docs} docs x {terms} terms, nnz {int(indices.shape[0])} (~{indices.shape[0] / docs:.0f}/doc), {nq} queries, k={k}",
           "build": {"generate_s": t_gen, "index_build_s": t_build}, "modes": {}}
    out = {}
    for mode in ("exact", "throughput"):
        t0 = time.time()
        idx.set_mode(mode)
        torch.cuda.synchronize()
        t_prep = time.time() - t0
        idx.search_device(d_ip, q_terms, d_qw, k)
        torch.cuda.synchronize()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            S, I = idx.search_device(d_ip, q_terms, d_qw, k)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        postings = idx.last_postings
        ms = float(np.median(ts))
        out[mode] = (S.cpu().numpy(), I.cpu().numpy())
        res["modes"][mode] = {"ms_per_batch": ms, "qps": nq / (ms * 1e-3), "algorithmic_gbs": 8.0 * postings / (ms * 1e-3) / 1e9,
                              "frac_hbm": 8.0 * postings / (ms * 1e-3) / 1e9 / peak, "prepare_s": t_prep}
    res["postings_touched"] = int(postings)
    res["modes_agree"] = {"scores_identical": bool(np.array_equal(out["exact"][0], out["throughput"][0])),
                          "id_positions_identical": float((out["exact"][1] == out["throughput"][1]).mean())}
    res["roofline"] = {"bound": "hbm", "achieved": res["modes"]["throughput"]["algorithmic_gbs"], "peak": peak, "unit": "GB/s",
                       "frac": res["modes"]["throughput"]["frac_hbm"], "kernel": "sparse_score_batched_kernel (+ merge + exact re-score)",
                       "note": "algorithmic bytes assume every query streams its own postings; the throughput kernel shares a posting between "
                               "the queries of a CTA and the L2 between CTAs, so its DRAM traffic is far below this figure"}
    # CPU beside it: vectorised scipy CSC column adds + canonical top-k (faster than rank_bm25's pure-Python loop), 4 queries
    M = sp.csr_matrix((values.astype(np.float64), indices, indptr), shape=(docs, terms)).tocsc()
    qt = q_terms.cpu().numpy()
    t0 = time.perf_counter()
    bad = 0
    for qi in range(4):
        sc = np.zeros(docs, np.float64)
        for t in qt[q_indptr[qi]:q_indptr[qi + 1]]:
            a0, a1 = M.indptr[t], M.indptr[t + 1]
            sc[M.indices[a0:a1]] += M.data[a0:a1]
        try:
            O.check_topk_against_scores(out["throughput"][1][qi], out["throughput"][0][qi], sc, k, True, rtol=1e-5, atol=1e-9)
        except AssertionError:
            bad += 1
    res["cpu_baseline"] = {"value": 4 / (time.perf_counter() - t0), "unit": "queries/s", "cores": 1, "kind": "port",
                           "sample": "first 4 queries, scipy CSC column adds + stable top-k (rank_bm25 itself is absent; this is faster than its Python loop)"}
    res["parity"] = {"queries_checked_against_cpu": 4, "mismatch": bad}
    return res


def measure_capacity(a, world, rank, local, dev, peak, sweep=False):
    import torch
    import torch.distributed as dist
    import persian_rag_system_b200 as P
    from persian_rag_system_b200.sharded import ShardedFlatIndex
    d, rows, B, k = 384, a.capacity_rows, 64, 10
    sh = ShardedFlatIndex(d, P.METRIC_INNER_PRODUCT, "fp16", device=local, exchange=a.exchange, nq_cap=1024 if sweep else 64,
                          k_cap=128 if sweep else 16, lanes=1)
    idx = sh.local
    idx.reserve(rows)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    slab = (512 << 20) // (d * 4)
    done = 0
    while done < rows:
        c = min(slab, rows - done)
        xb = torch.randn(c, d, generator=gen, device=dev)
        xb /= xb.norm(dim=1, keepdim=True)
        idx.add(xb.half())
        done += c
    del xb
    sh.offset, sh.ntotal_global = rank * rows, rows * world
    idx.set_id_offset(rank * rows)
    gq = torch.Generator(device=dev).manual_seed(77)
    q = torch.randn(B, d, generator=gq, device=dev)
    q /= q.norm(dim=1, keepdim=True)
    for _ in range(3):
        D, I = sh.search(q, k)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    iters = 20
    idx.set_timing(True); idx.scan_time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        D, I = sh.search(q, k)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1) / iters
    scan_ms, n = idx.scan_time()
    idx.set_timing(False)
    t = torch.tensor([ms, scan_ms / max(n, 1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sh.check_exchange()
    ms, scan = (float(v) for v in t.tolist())
    shard_bytes = rows * d * 2
    # sanity: the best hit of query 0 must be a valid global id
    ok = bool((I[:, 0] >= 0).all().item() and (I[:, 0] < rows * world).all().item())
    sweep_rows = None
    if sweep:
        # BASELINE configs[4]: batch x k sweep on the same per-GPU share (B in {1,16,256,4096} x k in {1,10,100,1024})
        sweep_rows = []
        for Bs in (1, 16, 256, 4096):
            qs = torch.randn(Bs, d, generator=gq, device=dev)
            qs /= qs.norm(dim=1, keepdim=True)
            for ks in (1, 10, 100, 1024):
                m, sc = _timed_search(sh, idx, qs, ks, 6 if Bs <= 256 else 2, world)
                sweep_rows.append({"batch": Bs, "k": ks, "ms_per_batch": round(m, 4), "qps": round(Bs / (m * 1e-3), 1),
                                   "scan_ms": round(sc, 4), "tflops": round(2.0 * rows * world * d * Bs / (m * 1e-3) / 1e12, 1)})
        if world > 1:
            sh.check_exchange()
    return {"batch_k_sweep": sweep_rows,
            "workload": f"configs[4] share: {rows} x {d} fp16 rows per GPU, global corpus {rows * world} rows, batch {B}, k={k}",
            "scaling": "weak (corpus grows with N)", "ms_per_batch": ms, "qps": B / (ms * 1e-3), "scan_ms": scan,
            "scan_gbs_per_gpu": shard_bytes / (scan * 1e-3) / 1e9, "frac_hbm": shard_bytes / (scan * 1e-3) / 1e9 / peak,
            "aggregate_scan_gbs": world * shard_bytes / (scan * 1e-3) / 1e9, "ids_valid": ok}


def _arm_watchdog(out, rank, budget_s):
    """After `budget_s` seconds: dump every thread's stack to stderr, print the JSON collected so far (rank 0) with
    `secondary_incomplete`, and end the process.  Cancelled by the normal end of the run."""
    def bail():
        try:
            faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
            if rank == 0:
                o = dict(out)
                o["secondary_incomplete"] = f"a secondary measurement did not finish within {budget_s} s; stacks on stderr"
                print(json.dumps(o))
                sys.stdout.flush()
        finally:
            os._exit(0)
    t = threading.Timer(float(budget_s), bail)
    t.daemon = True
    t.start()
    return t


def main():
    a = parse()
    faulthandler.enable()
    faulthandler.dump_traceback_later(240, repeat=True, exit=False)   # a default run takes about a minute: leave the stacks of a stuck one on stderr
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
    # The line is out: leave without the interpreter / CUDA / NCCL / NVML tear-down (two runs of this round ended with an
    # empty stdout file and a process that never exited; the line sits in the stdout buffer until it is flushed).
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()

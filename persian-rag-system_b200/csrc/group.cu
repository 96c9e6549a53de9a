// group.cu -- ONE process, several GPUs: a flat index whose rows are split over the devices of the box.
//
// The reference's retriever is a single Python process (RetrievalSystem, src/retrieval.py:13; the Gradio app
// at scripts/gradio_luncher.py:354-362 shares one instance between its threads), so the drop-in must be able
// to use every GPU of the box without torchrun.  A prs_group owns one shard index + one exchange buffer per
// device and one host worker thread per device; peer access is enabled between all pairs, so the exchange
// buffers are addressed directly (no CUDA IPC).  A search is the same as in the one-process-per-GPU layout
// (sharded.py): every device scans its contiguous row block and runs the fused merge + NVLink exchange kernel
// (xchg.cuh); the workers issue the launches of all devices in parallel.  Device 0 of the group writes the
// caller's D / I.  Rows are dealt in contiguous ascending blocks of ceil(n_total / G) rows (reserve first, like
// the index builders of the reference mirror do), so ties resolve on global ids exactly as in the single index.
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "host_common.h"

using namespace prs;

namespace {

struct Worker {
    std::thread th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<int()> job;
    bool has = false, quit = false;
    int rc = 0;
    std::string err;
    void loop() {
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv.wait(lk, [&] { return has || quit; });
            if (quit) return;
            lk.unlock();
            const int r = job();
            const std::string e = r ? std::string(prs_last_error()) : std::string();
            lk.lock();
            rc = r; err = e; has = false;
            cv.notify_all();
        }
    }
};

}  // namespace

struct prs_group {
    int G = 0, d = 0, metric = 0, storage = 0;
    long long nq_cap = 0, k_cap = 0, block = 0, n = 0;       // block: rows per shard (0 = not decided yet)
    std::vector<int> dev;
    std::vector<prs_index*> shard;
    std::vector<prs_xchg*> xchg;
    std::vector<cudaStream_t> stream;
    std::vector<DevBuf> sD, sI, sQ;                           // per-device scratch: results of the non-leading devices, staged queries
    std::vector<std::unique_ptr<Worker>> worker;
    cudaEvent_t ev_q = nullptr;
    std::mutex mu;
};

static int group_run(prs_group* g, const std::function<int(int)>& fn) {
    for (int i = 0; i < g->G; ++i) {
        Worker& w = *g->worker[i];
        std::lock_guard<std::mutex> lk(w.mu);
        w.job = [fn, i] { return fn(i); };
        w.has = true;
        w.cv.notify_all();
    }
    int rc = 0;
    for (int i = 0; i < g->G; ++i) {
        Worker& w = *g->worker[i];
        std::unique_lock<std::mutex> lk(w.mu);
        w.cv.wait(lk, [&] { return !w.has; });
        if (w.rc && !rc) { rc = w.rc; set_error("device %d: %s", g->dev[i], w.err.c_str()); }
    }
    return rc;
}

extern "C" {

int prs_group_create(int d, int metric, int storage, const int* devs, int ndev, int64_t nq_cap, int k_cap, prs_group** out) {
    if (!out) { set_error("group_create: out is null"); return PRS_EINVAL; }
    *out = nullptr;
    if (!devs || ndev < 1 || ndev > 16 || nq_cap < 1 || k_cap < 1 || k_cap > PRS_MAX_K) { set_error("group_create: bad arguments"); return PRS_EINVAL; }
    for (int a = 0; a < ndev; ++a) for (int b = 0; b < a; ++b) if (devs[a] == devs[b]) { set_error("group_create: device %d listed twice", devs[a]); return PRS_EINVAL; }
    prs_group* g = new (std::nothrow) prs_group();
    if (!g) { set_error("out of host memory"); return PRS_ENOMEM; }
    g->G = ndev; g->d = d; g->metric = metric; g->storage = storage; g->nq_cap = nq_cap; g->k_cap = k_cap;
    g->dev.assign(devs, devs + ndev);
    g->shard.assign(ndev, nullptr); g->xchg.assign(ndev, nullptr); g->stream.assign(ndev, nullptr);
    g->sD.resize(ndev); g->sI.resize(ndev); g->sQ.resize(ndev);
    int rc = 0;
    for (int i = 0; i < ndev && !rc; ++i) {
        rc = prs_index_create(d, metric, storage, devs[i], &g->shard[i]);
        if (rc) break;
        DeviceGuard dg(devs[i]);
        for (int j = 0; j < ndev && !rc; ++j) {
            if (j == i) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devs[i], devs[j]) != cudaSuccess || !can) {
                cudaGetLastError();
                set_error("group_create: device %d cannot address device %d (no peer access)", devs[i], devs[j]);
                rc = PRS_EUNSUP;
                break;
            }
            const cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", devs[i], devs[j], cudaGetErrorString(e)); rc = PRS_ECUDA; }
            cudaGetLastError();
        }
        if (!rc && cudaStreamCreateWithFlags(&g->stream[i], cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); set_error("cudaStreamCreate failed"); rc = PRS_ECUDA; }
        if (!rc && ndev > 1) rc = prs_xchg_create(devs[i], ndev, i, nq_cap, k_cap, &g->xchg[i]);
    }
    if (!rc && ndev > 1) rc = prs_xchg_link_local(g->xchg.data(), ndev);
    if (!rc) {
        DeviceGuard dg(devs[0]);
        if (cudaEventCreateWithFlags(&g->ev_q, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); set_error("cudaEventCreate failed"); rc = PRS_ECUDA; }
    }
    if (rc) { const std::string keep = prs_last_error(); prs_group_free(g); set_error("%s", keep.c_str()); return rc; }
    for (int i = 0; i < ndev; ++i) {
        g->worker.emplace_back(new Worker());
        Worker* w = g->worker.back().get();
        w->th = std::thread([w] { w->loop(); });
    }
    *out = g;
    return 0;
}

void prs_group_free(prs_group* g) {
    if (!g) return;
    for (auto& w : g->worker) {
        { std::lock_guard<std::mutex> lk(w->mu); w->quit = true; w->cv.notify_all(); }
        if (w->th.joinable()) w->th.join();
    }
    for (int i = 0; i < g->G; ++i) {
        DeviceGuard dg(g->dev[i]);
        cudaDeviceSynchronize();
        if (g->xchg[i]) prs_xchg_free(g->xchg[i]);
        if (g->shard[i]) prs_index_free(g->shard[i]);
        if (g->stream[i]) cudaStreamDestroy(g->stream[i]);
        g->sD[i].release(); g->sI[i].release(); g->sQ[i].release();
    }
    if (g->ev_q) { DeviceGuard dg(g->dev[0]); cudaEventDestroy(g->ev_q); }
    delete g;
}

int prs_group_ndev(const prs_group* g) { return g ? g->G : -1; }
int64_t prs_group_ntotal(const prs_group* g) { return g ? g->n : -1; }
int prs_group_d(const prs_group* g) { return g ? g->d : -1; }
int prs_group_metric(const prs_group* g) { return g ? g->metric : -1; }
int prs_group_storage(const prs_group* g) { return g ? g->storage : -1; }
int64_t prs_group_shard_rows(const prs_group* g, int i) { return (g && i >= 0 && i < g->G) ? prs_index_ntotal(g->shard[i]) : -1; }

int prs_group_reserve(prs_group* g, int64_t n_total) {
    if (!g || n_total < 0) { set_error("group_reserve: bad arguments"); return PRS_EINVAL; }
    std::lock_guard<std::mutex> lock(g->mu);
    if (g->n > 0) { set_error("group_reserve: rows were already added (the block size is fixed by the first reserve / add)"); return PRS_EINVAL; }
    g->block = (n_total + g->G - 1) / g->G;
    if (g->storage != PRS_F32) g->block = (g->block + 63) / 64 * 64;          // T64 blocks stay whole inside a shard
    if (g->block < 1) g->block = 1;
    return group_run(g, [g](int i) { return prs_index_reserve(g->shard[i], g->block); });
}

}  // extern "C"

// rows [r0, r0+n) of the global index -> (shard, count) pieces; rows past G*block go to the last shard
template <class F>
static int group_deal(prs_group* g, long long n, F&& piece) {
    long long done = 0;
    while (done < n) {
        const long long row = g->n + done;
        int s = (int)std::min<long long>(g->G - 1, row / g->block);
        const long long room = s == g->G - 1 ? n - done : std::min<long long>(n - done, (long long)(s + 1) * g->block - row);
        int rc = piece(s, done, room);
        if (rc) return rc;
        done += room;
    }
    return 0;
}

static void group_fix_offsets(prs_group* g) {
    long long off = 0;
    for (int i = 0; i < g->G; ++i) { prs_index_set_id_offset(g->shard[i], off); off += prs_index_ntotal(g->shard[i]); }
}

extern "C" {

int prs_group_add_host(prs_group* g, const float* x, int64_t n) {
    if (!g || n < 0 || (n > 0 && !x)) { set_error("group_add: bad arguments"); return PRS_EINVAL; }
    if (n == 0) return 0;
    if (g->block == 0) { int rc = prs_group_reserve(g, n); if (rc) return rc; }
    std::lock_guard<std::mutex> lock(g->mu);
    int rc = group_deal(g, n, [&](int s, long long first, long long cnt) {
        return prs_index_add_host(g->shard[s], x + (size_t)first * g->d, cnt);
    });
    if (rc) return rc;
    g->n += n;
    group_fix_offsets(g);
    return 0;
}

// x: [n, d] on ANY device of the group (the owning device's memory is read over NVLink by the others)
int prs_group_add_device(prs_group* g, const void* x, int dtype, int64_t n, void* stream) {
    if (!g || n < 0 || (n > 0 && !x)) { set_error("group_add: bad arguments"); return PRS_EINVAL; }
    if (n == 0) return 0;
    if (g->block == 0) { int rc = prs_group_reserve(g, n); if (rc) return rc; }
    std::lock_guard<std::mutex> lock(g->mu);
    cudaPointerAttributes a;
    PRS_CUDA(cudaPointerGetAttributes(&a, x));
    const int src_dev = a.type == cudaMemoryTypeDevice ? a.device : g->dev[0];
    { DeviceGuard dg(src_dev); PRS_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); }      // the rows exist before a peer reads them
    const size_t es = dtype == PRS_F32 ? 4 : 2;
    int rc = group_deal(g, n, [&](int s, long long first, long long cnt) {
        int r = prs_index_add_device(g->shard[s], (const unsigned char*)x + (size_t)first * g->d * es, dtype, cnt, nullptr);
        if (r) return r;
        DeviceGuard dg(g->dev[s]);
        PRS_CUDA(cudaStreamSynchronize(0));
        return 0;
    });
    if (rc) return rc;
    g->n += n;
    group_fix_offsets(g);
    return 0;
}

static int group_search_chunk(prs_group* g, const void* q, int qdtype, long long nq, int k, float* D, int64_t* I, cudaStream_t user) {
    // the queries may be produced on the caller's stream: the other devices wait for them
    { DeviceGuard dg(g->dev[0]); PRS_CUDA(cudaEventRecord(g->ev_q, user)); }
    return group_run(g, [=](int i) -> int {
        DeviceGuard dg(g->dev[i]);
        cudaStream_t st = i == 0 ? user : g->stream[i];
        float* Di = D; int64_t* Ii = I;
        if (i > 0) {
            PRS_CUDA(cudaStreamWaitEvent(st, g->ev_q, 0));
            int rc;
            if ((rc = g->sD[i].ensure((size_t)nq * k * 4))) return rc;
            if ((rc = g->sI[i].ensure((size_t)nq * k * 8))) return rc;
            Di = (float*)g->sD[i].p; Ii = (int64_t*)g->sI[i].p;
        }
        if (g->G == 1) return prs_index_search_device(g->shard[0], q, qdtype, nq, k, Di, Ii, st);
        return prs_index_search_sharded_device(g->shard[i], g->xchg[i], q, qdtype, nq, k, Di, Ii, st);
    });
}

// q, D, I: device pointers on the group's FIRST device (or page-locked host memory); asynchronous on `stream`
// (a stream of the first device).  Every device scans its block; the first device writes D / I.
int prs_group_search_device(prs_group* g, const void* q, int qdtype, int64_t nq, int k, float* D, int64_t* I, void* stream) {
    if (!g) { set_error("null group"); return PRS_EINVAL; }
    if (k < 1 || k > g->k_cap) { set_error("group_search: k=%d exceeds the group's k_cap=%lld", k, g->k_cap); return PRS_EINVAL; }
    if (nq < 0) { set_error("group_search: nq < 0"); return PRS_EINVAL; }
    if (nq == 0) return 0;
    if (!q || !D || !I) { set_error("group_search: null pointer"); return PRS_EINVAL; }
    std::lock_guard<std::mutex> lock(g->mu);
    const size_t qes = qdtype == PRS_F32 ? 4 : 2;
    const long long chunk = std::max<long long>(1, std::min<long long>(g->nq_cap, (g->nq_cap * g->k_cap) / k));
    for (long long q0 = 0; q0 < nq; q0 += chunk) {
        const long long c = std::min<long long>(chunk, nq - q0);
        int rc = group_search_chunk(g, (const unsigned char*)q + (size_t)q0 * g->d * qes, qdtype, c, k, D + (size_t)q0 * k, I + (size_t)q0 * k,
                                    (cudaStream_t)stream);
        if (rc) return rc;
    }
    return 0;
}

int prs_group_search_host(prs_group* g, const float* q, int64_t nq, int k, float* D, int64_t* I) {
    if (!g) { set_error("null group"); return PRS_EINVAL; }
    if (nq < 0) { set_error("group_search: nq < 0"); return PRS_EINVAL; }
    if (nq == 0) return 0;
    if (!q || !D || !I) { set_error("group_search: null pointer"); return PRS_EINVAL; }
    DeviceGuard dg(g->dev[0]);
    int rc;
    {
        std::lock_guard<std::mutex> lock(g->mu);
        if ((rc = g->sQ[0].ensure((size_t)nq * g->d * 4))) return rc;
        if ((rc = g->sD[0].ensure((size_t)nq * k * 4))) return rc;
        if ((rc = g->sI[0].ensure((size_t)nq * k * 8))) return rc;
    }
    cudaStream_t st = g->stream[0];
    PRS_CUDA(cudaMemcpyAsync(g->sQ[0].p, q, (size_t)nq * g->d * 4, cudaMemcpyHostToDevice, st));
    if ((rc = prs_group_search_device(g, g->sQ[0].p, PRS_F32, nq, k, (float*)g->sD[0].p, (int64_t*)g->sI[0].p, st))) return rc;
    PRS_CUDA(cudaMemcpyAsync(D, g->sD[0].p, (size_t)nq * k * 4, cudaMemcpyDeviceToHost, st));
    PRS_CUDA(cudaMemcpyAsync(I, g->sI[0].p, (size_t)nq * k * 8, cudaMemcpyDeviceToHost, st));
    PRS_CUDA(cudaStreamSynchronize(st));
    // a peer that timed out reports through its exchange context: surface it now rather than on the next call
    for (int i = 0; i < g->G && g->G > 1; ++i) if ((rc = prs_xchg_status(g->xchg[i]))) return rc;
    return 0;
}

int prs_group_reconstruct_host(prs_group* g, int64_t i0, int64_t n, float* out) {
    if (!g || i0 < 0 || n < 0 || i0 + n > g->n || (n > 0 && !out)) { set_error("group_reconstruct: range out of bounds"); return PRS_EINVAL; }
    std::lock_guard<std::mutex> lock(g->mu);
    long long first = 0;
    for (int i = 0; i < g->G && n > 0; ++i) {
        const long long rows = prs_index_ntotal(g->shard[i]);
        const long long a = std::max<long long>(i0, first), b = std::min<long long>(i0 + n, first + rows);
        if (a < b) {
            int rc = prs_index_reconstruct_host(g->shard[i], a - first, b - a, out + (size_t)(a - i0) * g->d);
            if (rc) return rc;
        }
        first += rows;
    }
    return 0;
}

}  // extern "C"

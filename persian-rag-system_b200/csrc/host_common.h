// host_common.h -- host-side helpers shared by the translation units of libprs.so
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/prs.h"

namespace prs {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

#define PRS_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            prs::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PRS_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)

#define PRS_LAUNCH_CHECK()                                                                    \
    do {                                                                                      \
        prs::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            prs::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return PRS_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)

// sets the device for the current scope and restores the previous one
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; prev = -1; }
        if (cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// grow-only device buffer
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        size_t want = need + need / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&p, need);
            want = need;
        }
        if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e)); p = nullptr; return PRS_ENOMEM; }
        bytes = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};

// grow-only page-locked host buffer the device can address (zero-copy staging of small host calls)
struct PinnedBuf {
    void* p = nullptr;       // host address
    void* dp = nullptr;      // the same memory as the device sees it
    size_t bytes = 0;
    int ensure(size_t need) {
        if (need <= bytes) return 0;
        release();
        const size_t want = need + need / 2;
        if (cudaHostAlloc(&p, want, cudaHostAllocMapped) != cudaSuccess || cudaHostGetDevicePointer(&dp, p, 0) != cudaSuccess) {
            cudaGetLastError(); release(); set_error("cudaHostAlloc(%zu) failed", want); return PRS_ENOMEM;
        }
        bytes = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; dp = nullptr; bytes = 0; }
};

// bench instrumentation: brackets the dominant (scan) kernel launches with CUDA events on the
// launching stream; collect() synchronises them and returns the summed device time.
struct ScanTimer {
    bool enabled = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
    std::vector<cudaEvent_t> pool;                 // events are reused: creating one costs microseconds of host time
    cudaEvent_t take() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    void begin(cudaStream_t st) {
        if (!enabled) return;
        cudaEvent_t a = take(), b = take();
        cudaEventRecord(a, st);
        ev.emplace_back(a, b);
    }
    void end(cudaStream_t st) { if (enabled && !ev.empty()) cudaEventRecord(ev.back().second, st); }
    int collect(double* total_ms, long long* launches) {
        double tot = 0.0;
        cudaError_t bad = cudaSuccess;
        for (auto& e : ev) {
            float ms = 0.f;
            cudaError_t r = cudaEventSynchronize(e.second);
            if (r == cudaSuccess) r = cudaEventElapsedTime(&ms, e.first, e.second);
            pool.push_back(e.first); pool.push_back(e.second);
            if (r != cudaSuccess) bad = r;
            tot += ms;
        }
        *total_ms = tot; *launches = (long long)ev.size();
        ev.clear();
        if (bad != cudaSuccess) { cudaGetLastError(); set_error("scan timer: %s", cudaGetErrorString(bad)); return PRS_ECUDA; }
        return 0;
    }
    ~ScanTimer() {
        for (auto& e : ev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto e : pool) cudaEventDestroy(e);
    }
};

}  // namespace prs

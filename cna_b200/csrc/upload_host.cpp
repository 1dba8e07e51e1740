// Host -> device upload of the caller's pageable buffers at PCIe speed.
//
// Reference: none — the reference never leaves the host.  The hot path's end-to-end cost from a host AnnData
// is dominated by moving .obsp['connectivities'] (src/cna/tools/_nam.py:12-19: 0.46 GB at 1 M cells) to the
// device.  scipy's CSR buffers are pageable; cudaMemcpyAsync stages pageable memory through the driver's
// own bounce buffer on ONE thread (~10 GB/s measured: 45 ms for config C, against 9 ms from pinned memory).
// Here a few threads copy 4 MB chunks into a ring of page-locked slots owned by the library and queue one
// asynchronous copy per chunk: the staging memcpy of chunk i+1 overlaps the DMA of chunk i, and the call
// returns (or its handle is released) once the last chunk has left the caller's buffer — the semantics of
// a pageable cudaMemcpyAsync, at a multiple of its speed.  Buffers that are already page-locked (pinned
// torch tensors, cudaHostRegister'ed arrays, cna_b200.read_h5ad) take a single cudaMemcpyAsync.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/cna_b200.h"

namespace cna {
int set_error(int code, const char *fmt, ...);  // api.cu
}

namespace {

constexpr size_t kChunk = size_t(4) << 20;
constexpr int kSlots = 16;

struct Ring {
    char *host = nullptr;
    cudaEvent_t done[kSlots] = {};
    bool used[kSlots] = {};
    std::mutex m;  // one upload at a time owns the ring
};
Ring g_ring;

int ensure_ring() {
    if (g_ring.host) return 0;
    if (cudaHostAlloc(reinterpret_cast<void **>(&g_ring.host), kChunk * kSlots, cudaHostAllocDefault) != cudaSuccess) {
        g_ring.host = nullptr;
        return 1;
    }
    for (int s = 0; s < kSlots; ++s)
        if (cudaEventCreateWithFlags(&g_ring.done[s], cudaEventDisableTiming) != cudaSuccess) return 1;
    return 0;
}

bool is_page_locked(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

int upload(void *dst, const void *src, int64_t bytes, cudaStream_t st, int n_threads) {
    if (bytes <= 0) return CNA_OK;
    if (is_page_locked(src) || size_t(bytes) <= kChunk) {
        if (cudaMemcpyAsync(dst, src, size_t(bytes), cudaMemcpyHostToDevice, st) != cudaSuccess)
            return cna::set_error(CNA_ERR_CUDA, "cna_host_upload: cudaMemcpyAsync failed: %s",
                                  cudaGetErrorString(cudaGetLastError()));
        return CNA_OK;
    }
    std::lock_guard<std::mutex> lk(g_ring.m);
    if (ensure_ring()) return cna::set_error(CNA_ERR_CUDA, "cna_host_upload: cannot allocate the staging ring");
    int dev = 0;
    cudaGetDevice(&dev);
    const int64_t n_chunks = (bytes + int64_t(kChunk) - 1) / int64_t(kChunk);
    int T = n_threads > 0 ? n_threads : 4;
    T = int(std::min<int64_t>(std::min(T, kSlots), n_chunks));
    std::atomic<int> failed{0};
    // thread t owns slots t, t + T, ... and chunks t, t + T, ...: a slot is reused by the thread that filled it,
    // after the copy that drained it
    auto work = [&](int t) {
        cudaSetDevice(dev);
        int turn = 0;
        for (int64_t c = t; c < n_chunks; c += T, ++turn) {
            const int per_thread = kSlots / T;
            const int slot = t + T * (turn % per_thread);
            if (g_ring.used[slot] && cudaEventSynchronize(g_ring.done[slot]) != cudaSuccess) failed = 1;
            const size_t off = size_t(c) * kChunk, len = std::min(kChunk, size_t(bytes) - off);
            char *stage = g_ring.host + size_t(slot) * kChunk;
            std::memcpy(stage, static_cast<const char *>(src) + off, len);
            if (cudaMemcpyAsync(static_cast<char *>(dst) + off, stage, len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                cudaEventRecord(g_ring.done[slot], st) != cudaSuccess)
                failed = 1;
            g_ring.used[slot] = true;
        }
    };
    std::vector<std::thread> threads;
    for (int t = 1; t < T; ++t) threads.emplace_back(work, t);
    work(0);
    for (auto &th : threads) th.join();
    if (failed) return cna::set_error(CNA_ERR_CUDA, "cna_host_upload: %s", cudaGetErrorString(cudaGetLastError()));
    return CNA_OK;
}

struct Job {
    std::thread th;
    int rc = 0;
};

}  // namespace

extern "C" int cna_host_upload(void *dst, const void *src, int64_t bytes, void *stream, int n_threads) {
    if (!dst || !src) return cna::set_error(CNA_ERR_INVALID, "cna_host_upload: null pointer");
    return upload(dst, src, bytes, reinterpret_cast<cudaStream_t>(stream), n_threads);
}

extern "C" void *cna_host_upload_async(void *dst, const void *src, int64_t bytes, void *stream, int n_threads) {
    if (!dst || !src) {
        cna::set_error(CNA_ERR_INVALID, "cna_host_upload_async: null pointer");
        return nullptr;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    Job *job = new Job;
    job->th = std::thread([=] {
        cudaSetDevice(dev);
        job->rc = upload(dst, src, bytes, reinterpret_cast<cudaStream_t>(stream), n_threads);
    });
    return job;
}

extern "C" int cna_host_upload_wait(void *handle) {
    if (!handle) return CNA_ERR_INVALID;
    Job *job = static_cast<Job *>(handle);
    job->th.join();
    const int rc = job->rc;
    delete job;
    return rc;
}

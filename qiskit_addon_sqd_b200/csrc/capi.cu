// Error plumbing and version of the C-ABI (include/sqd_b200.h).
#include <sched.h>
#include <stdarg.h>
#include <stdlib.h>
#include <unistd.h>

#include <chrono>

#include <atomic>

#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
static thread_local long long t_launches = 0;   // launches issued (or captured) by this host thread

long long thread_launches() { return t_launches; }
void add_launches(long long n) {
    g_launches.fetch_add(n, std::memory_order_relaxed);
    t_launches += n;
}

int check_launch(const char* what, int n_launched) {
    g_launches.fetch_add(n_launched, std::memory_order_relaxed);
    t_launches += n_launched;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

static thread_local void* g_pinned = nullptr;
static thread_local cudaEvent_t g_sync_event = nullptr;
static thread_local int g_sync_event_dev = -1;
constexpr size_t kPinnedBytes = 64 * 1024;

// Wait for everything enqueued on `st`.  cudaStreamSynchronize spins: fine for one process, but the host
// threads of concurrent solves (8 per process, 8 processes on a 32-core host when all GPUs are in use) then
// burn the cores that the launching threads need (round 1: 0.58 weak-scaling efficiency at 8 GPUs with no
// collective on the data path).  Modes (SQD_WAIT_MODE): 0 = cudaStreamSynchronize; 1 = blocking-sync event
// (the thread sleeps in the driver: no CPU, but ~100 us wake-up latency); 2 = hybrid (default when the
// process is one of several ranks): poll the event, yielding for the first 40 us, then sleeping 20 us a time.
int stream_wait_blocking(cudaStream_t st) {
    static const int knob = [] {
        const char* v = getenv("SQD_WAIT_MODE");
        if (v) return atoi(v);
        const char* w = getenv("WORLD_SIZE");
        return (w && atoi(w) > 2) ? 2 : 0;
    }();
    if (knob == 0) {
        SQD_CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    }
    int dev = 0;
    SQD_CUDA_OK(cudaGetDevice(&dev));
    if (g_sync_event == nullptr || g_sync_event_dev != dev) {
        if (g_sync_event != nullptr) cudaEventDestroy(g_sync_event);
        SQD_CUDA_OK(cudaEventCreateWithFlags(
            &g_sync_event, (knob == 1 ? cudaEventBlockingSync : cudaEventDefault) | cudaEventDisableTiming));
        g_sync_event_dev = dev;
    }
    SQD_CUDA_OK(cudaEventRecord(g_sync_event, st));
    if (knob == 1) {
        SQD_CUDA_OK(cudaEventSynchronize(g_sync_event));
        return 0;
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        const cudaError_t e = cudaEventQuery(g_sync_event);
        if (e == cudaSuccess) return 0;
        if (e != cudaErrorNotReady) {
            set_error("cudaEventQuery failed: %s", cudaGetErrorString(e));
            return -2;
        }
        const auto us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0)
                            .count();
        if (us < 40) sched_yield();
        else usleep(20);
    }
}

int read_back(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
    if (bytes > kPinnedBytes) {
        set_error("read_back: %zu bytes exceed the pinned staging buffer", bytes);
        return -1;
    }
    if (g_pinned == nullptr) SQD_CUDA_OK(cudaHostAlloc(&g_pinned, kPinnedBytes, cudaHostAllocPortable));
    SQD_CUDA_OK(cudaMemcpyAsync(g_pinned, d_src, bytes, cudaMemcpyDeviceToHost, st));
    if (stream_wait_blocking(st)) return -2;
    memcpy(h_dst, g_pinned, bytes);
    return 0;
}

}  // namespace sqd

extern "C" {

int sqd_version(void) { return SQD_B200_VERSION; }

const char* sqd_last_error(void) { return sqd::g_err; }

int sqd_stream_wait(void* stream) { return sqd::stream_wait_blocking((cudaStream_t)stream); }

int sqd_download(void* h_dst, const void* d_src, long long bytes, void* stream) {
    // device -> caller's host array through a per-thread pinned staging buffer, the final memcpy included: the
    // calling Python thread has released the interpreter lock for the duration of the call, so the K solver
    // threads of a batch copy their results concurrently (a numpy-side copy out of the staging buffer holds
    // the lock: 58 MB per bench step, serialised; a D2H straight into pageable memory was measured slower)
    if (bytes < 0) {
        sqd::set_error("sqd_download: negative size");
        return -1;
    }
    if (bytes == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    static thread_local char* stage = nullptr;
    static thread_local size_t stage_bytes = 0;
    const size_t chunk_max = (size_t)32 << 20;
    const size_t want = (size_t)bytes < chunk_max ? (size_t)bytes : chunk_max;
    if (stage_bytes < want) {
        if (stage) cudaFreeHost(stage);
        stage = nullptr;
        stage_bytes = 0;
        size_t cap = (size_t)1 << 20;
        while (cap < want) cap <<= 1;
        SQD_CUDA_OK(cudaHostAlloc((void**)&stage, cap, cudaHostAllocPortable));
        stage_bytes = cap;
    }
    for (size_t off = 0; off < (size_t)bytes; off += stage_bytes) {
        const size_t n = (size_t)bytes - off < stage_bytes ? (size_t)bytes - off : stage_bytes;
        SQD_CUDA_OK(cudaMemcpyAsync(stage, (const char*)d_src + off, n, cudaMemcpyDeviceToHost, st));
        if (sqd::stream_wait_blocking(st)) return -2;
        memcpy((char*)h_dst + off, stage, n);
    }
    return 0;
}

long long sqd_launch_count(int reset) {
    return reset ? sqd::g_launches.exchange(0) : sqd::g_launches.load();
}

}  // extern "C"

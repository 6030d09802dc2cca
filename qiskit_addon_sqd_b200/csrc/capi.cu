// Error plumbing and version of the C-ABI (include/sqd_b200.h).
#include <stdarg.h>
#include <stdlib.h>

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

int check_launch(const char* what, int n_launched) {
    g_launches.fetch_add(n_launched, std::memory_order_relaxed);
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

// Wait for everything enqueued on `st` WITHOUT spinning: the host threads of concurrent solves (8 per
// process, 8 processes on a 32-core host when all GPUs are in use) would otherwise burn the cores that the
// launching threads need.
int stream_wait_blocking(cudaStream_t st) {
    static const int knob = getenv("SQD_BLOCKING_SYNC") ? atoi(getenv("SQD_BLOCKING_SYNC")) : 0;
    if (knob == 0) {
        SQD_CUDA_OK(cudaStreamSynchronize(st));
        return 0;
    }
    int dev = 0;
    SQD_CUDA_OK(cudaGetDevice(&dev));
    if (g_sync_event == nullptr || g_sync_event_dev != dev) {
        if (g_sync_event != nullptr) cudaEventDestroy(g_sync_event);
        SQD_CUDA_OK(cudaEventCreateWithFlags(&g_sync_event, cudaEventBlockingSync | cudaEventDisableTiming));
        g_sync_event_dev = dev;
    }
    SQD_CUDA_OK(cudaEventRecord(g_sync_event, st));
    SQD_CUDA_OK(cudaEventSynchronize(g_sync_event));
    return 0;
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

long long sqd_launch_count(int reset) {
    return reset ? sqd::g_launches.exchange(0) : sqd::g_launches.load();
}

}  // extern "C"

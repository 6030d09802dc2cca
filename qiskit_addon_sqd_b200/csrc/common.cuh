// Shared helpers for the sqd_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace sqd {

// ---- error plumbing (C-ABI: 0 = ok, <0 = error, message via sqd_last_error) ------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what, int n_launched = 1);  // also counts kernel launches
long long thread_launches();        // kernels launched or captured by this host thread so far
void add_launches(long long n);     // replays of a captured graph: count the kernels they stand for
// Small device -> host read-back through a per-thread PINNED staging buffer, then a stream sync.  A copy
// into pageable memory makes the runtime hold a context-wide lock until the stream has drained, which
// stalls the launches of every other host thread (one thread per concurrent subspace solve).
int read_back(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st);

#define SQD_CUDA_OK(expr)                                                              \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            sqd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                           __FILE__, __LINE__);                                        \
            return -2;                                                                 \
        }                                                                              \
    } while (0)

#define SQD_REQUIRE(cond, ...)            \
    do {                                  \
        if (!(cond)) {                    \
            sqd::set_error(__VA_ARGS__);  \
            return -1;                    \
        }                                 \
    } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- device helpers -----------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block reduction of NV values per thread (fixed tree, fixed order).
// `red` is shared scratch of at least NV * (blockDim.x/32) doubles.  Result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[k * nwarp + warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0.0;
            for (int w = 0; w < nwarp; ++w) s += red[k * nwarp + w];
            v[k] = s;
        }
    }
    __syncthreads();
}

// mask of bits strictly between positions p and q
__host__ __device__ __forceinline__ uint64_t between_mask(int p, int q) {
    const int lo = p < q ? p : q, hi = p < q ? q : p;
    return ((1ull << hi) - 1ull) & ~((2ull << lo) - 1ull);
}

__host__ __device__ __forceinline__ uint64_t below_mask(int k) { return (1ull << k) - 1ull; }

#ifdef __CUDA_ARCH__
__device__ __forceinline__ int popc64(uint64_t x) { return __popcll(x); }
__device__ __forceinline__ int lowbit64(uint64_t x) { return __ffsll((long long)x) - 1; }
#else
inline int popc64(uint64_t x) { return __builtin_popcountll(x); }
inline int lowbit64(uint64_t x) { return __builtin_ctzll(x); }
#endif

// ---- TMA-class bulk copy global -> shared (cp.async.bulk + mbarrier; SASS: UBLKCP) -----------
// One elected thread issues the copy; everybody waits on the mbarrier phase.
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

}  // namespace sqd

// Configuration recovery: per-bitstring Hamming-weight correction driven by orbital occupancies.
// Reference: qiskit_addon_sqd/configuration_recovery.py:59-306 (a Python loop per bitstring with four
// numpy Generator.choice(replace=False, p=...) call sites, :247, :260, :284, :299).
//
// B200 design, two kernels:
//   recover_prepare : one thread per (row, half).  Evaluates the flip weights with the reference's
//                     exact operation order (no FMA contraction; numpy's pairwise summation order for
//                     np.sum, sequential order for np.cumsum), and stores for the half: n_diff, the
//                     candidate columns, the choice probabilities p and the first-round cdf.
//   recover_select  : weighted sampling without replacement, restating numpy's algorithm
//                     (draw size-n_uniq uniforms, searchsorted(cdf, x, 'right'), keep first
//                     occurrences, zero p[found], retry).  One warp evaluates searchsorted for a draw
//                     with two ballots (lane c holds cdf[c] and cdf[c+32]).
//       mode 0 (exact stream): ONE warp walks the rows in order and consumes the caller's PCG64
//               stream exactly as numpy would -- the stream offset of row i depends on the collisions
//               of every earlier row, a true sequential chain -- with the next row's cdf prefetched.
//               Output and final generator state are bit-identical to the seeded reference.
//       mode 1 (substreams): one warp per row, each with its own PCG64 stream derived from
//               (seed, row) -- same distribution, embarrassingly parallel, not stream-identical.
#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

typedef unsigned __int128 u128;

struct Pcg64 {
    u128 state, inc;
    __device__ __forceinline__ uint64_t next() {
        const u128 mult = ((u128)0x2360ED051FC65DA4ull << 64) | (u128)0x4385DF649FCCF645ull;
        state = state * mult + inc;
        const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
        const uint64_t x = hi ^ lo;
        const unsigned rot = (unsigned)(state >> 122);
        return (x >> rot) | (x << ((64u - rot) & 63u));
    }
    __device__ __forceinline__ double next_double() {
        return (double)(next() >> 11) * (1.0 / 9007199254740992.0);
    }
};

// _p_flip_0_to_1 (configuration_recovery.py:131-159), operation order preserved, never fused
__device__ __forceinline__ double p_flip_0_to_1(double ratio, double occ, double eps) {
    if (occ < ratio) return __ddiv_rn(__dmul_rn(occ, eps), ratio);
    if (ratio == 1.0) return eps;
    const double slope = __ddiv_rn(__dsub_rn(1.0, eps), __dsub_rn(1.0, ratio));
    const double intercept = __dsub_rn(1.0, slope);
    return __dadd_rn(__dmul_rn(occ, slope), intercept);
}
// _p_flip_1_to_0 (configuration_recovery.py:162-178)
__device__ __forceinline__ double p_flip_1_to_0(double ratio, double occ, double eps) {
    return p_flip_0_to_1(__dsub_rn(1.0, ratio), __dsub_rn(1.0, occ), eps);
}

// numpy pairwise_sum for n < 128 contiguous doubles (the order np.sum uses)
__device__ __forceinline__ double numpy_sum(const double* a, int n) {
    if (n < 8) {
        double res = -0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

struct HalfHeader {
    int32_t n_diff;  // popcount - target
    int32_t n_cand;  // 0: half skipped (no weight anywhere, or n_diff == 0); -1: numpy would raise
                     // "Fewer non-zero entries in p than size"
};

__global__ void recover_prepare_kernel(const uint64_t* __restrict__ left,
                                       const uint64_t* __restrict__ right, int64_t n, int norb,
                                       const double* __restrict__ occ_left,
                                       const double* __restrict__ occ_right, int hamming_left,
                                       int hamming_right, HalfHeader* __restrict__ hdr,
                                       uint8_t* __restrict__ cand, double* __restrict__ pc,
                                       double* __restrict__ cdf) {
    const int64_t rec = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (rec >= 2 * n) return;
    const int64_t row = rec >> 1;
    const int half = (int)(rec & 1);
    const uint64_t word = half ? right[row] : left[row];
    const double* occ = half ? occ_right : occ_left;
    const int target = half ? hamming_right : hamming_left;
    const double ratio = (double)target / (double)norb;
    const double eps = 0.01;

    double p[64];
    bool any = false;
    for (int j = 0; j < norb; ++j) {
        const bool bit = (word >> (norb - 1 - j)) & 1ull;
        double v = bit ? p_flip_1_to_0(ratio, occ[j], eps) : p_flip_0_to_1(ratio, occ[j], eps);
        v = fmin(1.0, fmax(0.0, v));  // np.minimum(1, np.maximum(0, p))
        p[j] = v;
        any |= (v != 0.0);
    }
    HalfHeader h;
    h.n_diff = popc64(word) - target;
    h.n_cand = 0;
    if (any && h.n_diff != 0) {
        const double tot = numpy_sum(p, norb);
        for (int j = 0; j < norb; ++j) p[j] = __ddiv_rn(p[j], tot);
        // candidates: occupied columns when there are too many electrons, empty ones otherwise
        const bool want = h.n_diff > 0;
        double q[64];
        int nc = 0;
        uint8_t* crow = cand + rec * norb;
        for (int j = 0; j < norb; ++j) {
            const bool bit = (word >> (norb - 1 - j)) & 1ull;
            if (bit == want) {
                crow[nc] = (uint8_t)j;
                q[nc++] = p[j];
            }
        }
        const double qs = numpy_sum(q, nc);
        double* prow = pc + rec * norb;
        double* crow_cdf = cdf + rec * norb;
        double acc = 0.0;
        for (int c = 0; c < nc; ++c) {
            q[c] = __ddiv_rn(q[c], qs);
            prow[c] = q[c];
            acc = c == 0 ? q[0] : __dadd_rn(acc, q[c]);
            crow_cdf[c] = acc;
        }
        for (int c = 0; c < nc; ++c) crow_cdf[c] = __ddiv_rn(crow_cdf[c], acc);  // cdf /= cdf[-1]
        int nz = 0;
        for (int c = 0; c < nc; ++c) nz += (q[c] > 0.0);
        const int k = h.n_diff > 0 ? h.n_diff : -h.n_diff;
        h.n_cand = nz < k ? -1 : nc;
    }
    hdr[rec] = h;
}

// Weighted choice without replacement by a full warp; every lane carries the same generator state.
// Returns the bit mask of chosen candidate ordinals.
__device__ __forceinline__ uint64_t warp_choice(Pcg64& rng, const double* __restrict__ prow, double c_lo,
                                                double c_hi, int n_cand, int k, double* s_cdf) {
    const int lane = threadIdx.x & 31;
    uint64_t found = 0ull;
    int n_uniq = 0;
    while (n_uniq < k) {
        if (n_uniq > 0) {
            // p[found] = 0; cdf = cumsum(p); cdf /= cdf[-1]   (sequential order, lane 0)
            if (lane == 0) {
                double acc = 0.0;
                for (int c = 0; c < n_cand; ++c) {
                    const double pv = ((found >> c) & 1ull) ? 0.0 : prow[c];
                    acc = c == 0 ? pv : __dadd_rn(acc, pv);
                    s_cdf[c] = acc;
                }
            }
            __syncwarp();
            const double last = s_cdf[n_cand - 1];
            c_lo = lane < n_cand ? __ddiv_rn(s_cdf[lane], last) : INFINITY;
            c_hi = lane + 32 < n_cand ? __ddiv_rn(s_cdf[lane + 32], last) : INFINITY;
            __syncwarp();
        }
        const int need = k - n_uniq;
        uint64_t round_found = 0ull;
        for (int dr = 0; dr < need; ++dr) {
            const double x = rng.next_double();
            int idx = __popc(__ballot_sync(0xffffffffu, c_lo <= x)) +
                      __popc(__ballot_sync(0xffffffffu, c_hi <= x));
            if (idx >= n_cand) idx = n_cand - 1;
            round_found |= 1ull << idx;  // first occurrence kept == set semantics for the outcome
        }
        found |= round_found;
        n_uniq = popc64(found);
    }
    return found;
}

__device__ __forceinline__ uint64_t toggle_mask(uint64_t chosen, int cand_lo, int cand_hi, int n_cand,
                                                int norb) {
    const int lane = threadIdx.x & 31;
    uint64_t t = 0ull;
    if (lane < n_cand && ((chosen >> lane) & 1ull)) t |= 1ull << (norb - 1 - cand_lo);
    if (lane + 32 < n_cand && ((chosen >> (lane + 32)) & 1ull)) t |= 1ull << (norb - 1 - cand_hi);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t |= __shfl_xor_sync(0xffffffffu, t, o);
    return t;
}

// mode 0: one warp, rows in order, caller's PCG64 stream
__global__ void __launch_bounds__(32)
recover_select_chain_kernel(const uint64_t* __restrict__ left, const uint64_t* __restrict__ right,
                            int64_t n, int norb, const HalfHeader* __restrict__ hdr,
                            const uint8_t* __restrict__ cand, const double* __restrict__ pc,
                            const double* __restrict__ cdf, uint64_t* __restrict__ rng_state,
                            uint64_t* __restrict__ left_out, uint64_t* __restrict__ right_out,
                            int32_t* __restrict__ status) {
    __shared__ double s_cdf[64];
    const int lane = threadIdx.x & 31;
    Pcg64 rng;
    rng.state = ((u128)rng_state[0] << 64) | (u128)rng_state[1];
    rng.inc = ((u128)rng_state[2] << 64) | (u128)rng_state[3];
    struct Rec {
        HalfHeader h;
        double c_lo, c_hi;
        int cand_lo, cand_hi;
        uint64_t word;
    };
    auto fetch = [&](int64_t rec) {
        Rec r;
        r.h = hdr[rec];
        const double* crow = cdf + rec * norb;
        const uint8_t* krow = cand + rec * norb;
        r.c_lo = lane < r.h.n_cand ? crow[lane] : INFINITY;
        r.c_hi = lane + 32 < r.h.n_cand ? crow[lane + 32] : INFINITY;
        r.cand_lo = lane < r.h.n_cand ? krow[lane] : 0;
        r.cand_hi = lane + 32 < r.h.n_cand ? krow[lane + 32] : 0;
        r.word = (rec & 1) ? right[rec >> 1] : left[rec >> 1];
        return r;
    };
    bool failed = false;
    Rec nxt = fetch(0);
    uint64_t word_l = 0ull;
    for (int64_t rec = 0; rec < 2 * n; ++rec) {
        const Rec cur = nxt;
        if (rec + 1 < 2 * n) nxt = fetch(rec + 1);  // prefetch: independent of the random stream
        const int64_t row = rec >> 1;
        uint64_t word = cur.word;
        if (!failed) {
            if (cur.h.n_cand < 0) {
                failed = true;  // numpy raises here; the host re-raises, nothing after this row counts
                if (lane == 0) {
                    status[0] = 1;
                    status[1] = (int32_t)row;
                }
            } else if (cur.h.n_cand > 0) {
                const int k = cur.h.n_diff > 0 ? cur.h.n_diff : -cur.h.n_diff;
                const uint64_t chosen =
                    warp_choice(rng, pc + rec * norb, cur.c_lo, cur.c_hi, cur.h.n_cand, k, s_cdf);
                word ^= toggle_mask(chosen, cur.cand_lo, cur.cand_hi, cur.h.n_cand, norb);
            }
        }
        if ((rec & 1) == 0) {
            word_l = word;
        } else if (lane == 0) {
            left_out[row] = word_l;
            right_out[row] = word;
        }
    }
    if (lane == 0) {
        rng_state[0] = (uint64_t)(rng.state >> 64);
        rng_state[1] = (uint64_t)rng.state;
    }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// mode 1: one warp per row, independent PCG64 substream per row
__global__ void recover_select_parallel_kernel(const uint64_t* __restrict__ left,
                                               const uint64_t* __restrict__ right, int64_t n, int norb,
                                               const HalfHeader* __restrict__ hdr,
                                               const uint8_t* __restrict__ cand,
                                               const double* __restrict__ pc,
                                               const double* __restrict__ cdf, uint64_t seed,
                                               uint64_t* __restrict__ left_out,
                                               uint64_t* __restrict__ right_out,
                                               int32_t* __restrict__ status) {
    __shared__ double s_cdf_all[8][64];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= n) return;
    uint64_t sm = seed ^ (0xD1B54A32D192ED03ull * (uint64_t)(row + 1));
    Pcg64 rng;
    rng.state = ((u128)splitmix64(sm) << 64) | (u128)splitmix64(sm);
    rng.inc = (((u128)splitmix64(sm) << 64) | (u128)splitmix64(sm)) | (u128)1;
    rng.next();
    uint64_t words[2] = {left[row], right[row]};
    for (int half = 0; half < 2; ++half) {
        const int64_t rec = 2 * row + half;
        const HalfHeader h = hdr[rec];
        if (h.n_cand < 0) {
            if (lane == 0) {
                atomicExch(&status[0], 1);
                atomicMin(&status[1], (int32_t)row);
            }
            continue;
        }
        if (h.n_cand == 0) continue;
        const double* crow = cdf + rec * norb;
        const uint8_t* krow = cand + rec * norb;
        const double c_lo = lane < h.n_cand ? crow[lane] : INFINITY;
        const double c_hi = lane + 32 < h.n_cand ? crow[lane + 32] : INFINITY;
        const int cand_lo = lane < h.n_cand ? krow[lane] : 0;
        const int cand_hi = lane + 32 < h.n_cand ? krow[lane + 32] : 0;
        const int k = h.n_diff > 0 ? h.n_diff : -h.n_diff;
        const uint64_t chosen =
            warp_choice(rng, pc + rec * norb, c_lo, c_hi, h.n_cand, k, s_cdf_all[warp]);
        words[half] ^= toggle_mask(chosen, cand_lo, cand_hi, h.n_cand, norb);
    }
    if (lane == 0) {
        left_out[row] = words[0];
        right_out[row] = words[1];
    }
}

}  // namespace sqd

using namespace sqd;

extern "C" {

int64_t sqd_recover_workspace_bytes(int64_t n, int norb) {
    if (n < 0 || norb < 1 || norb > 64) return -1;
    const int64_t recs = 2 * n;
    int64_t b = 0;
    b += ((recs * (int64_t)sizeof(HalfHeader) + 255) / 256) * 256;
    b += ((recs * norb + 255) / 256) * 256;                             // cand (u8)
    b += ((recs * norb * (int64_t)sizeof(double) + 255) / 256) * 256;  // p
    b += ((recs * norb * (int64_t)sizeof(double) + 255) / 256) * 256;  // cdf
    return b + 256;
}

int sqd_recover(const uint64_t* d_left, const uint64_t* d_right, int64_t n, int norb,
                const double* d_occ_left, const double* d_occ_right, int hamming_left,
                int hamming_right, int mode, uint64_t* d_rng_state, uint64_t seed,
                uint64_t* d_left_out, uint64_t* d_right_out, int32_t* d_status, void* d_workspace,
                int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(norb >= 1 && norb <= 64, "sqd_recover: norb=%d must be in [1, 64]", norb);
    SQD_REQUIRE(hamming_left >= 0 && hamming_right >= 0,
                "The numbers of electrons must be specified as non-negative integers.");
    SQD_REQUIRE(workspace_bytes >= sqd_recover_workspace_bytes(n, norb), "sqd_recover: workspace too small");
    SQD_REQUIRE(mode == 0 || mode == 1, "sqd_recover: mode must be 0 (exact stream) or 1 (substreams)");
    const int32_t init[2] = {0, 0x7fffffff};
    SQD_CUDA_OK(cudaMemcpyAsync(d_status, init, sizeof(init), cudaMemcpyHostToDevice, st));
    if (n == 0) return 0;
    const int64_t recs = 2 * n;
    char* p = (char*)d_workspace;
    HalfHeader* hdr = (HalfHeader*)p;
    p += ((recs * (int64_t)sizeof(HalfHeader) + 255) / 256) * 256;
    uint8_t* cand = (uint8_t*)p;
    p += ((recs * norb + 255) / 256) * 256;
    double* pc = (double*)p;
    p += ((recs * norb * (int64_t)sizeof(double) + 255) / 256) * 256;
    double* cdf = (double*)p;
    recover_prepare_kernel<<<(unsigned)((recs + 127) / 128), 128, 0, st>>>(
        d_left, d_right, n, norb, d_occ_left, d_occ_right, hamming_left, hamming_right, hdr, cand, pc,
        cdf);
    if (check_launch("recover_prepare_kernel")) return -2;
    if (mode == 0) {
        SQD_REQUIRE(d_rng_state != nullptr, "sqd_recover: mode 0 needs the PCG64 state");
        recover_select_chain_kernel<<<1, 32, 0, st>>>(d_left, d_right, n, norb, hdr, cand, pc, cdf,
                                                      d_rng_state, d_left_out, d_right_out, d_status);
        return check_launch("recover_select_chain_kernel");
    }
    recover_select_parallel_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(
        d_left, d_right, n, norb, hdr, cand, pc, cdf, seed, d_left_out, d_right_out, d_status);
    return check_launch("recover_select_parallel_kernel");
}

}  // extern "C"

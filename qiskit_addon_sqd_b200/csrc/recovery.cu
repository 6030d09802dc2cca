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
//       mode 0 (exact stream): the caller's PCG64 stream is consumed exactly as numpy would.  The
//               stream offset of half-row i depends on the collisions of every earlier half-row -- a
//               true sequential chain -- but the chain only has to carry ONE small integer (how many
//               draws beyond the collision-free count have been consumed so far).  So, per window of
//               kWinRecs half-rows (see "speculative exact stream" below): a parallel kernel tabulates,
//               for every half-row and for each of kHyp hypotheses about that integer, how many draws
//               the half-row would consume (PCG64 jump-ahead gives every thread its uniforms); one warp
//               then walks the window with one shared-memory lookup per half-row; a last parallel pass
//               redoes the sampling of every half-row at its now known offset.  Output and final
//               generator state are bit-identical to the seeded reference.  (The round-1 kernel, one
//               warp sampling the rows one after the other, is kept behind SQD_RECOVER_CHAIN=1.)
//       mode 1 (substreams): one warp per row, each with its own PCG64 stream derived from
//               (seed, row) -- same distribution, embarrassingly parallel, not stream-identical.
#include <stdlib.h>

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


// ---------------------------------------------------------------------------------------------
// speculative exact stream
// ---------------------------------------------------------------------------------------------
constexpr int kHyp = 256;        // hypotheses per half-row (threads of a table CTA)
constexpr int kWinRecs = 2048;   // half-rows per window
constexpr int kMaxDraws = 96;    // draws a tabulated hypothesis may consume (more: the walk samples it itself)
constexpr int kUBuf = kHyp + kMaxDraws;
constexpr uint8_t kTabOverflow = 255;   // hypothesis not tabulated: the walk samples the half-row itself
constexpr uint8_t kTabFail = 254;       // numpy raises at this half-row
constexpr int kChunksPerWin = kWinRecs / 32;
constexpr int16_t kCompSpecial = 0x7FFF;   // chunk not composable for this hypothesis: the walk steps through it

__device__ __forceinline__ u128 pcg_mult() {
    return ((u128)0x2360ED051FC65DA4ull << 64) | (u128)0x4385DF649FCCF645ull;
}
__device__ __forceinline__ uint64_t pcg_output(u128 state) {
    const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
    const uint64_t x = hi ^ lo;
    const unsigned rot = (unsigned)(state >> 122);
    return (x >> rot) | (x << ((64u - rot) & 63u));
}
// LCG skip-ahead: (mult, plus) with state_after_delta_steps = mult * state + plus
__device__ __forceinline__ void pcg_skip(u128 inc, uint64_t delta, u128* m_out, u128* p_out) {
    u128 acc_m = 1, acc_p = 0, cur_m = pcg_mult(), cur_p = inc;
    while (delta) {
        if (delta & 1ull) {
            acc_m *= cur_m;
            acc_p = acc_p * cur_m + cur_p;
        }
        cur_p = (cur_m + 1) * cur_p;
        cur_m *= cur_m;
        delta >>= 1;
    }
    *m_out = acc_m;
    *p_out = acc_p;
}
__device__ __forceinline__ u128 pcg_advance(u128 state, u128 inc, uint64_t delta) {
    u128 m, p;
    pcg_skip(inc, delta, &m, &p);
    return m * state + p;
}

// numpy's Generator.choice(replace=False, p=...) on one thread.  draw() returns the next uniform double of
// the stream; cdf0 is the normalised first-round cdf, prow the probabilities.  Returns the number of draws
// consumed (-1: more than max_draws would be needed) and the set of chosen candidate ordinals.
template <class Draw>
__device__ __forceinline__ int thread_choice(Draw& draw, const double* prow, const double* cdf0, int n_cand, int k,
                                             int max_draws, uint64_t* found_out) {
    uint64_t found = 0ull;
    int n_uniq = 0, used = 0;
    bool first = true;
    while (n_uniq < k) {
        const int need = k - n_uniq;
        if (used + need > max_draws) return -1;
        if (first) {
            for (int dr = 0; dr < need; ++dr) {
                const double x = draw();
                int lo = 0, hi = n_cand;   // searchsorted(cdf, x, 'right') = number of entries <= x
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (cdf0[mid] <= x) lo = mid + 1; else hi = mid;
                }
                if (lo >= n_cand) lo = n_cand - 1;
                found |= 1ull << lo;
            }
        } else {
            // p[found] = 0; cdf = cumsum(p); cdf /= cdf[-1]  (sequential order)
            double last = 0.0;
            for (int c = 0; c < n_cand; ++c) {
                const double pv = ((found >> c) & 1ull) ? 0.0 : prow[c];
                last = c == 0 ? pv : __dadd_rn(last, pv);
            }
            uint64_t round_found = 0ull;
            for (int dr = 0; dr < need; ++dr) {
                const double x = draw();
                double acc = 0.0;
                int idx = 0;
                for (int c = 0; c < n_cand; ++c) {
                    const double pv = ((found >> c) & 1ull) ? 0.0 : prow[c];
                    acc = c == 0 ? pv : __dadd_rn(acc, pv);
                    idx += (__ddiv_rn(acc, last) <= x) ? 1 : 0;
                }
                if (idx >= n_cand) idx = n_cand - 1;
                round_found |= 1ull << idx;
            }
            found |= round_found;
        }
        used += need;
        n_uniq = popc64(found);
        first = false;
    }
    *found_out = found;
    return used;
}

// Walk state of the exact stream (device memory; the host only reads `done`)
struct RecoverCtl {
    uint64_t state_hi, state_lo;   // generator state at the start of the current window
    int64_t cursor;                // first half-row of the current window
    int64_t offset;                // draws consumed before `cursor`
    int64_t extras_total;          // draws beyond the collision-free count so far (drift estimate)
    uint32_t rho;                  // expected extra draws per half-row, 16.16 fixed point
    int32_t done;                  // 1: all half-rows walked (or numpy's error met)
    int32_t failed;
    int32_t pad;
    // the window the walk has just finished, for the replay kernel that records the per-row offsets
    int64_t w_r0, w_off0;
    uint32_t w_rho;
    int32_t w_rows;
};

__device__ __forceinline__ int drift_of(uint32_t rho, int j) { return (int)(((uint64_t)rho * (uint64_t)j) >> 16); }

// draws of a half-row when nothing collides (0: nothing to sample; 255: numpy raises)
__global__ void recover_counts_kernel(const HalfHeader* __restrict__ hdr, int64_t recs, uint8_t* __restrict__ kq,
                                      int32_t* __restrict__ kcount) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= recs) return;
    const HalfHeader h = hdr[r];
    int k = 0;
    uint8_t q = 0;
    if (h.n_cand < 0) q = 255;
    else if (h.n_cand > 0) { k = h.n_diff > 0 ? h.n_diff : -h.n_diff; q = (uint8_t)k; }
    kq[r] = q;
    kcount[r] = k;
}

// skip tables: state after t+1 steps = M[t] * base + P[t]
__global__ void recover_skip_table_kernel(const uint64_t* __restrict__ rng_state, u128* __restrict__ skipM,
                                          u128* __restrict__ skipP, RecoverCtl* ctl,
                                          uint64_t* __restrict__ init_state) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const u128 inc = ((u128)rng_state[2] << 64) | (u128)rng_state[3];
    if (t < kUBuf) pcg_skip(inc, (uint64_t)t + 1, skipM + t, skipP + t);
    if (t == 0) {
        init_state[0] = rng_state[0];
        init_state[1] = rng_state[1];
        ctl->state_hi = rng_state[0];
        ctl->state_lo = rng_state[1];
        ctl->cursor = 0;
        ctl->offset = 0;
        ctl->extras_total = 0;
        ctl->rho = 0;
        ctl->done = 0;
        ctl->failed = 0;
    }
}

// One CTA per half-row of the window, one thread per hypothesis e: "the walk arrives here having consumed
// kprefix + drift_j + (e - kHyp/2) draws since the window start".  table[j][e] = extra draws (beyond k) the
// half-row then consumes, or kTabOverflow / kTabFail.
__global__ void __launch_bounds__(kHyp)
recover_table_kernel(const RecoverCtl* __restrict__ ctl, int64_t recs, int norb, const HalfHeader* __restrict__ hdr,
                     const double* __restrict__ pc, const double* __restrict__ cdf,
                     const int32_t* __restrict__ kprefix, const uint64_t* __restrict__ rng_state,
                     const u128* __restrict__ skipM, const u128* __restrict__ skipP, uint8_t* __restrict__ table) {
    if (ctl->done) return;
    const int j = blockIdx.x;
    const int64_t r0 = ctl->cursor;
    const int64_t rec = r0 + j;
    if (rec >= recs) return;
    const HalfHeader h = hdr[rec];
    const int e = threadIdx.x;
    const int cj = drift_of(ctl->rho, j);
    uint8_t* trow = table + (int64_t)j * kHyp;
    if (h.n_cand <= 0) {   // nothing to sample (or numpy's error): no draws
        trow[e] = h.n_cand < 0 ? kTabFail : (uint8_t)0;
        return;
    }
    __shared__ double s_cdf[64], s_p[64], s_u[kUBuf];
    __shared__ u128 s_base;
    const int e_lo = cj >= kHyp / 2 ? 0 : kHyp / 2 - cj;
    if (e < h.n_cand) {
        s_cdf[e] = cdf[rec * norb + e];
        s_p[e] = pc[rec * norb + e];
    }
    if (e == 0) {
        const u128 inc = ((u128)rng_state[2] << 64) | (u128)rng_state[3];
        const u128 st = ((u128)ctl->state_hi << 64) | (u128)ctl->state_lo;
        const uint64_t delta = (uint64_t)((int64_t)(kprefix[rec] - kprefix[r0]) + cj + e_lo - kHyp / 2);
        s_base = pcg_advance(st, inc, delta);
    }
    __syncthreads();
    const u128 base = s_base;
    for (int t = e; t < kUBuf; t += kHyp)
        s_u[t] = (double)(pcg_output(skipM[t] * base + skipP[t]) >> 11) * (1.0 / 9007199254740992.0);
    __syncthreads();
    uint8_t out = kTabOverflow;
    if (e >= e_lo) {
        const int k = h.n_diff > 0 ? h.n_diff : -h.n_diff;
        const double* up = s_u + (e - e_lo);
        int pos = 0;
        auto draw = [&]() { return up[pos++]; };
        uint64_t found;
        const int used = thread_choice(draw, s_p, s_cdf, h.n_cand, k, kMaxDraws, &found);
        if (used >= 0 && used - k < kTabFail) out = (uint8_t)(used - k);
    }
    trow[e] = out;
}

// Composition of the 32 transitions of a chunk, for all hypotheses at once: comp[chunk][e] = the index the
// walk holds after the chunk's last half-row when it enters the chunk with e (may lie outside [0, kHyp): the
// window then ends at the chunk boundary), or kCompSpecial when the path meets numpy's error, an untabulated
// hypothesis or leaves the band inside the chunk.  One CTA per chunk, table slice in shared memory.
__global__ void __launch_bounds__(kHyp)
recover_compose_kernel(const RecoverCtl* __restrict__ ctl, int64_t recs, const uint8_t* __restrict__ table,
                       int16_t* __restrict__ comp) {
    if (ctl->done) return;
    __shared__ __align__(16) uint8_t s_tab[32 * kHyp];
    __shared__ int s_dc[32];
    const int chunk = blockIdx.x, e0 = threadIdx.x;
    const int64_t r0 = ctl->cursor;
    const uint32_t rho = ctl->rho;
    const int nwin = (int)((recs - r0) < kWinRecs ? (recs - r0) : kWinRecs);
    const int rows = nwin - chunk * 32 < 32 ? nwin - chunk * 32 : 32;
    if (rows <= 0) return;
    const uint4* src = reinterpret_cast<const uint4*>(table + (int64_t)chunk * 32 * kHyp);
    for (int i = e0; i < rows * (kHyp / 16); i += kHyp) reinterpret_cast<uint4*>(s_tab)[i] = src[i];
    if (e0 < 32) s_dc[e0] = drift_of(rho, chunk * 32 + e0 + 1) - drift_of(rho, chunk * 32 + e0);
    __syncthreads();
    int e = e0;
    bool special = false;
    for (int jj = 0; jj < rows; ++jj) {
        const int t = s_tab[jj * kHyp + e];
        if (t >= kTabFail) {
            special = true;
            break;
        }
        e += t - s_dc[jj];
        if ((unsigned)e >= (unsigned)kHyp && jj + 1 < rows) {
            special = true;
            break;
        }
    }
    comp[chunk * kHyp + e0] = special ? kCompSpecial : (int16_t)e;
}

// The sequential part: one warp, ONE shared-memory lookup per CHUNK of 32 half-rows (the composed transitions);
// chunks marked special for the current hypothesis are stepped through row by row.  Ends the window when the
// hypothesis index leaves [0, kHyp) and re-bases.  e_start[chunk] = index at the chunk's first row, for the
// replay kernel (-1: the walk recorded the offsets of that chunk itself, -2: not walked).
__global__ void __launch_bounds__(32)
recover_walk_kernel(RecoverCtl* ctl, int64_t recs, int norb, const HalfHeader* __restrict__ hdr,
                    const double* __restrict__ pc, const double* __restrict__ cdf,
                    const int32_t* __restrict__ kprefix, const uint8_t* __restrict__ table,
                    const int16_t* __restrict__ comp, int16_t* __restrict__ e_start,
                    uint64_t* __restrict__ rng_state, int64_t* __restrict__ off_abs, int32_t* __restrict__ status) {
    if (ctl->done) return;
    __shared__ __align__(16) int16_t s_comp[kChunksPerWin * kHyp];
    __shared__ double s_cdf[64];
    const int lane = threadIdx.x & 31;
    const int64_t r0 = ctl->cursor;
    const uint32_t rho = ctl->rho;
    const u128 inc = ((u128)rng_state[2] << 64) | (u128)rng_state[3];
    const u128 win_state = ((u128)ctl->state_hi << 64) | (u128)ctl->state_lo;
    const int64_t off0 = ctl->offset;
    const int nwin = (int)((recs - r0) < kWinRecs ? (recs - r0) : kWinRecs);
    const int nchunks = (nwin + 31) / 32;
    for (int i = lane; i < nchunks * (kHyp / 8); i += 32)
        reinterpret_cast<uint4*>(s_comp)[i] = reinterpret_cast<const uint4*>(comp)[i];
    for (int c = lane; c < kChunksPerWin; c += 32) e_start[c] = -2;
    __syncwarp();
    int e = kHyp / 2;        // hypothesis index of the current half-row
    int j = 0;               // half-rows walked
    bool failed = false, out_of_band = false;
    for (int chunk = 0; chunk < nchunks && !out_of_band && !failed; ++chunk) {
        const int rows = nwin - chunk * 32 < 32 ? nwin - chunk * 32 : 32;
        const int v = s_comp[chunk * kHyp + e];
        if (v != kCompSpecial) {
            if (lane == 0) e_start[chunk] = (int16_t)e;
            e = v;
            j += rows;
            if ((unsigned)e >= (unsigned)kHyp) out_of_band = true;
            continue;
        }
        // special chunk: row by row, table entries straight from global memory
        if (lane == 0) e_start[chunk] = -1;
        int e_mine = 0, done_rows = 0;
        for (int jj = 0; jj < rows; ++jj) {
            if (lane == jj) e_mine = e;
            done_rows = jj + 1;
            const int jg = chunk * 32 + jj;
            const int64_t rec = r0 + jg;
            int extra = table[(int64_t)jg * kHyp + e];
            if (extra >= kTabFail) {
                if (extra == kTabFail) {   // numpy raises here: the generator has consumed everything before it
                    failed = true;
                    if (lane == 0) {
                        status[0] = 1;
                        status[1] = (int32_t)(rec >> 1);
                    }
                    break;
                }
                // not tabulated (many retries): sample this half-row here, as the round-1 kernel did
                const HalfHeader h = hdr[rec];
                const int k = h.n_diff > 0 ? h.n_diff : -h.n_diff;
                const int64_t off_rel = (int64_t)(kprefix[rec] - kprefix[r0]) + drift_of(rho, jg) + e - kHyp / 2;
                Pcg64 rng;
                rng.inc = inc;
                rng.state = pcg_advance(win_state, inc, (uint64_t)off_rel);
                const double* crow = cdf + rec * norb;
                const double c_lo = lane < h.n_cand ? crow[lane] : INFINITY;
                const double c_hi = lane + 32 < h.n_cand ? crow[lane + 32] : INFINITY;
                const u128 before = rng.state;
                warp_choice(rng, pc + rec * norb, c_lo, c_hi, h.n_cand, k, s_cdf);
                Pcg64 probe;   // draws used = steps between the two states (rare path: count them)
                probe.inc = inc;
                probe.state = before;
                int used = 0;
                while (probe.state != rng.state) {
                    probe.next();
                    ++used;
                }
                extra = used - k;
            }
            e += extra - (drift_of(rho, jg + 1) - drift_of(rho, jg));
            ++j;
            if ((unsigned)e >= (unsigned)kHyp) {
                out_of_band = true;
                break;
            }
        }
        if (lane < done_rows) {
            const int jg = chunk * 32 + lane;
            const int64_t rec = r0 + jg;
            off_abs[rec] = off0 + (int64_t)(kprefix[rec] - kprefix[r0]) + drift_of(rho, jg) + e_mine - kHyp / 2;
        }
        __syncwarp();
    }
    if (lane == 0) {
        ctl->w_r0 = r0;
        ctl->w_off0 = off0;
        ctl->w_rho = rho;
        ctl->w_rows = j;
    }
    // j half-rows walked (a failing one not counted); e is the hypothesis index AT half-row r0 + j in every
    // exit path, i.e. draws consumed beyond the collision-free count = drift_j + e - kHyp/2
    const int64_t next = r0 + j;
    const int64_t extras_win = (int64_t)drift_of(rho, j) + e - kHyp / 2;
    const int64_t off_next = off0 + (int64_t)(kprefix[next] - kprefix[r0]) + extras_win;
    if (lane == 0) {
        const u128 st = pcg_advance(win_state, inc, (uint64_t)(off_next - off0));
        ctl->state_hi = (uint64_t)(st >> 64);
        ctl->state_lo = (uint64_t)st;
        ctl->cursor = next;
        ctl->offset = off_next;
        ctl->extras_total += extras_win;
        if (next > 0) {
            const double r = (double)ctl->extras_total / (double)next;
            ctl->rho = (uint32_t)(r * 65536.0 + 0.5);
        }
        if (failed || next >= recs) {
            ctl->done = 1;
            ctl->failed = failed ? 1 : 0;
            rng_state[0] = (uint64_t)(st >> 64);   // the caller's generator continues from here
            rng_state[1] = (uint64_t)st;
        }
    }
}

// Stream offsets of the half-rows of the chunks the walk crossed in one lookup: one warp per chunk repeats the
// chunk's 32 transitions from the recorded entry index (64 chunks in parallel, table slice in shared memory).
__global__ void __launch_bounds__(32)
recover_replay_kernel(const RecoverCtl* __restrict__ ctl, const int32_t* __restrict__ kprefix,
                      const uint8_t* __restrict__ table, const int16_t* __restrict__ e_start,
                      int64_t* __restrict__ off_abs) {
    __shared__ __align__(16) uint8_t s_tab[32 * kHyp];
    const int chunk = blockIdx.x, lane = threadIdx.x & 31;
    const int es = e_start[chunk];
    const int rows_walked = ctl->w_rows;
    if (es < 0 || chunk * 32 >= rows_walked) return;
    const int64_t r0 = ctl->w_r0, off0 = ctl->w_off0;
    const uint32_t rho = ctl->w_rho;
    const int rows = rows_walked - chunk * 32 < 32 ? rows_walked - chunk * 32 : 32;
    const uint4* src = reinterpret_cast<const uint4*>(table + (int64_t)chunk * 32 * kHyp);
    for (int i = lane; i < rows * (kHyp / 16); i += 32) reinterpret_cast<uint4*>(s_tab)[i] = src[i];
    __syncwarp();
    int e = es, e_mine = 0;
    for (int jj = 0; jj < rows; ++jj) {
        if (lane == jj) e_mine = e;
        const int jg = chunk * 32 + jj;
        e += (int)s_tab[jj * kHyp + e] - (drift_of(rho, jg + 1) - drift_of(rho, jg));
    }
    if (lane < rows) {
        const int jg = chunk * 32 + lane;
        const int64_t rec = r0 + jg;
        off_abs[rec] = off0 + (int64_t)(kprefix[rec] - kprefix[r0]) + drift_of(rho, jg) + e_mine - kHyp / 2;
    }
}

// Every half-row again, now at its known stream offset: the flips themselves.  One thread per half-row.
__global__ void recover_apply_kernel(const RecoverCtl* __restrict__ ctl, const uint64_t* __restrict__ left,
                                     const uint64_t* __restrict__ right, int64_t recs, int norb,
                                     const HalfHeader* __restrict__ hdr, const uint8_t* __restrict__ cand,
                                     const double* __restrict__ pc, const double* __restrict__ cdf,
                                     const int64_t* __restrict__ off_abs, const uint64_t* __restrict__ init_state,
                                     const uint64_t* __restrict__ rng_state, uint64_t* __restrict__ left_out,
                                     uint64_t* __restrict__ right_out) {
    const int64_t rec = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (rec >= recs) return;
    uint64_t word = (rec & 1) ? right[rec >> 1] : left[rec >> 1];
    const HalfHeader h = hdr[rec];
    const bool live = !(ctl->failed && rec >= ctl->cursor);   // nothing after numpy's error counts
    if (live && h.n_cand > 0) {
        const int k = h.n_diff > 0 ? h.n_diff : -h.n_diff;
        Pcg64 rng;
        rng.inc = ((u128)rng_state[2] << 64) | (u128)rng_state[3];
        rng.state = pcg_advance(((u128)init_state[0] << 64) | (u128)init_state[1], rng.inc,
                                (uint64_t)off_abs[rec]);
        auto draw = [&]() { return rng.next_double(); };
        uint64_t found = 0ull;
        thread_choice(draw, pc + rec * norb, cdf + rec * norb, h.n_cand, k, 1 << 30, &found);
        const uint8_t* krow = cand + rec * norb;
        for (int c = 0; c < h.n_cand; ++c)
            if ((found >> c) & 1ull) word ^= 1ull << (norb - 1 - krow[c]);
    }
    if (rec & 1) right_out[rec >> 1] = word; else left_out[rec >> 1] = word;
}


// ---------------------------------------------------------------------------------------------
// duplicate merge (configuration_recovery.py:112-126): distinct repaired rows in first-seen order, the
// probabilities of equal rows added IN INPUT ORDER (the reference's row-by-row `freqs[idx] += p`)
// ---------------------------------------------------------------------------------------------
constexpr int32_t kMergeEmpty = 0x7fffffff;

__global__ void recover_fill_i32_kernel(int32_t* __restrict__ a, int64_t n, int32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

__device__ __forceinline__ uint64_t merge_hash(uint64_t lo, uint64_t ro) {
    uint64_t h = lo * 0x9E3779B97F4A7C15ull ^ (ro + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 32;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 29;
    return h;
}

// open addressing over the rows themselves: a slot holds the index of a representative row; equal rows meet
// in the same slot, which remembers their smallest index (first occurrence) and their number
__global__ void merge_insert_kernel(const uint64_t* __restrict__ lo, const uint64_t* __restrict__ ro, int n,
                                    int32_t* slot_rep, int32_t* slot_min, int32_t* slot_cnt,
                                    int32_t* __restrict__ row_slot, uint32_t mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t l = lo[i], r = ro[i];
    uint32_t s = (uint32_t)merge_hash(l, r) & mask;
    for (;;) {
        int32_t cur = *((volatile int32_t*)(slot_rep + s));
        if (cur == kMergeEmpty) {
            const int32_t old = atomicCAS(slot_rep + s, kMergeEmpty, (int32_t)i);
            cur = old == kMergeEmpty ? i : old;
        }
        if (lo[cur] == l && ro[cur] == r) {
            row_slot[i] = (int32_t)s;
            atomicMin(slot_min + s, (int32_t)i);
            atomicAdd(slot_cnt + s, 1);
            return;
        }
        s = (s + 1) & mask;
    }
}

__global__ void merge_flag_kernel(int n, const int32_t* __restrict__ row_slot, const int32_t* __restrict__ slot_min,
                                  int32_t* __restrict__ is_first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) is_first[i] = slot_min[row_slot[i]] == i ? 1 : 0;
}

// first occurrences, in input order = output order: rank r = number of first occurrences before row i
__global__ void merge_groups_kernel(int n, const uint64_t* __restrict__ lo, const uint64_t* __restrict__ ro,
                                    const int32_t* __restrict__ row_slot, const int32_t* __restrict__ is_first,
                                    const int32_t* __restrict__ rank, const int32_t* __restrict__ slot_cnt,
                                    int32_t* __restrict__ slot_rank, int32_t* __restrict__ gcnt,
                                    uint64_t* __restrict__ out_lo, uint64_t* __restrict__ out_ro) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !is_first[i]) return;
    const int r = rank[i], s = row_slot[i];
    slot_rank[s] = r;
    gcnt[r] = slot_cnt[s];
    out_lo[r] = lo[i];
    out_ro[r] = ro[i];
}

__global__ void merge_scatter_kernel(int n, const int32_t* __restrict__ row_slot, const int32_t* __restrict__ slot_rank,
                                     const int32_t* __restrict__ goff, int32_t* gcur, int32_t* __restrict__ members) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int r = slot_rank[row_slot[i]];
    members[goff[r] + atomicAdd(gcur + r, 1)] = i;   // arrival order: sorted by the summing kernel
}

// one warp per group: members sorted ascending (rank by counting), then ONE lane adds the probabilities in
// that order, starting from 0.0 -- np.bincount's / the reference loop's floating-point sums
__global__ void merge_sum_kernel(int n_groups, const int32_t* __restrict__ goff, const int32_t* __restrict__ gcnt,
                                 const int32_t* __restrict__ members, int32_t* __restrict__ sorted,
                                 const double* __restrict__ prob, double* __restrict__ sums) {
    __shared__ double s_p[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = gridDim.x * (blockDim.x >> 5);
    for (int g = blockIdx.x * (blockDim.x >> 5) + warp; g < n_groups; g += nwarp) {
        const int off = goff[g], L = gcnt[g];
        double acc = 0.0;
        if (L <= 32) {
            const int mine = lane < L ? members[off + lane] : 0x7fffffff;
            int rk = 0;
            for (int l = 0; l < L; ++l) rk += __shfl_sync(0xffffffffu, mine, l) < mine;
            if (lane < L) s_p[warp][rk] = prob[mine];
            __syncwarp();
            if (lane == 0)
                for (int k = 0; k < L; ++k) acc = __dadd_rn(acc, s_p[warp][k]);
            __syncwarp();
        } else {
            for (int k = lane; k < L; k += 32) {
                const int x = members[off + k];
                int rk = 0;
                for (int l = 0; l < L; ++l) rk += members[off + l] < x;
                sorted[off + rk] = x;
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0)
                for (int k = 0; k < L; ++k) acc = __dadd_rn(acc, prob[sorted[off + k]]);
        }
        if (lane == 0) sums[g] = acc;
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
    // speculative exact stream: kq, kcount, kprefix, off_abs, window table, skip tables, walk state
    b += ((recs + 255) / 256) * 256;
    b += 2 * (((recs + 1) * 4 + 255) / 256) * 256;
    b += ((recs * 8 + 255) / 256) * 256;
    b += (int64_t)kWinRecs * kHyp + 2 * (int64_t)kUBuf * 16 + 512;
    b += (int64_t)kChunksPerWin * kHyp * 2 + 256;   // composed transitions + chunk entry indices
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
    p += ((recs * norb * (int64_t)sizeof(double) + 255) / 256) * 256;
    if (mode == 0) {
        SQD_REQUIRE(d_rng_state != nullptr, "sqd_recover: mode 0 needs the PCG64 state");
        static const int knob_chain = getenv("SQD_RECOVER_CHAIN") ? atoi(getenv("SQD_RECOVER_CHAIN")) : 0;
        const bool speculative = !knob_chain && recs * 64 < 2147483647LL;
        if (!speculative) {
            recover_select_chain_kernel<<<1, 32, 0, st>>>(d_left, d_right, n, norb, hdr, cand, pc, cdf,
                                                          d_rng_state, d_left_out, d_right_out, d_status);
            return check_launch("recover_select_chain_kernel");
        }
        uint8_t* kq = (uint8_t*)p;
        p += ((recs + 255) / 256) * 256;
        int32_t* kcount = (int32_t*)p;
        p += (((recs + 1) * 4 + 255) / 256) * 256;
        int32_t* kprefix = (int32_t*)p;
        p += (((recs + 1) * 4 + 255) / 256) * 256;
        int64_t* off_abs = (int64_t*)p;
        p += ((recs * 8 + 255) / 256) * 256;
        uint8_t* table = (uint8_t*)p;
        p += (int64_t)kWinRecs * kHyp;
        u128* skipM = (u128*)p;
        p += (int64_t)kUBuf * 16;
        u128* skipP = (u128*)p;
        p += (int64_t)kUBuf * 16;
        RecoverCtl* ctl = (RecoverCtl*)p;
        uint64_t* init_state = (uint64_t*)(p + 256);
        int16_t* comp = (int16_t*)(p + 512);
        int16_t* e_start = comp + (int64_t)kChunksPerWin * kHyp;
        recover_counts_kernel<<<(unsigned)((recs + 255) / 256), 256, 0, st>>>(hdr, recs, kq, kcount);
        if (check_launch("recover_counts_kernel")) return -2;
        if (sqd_exclusive_scan(kcount, kprefix, (int)recs, nullptr, stream)) return -2;
        recover_skip_table_kernel<<<(kUBuf + 127) / 128, 128, 0, st>>>(d_rng_state, skipM, skipP, ctl, init_state);
        if (check_launch("recover_skip_table_kernel")) return -2;
        // a window ends early when the walk drifts out of its band of hypotheses: launch the minimum number of
        // windows, then look at the done flag every few windows
        const int64_t min_windows = (recs + kWinRecs - 1) / kWinRecs;
        int64_t launched = 0;
        for (;;) {
            const int64_t batch = launched == 0 ? min_windows : 8;
            for (int64_t w = 0; w < batch; ++w) {
                recover_table_kernel<<<kWinRecs, kHyp, 0, st>>>(ctl, recs, norb, hdr, pc, cdf, kprefix, d_rng_state,
                                                               skipM, skipP, table);
                recover_compose_kernel<<<kChunksPerWin, kHyp, 0, st>>>(ctl, recs, table, comp);
                recover_walk_kernel<<<1, 32, 0, st>>>(ctl, recs, norb, hdr, pc, cdf, kprefix, table, comp, e_start,
                                                      d_rng_state, off_abs, d_status);
                recover_replay_kernel<<<kChunksPerWin, 32, 0, st>>>(ctl, kprefix, table, e_start, off_abs);
            }
            if (check_launch("recover table/compose/walk/replay kernels", (int)(4 * batch))) return -2;
            launched += batch;
            RecoverCtl h_ctl;
            if (read_back(&h_ctl, ctl, sizeof(RecoverCtl), st)) return -2;
            if (h_ctl.done) break;
            SQD_REQUIRE(launched < 64 * min_windows + 1024, "sqd_recover: the exact-stream walk does not advance");
        }
        recover_apply_kernel<<<(unsigned)((recs + 127) / 128), 128, 0, st>>>(
            ctl, d_left, d_right, recs, norb, hdr, cand, pc, cdf, off_abs, init_state, d_rng_state, d_left_out,
            d_right_out);
        return check_launch("recover_apply_kernel");
    }
    recover_select_parallel_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(
        d_left, d_right, n, norb, hdr, cand, pc, cdf, seed, d_left_out, d_right_out, d_status);
    return check_launch("recover_select_parallel_kernel");
}

int64_t sqd_merge_rows_workspace_bytes(int64_t n) {
    if (n < 0 || n >= 2147483647LL / 4) return -1;
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    // slot_rep, slot_min, slot_cnt, slot_rank | row_slot, is_first, rank(+1), gcnt(+1), goff(+1), gcur, members, sorted
    return 4 * cap * 4 + (int64_t)(8 * (n + 1)) * 4 + 1024;
}

int sqd_merge_rows(const uint64_t* d_left, const uint64_t* d_right, int64_t n64, const double* d_prob,
                   uint64_t* d_out_left, uint64_t* d_out_right, double* d_out_sum, int32_t* h_n_unique,
                   void* d_workspace, int64_t workspace_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(n64 >= 0 && workspace_bytes >= sqd_merge_rows_workspace_bytes(n64) &&
                    sqd_merge_rows_workspace_bytes(n64) >= 0,
                "sqd_merge_rows: bad size or workspace too small");
    *h_n_unique = 0;
    if (n64 == 0) return 0;
    const int n = (int)n64;
    int64_t cap = 1024;
    while (cap < 2 * n64) cap <<= 1;
    int32_t* p = (int32_t*)d_workspace;
    int32_t* slot_rep = p;   p += cap;
    int32_t* slot_min = p;   p += cap;
    int32_t* slot_cnt = p;   p += cap;
    int32_t* slot_rank = p;  p += cap;
    int32_t* row_slot = p;   p += n + 1;
    int32_t* is_first = p;   p += n + 1;
    int32_t* rank = p;       p += n + 1;
    int32_t* gcnt = p;       p += n + 1;
    int32_t* goff = p;       p += n + 1;
    int32_t* gcur = p;       p += n + 1;
    int32_t* members = p;    p += n + 1;
    int32_t* sorted = p;
    // slot_rep and slot_min (adjacent) start at kMergeEmpty = INT_MAX
    recover_fill_i32_kernel<<<(unsigned)((2 * cap + 255) / 256), 256, 0, st>>>(slot_rep, 2 * cap, kMergeEmpty);
    SQD_CUDA_OK(cudaMemsetAsync(slot_cnt, 0, (size_t)cap * 4, st));
    SQD_CUDA_OK(cudaMemsetAsync(gcur, 0, (size_t)(n + 1) * 4, st));
    const unsigned nb = (unsigned)((n + 255) / 256);
    merge_insert_kernel<<<nb, 256, 0, st>>>(d_left, d_right, n, slot_rep, slot_min, slot_cnt, row_slot,
                                            (uint32_t)(cap - 1));
    merge_flag_kernel<<<nb, 256, 0, st>>>(n, row_slot, slot_min, is_first);
    if (check_launch("merge insert/flag kernels", 3)) return -2;
    int n_unique = 0;
    if (sqd_exclusive_scan(is_first, rank, n, &n_unique, stream)) return -2;
    SQD_REQUIRE(n_unique > 0 && n_unique <= n, "sqd_merge_rows: inconsistent group count %d", n_unique);
    merge_groups_kernel<<<nb, 256, 0, st>>>(n, d_left, d_right, row_slot, is_first, rank, slot_cnt, slot_rank, gcnt,
                                            d_out_left, d_out_right);
    if (check_launch("merge_groups_kernel")) return -2;
    if (sqd_exclusive_scan(gcnt, goff, n_unique, nullptr, stream)) return -2;
    merge_scatter_kernel<<<nb, 256, 0, st>>>(n, row_slot, slot_rank, goff, gcur, members);
    const int blocks = (int)min((int64_t)kNumSMs * 8, ((int64_t)n_unique + 7) / 8);
    merge_sum_kernel<<<blocks, 256, 0, st>>>(n_unique, goff, gcnt, members, sorted, d_prob, d_out_sum);
    if (check_launch("merge scatter/sum kernels", 2)) return -2;
    *h_n_unique = n_unique;
    return 0;
}

}  // extern "C"

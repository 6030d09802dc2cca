// CI sigma-vector build  sigma = O c  in the product space of sampled alpha x beta strings.
//
// Replaces pyscf selected_ci.contract_2e (SCIcontract_2e_aaaa + SCIcontract_2e_bbaa, recalled) reached
// from qiskit_addon_sqd/fermion.py:721-723, 810-818.  pyscf resolves the two-body operator through
// (N-2)-electron intermediates and dense dgemms, O(n_det * npair^2).  The B200 design applies the
// projected operator directly from the in-set excitation tables, O(n_det * links):
//
//   sigma[a,b] = diag[a,b] c[a,b] + sum_{b' in T_b(b)} Hb[b,b'] c[a, b']                    (kernel B)
//              + sum_{b' in S_b(b)} sgn_b Wa[a, rs] c[a, b']                                 (kernel A, self item)
//              + sum_{a' in D_a(a)}    Ha[a,a'] c[a', b]                                     (kernel A, phase C)
//              + sum_{a' in S_a(a)}  ( Ha[a,a'] + sgn_a Wb[pq, b] ) c[a', b]                 (kernel A, phase D)
//              + sum_{a' in S_a(a)} sgn_a sum_{b' in S_b(b)} sgn_b g_ab[pq, rs] c[a', b']    (kernel A, phase D)
//
// S = in-set single excitations, D = in-set doubles, T = S u D.
//
// Excitation lists are extremely skewed (the Hartree-Fock string has 10-20x the partners of the median
// string), so the data layout is built for balance:
//   * beta lists are stored as SELL-32 slices: columns sorted by list length, 32 per slice, entries
//     interleaved so that a warp reads one coalesced 128-byte line per step and all 32 lanes run the
//     same trip count (no divergence, no per-thread pointer chasing);
//   * the few outlier columns ("long columns") are reduced by a whole warp each;
//   * alpha rows are cut into chunks of bounded cost, one CTA per chunk; rows with several chunks sum
//     their partial vectors in chunk order.
// Every sigma element is accumulated in a fixed order by a fixed thread: results are bit-reproducible.
// Kernel A stages, per alpha single excitation, the row c[a',:] and the integral row g_ab[pq,:] in
// shared memory with the bulk-copy engine (cp.async.bulk + mbarrier; SASS UBLKCP/SYNCS) through a ring
// of stages filled by a dedicated producer warp.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

struct SigmaArgs {
    sqd_operator op;
    const double* c;
    double* sigma;
    const int* done;  // optional device flag: non-zero -> the launch is a no-op (Davidson finished)
    long long* prof;  // optional diagnostics: 8 clock64() stamps per kernel-A CTA (sqd_sigma_profile)
    int row_begin, row_end;  // rows of sigma this launch owns (sharded builds); other rows are untouched
    // device-driven Davidson loop (CUDA graph replay): c and sigma are BASE pointers and the vector slot is
    // read from device memory, so that every cycle launches with identical arguments
    const int* slot_ptr;
    long long vec_stride;
};
__device__ __forceinline__ size_t slot_offset(const int* slot_ptr, long long stride) {
    return slot_ptr != nullptr ? (size_t)(*slot_ptr) * (size_t)stride : 0;
}

// try_wait with a suspend-time hint: a warp whose barrier phase is not complete is parked by the hardware
// (no issue slots consumed) until the phase completes or ~the hint elapses.  Busy-polling here is
// poisonous: dozens of waiting warps would take the issue slots of the one warp everybody waits for.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(64);
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kMaxStages = 8;
constexpr int kUnroll = 4;
constexpr int kUnrollC = 8;   // row loads in flight per thread in phase C
constexpr int kTileC = 256;   // alpha doubles staged per tile in phase C
constexpr int kMaxLong = SQD_MAX_LONG_COLUMNS;
constexpr uint32_t kPadMarker = 0x7ffffu;  // SELL mode-0 padding entry
constexpr int kLongA = 4;   // long columns actually used (registers per thread in kernel A)
constexpr int kRowsB = 4;      // rows of c per CTA in kernel B
constexpr int kWarpsB = 8;     // warps per CTA in kernel B (they split one SELL slice)
constexpr int kSingleCost = 16;  // plan cost of one alpha single excitation, in units of one double

// ---------------------------------------------------------------------------------------------------
// Work decomposition.  The excitation lists are extremely skewed (the Hartree-Fock string of a sampled
// set typically has 10-20x more in-set partners than the median string), so neither "one CTA per row"
// nor "one thread per column" is balanced on its own:
//   * rows are cut into CHUNKS of bounded cost (4 units per single excitation, 1 per double); a row
//     that needs several chunks writes per-chunk partial vectors that sigma_combine_kernel adds in chunk
//     order (deterministic);
//   * columns whose beta single-excitation list is long are taken out of the thread-per-column loops and
//     reduced cooperatively by one warp each (lane-strided, fixed shuffle tree).
// ---------------------------------------------------------------------------------------------------
// cost class 0 .. 63 of a single-chunk row (cost <= cost_per_chunk < 2^24: 32-bit arithmetic; a 64-bit
// division costs the single planning thread ~150 cycles per row and pass)
__device__ __forceinline__ int cost_class(int cost, int cost_per_chunk) {
    return (int)(((unsigned)cost * 64u) / (unsigned)(cost_per_chunk + 1));
}

__global__ void sigma_plan_kernel(const sqd_spin_table A, const sqd_spin_table B, int cost_per_chunk,
                                  int long_threshold, int max_chunks, int* __restrict__ chunk_row,
                                  int* __restrict__ chunk_beg, int* __restrict__ chunk_end,
                                  int* __restrict__ chunk_slot, int* __restrict__ split_row,
                                  int* __restrict__ split_slot_beg, int* __restrict__ split_n,
                                  int* __restrict__ long_idx, int* __restrict__ long_cols,
                                  int* __restrict__ counts, int stage_in_smem) {
    // setup-time, O(na + nb) work.  The ordering logic is sequential (one thread keeps the order trivially
    // deterministic) but it works from shared-memory copies of the row pointers / list lengths that the
    // whole CTA loads with coalesced reads first -- the single thread then never waits for global memory.
    // Two passes: rows that need several chunks are emitted first so that the heaviest CTAs start first.
    extern __shared__ int plan_smem[];
    if (blockIdx.x != 0) return;
    const bool staged = stage_in_smem != 0;
    int* rp_s = plan_smem;                    // [na + 1]
    int* nsa_s = rp_s + (A.n + 1);            // [na]
    int* nsb_s = nsa_s + A.n;                 // [nb]
    if (staged) {
        for (int i = threadIdx.x; i <= A.n; i += blockDim.x) rp_s[i] = A.row_ptr[i];
        for (int i = threadIdx.x; i < A.n; i += blockDim.x) nsa_s[i] = A.n_single[i];
        for (int i = threadIdx.x; i < B.n; i += blockDim.x) nsb_s[i] = B.n_single[i];
    }
    for (int b = threadIdx.x; b < B.n; b += blockDim.x) long_idx[b] = -1;
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int* a_row_ptr = staged ? rp_s : A.row_ptr;
    const int* a_n_single = staged ? nsa_s : A.n_single;
    const int* b_n_single = staged ? nsb_s : B.n_single;
    int nc = 0, nslots = 0, nsplit = 0;
    // Order of the chunk list = order in which the persistent CTAs of kernel A pick their work: rows that
    // need several chunks first (every such chunk is full), then the single-chunk rows by descending cost
    // (counting sort into 64 cost classes, row order inside a class) -- the heaviest CTAs start first and
    // the snake assignment of kernel A pairs a heavy chunk with a light one.
    constexpr int kClasses = 64;
    int class_pos[kClasses + 1];
    for (int k = 0; k <= kClasses; ++k) class_pos[k] = 0;
    int n_multi_chunks = 0;
    for (int a = 0; a < A.n; ++a) {
        const int beg = a_row_ptr[a], ns = a_n_single[a], end = a_row_ptr[a + 1];
        const int cost = kSingleCost * ns + (end - beg - ns);
        int k = (cost + cost_per_chunk - 1) / cost_per_chunk;
        if (k < 1) k = 1;
        if (k > 1) {
            n_multi_chunks += k;
        } else {
            int cls = cost_class(cost, cost_per_chunk);
            cls = kClasses - 1 - (cls < kClasses ? cls : kClasses - 1);  // class 0 = most expensive
            ++class_pos[cls + 1];
        }
    }
    bool capped = n_multi_chunks + A.n > max_chunks;  // capacity guard, never hit with the documented bound
    for (int k = 0; k < kClasses; ++k) class_pos[k + 1] += class_pos[k];
    const int single_base = capped ? 0 : n_multi_chunks;
    for (int pass = 0; pass < 2; ++pass) {
        for (int a = 0; a < A.n; ++a) {
            const int beg = a_row_ptr[a], ns = a_n_single[a], end = a_row_ptr[a + 1];
            const int nd = end - beg - ns;
            const int cost = kSingleCost * ns + nd;
            int k = (cost + cost_per_chunk - 1) / cost_per_chunk;
            if (k < 1) k = 1;
            if (capped) k = 1;
            if ((k > 1) != (pass == 0)) continue;
            if (k > 1) {
                split_row[nsplit] = a;
                split_slot_beg[nsplit] = nslots;
                split_n[nsplit] = k;
                ++nsplit;
            }
            for (int c = 0; c < k; ++c) {
                int p0 = c * cost_per_chunk, p1 = (c + 1) * cost_per_chunk;
                if (k == 1) p1 = cost;
                if (p0 > cost) p0 = cost;
                if (p1 > cost) p1 = cost;
                const int e0 = p0 < kSingleCost * ns ? p0 / kSingleCost : ns + (p0 - kSingleCost * ns);
                const int e1 = p1 < kSingleCost * ns ? p1 / kSingleCost : ns + (p1 - kSingleCost * ns);
                int at = nc;
                if (k == 1 && !capped) {
                    int cls = cost_class(cost, cost_per_chunk);
                    cls = kClasses - 1 - (cls < kClasses ? cls : kClasses - 1);
                    at = single_base + class_pos[cls]++;
                }
                chunk_row[at] = a;
                chunk_beg[at] = beg + e0;
                chunk_end[at] = beg + e1;
                chunk_slot[at] = k > 1 ? nslots++ : -1;
                ++nc;
            }
        }
    }
    // long columns: the (at most kLongA) beta strings with the longest single-excitation lists above the
    // threshold; their links are spread over all threads of a CTA instead of living in one SELL lane
    int nlong = 0;
    int taken[kLongA];  // long_idx was set to -1 by the whole CTA above
    for (int round = 0; round < kLongA; ++round) {
        int best = -1, best_len = long_threshold;
        for (int b = 0; b < B.n; ++b)
            if (b_n_single[b] > best_len) {
                bool used = false;
                for (int r = 0; r < nlong; ++r) used = used || taken[r] == b;
                if (!used) {
                    best = b;
                    best_len = b_n_single[b];
                }
            }
        if (best < 0) break;
        long_idx[best] = nlong;
        taken[nlong] = best;
        long_cols[nlong++] = best;
    }
    counts[0] = nc;
    counts[1] = nslots;
    counts[2] = nsplit;
    counts[3] = nlong;
}


// ---------------------------------------------------------------------------------------------------
// SELL-32 builder (setup time).  mode 0: single excitations only, columns longer than long_threshold get
// length 0 (they are handled cooperatively); mode 1: every entry (singles then doubles) with values.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int sell_len(const sqd_spin_table& T, int b, int mode,
                                        const int* __restrict__ long_idx) {
    if (mode == 0) return (long_idx != nullptr && long_idx[b] >= 0) ? 0 : T.n_single[b];
    return T.row_ptr[b + 1] - T.row_ptr[b];
}

// rank of every column in (length descending, index ascending) order; perm[rank] = column
__global__ void sell_rank_kernel(const sqd_spin_table T, int mode, const int* __restrict__ long_idx,
                                 int* __restrict__ perm, int* __restrict__ len_sorted) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= T.n) return;
    const int lb = sell_len(T, b, mode, long_idx);
    int rank = 0;
    for (int o = 0; o < T.n; ++o) {
        const int lo = sell_len(T, o, mode, long_idx);
        rank += (lo > lb) || (lo == lb && o < b);
    }
    perm[rank] = b;
    len_sorted[rank] = lb;
}

// slice_ptr (exclusive scan of 32 * slice length) by one thread: n/32 slices, setup time
__global__ void sell_slice_kernel(int n, const int* __restrict__ len_sorted, int* __restrict__ slice_ptr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int ns = (n + 31) / 32;
    int off = 0;
    for (int s = 0; s < ns; ++s) {
        slice_ptr[s] = off;
        // sorted descending: the first lane of a slice is its longest; lengths are padded to a multiple
        // of kUnroll so the unrolled gather loops need no tail handling
        off += 32 * ((len_sorted[32 * s] + kUnroll - 1) / kUnroll * kUnroll);
    }
    slice_ptr[ns] = off;
}

__global__ void sell_fill_kernel(const sqd_spin_table T, int mode, const int* __restrict__ perm,
                                 const int* __restrict__ len_sorted, const int* __restrict__ slice_ptr,
                                 uint32_t* __restrict__ pack, double* __restrict__ val) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    const int s = pos >> 5, lane = pos & 31;
    const int ns = (T.n + 31) / 32;
    if (s >= ns) return;
    const int base = slice_ptr[s];
    const int slen = (slice_ptr[s + 1] - base) >> 5;
    const int mylen = pos < T.n ? len_sorted[pos] : 0;
    const int src = pos < T.n ? T.row_ptr[perm[pos]] : 0;
    // padding: mode 0 -> marker (partner field all ones, never a valid index); mode 1 -> partner 0, value 0
    const uint32_t pad = mode == 0 ? kPadMarker : 0u;
    for (int k = 0; k < slen; ++k) {
        const bool ok = k < mylen;
        pack[base + k * 32 + lane] = ok ? T.pack[src + k] : pad;
        if (mode == 1) val[base + k * 32 + lane] = ok ? T.val[src + k] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------------
// Kernel B:  sigma[a,b] = diag[a,b] c[a,b] + sum_{b'} Hb[b,b'] c[a,b']
// CTA = kRowsB rows x ONE SELL slice (32 columns); its kWarpsB warps split the slice's entry range and
// their partial sums meet in shared memory in warp order.  The rows of c are staged with bulk copies.  Writes every element of sigma (pads = 0); kernel A then adds its part.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsB * 32)
sigma_b_kernel(const SigmaArgs P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (P.done != nullptr && *P.done != 0) return;
    const double* Pc = P.c + slot_offset(P.slot_ptr, P.vec_stride);
    double* Psig = P.sigma + slot_offset(P.slot_ptr, P.vec_stride);
    const sqd_operator& op = P.op;
    const sqd_sell& L = op.bb;
    const int na = op.a.n, nb = op.b.n, ldc = op.ldc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a0 = P.row_begin + blockIdx.x * kRowsB;
    const int nrows = min(kRowsB, min(na, P.row_end) - a0);
    const bool ham = op.use_same_spin != 0;
    const int slice = blockIdx.y;
    const int pos = slice * 32 + lane;

    double* Cs = reinterpret_cast<double*>(smem_raw);          // [kRowsB][ldc]
    double* red = Cs + (size_t)kRowsB * ldc;                    // [kWarpsB][kRowsB][32]
    uint64_t* bar = reinterpret_cast<uint64_t*>(red + kWarpsB * kRowsB * 32);
    if (slice >= L.n_slices) {  // only pad positions live here
        if (warp == 0 && pos >= nb && pos < ldc)
            for (int r = 0; r < nrows; ++r) Psig[(size_t)(a0 + r) * ldc + pos] = 0.0;
        return;
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)(nrows * ldc * sizeof(double)));
        for (int r = 0; r < nrows; ++r)
            bulk_g2s(Cs + (size_t)r * ldc, Pc + (size_t)(a0 + r) * ldc, (uint32_t)(ldc * sizeof(double)), bar);
    }
    double acc[kRowsB];
#pragma unroll
    for (int r = 0; r < kRowsB; ++r) acc[r] = 0.0;
    const int base = L.slice_ptr[slice];
    const int slen = (L.slice_ptr[slice + 1] - base) >> 5;  // multiple of kUnroll; padding has value 0
    // this warp's share of the entry range, a multiple of kUnroll long
    int per = (slen + kWarpsB - 1) / kWarpsB;
    per = (per + kUnroll - 1) / kUnroll * kUnroll;
    const int k_beg = min(slen, warp * per), k_end = min(slen, k_beg + per);
    mbar_wait_parity(bar, 0);
    if (ham) {
        const uint32_t* pp = L.pack + base + lane;
        const double* vp = L.val + base + lane;
        for (int k0 = k_beg; k0 < k_end; k0 += kUnroll) {
            uint32_t pk[kUnroll];
            double v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                pk[u] = __ldg(pp + (k0 + u) * 32);
                v[u] = __ldg(vp + (k0 + u) * 32);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const uint32_t bp = pk[u] & 0x7ffffu;
#pragma unroll
                for (int r = 0; r < kRowsB; ++r)
                    if (r < nrows) acc[r] = fma(v[u], Cs[r * ldc + bp], acc[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kRowsB; ++r) red[(warp * kRowsB + r) * 32 + lane] = acc[r];
    __syncthreads();
    if (warp == 0) {
        if (pos < nb) {
            const int b = L.perm[pos];
#pragma unroll
            for (int r = 0; r < kRowsB; ++r) {
                if (r < nrows) {
                    double t = red[r * 32 + lane];
#pragma unroll
                    for (int w = 1; w < kWarpsB; ++w) t += red[(w * kRowsB + r) * 32 + lane];
                    Psig[(size_t)(a0 + r) * ldc + b] =
                        fma(op.diag[(size_t)(a0 + r) * ldc + b], Cs[r * ldc + b], t);
                }
            }
        } else if (pos < ldc) {
            for (int r = 0; r < nrows; ++r) Psig[(size_t)(a0 + r) * ldc + pos] = 0.0;  // pads
        }
    }
}

// sigma[a,:] += sum over the row's chunk partials, in chunk order
__global__ void sigma_combine_kernel(const int* __restrict__ done, const sqd_sigma_plan pl, int ldc,
                                     double* __restrict__ sigma, int row_begin, int row_end,
                                     const int* __restrict__ slot_ptr, long long vec_stride) {
    if (done != nullptr && *done != 0) return;
    sigma += slot_offset(slot_ptr, vec_stride);
    const int j = blockIdx.x;
    const int a = pl.split_row[j], s0 = pl.split_slot_beg[j], k = pl.split_n[j];
    if (a < row_begin || a >= row_end) return;
    for (int b = threadIdx.x; b < ldc; b += blockDim.x) {
        double acc = pl.part[(size_t)s0 * ldc + b];
        for (int c = 1; c < k; ++c) acc += pl.part[(size_t)(s0 + c) * ldc + b];
        sigma[(size_t)a * ldc + b] += acc;
    }
}

// ---------------------------------------------------------------------------------------------------
// Kernel A: alpha excitations.  One CTA per chunk of one alpha row.  Warp-specialised: the LAST warp is
// the producer (one lane drives the bulk-copy engine through an NST-deep ring of {c[a',:], g_ab[pq,:]}
// stages guarded by full/empty mbarriers); all other warps are consumers.
//   natural mapping  (thread t <-> column t)        : phase C streaming and the c[a',b] terms, coalesced
//   sorted mapping   (thread t <-> column perm[t])  : the beta-single gathers through SELL slices
// The two partial results meet in a shared-memory exchange buffer before one coalesced update of sigma.
// ---------------------------------------------------------------------------------------------------
// Chunk of round k for this CTA: even rounds walk the (cost-descending) list forwards, odd rounds
// backwards, so the CTA that got the heaviest chunk of one round gets the lightest of the next.
__device__ __forceinline__ int snake_chunk(int k) {
    return (k & 1) ? (k + 1) * (int)gridDim.x - 1 - (int)blockIdx.x : k * (int)gridDim.x + (int)blockIdx.x;
}

// ---------------------------------------------------------------------------------------------------
// "Wide" kernel: the same sum as kernels B + A, for subspaces whose rows do not fit the shared-memory
// staging of the kernels above (nb > 5760 for v1, nb > 8192 for v2) -- up to the 2^19 strings per spin of
// the tables.  No row of c is staged: CTA = one alpha row x kWideCols consecutive columns, thread = one
// column, the c[a', b'] operands are gathered through L1/L2 (the rows c[a',:] are shared by the CTAs of one
// alpha row).  kWideBlock alpha single excitations share one pass over the column's beta links (one load
// of the link word for kWideBlock FMAs, kWideBlock independent accumulation chains); their integral rows
// g_ab[pq,:] are copied to shared memory first (kWideBlock * ldg doubles: a random 64-bit gather costs a few
// bank cycles there against one L1 tag lookup per distinct 128-byte line -- ncu of the first version, with
// both gathers in L1: 74 % L1 throughput, 0.74 IPC).  No plan, no SELL copies; every element is summed by one
// thread in table order, so builds restricted to row blocks are bit-equal to the full build.
// ---------------------------------------------------------------------------------------------------
constexpr int kWideCols = 256;

template <int kWideBlock>
__global__ void __launch_bounds__(kWideCols)
sigma_wide_kernel(const SigmaArgs P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (P.done != nullptr && *P.done != 0) return;
    const double* __restrict__ c = P.c + slot_offset(P.slot_ptr, P.vec_stride);
    double* __restrict__ sig = P.sigma + slot_offset(P.slot_ptr, P.vec_stride);
    const sqd_operator& op = P.op;
    const int nb = op.b.n, ldc = op.ldc, ldg = op.ldg;
    // column tile = fastest index: the CTAs in flight at any time cover a few alpha rows completely, so the rows
    // c[a',:] they gather from stay in L2 (row-major CTA order spread 1000 resident CTAs over 1000 alpha rows:
    // 53 GB of DRAM reads per build at 8.1e7 determinants, ncu)
    const int n_tiles = (ldc + kWideCols - 1) / kWideCols;
    const int a = P.row_begin + (int)(blockIdx.x / (unsigned)n_tiles);
    const int b = (int)(blockIdx.x % (unsigned)n_tiles) * kWideCols + (int)threadIdx.x;
    if (a >= P.row_end) return;                 // whole CTA
    double* g_s = reinterpret_cast<double*>(smem_raw);   // [kWideBlock][ldg] integral rows of the current block
    const bool live = b < nb;                   // pad / out-of-range columns only help with the staging
    const bool ham = op.use_same_spin != 0;
    const double* __restrict__ crow = c + (size_t)a * ldc;
    double acc = 0.0;
    int fb = 0, fs = 0;
    const int eb = __ldg(op.a.row_ptr + a), es = eb + __ldg(op.a.n_single + a), ee = __ldg(op.a.row_ptr + a + 1);
    if (live) {
        acc = __ldg(op.diag + (size_t)a * ldc + b) * crow[b];
        // beta excitations of column b: singles carry Hb + sgn_b Wa[a, rs], doubles Hb
        fb = __ldg(op.b.row_ptr + b);
        fs = fb + __ldg(op.b.n_single + b);
        const int fe = __ldg(op.b.row_ptr + b + 1);
        const double* __restrict__ wa = op.Wa != nullptr ? op.Wa + (size_t)a * ldg : nullptr;
        for (int f = fb; f < fs; ++f) {
            const uint32_t pv = __ldg(op.b.pack + f);
            double v = ham ? __ldg(op.b.val + f) : 0.0;
            if (wa != nullptr) {
                const double w = __ldg(wa + ((pv >> 19) & 0xfffu));
                v += (pv >> 31) ? -w : w;
            }
            acc = fma(v, crow[pv & 0x7ffffu], acc);
        }
        if (ham) {
            for (int f = fs; f < fe; ++f) acc = fma(__ldg(op.b.val + f), crow[__ldg(op.b.col + f)], acc);
            // alpha doubles
            for (int e = es; e < ee; ++e)
                acc = fma(__ldg(op.a.val + e), c[(size_t)__ldg(op.a.col + e) * ldc + b], acc);
        }
    }
    // alpha singles: (Ha + sgn_a Wb[pq, b]) c[a', b] + sgn_a sum_{b'} sgn_b g_ab[pq, rs] c[a', b']
    for (int e0 = eb; e0 < es; e0 += kWideBlock) {
        const double* cr[kWideBlock];
        double sa[kWideBlock], sum[kWideBlock];
        __syncthreads();   // the previous block's rows are no longer read
#pragma unroll
        for (int j = 0; j < kWideBlock; ++j) {
            const int e = e0 + j;
            const bool ok = e < es;
            const uint32_t m = ok ? __ldg(op.a.meta + e) : 0u;
            const uint32_t pq = m & 0x7fffffffu;
            sa[j] = ok ? ((m >> 31) ? -1.0 : 1.0) : 0.0;
            cr[j] = c + (size_t)(ok ? __ldg(op.a.col + e) : (uint32_t)a) * ldc;
            sum[j] = 0.0;
            const double* __restrict__ grow = op.gab + (size_t)pq * ldg;
            for (int t = threadIdx.x; t < ldg; t += kWideCols) g_s[j * ldg + t] = __ldg(grow + t);
            if (ok && live) {
                const double va = ham ? __ldg(op.a.val + e) : 0.0;
                const double wb = op.Wb != nullptr ? __ldg(op.Wb + (size_t)pq * ldc + b) : 0.0;
                acc = fma(fma(sa[j], wb, va), cr[j][b], acc);
            }
        }
        __syncthreads();
        for (int f = fb; f < fs; ++f) {   // empty for columns that are not live
            const uint32_t pv = __ldg(op.b.pack + f);
            const uint32_t bp = pv & 0x7ffffu, rs = (pv >> 19) & 0xfffu;
            const bool neg = (pv >> 31) != 0;
#pragma unroll
            for (int j = 0; j < kWideBlock; ++j) {
                const double t = g_s[j * ldg + rs] * cr[j][bp];
                sum[j] += neg ? -t : t;
            }
        }
#pragma unroll
        for (int j = 0; j < kWideBlock; ++j) acc = fma(sa[j], sum[j], acc);
    }
    if (b < ldc) sig[(size_t)a * ldc + b] = live ? acc : 0.0;   // pad column: zero
}


struct ChunkInfo {
    int a, slot, it_beg, db_beg, db_end, n_total;
    bool self_item, owned;
};

__device__ __forceinline__ ChunkInfo load_chunk(const SigmaArgs& P, int chunk) {
    const sqd_operator& op = P.op;
    const sqd_sigma_plan& pl = op.plan;
    ChunkInfo c;
    c.a = pl.chunk_row[chunk];
    c.owned = c.a >= P.row_begin && c.a < P.row_end;  // sharded build: another rank may own this row
    const int cbeg = pl.chunk_beg[chunk], cend = pl.chunk_end[chunk];
    c.slot = pl.chunk_slot[chunk];
    const int single_end = op.a.row_ptr[c.a] + op.a.n_single[c.a];
    c.it_beg = min(cbeg, single_end);                                 // phase D range
    const int it_end = min(cend, single_end);
    c.db_beg = max(cbeg, single_end);                                 // phase C range
    c.db_end = max(cend, single_end);
    // The first chunk of a row also owns the row's "self item": the term
    //   sum_{b' in S_b(b)} sgn_b Wa[a, rs] c[a, b']
    // has exactly the shape of an alpha single excitation with a' = a, sgn_a = +1 and the integral row
    // replaced by Wa[a, :], so it rides the same ring and the same gather loop.
    c.self_item = (cbeg == op.a.row_ptr[c.a]) && (op.Wa != nullptr);
    c.n_total = (it_end - c.it_beg) + (c.self_item ? 1 : 0);
    return c;
}

template <int CPT, bool STAGE_PACK, int TB, int MINB>
__global__ void __launch_bounds__(TB, MINB)
sigma_a_kernel(const SigmaArgs P, const int NST) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (P.done != nullptr && *P.done != 0) return;
    const double* Pc = P.c + slot_offset(P.slot_ptr, P.vec_stride);
    double* Psig = P.sigma + slot_offset(P.slot_ptr, P.vec_stride);
    const sqd_operator& op = P.op;
    const sqd_sigma_plan& pl = op.plan;
    const sqd_sell& L = op.bd;
    const int nb = op.b.n, ldc = op.ldc, ldg = op.ldg;
    const int ncons = blockDim.x - 32;  // consumer threads
    const int nwarp_c = ncons >> 5;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int n_long = pl.n_long;
    const bool ham = op.use_same_spin != 0;

    double* xbuf = reinterpret_cast<double*>(smem_raw);  // [ldc] exchange sorted -> natural mapping
    double* stage = xbuf + ldc;                           // NST x ([ldc] row c[a',:] + [ldg] row g_ab[pq,:])
    const int stage_len = ldc + ldg;
    double* acc_long = stage + (size_t)NST * stage_len;   // [kLongA][32 warps] (kMaxLong >= 4*32/... sized)
    uint64_t* bars = reinterpret_cast<uint64_t*>(acc_long + kLongA * 32);
    uint64_t* full = bars;                  // [NST]
    uint64_t* empty = bars + kMaxStages;    // [NST]
    double* dval_s = reinterpret_cast<double*>(bars + 2 * kMaxStages);    // [kTileC]
    uint32_t* dcol_s = reinterpret_cast<uint32_t*>(dval_s + kTileC);      // [kTileC]
    uint32_t* pk_s = dcol_s + kTileC;                                     // [n_entries] when staged

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nwarp_c);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // =========================== producer warp ===========================
    if (tid >= ncons) {
        if (tid == ncons) {
            int s = 0, round = 0;  // ring position: runs on across the chunks of this CTA
            for (int k = 0; k * (int)gridDim.x < pl.n_chunks; ++k) {
                const int chunk = snake_chunk(k);
                if (chunk >= pl.n_chunks) continue;
                const ChunkInfo ck = load_chunk(P, chunk);
                if (!ck.owned) continue;
                for (int t = 0; t < ck.n_total; ++t) {
                    if (round > 0) mbar_wait_parity(&empty[s], (uint32_t)((round - 1) & 1));
                    const double* crow;
                    const double* grow;
                    if (ck.self_item && t == 0) {
                        crow = Pc + (size_t)ck.a * ldc;
                        grow = op.Wa + (size_t)ck.a * ldg;
                    } else {
                        const int e = ck.it_beg + t - (ck.self_item ? 1 : 0);
                        crow = Pc + (size_t)op.a.col[e] * ldc;
                        grow = op.gab + (size_t)(op.a.meta[e] & 0x7fffffffu) * ldg;
                    }
                    double* dst = stage + (size_t)s * stage_len;
                    mbar_expect_tx(&full[s], (uint32_t)((ldc + ldg) * sizeof(double)));
                    bulk_g2s(dst, crow, (uint32_t)(ldc * sizeof(double)), &full[s]);
                    bulk_g2s(dst + ldc, grow, (uint32_t)(ldg * sizeof(double)), &full[s]);
                    if (++s == NST) {
                        s = 0;
                        ++round;
                    }
                }
            }
        }
        return;
    }

    // =========================== consumer warps ==========================
#define SQD_STAMP(k) \
    if (P.prof != nullptr && tid == 0) P.prof[(size_t)chunk * 8 + (k)] = clock64();
    // The beta SELL table is re-read for every alpha excitation of the chunk: when it is small enough
    // (host decision, STAGE_PACK) it is kept in shared memory, pre-decoded into byte offsets:
    //   bits 0-15 = 8*b' (into the staged c row), bits 16-30 = 8*rs (into the staged integral row),
    //   bit 31 = sign; padding entries point at the zero pad of the integral row (rs = norb^2).
    if (STAGE_PACK) {
        const uint32_t zero_slot = (uint32_t)(op.norb * op.norb * 8) << 16;
        for (int i = tid; i < L.n_entries; i += ncons) {
            const uint32_t pv = __ldg(L.pack + i);
            const uint32_t col = pv & 0x7ffffu;
            pk_s[i] = col == kPadMarker ? zero_slot
                                        : ((col << 3) | (((pv >> 19) & 0xfffu) << 19) | (pv & 0x80000000u));
        }
        asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
    }
    // long columns: this thread's share of their links lives in registers for the whole chunk
    uint32_t lpk[kLongA][CPT];
    double lacc[kLongA];
#pragma unroll
    for (int li = 0; li < kLongA; ++li) {
        lacc[li] = 0.0;
        const int lb = li < n_long ? pl.long_cols[li] : 0;
        const int lbeg = op.b.row_ptr[lb], llen = li < n_long ? op.b.n_single[lb] : 0;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int idx = tid + c * ncons;
            lpk[li][c] = idx < llen ? __ldg(op.b.pack + lbeg + idx) : 0xffffffffu;  // rs=0xfff: never real
        }
    }
    int s_base[CPT], s_len[CPT], my_len[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int pos = tid + c * ncons;
        const int slice = pos >> 5;
        const bool ok = slice < L.n_slices;
        s_base[c] = ok ? L.slice_ptr[slice] : 0;
        s_len[c] = ok ? (L.slice_ptr[slice + 1] - s_base[c]) >> 5 : 0;
        my_len[c] = pos < nb ? L.len[pos] : 0;
    }

    // Persistent CTA: the grid is capped at what is resident at once and every CTA walks the chunk list
    // in snake order, so a sigma build never leaves CTAs waiting in the hardware queue in front of
    // the kernels of the other solves that share the GPU.
    int s = 0, round = 0;
    for (int k = 0; k * (int)gridDim.x < pl.n_chunks; ++k) {
    const int chunk = snake_chunk(k);
    if (chunk >= pl.n_chunks) continue;
    const ChunkInfo ck = load_chunk(P, chunk);
    if (!ck.owned) continue;
    const int a = ck.a, slot = ck.slot, it_beg = ck.it_beg, db_beg = ck.db_beg, db_end = ck.db_end;
    const int n_total = ck.n_total;
    const bool self_item = ck.self_item;
    SQD_STAMP(0)
    double acc_nat[CPT], acc_srt[CPT];
#pragma unroll
    for (int li = 0; li < kLongA; ++li) lacc[li] = 0.0;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        acc_nat[c] = 0.0;
        acc_srt[c] = 0.0;
    }

    SQD_STAMP(1)
    // ---- phase C: alpha doubles.  The (partner, value) list of the chunk is first copied to shared
    // memory with coalesced loads, so that the row reads c[a', b] -- the only long-latency accesses left
    // -- are independent and kUnrollC of them are in flight per thread.
    if (ham) {
        for (int t0 = db_beg; t0 < db_end; t0 += kTileC) {
            const int tn = min(kTileC, db_end - t0);
            asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
            for (int i = tid; i < tn; i += ncons) {
                dcol_s[i] = __ldg(op.a.col + t0 + i);
                dval_s[i] = __ldg(op.a.val + t0 + i);
            }
            asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
            for (int e0 = 0; e0 < tn; e0 += kUnrollC) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int b = tid + c * ncons;
                    if (b < nb) {
                        double x[kUnrollC];
#pragma unroll
                        for (int u = 0; u < kUnrollC; ++u)
                            x[u] = __ldg(Pc + (size_t)dcol_s[min(e0 + u, tn - 1)] * ldc + b);
#pragma unroll
                        for (int u = 0; u < kUnrollC; ++u)
                            if (e0 + u < tn) acc_nat[c] = fma(dval_s[e0 + u], x[u], acc_nat[c]);
                    }
                }
            }
        }
    }

    SQD_STAMP(2)
    // ---- phase D: alpha singles (and the self item) through the staged ring -------------------------
    long long t_wait = 0, t_pre = 0, t_loop = 0;
    for (int t = 0; t < n_total; ++t) {
        const long long c0 = P.prof ? clock64() : 0;
        const bool is_self = self_item && t == 0;
        const int e = it_beg + t - (self_item ? 1 : 0);
        double sa = 1.0, va = 0.0;
        uint32_t pq = 0;
        if (!is_self) {
            const uint32_t m = __ldg(op.a.meta + e);
            pq = m & 0x7fffffffu;
            sa = (m >> 31) ? -1.0 : 1.0;
            va = ham ? __ldg(op.a.val + e) : 0.0;
        }
        const double* Cn = stage + (size_t)s * stage_len;
        const double* gs = Cn + ldc;
        // global loads that do not depend on the stage are issued before the wait
        double wb[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int b = tid + c * ncons;
            wb[c] = (!is_self && op.Wb && b < nb) ? __ldg(op.Wb + (size_t)pq * ldc + b) : 0.0;
        }
        const long long c1 = P.prof ? clock64() : 0;
        mbar_wait_parity(&full[s], (uint32_t)(round & 1));
        const long long c2 = P.prof ? clock64() : 0;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            // sorted mapping: SELL slice, warp-uniform trip count (a multiple of kUnroll), no branches
            double sum = 0.0;
            const int sl = s_len[c];
            if (STAGE_PACK) {
                const char* CnB = reinterpret_cast<const char*>(Cn);
                const char* gsB = reinterpret_cast<const char*>(gs);
                const uint32_t* src = pk_s + s_base[c] + lane;
                for (int k0 = 0; k0 < sl; k0 += kUnroll) {
                    uint32_t pk[kUnroll];
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) pk[u] = src[(k0 + u) * 32];
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        const double cv = *reinterpret_cast<const double*>(CnB + (pk[u] & 0xffffu));
                        const double gv = *reinterpret_cast<const double*>(gsB + ((pk[u] >> 16) & 0x7fffu));
                        // the sign bit of the link goes straight into the sign bit of the integral
                        const double gsg = __hiloint2double(__double2hiint(gv) ^ (int)(pk[u] & 0x80000000u),
                                                            __double2loint(gv));
                        sum = fma(gsg, cv, sum);
                    }
                }
            } else {
                const int ml = my_len[c];
                const uint32_t* src = L.pack + s_base[c] + lane;
                for (int k0 = 0; k0 < sl; k0 += kUnroll) {
                    uint32_t pk[kUnroll];
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) pk[u] = __ldg(src + (k0 + u) * 32);
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        const bool ok = k0 + u < ml;
                        const uint32_t pv = ok ? pk[u] : 0u;
                        const double tv = gs[(pv >> 19) & 0xfffu] * Cn[pv & 0x7ffffu];
                        sum += ok ? ((pv >> 31) ? -tv : tv) : 0.0;
                    }
                }
            }
            acc_srt[c] = fma(sa, sum, acc_srt[c]);
            // natural mapping: the c[a', b] terms (none for the self item: sa*wb + va == 0 there)
            const int b = tid + c * ncons;
            if (b < nb) acc_nat[c] = fma(fma(sa, wb[c], va), Cn[b], acc_nat[c]);
        }
#pragma unroll
        for (int li = 0; li < kLongA; ++li) {
            if (li < n_long) {
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const uint32_t pk = lpk[li][c];
                    const bool ok = pk != 0xffffffffu;
                    const uint32_t pv = ok ? pk : 0u;
                    const double tv = sa * gs[(pv >> 19) & 0xfffu] * Cn[pv & 0x7ffffu];
                    lacc[li] += ok ? ((pv >> 31) ? -tv : tv) : 0.0;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with the stage
        if (++s == NST) {
            s = 0;
            ++round;
        }
        if (P.prof) {
            const long long c3 = clock64();
            t_pre += c1 - c0;
            t_wait += c2 - c1;
            t_loop += c3 - c2;
        }
    }

    SQD_STAMP(3)
    // ---- exchange sorted -> natural, then one coalesced update -------------------------------------
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int pos = tid + c * ncons;
        if (pos < nb) xbuf[L.perm[pos]] = acc_srt[c];
    }
#pragma unroll
    for (int li = 0; li < kLongA; ++li) {
        const double v = warp_sum(lacc[li]);  // fixed tree inside the warp, fixed warp order below
        if (lane == 0) acc_long[li * 32 + warp] = v;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int b = tid + c * ncons;
        if (b < nb) {
            const int li = pl.long_idx[b];
            double v = acc_nat[c] + xbuf[b];
            if (li >= 0)
                for (int w = 0; w < nwarp_c; ++w) v += acc_long[li * 32 + w];
            if (slot < 0)
                Psig[(size_t)a * ldc + b] += v;  // kernel B wrote the element earlier in the stream
            else
                pl.part[(size_t)slot * ldc + b] = v;
        } else if (b < ldc && slot >= 0) {
            pl.part[(size_t)slot * ldc + b] = 0.0;
        }
    }
    SQD_STAMP(4)
    if (P.prof != nullptr && tid == 0) {
        P.prof[(size_t)chunk * 8 + 5] = n_total;
        P.prof[(size_t)chunk * 8 + 6] = db_end - db_beg;
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        P.prof[(size_t)chunk * 8 + 7] = smid;
        P.prof[(size_t)chunk * 8 + 0] = t_pre;   // diagnostics layout: see tests/gpu_sigma_phases.py
        P.prof[(size_t)chunk * 8 + 1] = t_wait;
        P.prof[(size_t)chunk * 8 + 4] = t_loop;
    }
    // xbuf / acc_long are reused by the next chunk
    asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
    }  // chunk loop
#undef SQD_STAMP
}

struct SigmaPlan {
    int CPT, threads, stages;
    bool stage_pack;
    size_t smem;
};

// tuning knobs (environment, read once): SQD_SIGMA_STAGES = ring depth (0 = auto),
// SQD_SIGMA_PACK_BYTES = largest beta SELL table that is staged in shared memory
static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// columns per thread of the alpha kernel for rows of ldc doubles; 0 = beyond the largest instance
static int v1_columns_per_thread(int ldc) {
    if (ldc <= 992) return 1;
    const int cpt = (ldc + 511) / 512;
    return cpt <= 2 ? 2 : cpt <= 4 ? 4 : cpt <= 8 ? 8 : cpt <= 12 ? 12 : 0;
}

// true when rows of ldc doubles (and integral rows of ldg doubles) fit the shared-memory staging of
// kernels A and B with the smallest ring; otherwise the wide kernel is the only v1-family choice
static bool v1_shape_fits(int ldc, int ldg) {
    if (v1_columns_per_thread(ldc) == 0) return false;
    const size_t smem_a = (size_t)(ldc + 2 * (ldc + ldg) + kLongA * 32) * sizeof(double) +
                          (2 * kMaxStages) * sizeof(uint64_t) + kTileC * 12;
    const size_t smem_b = (size_t)(kRowsB * ldc + kWarpsB * kRowsB * 32) * sizeof(double) + 16;
    return smem_a <= 224 * 1024 && smem_b <= 224 * 1024;
}

static bool plan_sigma(const sqd_operator* op, SigmaPlan* pl) {
    static const int knob_stages = env_int("SQD_SIGMA_STAGES", 0);
    static const int knob_pack = env_int("SQD_SIGMA_PACK_BYTES", 24 * 1024);
    const int ldc = op->ldc;
    int threads, cpt;
    // `threads` counts the consumer threads; one more warp (the producer) is added at launch.  One column
    // per thread up to 992 columns; beyond that CPT columns per thread with at most 512 consumers (the
    // register budget of the multi-column instances: __launch_bounds__(544), 120 registers)
    if (ldc <= 992) {
        threads = ((ldc + 31) / 32) * 32;
        if (threads < 32) threads = 32;
        cpt = 1;
    } else {
        cpt = (ldc + 511) / 512;
        cpt = cpt <= 2 ? 2 : cpt <= 4 ? 4 : cpt <= 8 ? 8 : cpt <= 12 ? 12 : 0;
        if (cpt == 0) return false;
        threads = (((ldc + cpt - 1) / cpt + 31) / 32) * 32;
    }
    const int cpt_t = cpt;
    // beta SELL table staged in shared memory: up to 24 KB for the one-column instances (their CTAs share
    // an SM, see the ring budget below); the multi-column instances are alone on their SM (96 registers x
    // 544 threads), so the table may take whatever two ring stages leave free
    const size_t pack_bytes = ((size_t)op->bd.n_entries * 4 + 15) / 16 * 16;
    const int n2 = op->norb * op->norb;
    auto smem_with = [&](int nst, bool staged) {
        return (size_t)(ldc + nst * (ldc + op->ldg) + kLongA * 32) * sizeof(double) +
               (2 * kMaxStages) * sizeof(uint64_t) + kTileC * 12 + (staged ? pack_bytes : 0);
    };
    const size_t budget = cpt == 1 ? 56 * 1024 : 200 * 1024;
    const size_t pack_limit = cpt == 1 ? (size_t)knob_pack : (size_t)160 * 1024;
    const bool stage_pack = op->bd.n_entries > 0 && pack_bytes <= pack_limit && ldc <= 8191 && n2 <= 4095 &&
                            op->ldg >= n2 + 1 && smem_with(2, true) <= 224 * 1024;
    auto smem_of = [&](int nst) { return smem_with(nst, stage_pack); };
    if (smem_of(2) > 224 * 1024) return false;
    // kernel B: kRowsB rows of c + the cross-warp reduction buffer
    if ((size_t)(kRowsB * ldc + kWarpsB * kRowsB * 32) * sizeof(double) + 16 > 224 * 1024) return false;
    // ring depth: as deep as fits in the budget (56 KB keeps 4 one-column CTAs per SM), 2 to 6 stages
    int nst = 2;
    while (nst < 6 && smem_of(nst + 1) <= budget) ++nst;
    if (knob_stages >= 2 && knob_stages <= kMaxStages && smem_of(knob_stages) <= 224 * 1024)
        nst = knob_stages;
    pl->CPT = cpt_t;
    pl->threads = threads;
    pl->stages = nst;
    pl->stage_pack = stage_pack;
    pl->smem = smem_of(nst);
    return true;
}

template <typename K>
static int opt_in_smem(K kern, size_t smem, bool* configured) {
    if (smem > 48 * 1024) {
        int dev = 0;
        SQD_CUDA_OK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            SQD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(227 * 1024)));
            // as many CTAs per SM as the registers allow: do not let a smaller carve-out cap them
            SQD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                             (int)cudaSharedmemCarveoutMaxShared));
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    return 0;
}

template <int CPT, bool STAGE_PACK, int TB, int MINB>
static int launch_sigma_a_tb(const SigmaArgs& args, const SigmaPlan& pl, cudaStream_t st) {
    auto kern = sigma_a_kernel<CPT, STAGE_PACK, TB, MINB>;
    static bool cfg_a[64] = {false};
    if (opt_in_smem(kern, pl.smem, cfg_a)) return -2;
    // grid = the CTAs that are resident at once (occupancy query, cached per kernel instance and shape)
    static thread_local int cached_key = -1, cached_ctas = 0;
    const int key = (pl.threads + 32) * 1024 + (int)(pl.smem / 256);
    if (key != cached_key) {
        int per_sm = 0;
        SQD_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, pl.threads + 32, pl.smem));
        cached_ctas = (per_sm > 0 ? per_sm : 1) * kNumSMs;
        cached_key = key;
    }
    // Persistent CTAs (grid capped at residency) only when several solves share the GPU: there they keep
    // a sigma build from parking CTAs in the hardware queue in front of the other solves' kernels.  A lone
    // solve is better off with one CTA per chunk -- the hardware scheduler balances the uneven chunks
    // (1e6 determinants: 0.91 ms against 1.18 ms per build).  SQD_SIGMA_PERSISTENT=0/1 overrides.
    static const int knob_persistent = env_int("SQD_SIGMA_PERSISTENT", -1);
    const bool persistent = knob_persistent >= 0 ? knob_persistent != 0 : args.op.throughput_mode != 0;
    const int grid = (!persistent || args.op.plan.n_chunks < cached_ctas) ? args.op.plan.n_chunks
                                                                            : cached_ctas;
    kern<<<grid, pl.threads + 32, pl.smem, st>>>(args, pl.stages);
    return check_launch("sigma_a_kernel");
}

// CTAs of at most 384 threads are compiled under a register cap that lets three of them share an SM
template <int CPT, bool STAGE_PACK>
static int launch_sigma_a(const SigmaArgs& args, const SigmaPlan& pl, cudaStream_t st) {
    if constexpr (CPT == 1) {
        // SQD_SIGMA_MINB overrides the choice made from the operator's throughput_mode (1 or 3 CTAs per SM)
        static const int knob_env = env_int("SQD_SIGMA_MINB", 0);
        const int knob_minb = knob_env > 0 ? knob_env : (args.op.throughput_mode ? 3 : 1);
        if (pl.threads + 32 <= 384 && knob_minb == 3)
            return launch_sigma_a_tb<1, STAGE_PACK, 384, 3>(args, pl, st);
        if (pl.threads + 32 <= 384 && knob_minb == 4)
            return launch_sigma_a_tb<1, STAGE_PACK, 384, 4>(args, pl, st);
        return launch_sigma_a_tb<1, STAGE_PACK, 1024, 1>(args, pl, st);
    } else {
        return launch_sigma_a_tb<CPT, STAGE_PACK, 544, 1>(args, pl, st);
    }
}

template <int CPT>
static int launch_sigma(const SigmaArgs& args, const SigmaPlan& pl, cudaStream_t st) {
    const sqd_operator& op = args.op;
    // kernel B
    static bool cfg_b[64] = {false};
    const size_t smem_b = (size_t)(kRowsB * op.ldc + kWarpsB * kRowsB * 32) * sizeof(double) + 16;
    if (opt_in_smem(sigma_b_kernel, smem_b, cfg_b)) return -2;
    const int n_pos_slices = (op.ldc + 31) / 32;  // slices incl. the pad positions
    const int rows_owned = args.row_end - args.row_begin;
    if (rows_owned <= 0) return 0;
    dim3 grid_b((rows_owned + kRowsB - 1) / kRowsB, n_pos_slices);
    sigma_b_kernel<<<grid_b, kWarpsB * 32, smem_b, st>>>(args);
    if (check_launch("sigma_b_kernel")) return -2;
    // kernel A
    if (pl.stage_pack ? launch_sigma_a<CPT, true>(args, pl, st) : launch_sigma_a<CPT, false>(args, pl, st))
        return -2;
    if (op.plan.n_split > 0) {
        sigma_combine_kernel<<<op.plan.n_split, 256, 0, st>>>(args.done, op.plan, op.ldc, args.sigma,
                                                              args.row_begin, args.row_end, args.slot_ptr,
                                                              args.vec_stride);
        return check_launch("sigma_combine_kernel");
    }
    return 0;
}

static int launch_sigma_wide(const SigmaArgs& args, cudaStream_t st) {
    const int rows_owned = args.row_end - args.row_begin;
    if (rows_owned <= 0) return 0;
    // alpha single excitations per pass over a column's beta links: SQD_WIDE_BLOCK = 2 (default), 4 or 8
    // (1e8 determinants, (30e,30o): 228 / 249 / 410 ms per build -- more rows of c in flight thrash L1)
    static const int knob_block = env_int("SQD_WIDE_BLOCK", 2);
    const int wb = (knob_block == 4 || knob_block == 8) ? knob_block : 2;
    static bool cfg_w[3][64] = {};
    const size_t smem = (size_t)wb * args.op.ldg * sizeof(double);
    const unsigned grid = (unsigned)rows_owned * (unsigned)((args.op.ldc + kWideCols - 1) / kWideCols);
    if (wb == 2) {
        if (opt_in_smem(sigma_wide_kernel<2>, smem, cfg_w[0])) return -2;
        sigma_wide_kernel<2><<<grid, kWideCols, smem, st>>>(args);
    } else if (wb == 8) {
        if (opt_in_smem(sigma_wide_kernel<8>, smem, cfg_w[2])) return -2;
        sigma_wide_kernel<8><<<grid, kWideCols, smem, st>>>(args);
    } else {
        if (opt_in_smem(sigma_wide_kernel<4>, smem, cfg_w[1])) return -2;
        sigma_wide_kernel<4><<<grid, kWideCols, smem, st>>>(args);
    }
    return check_launch("sigma_wide_kernel");
}

static thread_local long long* g_prof = nullptr;

int sigma_dispatch_rows(const sqd_operator* op, const double* d_c, double* d_sigma, const int* d_done,
                        int row_begin, int row_end, cudaStream_t st);
int sigma_dispatch_ctl(const sqd_operator* op, const double* d_cbase, double* d_sbase, const int* d_done,
                       const int* d_slot, long long stride, int in_graph, cudaStream_t st);
// v2 kernels (fermion_sigma2.cu)
int sigma2_dispatch_rows(const sqd_operator* op, const double* d_c, double* d_sigma, const int* d_done,
                         int row_begin, int row_end, const int* d_slot, long long stride, int in_graph,
                         cudaStream_t st);
int64_t sigma2_smem_bytes(const sqd_operator* op);

int sigma_dispatch_flag(const sqd_operator* op, const double* d_c, double* d_sigma,
                        const int* d_done, cudaStream_t st) {
    return sigma_dispatch_rows(op, d_c, d_sigma, d_done, 0, op->a.n, st);
}

static int sigma_dispatch_impl(const sqd_operator* op, const double* d_c, double* d_sigma, const int* d_done,
                               int row_begin, int row_end, const int* d_slot, long long stride, int in_graph,
                               cudaStream_t st);

int sigma_dispatch_rows(const sqd_operator* op, const double* d_c, double* d_sigma, const int* d_done,
                        int row_begin, int row_end, cudaStream_t st) {
    return sigma_dispatch_impl(op, d_c, d_sigma, d_done, row_begin, row_end, nullptr, 0, 0, st);
}

// device-driven loop: vector = base + (*d_slot) * stride, for both the input and the output
int sigma_dispatch_ctl(const sqd_operator* op, const double* d_cbase, double* d_sbase, const int* d_done,
                       const int* d_slot, long long stride, int in_graph, cudaStream_t st) {
    return sigma_dispatch_impl(op, d_cbase, d_sbase, d_done, 0, op->a.n, d_slot, stride, in_graph, st);
}

static int sigma_dispatch_impl(const sqd_operator* op, const double* d_c, double* d_sigma, const int* d_done,
                               int row_begin, int row_end, const int* d_slot, long long stride, int in_graph,
                               cudaStream_t st) {
    if (op->v2.enabled)
        return sigma2_dispatch_rows(op, d_c, d_sigma, d_done, row_begin, row_end, d_slot, stride, in_graph, st);
    SigmaPlan pl;
    SQD_REQUIRE(op->ldc % 2 == 0 && op->ldg % 2 == 0 && op->ldc >= op->b.n,
                "sqd_sigma: ldc/ldg must be even and ldc >= nb");
    if (op->wide) {
        SQD_REQUIRE(row_begin >= 0 && row_end <= op->a.n && row_begin <= row_end, "sqd_sigma: bad row range");
        SigmaArgs wargs{*op, d_c, d_sigma, d_done, nullptr, row_begin, row_end, d_slot, stride};
        return launch_sigma_wide(wargs, st);
    }
    SQD_REQUIRE(op->plan.n_chunks >= op->a.n && op->plan.chunk_row != nullptr,
                "sqd_sigma: the operator has no work plan (call sqd_sigma_plan_build first)");
    SQD_REQUIRE(op->plan.n_slots == 0 || op->plan.part != nullptr,
                "sqd_sigma: the plan needs a partial buffer of n_slots*ldc doubles");
    SQD_REQUIRE(op->bd.pack != nullptr && op->bb.pack != nullptr,
                "sqd_sigma: the operator has no SELL tables (call sqd_sell_build first)");
    SQD_REQUIRE(plan_sigma(op, &pl),
                "sqd_sigma: nb=%d (ldc=%d) with norb=%d does not fit the shared-memory row staging "
                "(limits: nb <= 5760 and 3*ldc + 2*ldg doubles <= 227 KB): set sqd_operator.wide",
                op->b.n, op->ldc, op->norb);
    SQD_REQUIRE(row_begin >= 0 && row_end <= op->a.n && row_begin <= row_end, "sqd_sigma: bad row range");
    SigmaArgs args{*op, d_c, d_sigma, d_done, g_prof, row_begin, row_end, d_slot, stride};
    switch (pl.CPT) {
        case 1: return launch_sigma<1>(args, pl, st);
        case 2: return launch_sigma<2>(args, pl, st);
        case 4: return launch_sigma<4>(args, pl, st);
        case 8: return launch_sigma<8>(args, pl, st);
        case 12: return launch_sigma<12>(args, pl, st);
    }
    set_error("sqd_sigma: no kernel instance for CPT=%d", pl.CPT);
    return -1;
}

}  // namespace sqd

using namespace sqd;

extern "C" {

int sqd_sigma_v1_supported(int ldc, int ldg) { return v1_shape_fits(ldc, ldg) ? 1 : 0; }

int64_t sqd_sigma_smem_bytes(const sqd_operator* op) {
    if (op->v2.enabled) return sigma2_smem_bytes(op);
    if (op->wide) return 0;
    SigmaPlan pl;
    if (!plan_sigma(op, &pl)) return -1;
    return (int64_t)pl.smem;
}

int sqd_sigma(const sqd_operator* op, const double* d_c, double* d_sigma, void* stream) {
    return sigma_dispatch_flag(op, d_c, d_sigma, nullptr, (cudaStream_t)stream);
}

int sqd_sigma_rows(const sqd_operator* op, const double* d_c, double* d_sigma, int row_begin,
                   int row_end, void* stream) {
    return sigma_dispatch_rows(op, d_c, d_sigma, nullptr, row_begin, row_end, (cudaStream_t)stream);
}

/* diagnostics: one sigma build with 8 clock64() stamps per kernel-A CTA written to d_prof
 * (int64[8 * n_chunks]: start, after table staging, after phase C, after phase D, end, #singles,
 * #doubles, SM id) */
int sqd_sigma_profile(const sqd_operator* op, const double* d_c, double* d_sigma, long long* d_prof,
                      void* stream) {
    g_prof = d_prof;
    const int rc = sigma_dispatch_flag(op, d_c, d_sigma, nullptr, (cudaStream_t)stream);
    g_prof = nullptr;
    return rc;
}

int sqd_sigma_plan_build(const sqd_spin_table* a, const sqd_spin_table* b, int cost_per_chunk,
                         int long_threshold, int max_chunks, int* d_chunk_row, int* d_chunk_beg,
                         int* d_chunk_end, int* d_chunk_slot, int* d_split_row, int* d_split_slot_beg,
                         int* d_split_n, int* d_long_idx, int* d_long_cols, int* d_counts,
                         int* h_counts, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(cost_per_chunk >= kSingleCost && cost_per_chunk % kSingleCost == 0,
                "sqd_sigma_plan_build: cost_per_chunk must be a positive multiple of %d", kSingleCost);
    SQD_REQUIRE(max_chunks >= a->n, "sqd_sigma_plan_build: max_chunks must be at least the number of rows");
    const size_t plan_smem = (size_t)(2 * a->n + 1 + b->n) * sizeof(int);
    const int stage = plan_smem <= 200 * 1024;
    static bool cfg_plan[64] = {false};
    if (stage && opt_in_smem(sigma_plan_kernel, plan_smem, cfg_plan)) return -2;
    sigma_plan_kernel<<<1, 256, stage ? plan_smem : 0, st>>>(
        *a, *b, cost_per_chunk, long_threshold, max_chunks, d_chunk_row, d_chunk_beg, d_chunk_end,
        d_chunk_slot, d_split_row, d_split_slot_beg, d_split_n, d_long_idx, d_long_cols, d_counts, stage);
    if (check_launch("sigma_plan_kernel")) return -2;
    if (h_counts == nullptr) return 0;  // the caller reads d_counts itself
    return read_back(h_counts, d_counts, 4 * sizeof(int), st);
}

int sqd_sell_build(const sqd_spin_table* t, int mode, const int* d_long_idx, int capacity, int* d_perm,
                   int* d_len, int* d_slice_ptr, uint32_t* d_pack, double* d_val, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(mode == 0 || mode == 1, "sqd_sell_build: mode must be 0 or 1");
    SQD_REQUIRE(mode == 0 || d_val != nullptr, "sqd_sell_build: mode 1 needs the value array");
    const int n = t->n;
    SQD_REQUIRE(n > 0 && capacity >= 32 * n, "sqd_sell_build: capacity must be at least nnz + 32*n");
    sell_rank_kernel<<<(n + 127) / 128, 128, 0, st>>>(*t, mode, d_long_idx, d_perm, d_len);
    sell_slice_kernel<<<1, 32, 0, st>>>(n, d_len, d_slice_ptr);
    const int ns = (n + 31) / 32;
    sell_fill_kernel<<<(ns * 32 + 127) / 128, 128, 0, st>>>(*t, mode, d_perm, d_len, d_slice_ptr, d_pack,
                                                           d_val);
    return check_launch("sell kernels", 3);
}

}  // extern "C"

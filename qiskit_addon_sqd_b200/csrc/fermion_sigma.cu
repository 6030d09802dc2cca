// CI sigma-vector build  sigma = O c  in the product space of sampled alpha x beta strings.
//
// Replaces pyscf selected_ci.contract_2e (SCIcontract_2e_aaaa + SCIcontract_2e_bbaa, recalled) reached
// from qiskit_addon_sqd/fermion.py:721-723, 810-818.  pyscf resolves the two-body operator through
// (N-2)-electron intermediates and dense dgemms, O(n_det * npair^2).  The B200 design applies the
// projected operator directly from the in-set excitation tables, O(n_det * links):
//
//   sigma[a,b] = diag[a,b] c[a,b]
//              + sum_{b' in T_b(b)}  ( Hb[b,b'] + [single] sgn_b Wa[a, rs] ) c[a, b']      (phase B)
//              + sum_{a' in D_a(a)}    Ha[a,a'] c[a', b]                                     (phase C)
//              + sum_{a' in S_a(a)}  ( Ha[a,a'] + sgn_a Wb[pq, b] ) c[a', b]                 (phase D)
//              + sum_{a' in S_a(a)} sgn_a sum_{b' in S_b(b)} sgn_b g_ab[pq, rs] c[a', b']    (phase D)
//
// S = in-set single excitations, D = in-set doubles, T = S u D.  One CTA owns R rows `a` and every
// column; a thread owns CPT columns and keeps R*CPT accumulators in registers, so each sigma element
// is produced by exactly one thread in a fixed order (bit-reproducible, no atomics).
// The rows c[a,:] and, per alpha single excitation, the row c[a',:] together with the integral row
// g_ab[pq,:] are staged in shared memory by the bulk-copy engine (cp.async.bulk + mbarrier, SASS
// UBLKCP), double-buffered so the copy of excitation k+1 overlaps the gathers of excitation k.
#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

struct SigmaArgs {
    sqd_operator op;
    const double* c;
    double* sigma;
    const int* done;  // optional device flag: non-zero -> the launch is a no-op (Davidson finished)
};

__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kMaxStages = 8;
constexpr int kUnroll = 4;
constexpr int kMaxLong = SQD_MAX_LONG_COLUMNS;

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
__global__ void sigma_plan_kernel(const sqd_spin_table A, const sqd_spin_table B, int cost_per_chunk,
                                  int long_threshold, int max_chunks, int* __restrict__ chunk_row,
                                  int* __restrict__ chunk_beg, int* __restrict__ chunk_end,
                                  int* __restrict__ chunk_slot, int* __restrict__ split_row,
                                  int* __restrict__ split_slot_beg, int* __restrict__ split_n,
                                  int* __restrict__ long_idx, int* __restrict__ long_cols,
                                  int* __restrict__ counts) {
    // setup-time, O(na + nb) sequential work: a single thread keeps the order trivially deterministic
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int nc = 0, nslots = 0, nsplit = 0;
    for (int a = 0; a < A.n; ++a) {
        const int beg = A.row_ptr[a], ns = A.n_single[a], end = A.row_ptr[a + 1];
        const int nd = end - beg - ns;
        const int cost = 4 * ns + nd;
        int k = (cost + cost_per_chunk - 1) / cost_per_chunk;
        if (k < 1) k = 1;
        if (nc + k > max_chunks) k = 1;  // cannot happen with the documented bound; stay correct anyway
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
            const int e0 = p0 < 4 * ns ? p0 / 4 : ns + (p0 - 4 * ns);
            const int e1 = p1 < 4 * ns ? p1 / 4 : ns + (p1 - 4 * ns);
            chunk_row[nc] = a;
            chunk_beg[nc] = beg + e0;
            chunk_end[nc] = beg + e1;
            chunk_slot[nc] = k > 1 ? nslots++ : -1;
            ++nc;
        }
    }
    int nlong = 0;
    for (int b = 0; b < B.n; ++b) {
        const bool is_long = B.n_single[b] > long_threshold && nlong < kMaxLong;
        long_idx[b] = is_long ? nlong : -1;
        if (is_long) long_cols[nlong++] = b;
    }
    counts[0] = nc;
    counts[1] = nslots;
    counts[2] = nsplit;
    counts[3] = nlong;
}

// sigma[a,:] = sum over the row's chunk partials, in chunk order
__global__ void sigma_combine_kernel(const int* __restrict__ done, const sqd_sigma_plan pl, int ldc,
                                     double* __restrict__ sigma) {
    if (done != nullptr && *done != 0) return;
    const int j = blockIdx.x;
    const int a = pl.split_row[j], s0 = pl.split_slot_beg[j], k = pl.split_n[j];
    for (int b = threadIdx.x; b < ldc; b += blockDim.x) {
        double acc = pl.part[(size_t)s0 * ldc + b];
        for (int c = 1; c < k; ++c) acc += pl.part[(size_t)(s0 + c) * ldc + b];
        sigma[(size_t)a * ldc + b] = acc;
    }
}

// Warp-specialised: the LAST warp of the CTA is the producer (one lane drives the bulk-copy engine
// through an NST-deep ring of {c[a',:], g_ab[pq,:]} stages guarded by full/empty mbarriers); all other
// warps are consumers.  One CTA per chunk.
template <int CPT>
__global__ void sigma_kernel(const SigmaArgs P, const int NST) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (P.done != nullptr && *P.done != 0) return;
    const sqd_operator& op = P.op;
    const sqd_sigma_plan& pl = op.plan;
    const int nb = op.b.n, ldc = op.ldc, ldg = op.ldg;
    const int ncons = blockDim.x - 32;  // consumer threads
    const int nwarp_c = ncons >> 5;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int a = pl.chunk_row[blockIdx.x];
    const int cbeg = pl.chunk_beg[blockIdx.x], cend = pl.chunk_end[blockIdx.x];
    const int slot = pl.chunk_slot[blockIdx.x];
    const int row_beg = op.a.row_ptr[a];
    const int single_end = row_beg + op.a.n_single[a];
    const bool first = cbeg == row_beg;          // the first chunk of a row also owns phase B
    const int it_beg = min(cbeg, single_end), it_end = min(cend, single_end);   // phase D range
    const int db_beg = max(cbeg, single_end), db_end = max(cend, single_end);   // phase C range
    const int n_items = it_end - it_beg;
    const int n_long = pl.n_long;

    double* Cs = reinterpret_cast<double*>(smem_raw);   // [ldc]   row a of c
    double* stage = Cs + ldc;                            // NST x ([ldc] row c[a',:] + [ldg] row g_ab[pq,:])
    const int stage_len = ldc + ldg;
    double* acc_long = stage + (size_t)NST * stage_len;  // [kMaxLong]
    uint64_t* bars = reinterpret_cast<uint64_t*>(acc_long + kMaxLong);
    uint64_t* full = bars;                  // [NST]
    uint64_t* empty = bars + kMaxStages;    // [NST]
    uint64_t* rows_bar = bars + 2 * kMaxStages;

    const bool ham = op.use_same_spin != 0;

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nwarp_c);
        }
        mbar_init(rows_bar, 1);
        mbar_fence_init();
    }
    if (tid < kMaxLong) acc_long[tid] = 0.0;
    __syncthreads();

    // =========================== producer warp ===========================
    if (tid >= ncons) {
        if (tid == ncons) {
            if (first) {
                mbar_expect_tx(rows_bar, (uint32_t)(ldc * sizeof(double)));
                bulk_g2s(Cs, P.c + (size_t)a * ldc, (uint32_t)(ldc * sizeof(double)), rows_bar);
            }
            for (int item = 0; item < n_items; ++item) {
                const int s = item % NST;
                if (item >= NST) mbar_wait_parity(&empty[s], (uint32_t)(((item / NST) - 1) & 1));
                const uint32_t ap = op.a.col[it_beg + item];
                const uint32_t pq = op.a.meta[it_beg + item] & 0x7fffffffu;
                double* dst = stage + (size_t)s * stage_len;
                mbar_expect_tx(&full[s], (uint32_t)((ldc + ldg) * sizeof(double)));
                bulk_g2s(dst, P.c + (size_t)ap * ldc, (uint32_t)(ldc * sizeof(double)), &full[s]);
                bulk_g2s(dst + ldc, op.gab + (size_t)pq * ldg, (uint32_t)(ldg * sizeof(double)),
                         &full[s]);
            }
        }
        return;
    }

    // =========================== consumer warps ==========================
    double acc[CPT];
    int bs_beg[CPT], bs_n[CPT], lidx[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int b = tid + c * ncons;
        acc[c] = 0.0;
        bs_beg[c] = b < nb ? op.b.row_ptr[b] : 0;
        bs_n[c] = b < nb ? op.b.n_single[b] : 0;
        lidx[c] = b < nb ? pl.long_idx[b] : -1;
    }

    // ---- phase C (no shared memory): alpha doubles, coalesced row streaming, kUnroll loads in flight
    if (ham) {
        for (int e0 = db_beg; e0 < db_end; e0 += kUnroll) {
            double v[kUnroll];
            const double* crow[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int e = min(e0 + u, db_end - 1);
                v[u] = (e0 + u < db_end) ? __ldg(op.a.val + e) : 0.0;
                crow[u] = P.c + (size_t)__ldg(op.a.col + e) * ldc;
            }
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int b = tid + c * ncons;
                if (b < nb) {
                    double x[kUnroll];
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) x[u] = __ldg(crow[u] + b);
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) acc[c] = fma(v[u], x[u], acc[c]);
                }
            }
        }
    }

    // ---- phase B (first chunk of the row): diagonal + beta excitations inside the staged row --------
    if (first) {
        mbar_wait_parity(rows_bar, 0);
        const double* wa = op.Wa ? op.Wa + (size_t)a * ldg : nullptr;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int b = tid + c * ncons;
            if (b < nb) {
                acc[c] = fma(op.diag[(size_t)a * ldc + b], Cs[b], acc[c]);
                if (lidx[c] < 0) {
                    const int beg = bs_beg[c], ns = bs_n[c], end = op.b.row_ptr[b + 1];
#pragma unroll 2
                    for (int e = beg; e < beg + ns; ++e) {
                        const uint32_t pk = __ldg(op.b.pack + e);
                        double coef = ham ? __ldg(op.b.val + e) : 0.0;
                        if (wa) {
                            const double w = __ldg(wa + ((pk >> 19) & 0xfffu));
                            coef += (pk >> 31) ? -w : w;
                        }
                        acc[c] = fma(coef, Cs[pk & 0x7ffffu], acc[c]);
                    }
                    if (ham) {
                        for (int e0 = beg + ns; e0 < end; e0 += kUnroll) {
                            uint32_t bp[kUnroll];
                            double v[kUnroll];
#pragma unroll
                            for (int u = 0; u < kUnroll; ++u) {
                                const int e = min(e0 + u, end - 1);
                                bp[u] = __ldg(op.b.col + e);
                                v[u] = (e0 + u < end) ? __ldg(op.b.val + e) : 0.0;
                            }
#pragma unroll
                            for (int u = 0; u < kUnroll; ++u) acc[c] = fma(v[u], Cs[bp[u]], acc[c]);
                        }
                    }
                }
            }
        }
        // long columns: one warp per column, lanes stride the whole list (singles then doubles)
        for (int li = warp; li < n_long; li += nwarp_c) {
            const int b = pl.long_cols[li];
            const int beg = op.b.row_ptr[b], ns = op.b.n_single[b], end = op.b.row_ptr[b + 1];
            double part = 0.0;
            for (int e = beg + lane; e < end; e += 32) {
                const uint32_t pk = __ldg(op.b.pack + e);
                double coef = ham ? __ldg(op.b.val + e) : 0.0;
                if (e < beg + ns) {
                    if (wa) {
                        const double w = __ldg(wa + ((pk >> 19) & 0xfffu));
                        coef += (pk >> 31) ? -w : w;
                    }
                    part = fma(coef, Cs[pk & 0x7ffffu], part);
                } else if (ham) {
                    part = fma(coef, Cs[pk], part);
                }
            }
            part = warp_sum(part);
            if (lane == 0) acc_long[li] += part;
        }
    }

    // ---- phase D: alpha singles through the staged ring, beta singles gathered from shared memory ---
    for (int item = 0; item < n_items; ++item) {
        const int s = item % NST;
        const uint32_t m = __ldg(op.a.meta + it_beg + item);
        const uint32_t pq = m & 0x7fffffffu;
        const double sa = (m >> 31) ? -1.0 : 1.0;
        const double va = ham ? __ldg(op.a.val + it_beg + item) : 0.0;
        const double* Cn = stage + (size_t)s * stage_len;
        const double* gs = Cn + ldc;
        mbar_wait_parity(&full[s], (uint32_t)((item / NST) & 1));
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int b = tid + c * ncons;
            if (b < nb) {
                double sum = 0.0;
                if (lidx[c] < 0) {
                    const int e1 = bs_beg[c] + bs_n[c];
#pragma unroll 4
                    for (int e = bs_beg[c]; e < e1; ++e) {
                        const uint32_t pk = __ldg(op.b.pack + e);
                        const double t = gs[(pk >> 19) & 0xfffu] * Cn[pk & 0x7ffffu];
                        sum += (pk >> 31) ? -t : t;
                    }
                }
                double coef = va;
                if (op.Wb) coef = fma(sa, __ldg(op.Wb + (size_t)pq * ldc + b), coef);
                acc[c] += sa * sum + coef * Cn[b];
            }
        }
        for (int li = warp; li < n_long; li += nwarp_c) {
            const int b = pl.long_cols[li];
            const int beg = op.b.row_ptr[b], e1 = beg + op.b.n_single[b];
            double part = 0.0;
            for (int e = beg + lane; e < e1; e += 32) {
                const uint32_t pk = __ldg(op.b.pack + e);
                const double t = gs[(pk >> 19) & 0xfffu] * Cn[pk & 0x7ffffu];
                part += (pk >> 31) ? -t : t;
            }
            part = warp_sum(part);
            if (lane == 0) acc_long[li] += sa * part;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);  // this warp is done with the stage
    }

    // ---- store: fold the cooperatively reduced long columns back into their owner threads ----------
    if (n_long > 0) asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
    double* out = slot < 0 ? P.sigma + (size_t)a * ldc : pl.part + (size_t)slot * ldc;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int b = tid + c * ncons;
        if (b < ldc) {
            double v = 0.0;
            if (b < nb) v = acc[c] + (lidx[c] >= 0 ? acc_long[lidx[c]] : 0.0);
            out[b] = v;
        }
    }
}

struct SigmaPlan {
    int CPT, threads, stages;
    size_t smem;
};

static bool plan_sigma(const sqd_operator* op, SigmaPlan* pl) {
    const int ldc = op->ldc;
    int threads, cpt;
    // `threads` counts the consumer threads; one more warp (the producer) is added at launch
    if (ldc <= 992) {
        threads = ((ldc + 31) / 32) * 32;
        if (threads < 32) threads = 32;
        cpt = 1;
        // fewer, fatter threads once the row is long: keeps more CTAs resident per SM
        if (ldc > 512) {
            threads = ((ldc / 2 + 31) / 32) * 32;
            cpt = 2;
        }
    } else {
        threads = 992;
        cpt = (ldc + threads - 1) / threads;
    }
    const int cpt_t = cpt <= 1 ? 1 : cpt <= 2 ? 2 : cpt <= 4 ? 4 : cpt <= 8 ? 8 : cpt <= 12 ? 12 : 0;
    if (cpt_t == 0) return false;
    auto smem_of = [&](int nst) {
        return (size_t)(ldc + nst * (ldc + op->ldg) + kMaxLong) * sizeof(double) +
               (2 * kMaxStages + 1) * sizeof(uint64_t);
    };
    if (smem_of(2) > 227 * 1024) return false;
    // ring depth: as deep as fits in ~56 KB (keeps 4 CTAs per SM resident), at least 2, at most 8
    int nst = 2;
    while (nst < kMaxStages && smem_of(nst + 1) <= 56 * 1024) ++nst;
    pl->CPT = cpt_t;
    pl->threads = threads;
    pl->stages = nst;
    pl->smem = smem_of(nst);
    return true;
}

template <int CPT>
static int launch_sigma(const SigmaArgs& args, const SigmaPlan& pl, cudaStream_t st) {
    auto kern = sigma_kernel<CPT>;
    static bool configured[64] = {false};  // per device: opt in to > 48 KB dynamic shared memory
    if (pl.smem > 48 * 1024) {
        int dev = 0;
        SQD_CUDA_OK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !configured[dev]) {
            SQD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(227 * 1024)));
            if (dev >= 0 && dev < 64) configured[dev] = true;
        }
    }
    kern<<<args.op.plan.n_chunks, pl.threads + 32, pl.smem, st>>>(args, pl.stages);
    if (check_launch("sigma_kernel")) return -2;
    if (args.op.plan.n_split > 0) {
        sigma_combine_kernel<<<args.op.plan.n_split, 256, 0, st>>>(args.done, args.op.plan, args.op.ldc,
                                                                  args.sigma);
        return check_launch("sigma_combine_kernel");
    }
    return 0;
}

int sigma_dispatch_flag(const sqd_operator* op, const double* d_c, double* d_sigma,
                        const int* d_done, cudaStream_t st) {
    SigmaPlan pl;
    SQD_REQUIRE(op->ldc % 2 == 0 && op->ldg % 2 == 0 && op->ldc >= op->b.n,
                "sqd_sigma: ldc/ldg must be even and ldc >= nb");
    SQD_REQUIRE(op->plan.n_chunks >= op->a.n && op->plan.chunk_row != nullptr,
                "sqd_sigma: the operator has no work plan (call sqd_sigma_plan_build first)");
    SQD_REQUIRE(op->plan.n_slots == 0 || op->plan.part != nullptr,
                "sqd_sigma: the plan needs a partial buffer of n_slots*ldc doubles");
    SQD_REQUIRE(plan_sigma(op, &pl),
                "sqd_sigma: nb=%d (ldc=%d) with norb=%d does not fit the shared-memory row staging "
                "(limit: 3*ldc + 2*ldg doubles <= 227 KB, ldc <= 11904)",
                op->b.n, op->ldc, op->norb);
    SigmaArgs args{*op, d_c, d_sigma, d_done};
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

int64_t sqd_sigma_smem_bytes(const sqd_operator* op) {
    SigmaPlan pl;
    if (!plan_sigma(op, &pl)) return -1;
    return (int64_t)pl.smem;
}

int sqd_sigma(const sqd_operator* op, const double* d_c, double* d_sigma, void* stream) {
    return sigma_dispatch_flag(op, d_c, d_sigma, nullptr, (cudaStream_t)stream);
}

int sqd_sigma_plan_build(const sqd_spin_table* a, const sqd_spin_table* b, int cost_per_chunk,
                   int long_threshold, int max_chunks, int* d_chunk_row, int* d_chunk_beg,
                   int* d_chunk_end, int* d_chunk_slot, int* d_split_row, int* d_split_slot_beg,
                   int* d_split_n, int* d_long_idx, int* d_long_cols, int* d_counts, int* h_counts,
                   void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(cost_per_chunk >= 4 && cost_per_chunk % 4 == 0,
                "sqd_sigma_plan_build: cost_per_chunk must be a positive multiple of 4");
    SQD_REQUIRE(max_chunks >= a->n, "sqd_sigma_plan_build: max_chunks must be at least the number of rows");
    sigma_plan_kernel<<<1, 32, 0, st>>>(*a, *b, cost_per_chunk, long_threshold, max_chunks, d_chunk_row,
                                        d_chunk_beg, d_chunk_end, d_chunk_slot, d_split_row,
                                        d_split_slot_beg, d_split_n, d_long_idx, d_long_cols, d_counts);
    if (check_launch("sigma_plan_kernel")) return -2;
    SQD_CUDA_OK(cudaMemcpyAsync(h_counts, d_counts, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    SQD_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"

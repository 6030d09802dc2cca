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

template <int R, int CPT>
__global__ void sigma_kernel(const SigmaArgs P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (P.done != nullptr && *P.done != 0) return;
    const sqd_operator& op = P.op;
    const int na = op.a.n, nb = op.b.n, ldc = op.ldc, ldg = op.ldg, norb = op.norb;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int a0 = blockIdx.x * R;
    const int nrows = min(R, na - a0);

    double* Cs = reinterpret_cast<double*>(smem_raw);   // [R][ldc]   rows a0..a0+R-1 of c
    double* stage = Cs + (size_t)R * ldc;                // 2 x ([ldc] row c[a',:] + [ldg] row g_ab[pq,:])
    const int stage_len = ldc + ldg;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 2 * (size_t)stage_len);  // [3]

    const bool ham = op.use_same_spin != 0;

    // ---- producer state (thread 0 only): cursor over the alpha single excitations of the R rows ----
    int pr = 0;   // row cursor
    int pk = 0;   // link cursor inside row
    int n_items = 0;
    for (int r = 0; r < nrows; ++r) n_items += op.a.n_single[a0 + r];

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue_next = [&](int item) {
        // advance cursor to the next existing link
        while (pr < nrows && pk >= op.a.n_single[a0 + pr]) {
            ++pr;
            pk = 0;
        }
        const int e = op.a.row_ptr[a0 + pr] + pk;
        const uint32_t ap = op.a.col[e];
        const uint32_t pq = op.a.meta[e] & 0x7fffffffu;
        uint64_t* bar = &bars[item & 1];
        double* dst = stage + (size_t)(item & 1) * stage_len;
        mbar_expect_tx(bar, (uint32_t)((ldc + ldg) * sizeof(double)));
        bulk_g2s(dst, P.c + (size_t)ap * ldc, (uint32_t)(ldc * sizeof(double)), bar);
        bulk_g2s(dst + ldc, op.gab + (size_t)pq * ldg, (uint32_t)(ldg * sizeof(double)), bar);
        ++pk;
    };

    if (tid == 0) {
        mbar_expect_tx(&bars[2], (uint32_t)(nrows * ldc * sizeof(double)));
        for (int r = 0; r < nrows; ++r)
            bulk_g2s(Cs + (size_t)r * ldc, P.c + (size_t)(a0 + r) * ldc,
                     (uint32_t)(ldc * sizeof(double)), &bars[2]);
        if (n_items > 0) issue_next(0);
    }

    double acc[R][CPT];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[r][c] = 0.0;

    // ---- phase C first (needs no shared memory): alpha doubles, coalesced row streaming ----------
    if (ham) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r < nrows) {
                const int a = a0 + r;
                const int beg = op.a.row_ptr[a] + op.a.n_single[a], end = op.a.row_ptr[a + 1];
                for (int e = beg; e < end; ++e) {
                    const double v = op.a.val[e];
                    const double* crow = P.c + (size_t)op.a.col[e] * ldc;
#pragma unroll
                    for (int c = 0; c < CPT; ++c) {
                        const int b = tid + c * nthr;
                        if (b < nb) acc[r][c] = fma(v, __ldg(crow + b), acc[r][c]);
                    }
                }
            }
        }
    }

    // ---- phase B: diagonal + beta excitations inside the staged rows ------------------------------
    while (!mbar_try_wait(&bars[2], 0)) {
    }
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int b = tid + c * nthr;
        if (b < nb) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (r < nrows)
                    acc[r][c] = fma(op.diag[(size_t)(a0 + r) * ldc + b], Cs[r * ldc + b], acc[r][c]);
            const int beg = op.b.row_ptr[b], ns = op.b.n_single[b], end = op.b.row_ptr[b + 1];
            for (int e = beg; e < beg + ns; ++e) {
                const uint32_t bp = op.b.col[e];
                const uint32_t m = op.b.meta[e];
                const uint32_t rs = m & 0x7fffffffu;
                const double sb = (m >> 31) ? -1.0 : 1.0;
                const double v = ham ? op.b.val[e] : 0.0;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (r < nrows) {
                        double coef = v;
                        if (op.Wa) coef = fma(sb, __ldg(op.Wa + (size_t)(a0 + r) * ldg + rs), coef);
                        acc[r][c] = fma(coef, Cs[r * ldc + bp], acc[r][c]);
                    }
                }
            }
            if (ham) {
                for (int e = beg + ns; e < end; ++e) {
                    const uint32_t bp = op.b.col[e];
                    const double v = op.b.val[e];
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (r < nrows) acc[r][c] = fma(v, Cs[r * ldc + bp], acc[r][c]);
                }
            }
        }
    }

    // ---- phase D: alpha singles, staged rows, beta singles gathered from shared memory -----------
    int item = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r < nrows) {
            const int a = a0 + r;
            const int beg = op.a.row_ptr[a], ns = op.a.n_single[a];
            for (int k = 0; k < ns; ++k, ++item) {
                if (tid == 0 && item + 1 < n_items) issue_next(item + 1);
                const uint32_t m = op.a.meta[beg + k];
                const uint32_t pq = m & 0x7fffffffu;
                const double sa = (m >> 31) ? -1.0 : 1.0;
                const double va = ham ? op.a.val[beg + k] : 0.0;
                const double* Cn = stage + (size_t)(item & 1) * stage_len;
                const double* gs = Cn + ldc;
                while (!mbar_try_wait(&bars[item & 1], (uint32_t)((item >> 1) & 1))) {
                }
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const int b = tid + c * nthr;
                    if (b < nb) {
                        const int bb = op.b.row_ptr[b], bn = op.b.n_single[b];
                        double s = 0.0;
                        for (int e = bb; e < bb + bn; ++e) {
                            const uint32_t mb = op.b.meta[e];
                            const double t = gs[mb & 0x7fffffffu] * Cn[op.b.col[e]];
                            s += (mb >> 31) ? -t : t;
                        }
                        double coef = va;
                        if (op.Wb) coef = fma(sa, __ldg(op.Wb + (size_t)pq * ldc + b), coef);
                        acc[r][c] += sa * s + coef * Cn[b];
                    }
                }
                __syncthreads();  // everybody is done with this stage before it is refilled
            }
        }
    }

    // ---- store -------------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (r < nrows) {
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int b = tid + c * nthr;
                if (b < ldc) P.sigma[(size_t)(a0 + r) * ldc + b] = (b < nb) ? acc[r][c] : 0.0;
            }
        }
    }
}

struct SigmaPlan {
    int R, CPT, threads;
    size_t smem;
};

static bool plan_sigma(const sqd_operator* op, SigmaPlan* pl) {
    const int nb = op->b.n, na = op->a.n, ldc = op->ldc;
    int threads, cpt;
    if (ldc <= 1024) {
        threads = ((ldc + 31) / 32) * 32;
        if (threads < 64) threads = 64;
        cpt = 1;
        // fewer, fatter threads once the row is long: keeps more CTAs resident per SM
        if (ldc > 512) {
            threads = ((ldc / 2 + 31) / 32) * 32;
            cpt = 2;
        }
    } else {
        threads = 1024;
        cpt = (ldc + threads - 1) / threads;
    }
    int cpt_t = cpt <= 1 ? 1 : cpt <= 2 ? 2 : cpt <= 4 ? 4 : cpt <= 8 ? 8 : cpt <= 12 ? 12 : 0;
    if (cpt_t == 0) return false;
    (void)nb;
    // rows per CTA: reuse of the beta tables across rows, but keep >= ~2 waves of CTAs
    int R = 1;
    const int cands[2] = {4, 2};
    for (int ci = 0; ci < 2; ++ci) {
        const int cand = cands[ci];
        if (cand * cpt_t <= 8 && (na + cand - 1) / cand >= 2 * kNumSMs) {
            R = cand;
            break;
        }
    }
    auto smem_of = [&](int r) {
        return (size_t)(r * ldc + 2 * (ldc + op->ldg)) * sizeof(double) + 3 * sizeof(uint64_t);
    };
    while (R > 1 && smem_of(R) > 200 * 1024) R >>= 1;
    if (smem_of(R) > 227 * 1024) return false;
    pl->R = R;
    pl->CPT = cpt_t;
    pl->threads = threads;
    pl->smem = smem_of(R);
    return true;
}

template <int R, int CPT>
static int launch_sigma(const SigmaArgs& args, const SigmaPlan& pl, cudaStream_t st) {
    auto kern = sigma_kernel<R, CPT>;
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
    const int grid = (args.op.a.n + R - 1) / R;
    kern<<<grid, pl.threads, pl.smem, st>>>(args);
    return check_launch("sigma_kernel");
}

int sigma_dispatch_flag(const sqd_operator* op, const double* d_c, double* d_sigma,
                        const int* d_done, cudaStream_t st) {
    SigmaPlan pl;
    SQD_REQUIRE(op->ldc % 2 == 0 && op->ldg % 2 == 0 && op->ldc >= op->b.n,
                "sqd_sigma: ldc/ldg must be even and ldc >= nb");
    SQD_REQUIRE(plan_sigma(op, &pl),
                "sqd_sigma: nb=%d (ldc=%d) with norb=%d does not fit the shared-memory row staging "
                "(limit: 2*(ldc+ldg)+ldc doubles <= 227 KB, ldc <= 12288)",
                op->b.n, op->ldc, op->norb);
    SigmaArgs args{*op, d_c, d_sigma, d_done};
#define SQD_SIGMA_CASE(r, c) \
    if (pl.R == r && pl.CPT == c) return launch_sigma<r, c>(args, pl, st);
    SQD_SIGMA_CASE(1, 1)
    SQD_SIGMA_CASE(2, 1)
    SQD_SIGMA_CASE(4, 1)
    SQD_SIGMA_CASE(1, 2)
    SQD_SIGMA_CASE(2, 2)
    SQD_SIGMA_CASE(4, 2)
    SQD_SIGMA_CASE(1, 4)
    SQD_SIGMA_CASE(2, 4)
    SQD_SIGMA_CASE(1, 8)
    SQD_SIGMA_CASE(1, 12)
#undef SQD_SIGMA_CASE
    set_error("sqd_sigma: no kernel instance for R=%d CPT=%d", pl.R, pl.CPT);
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

}  // extern "C"

// Single-root Davidson eigensolver with every step on the device.
//
// Replaces pyscf.lib.davidson1 as driven by selected_ci.kernel_fixed_space (recalled semantics:
// unit start vector at argmin(hdiag), preconditioner r/(hdiag - theta + 1e-4), max_space 12 with a
// collapse onto the Ritz vector, converged when |d theta| < tol and |r| < sqrt(tol); a stalled run
// returns the current vector and is not an error) -- reference call sites
// qiskit_addon_sqd/fermion.py:721-723 and :810-818.
//
// B200 design: the host only enqueues kernels; the Rayleigh-Ritz problem (parallel-order Jacobi in one
// CTA), the convergence test and the Gram-Schmidt coefficients live in a device-side state block, and
// every kernel returns immediately once the state says "done".  The host polls that flag every
// `check_every` cycles, so a cycle costs no host round trip.  The O(n_det) work is three fused
// streaming passes per cycle:
//   gram      : new column of V^T W                                  reads (m+1) vectors
//   residual  : x = V y, Hx = W y, r = Hx - theta x, t = r/(hdiag - theta + shift),
//               <v_i,t>, |r|^2, |t|^2 (+ collapse of V,W on restart)    reads (2m+1), writes 2..4
//   ortho1/2  : two classical Gram-Schmidt passes (the second fused with normalisation)
// All reductions are two-stage with a fixed order -> results are bit-reproducible run to run.
#include <math.h>
#include <stdlib.h>

#include <functional>
#include <vector>
#include <type_traits>

#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

int sigma_dispatch_flag(const sqd_operator* op, const double* d_c, double* d_sigma,
                        const int* d_done, cudaStream_t st);
int sigma_dispatch_rows(const sqd_operator* op, const double* d_c, double* d_sigma, const int* d_done,
                        int row_begin, int row_end, cudaStream_t st);
int sigma_dispatch_ctl(const sqd_operator* op, const double* d_cbase, double* d_sbase, const int* d_done,
                       const int* d_slot, long long stride, int in_graph, cudaStream_t st);
int nccl_allreduce_sum_f64(void* comm, double* buf, int64_t n, cudaStream_t st);
int nccl_allgather_blocks(void* comm, double* buf, const long long* offs, int world, cudaStream_t st);
int csr_matvec_flag(const int* d_done, int64_t d, const int32_t* row_ptr, const int32_t* col,
                    const double* val, const double* x, double* y, cudaStream_t st);
int csr_diag_embed(int64_t d, const int32_t* row_ptr, const int32_t* col, const double* val,
                   double* out, cudaStream_t st);

constexpr int kMaxS = SQD_MAX_SPACE;
constexpr int kRedBlocks = 2 * kNumSMs;  // CTAs of every streaming reduction pass
constexpr int kRedThreads = 256;
constexpr int kPartialRows = kMaxS + 4;
constexpr int kKeep = 4;  // Ritz vectors kept by a thick restart

// Restart policy when the basis is full (m == max_space):
//   mode 1 (default): collapse onto TWO vectors -- the current Ritz vector and the previous cycle's Ritz vector
//     orthogonalised against it ("GD+1" / locally optimal restart: the pair spans the conjugate-gradient-like
//     recurrence, so convergence barely notices the restart).  Needs only the lowest Ritz pair, which the
//     Rayleigh-quotient iteration delivers in a few microseconds;
//   mode 0: thick restart on the min(kKeep, M/3) lowest Ritz vectors, which needs the full decomposition of the
//     projected matrix (parallel Jacobi by one warp: ~150 us at m = 12, as long as a whole cycle).
static int env_knob(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
inline int restart_mode_knob() {
    static const int mode = env_knob("SQD_RESTART_MODE", 1);
    return mode;
}
inline int restart_keep(int M) {
    if (restart_mode_knob() == 1) return M >= 3 ? 2 : 1;
    return M / 3 < 1 ? 1 : (M / 3 > kKeep ? kKeep : M / 3);
}

struct DavState {
    int status;  // 0 running, 1 converged, 2 linear dependency, 3 (host) max_cycle
    int cycles;
    double theta, theta_prev, rnorm, tnorm2, inv_norm;
    int best;
    int ord[kMaxS];             // eigenvalue order (ascending) of the current decomposition
    double lam[kMaxS];          // eigenvalues of the projected matrix V^T H V
    double Q[kMaxS * kMaxS];    // its eigenvectors (column j <-> lam[j])
    double y[kMaxS];
    double c1[kMaxS];
    double c2[kMaxS];
    double gcol[kMaxS];         // newest column of V^T H V
    double G[kMaxS * kMaxS];    // projected matrix V^T H V of the current basis (row stride kMaxS)
    int ticket[4];              // arrival counters of the "last CTA finishes the step" tails
    // device-driven loop (one CUDA graph whose WHILE node replays the cycle): the cycle's control variables
    // live here and every kernel of the cycle is launched with the same arguments
    int m, slot;                // size of the basis, slot of the newest basis vector
    int M, q_keep, max_cycle;   // max_space, Ritz vectors kept by a thick restart, cycle limit
    int n_jacobi, n_rqi_iter;   // diagnostics: full decompositions, Rayleigh-quotient iterations
    int restart_mode;           // 1: keep {Ritz vector, previous Ritz vector} (no decomposition), 0: thick restart
};

// Control variables of a cycle.  Host-driven loop: passed as kernel arguments (m >= 0).  Device-driven
// loop: m < 0 and they are read from the state (the advance kernel at the end of the cycle updates them).
struct Ctl {
    int m, restart, me, slot;
};
__device__ __forceinline__ Ctl read_ctl(const DavState* st, int m_arg, int restart_arg) {
    Ctl c;
    if (m_arg >= 0) {
        c.m = m_arg;
        c.restart = restart_arg;
        c.slot = -1;
    } else {
        c.m = st->m;
        c.restart = (c.m == st->M) ? st->q_keep : 0;
        c.slot = st->slot;
    }
    c.me = c.restart ? c.restart : c.m;
    return c;
}

// The scalar steps of a cycle (Rayleigh-Ritz, convergence test, Gram-Schmidt coefficients) need the sums
// over all CTAs of the streaming pass before them.  Instead of a separate single-CTA launch, the LAST CTA
// of the streaming kernel to arrive (ticket counter, after a __threadfence) runs the step in its tail:
// same fixed summation order as before (the per-CTA partials are reduced in index order), three launches
// and three launch gaps fewer per cycle.
template <int MR>
__device__ void rayleigh_ritz_body(DavState* st, const double* partials, int nblk, int m, int need_full);
// the same bookkeeping for a cycle graph that the HOST replays (no conditional node)
__global__ void advance_plain_kernel(DavState* st) {
    if (threadIdx.x != 0 || st->status != 0) return;
    const int m = st->m;
    const int me = (m == st->M) ? st->q_keep : m;
    st->m = me + 1;
    st->slot = me;
}

__device__ void convergence_body(DavState* st, const double* partials, int nblk, int m, int restart,
                                 double tol, double tol_residual);
__device__ void norm_body(DavState* st, const double* partials, int nblk, int m, double lindep);

// true on every thread of exactly one CTA of the grid: the last one to get here.  Thread 0 must have
// issued the CTA's global writes before the call.
__device__ __forceinline__ bool last_block(int* ticket) {
    __shared__ int is_last;
    if (threadIdx.x == 0) {
        __threadfence();
        const int t = atomicAdd(ticket, 1);
        is_last = t == (int)gridDim.x - 1;
        if (is_last) *ticket = 0;  // ready for the next launch
    }
    __syncthreads();
    if (is_last) __threadfence();
    return is_last != 0;
}

// ---------------------------------------------------------------------------------------------
// streaming passes
// ---------------------------------------------------------------------------------------------
template <int MV>
__global__ void __launch_bounds__(kRedThreads)
gram_kernel(DavState* st, const double* __restrict__ V, const double* __restrict__ w, int64_t n, int m_arg,
            double* partials, int ritz_mode) {
    if (st->status != 0) return;
    // device-driven loop: w is the base of W and the newest vector sits in slot st->slot; a restart cycle
    // needs the full decomposition before the residual pass, so its tail does the whole Rayleigh-Ritz step
    // ritz_mode: != 0 on a restart cycle of the host-driven loop (the residual pass then needs the full
    // decomposition); the device-driven loop derives it from the state
    const Ctl ctl = read_ctl(st, m_arg, 0);
    const int m = ctl.m;
    if (m_arg < 0) {
        w += (int64_t)ctl.slot * n;
        ritz_mode = ctl.restart != 0;
    }
    __shared__ double red[MV * (kRedThreads / 32)];
    double acc[MV];
#pragma unroll
    for (int i = 0; i < MV; ++i) acc[i] = 0.0;
    // 16-byte loads, all basis vectors of an element pair in flight before the first FMA (n is even and
    // every vector is 16-byte aligned): at 1e7 determinants these passes stream from HBM
    const int64_t n2 = n >> 1;
    const double2* w2 = reinterpret_cast<const double2*>(w);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n2;
         j += (int64_t)gridDim.x * blockDim.x) {
        const double2 wj = w2[j];
        double2 v[MV];
#pragma unroll
        for (int i = 0; i < MV; ++i)
            if (i < m) v[i] = reinterpret_cast<const double2*>(V + (int64_t)i * n)[j];
#pragma unroll
        for (int i = 0; i < MV; ++i)
            if (i < m) acc[i] = fma(v[i].y, wj.y, fma(v[i].x, wj.x, acc[i]));
    }
    block_sum<MV>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < MV; ++i)
            if (i < m) partials[i * gridDim.x + blockIdx.x] = acc[i];
    }
    // tail: Rayleigh-Ritz on the finished Gram column.  ritz_mode 1: the whole step (decomposition and
    // lowest pair); 2: lowest pair only, the decomposition follows on the side stream
    if (last_block(&st->ticket[0])) rayleigh_ritz_body<(MV <= 16 ? MV : 16)>(st, partials, gridDim.x, m, ritz_mode);
}

template <int MV>
__global__ void __launch_bounds__(kRedThreads)
residual_kernel(DavState* st, double* __restrict__ V, double* __restrict__ W,
                const double* __restrict__ hdiag, int64_t n, int m_arg, int restart_arg, double level_shift,
                double* __restrict__ X, double* __restrict__ T, double* partials, double tol,
                double tol_residual) {
    if (st->status != 0) return;
    const Ctl ctl = read_ctl(st, m_arg, restart_arg);
    const int m = ctl.m, restart = ctl.restart;
    __shared__ double red[(MV + 2) * (kRedThreads / 32)];
    __shared__ double ys[MV];
    __shared__ double yk[kKeep][MV];  // thick restart: coefficient columns of the kept Ritz vectors 1..q-1
    if (threadIdx.x < MV) ys[threadIdx.x] = threadIdx.x < m ? st->y[threadIdx.x] : 0.0;
    if (restart > 1) {
        for (int idx = threadIdx.x; idx < kKeep * MV; idx += blockDim.x) {
            const int j = idx / MV, i = idx % MV;
            yk[j][i] = (j < restart && i < m) ? st->Q[i * kMaxS + st->ord[j]] : 0.0;
        }
    }
    __syncthreads();
    const double theta = st->theta;
    double acc[MV + 2];
#pragma unroll
    for (int i = 0; i < MV + 2; ++i) acc[i] = 0.0;
    // two elements per thread and iteration, 16-byte loads; the basis values stay in registers for the
    // projections, the W values are consumed as they arrive (a restart re-reads them: once per sweep)
    const int64_t n2 = n >> 1;
    const double2* hd2 = reinterpret_cast<const double2*>(hdiag);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n2;
         j += (int64_t)gridDim.x * blockDim.x) {
        double2 v[MV];
        double2 x = make_double2(0.0, 0.0), hx = make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < MV; ++i)
            v[i] = i < m ? reinterpret_cast<const double2*>(V + (int64_t)i * n)[j] : make_double2(0.0, 0.0);
#pragma unroll
        for (int i = 0; i < MV; ++i) {
            if (i < m) {
                const double2 wi = reinterpret_cast<const double2*>(W + (int64_t)i * n)[j];
                x.x = fma(ys[i], v[i].x, x.x);
                x.y = fma(ys[i], v[i].y, x.y);
                hx.x = fma(ys[i], wi.x, hx.x);
                hx.y = fma(ys[i], wi.y, hx.y);
            }
        }
        const double2 hd = hd2[j];
        const double rx = hx.x - theta * x.x, ry = hx.y - theta * x.y;
        double dx = hd.x - theta + level_shift, dy = hd.y - theta + level_shift;
        if (fabs(dx) < 1e-8) dx = dx < 0.0 ? -1e-8 : 1e-8;
        if (fabs(dy) < 1e-8) dy = dy < 0.0 ? -1e-8 : 1e-8;
        const double tx = rx / dx, ty = ry / dy;
        reinterpret_cast<double2*>(X)[j] = x;
        reinterpret_cast<double2*>(T)[j] = make_double2(tx, ty);
        if (restart) {
            // thick restart: the basis collapses onto the `restart` lowest Ritz vectors (element-wise; this
            // thread is the only reader and writer of elements 2j, 2j+1 of every basis vector).  All kept
            // vectors are formed before the first one is stored: they read the slots they overwrite.
            double2 xk[kKeep], hk[kKeep];
            for (int k = 1; k < restart; ++k) {
                xk[k] = make_double2(0.0, 0.0);
                hk[k] = make_double2(0.0, 0.0);
                for (int i = 0; i < m; ++i) {
                    const double2 wi = reinterpret_cast<const double2*>(W + (int64_t)i * n)[j];
                    const double2 vi = reinterpret_cast<const double2*>(V + (int64_t)i * n)[j];
                    const double c = yk[k][i];
                    xk[k].x = fma(c, vi.x, xk[k].x);
                    xk[k].y = fma(c, vi.y, xk[k].y);
                    hk[k].x = fma(c, wi.x, hk[k].x);
                    hk[k].y = fma(c, wi.y, hk[k].y);
                }
            }
            reinterpret_cast<double2*>(V)[j] = x;
            reinterpret_cast<double2*>(W)[j] = hx;
            for (int k = 1; k < restart; ++k) {
                reinterpret_cast<double2*>(V + (int64_t)k * n)[j] = xk[k];
                reinterpret_cast<double2*>(W + (int64_t)k * n)[j] = hk[k];
            }
        }
#pragma unroll
        for (int i = 0; i < MV; ++i)
            if (i < m) acc[i] = fma(v[i].y, ty, fma(v[i].x, tx, acc[i]));
        acc[MV] = fma(ry, ry, fma(rx, rx, acc[MV]));
        acc[MV + 1] = fma(ty, ty, fma(tx, tx, acc[MV + 1]));
    }
    block_sum<MV + 2>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < MV; ++i)
            if (i < m) partials[i * gridDim.x + blockIdx.x] = acc[i];
        partials[kMaxS * gridDim.x + blockIdx.x] = acc[MV];
        partials[(kMaxS + 1) * gridDim.x + blockIdx.x] = acc[MV + 1];
    }
    // tail: residual norm, convergence flag, first Gram-Schmidt coefficients (and the restart bookkeeping)
    if (last_block(&st->ticket[1])) convergence_body(st, partials, gridDim.x, m, restart, tol, tol_residual);
}

// T <- T - sum c1_i V_i ; partial <V_i, T>, |T|^2
template <int MV>
__global__ void __launch_bounds__(kRedThreads)
ortho1_kernel(DavState* st, const double* __restrict__ V, int64_t n, int m_arg, double* __restrict__ T,
              double* partials, double lindep) {
    if (st->status != 0) return;
    const int m = m_arg >= 0 ? m_arg : read_ctl(st, -1, 0).me;
    __shared__ double red[(MV + 1) * (kRedThreads / 32)];
    __shared__ double cs[MV];
    if (threadIdx.x < MV) cs[threadIdx.x] = threadIdx.x < m ? st->c1[threadIdx.x] : 0.0;
    __syncthreads();
    double acc[MV + 1];
#pragma unroll
    for (int i = 0; i < MV + 1; ++i) acc[i] = 0.0;
    const int64_t n2 = n >> 1;
    double2* T2 = reinterpret_cast<double2*>(T);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n2;
         j += (int64_t)gridDim.x * blockDim.x) {
        double2 v[MV];
        double2 t = T2[j];
#pragma unroll
        for (int i = 0; i < MV; ++i)
            if (i < m) v[i] = reinterpret_cast<const double2*>(V + (int64_t)i * n)[j];
#pragma unroll
        for (int i = 0; i < MV; ++i) {
            if (i < m) {
                t.x = fma(-cs[i], v[i].x, t.x);
                t.y = fma(-cs[i], v[i].y, t.y);
            }
        }
        T2[j] = t;
#pragma unroll
        for (int i = 0; i < MV; ++i)
            if (i < m) acc[i] = fma(v[i].y, t.y, fma(v[i].x, t.x, acc[i]));
        acc[MV] = fma(t.y, t.y, fma(t.x, t.x, acc[MV]));
    }
    block_sum<MV + 1>(acc, red);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < MV; ++i)
            if (i < m) partials[i * gridDim.x + blockIdx.x] = acc[i];
        partials[kMaxS * gridDim.x + blockIdx.x] = acc[MV];
    }
    // tail: second-pass coefficients, norm of the new basis vector, linear-dependency flag
    if (last_block(&st->ticket[2])) norm_body(st, partials, gridDim.x, m, lindep);
}

// V_new <- (T - sum c2_i V_i) * inv_norm
template <int MV>
__global__ void __launch_bounds__(kRedThreads)
ortho2_kernel(const DavState* __restrict__ st, const double* __restrict__ V, int64_t n, int m_arg,
              const double* __restrict__ T, double* __restrict__ vnew) {
    if (st->status != 0) return;
    // device-driven loop: vnew is the base of V, the new vector goes to slot `me`
    const int m = m_arg >= 0 ? m_arg : read_ctl(st, -1, 0).me;
    if (m_arg < 0) vnew += (int64_t)m * n;
    __shared__ double cs[MV];
    if (threadIdx.x < MV) cs[threadIdx.x] = threadIdx.x < m ? st->c2[threadIdx.x] : 0.0;
    __syncthreads();
    const double inv = st->inv_norm;
    const int64_t n2 = n >> 1;
    const double2* T2 = reinterpret_cast<const double2*>(T);
    double2* out2 = reinterpret_cast<double2*>(vnew);
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n2;
         j += (int64_t)gridDim.x * blockDim.x) {
        double2 t = T2[j];
        double2 v[MV];
#pragma unroll
        for (int i = 0; i < MV; ++i)
            if (i < m) v[i] = reinterpret_cast<const double2*>(V + (int64_t)i * n)[j];
#pragma unroll
        for (int i = 0; i < MV; ++i) {
            if (i < m) {
                t.x = fma(-cs[i], v[i].x, t.x);
                t.y = fma(-cs[i], v[i].y, t.y);
            }
        }
        out2[j] = make_double2(t.x * inv, t.y * inv);
    }
}

// ---------------------------------------------------------------------------------------------
// small (single CTA) steps
// ---------------------------------------------------------------------------------------------
// Sum of one row of per-CTA partials by a full warp: lane-strided loads (independent, pipelined) and a
// fixed xor-shuffle tree -> the same order on every run.  Result on all lanes.
__device__ __forceinline__ double reduce_partials(const double* partials, int row, int nblk) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int b = lane; b < nblk; b += 32) s += __ldcg(partials + row * nblk + b);  // written by other CTAs
    return warp_sum(s);
}

// Rayleigh-Ritz step in the tail of gram_kernel (one CTA, the last to arrive).
//
// Every cycle needs the LOWEST eigenpair of the projected matrix G (m x m, m <= max_space); only a restart
// cycle needs more (the q lowest Ritz vectors the basis collapses onto).  So:
//   * every cycle: Rayleigh-quotient iteration on G by one warp, started from the previous Ritz vector
//     padded with a zero (its Rayleigh quotient is the previous Ritz value theta_prev).  Two or three
//     solves of an m x m system (Gaussian elimination with partial pivoting, one row per lane) -- a few
//     microseconds.  Cauchy interlacing gives the acceptance test for free: lambda_1(G_m) <= theta_prev <=
//     lambda_2(G_m), so a converged eigenvalue that is not above theta_prev IS the lowest one;
//   * a restart cycle, or a rejected iteration: the full decomposition by parallel-order Jacobi (all
//     disjoint pairs of a round rotated at once), the robust slow path -- once per max_space - q cycles.
// (Round 1 kept a decomposition up to date every cycle -- arrowhead + Jacobi, 40-130 us of one-SM latency on
// the critical path of every cycle: a third of the summed kernel time of a bench step.)
__device__ void jacobi_full(DavState* st, double (*A)[kMaxS + 1], double (*J)[kMaxS + 1], int m, int lane,
                            double* cs_c, double* cs_s, int* pr_p, int* pr_q) {
    // A: symmetric m x m (destroyed), J: receives the eigenvectors (columns).  One warp.  The index maps use
    // shifts only: an integer division by a run-time value costs more than the rotation it addresses.
    const int np = (m + 1) / 2;   // pairs per round
    const int nplayers = 2 * np;  // even
    // lane -> (pair k, row chunk c): kw pairs per "row of lanes" (8 for m <= 16, else 16)
    const int kw_log = np <= 8 ? 3 : 4, kw = 1 << kw_log, nchunk = 32 >> kw_log;
    const int k_of = lane & (kw - 1), c_of = lane >> kw_log;
    for (int i = c_of; i < m; i += nchunk)
        for (int j = k_of; j < m; j += kw) J[i][j] = i == j ? 1.0 : 0.0;
    __syncwarp();
    for (int sweep = 0; sweep < 30 && m > 1; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = c_of; i < m; i += nchunk)
            for (int j = k_of; j < m; j += kw) {
                const double v = A[i][j] * A[i][j];
                if (i == j) dia += v; else off += v;
            }
        off = warp_sum(off);
        dia = warp_sum(dia);
        if (off <= 1e-26 * dia) break;  // identical on all lanes
        for (int step = 0; step < nplayers - 1; ++step) {
            if (lane < np) {
                const int k = lane;
                int p, q;
                if (k == 0) {
                    p = nplayers - 1;
                    q = step;
                } else {
                    p = step + k;
                    if (p >= nplayers - 1) p -= nplayers - 1;
                    q = step - k;
                    if (q < 0) q += nplayers - 1;
                }
                if (p > q) {
                    const int t = p;
                    p = q;
                    q = t;
                }
                double c = 1.0, sn = 0.0;
                if (q < m) {
                    const double apq = A[p][q];
                    if (fabs(apq) > 1e-300) {
                        const double tau = (A[q][q] - A[p][p]) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = rsqrt(1.0 + t * t);
                        sn = t * c;
                    } else {
                        q = p;  // nothing to rotate
                    }
                } else {
                    q = p;  // dummy pair: identity
                }
                pr_p[k] = p;
                pr_q[k] = q;
                cs_c[k] = c;
                cs_s[k] = sn;
            }
            __syncwarp();
            const bool have = k_of < np;
            const int p = have ? pr_p[k_of] : 0, q = have ? pr_q[k_of] : 0;
            const double c = have ? cs_c[k_of] : 1.0, sn = have ? cs_s[k_of] : 0.0;
            // column update: A <- A R, J <- J R  (pairs are disjoint: no two lanes touch the same column)
            if (p != q) {
                for (int i = c_of; i < m; i += nchunk) {
                    const double xp = A[i][p], xq = A[i][q];
                    A[i][p] = c * xp - sn * xq;
                    A[i][q] = sn * xp + c * xq;
                    const double yp = J[i][p], yq = J[i][q];
                    J[i][p] = c * yp - sn * yq;
                    J[i][q] = sn * yp + c * yq;
                }
            }
            __syncwarp();
            // row update: A <- R^T A
            if (p != q) {
                for (int j = c_of; j < m; j += nchunk) {
                    const double xp = A[p][j], xq = A[q][j];
                    A[p][j] = c * xp - sn * xq;
                    A[q][j] = sn * xp + c * xq;
                }
            }
            __syncwarp();
        }
    }
    // decomposition -> state: eigenvectors, eigenvalues, ascending order
    if (lane == 0) {  // selection sort, m <= 32
        unsigned used = 0u;
        for (int r = 0; r < m; ++r) {
            int b = -1;
            for (int i = 0; i < m; ++i)
                if (!((used >> i) & 1u) && (b < 0 || A[i][i] < A[b][b])) b = i;
            used |= 1u << b;
            st->ord[r] = b;
            if (r == 0) st->best = b;
        }
    }
    for (int i = c_of; i < m; i += nchunk)
        for (int j = k_of; j < m; j += kw) st->Q[i * kMaxS + j] = J[i][j];
    for (int i = lane; i < m; i += 32) st->lam[i] = A[i][i];
    __syncwarp();
}

// publish the Ritz pair (theta, y): unit norm, component of largest magnitude (first such) positive
__device__ __forceinline__ void publish_pair(DavState* st, double yi, double theta, int m, int lane) {
    const double nrm2 = warp_sum(lane < m ? yi * yi : 0.0);
    double mag = lane < m ? fabs(yi) : -1.0;
    int who = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double m2 = __shfl_xor_sync(0xffffffffu, mag, o);
        const int w2 = __shfl_xor_sync(0xffffffffu, who, o);
        if (m2 > mag || (m2 == mag && w2 < who)) { mag = m2; who = w2; }
    }
    const double lead = __shfl_sync(0xffffffffu, yi, who);
    const double f = (lead < 0.0 ? -1.0 : 1.0) / sqrt(nrm2);
    if (lane < m) st->y[lane] = yi * f;
    if (lane == 0) {
        st->theta_prev = st->theta;
        st->theta = theta;
    }
}

// Rayleigh-quotient iteration with the whole problem in registers: lane i owns row i of G (MR >= m doubles) and
// component i of the iterate.
//   * start: the exact lowest Ritz pair of the 2-D space spanned by the previous Ritz vector (padded with a
//     zero) and the new basis vector -- a closed form, and already below theta_prev;
//   * each iteration solves (G - mu I) z' = z by elimination WITHOUT pivot search, from the last row upwards:
//     the rows of the newer basis vectors (Rayleigh quotients well above mu) are the pivots, the row of the
//     vector that carries the Ritz vector is eliminated last and takes the near-singular pivot -- exactly where
//     inverse iteration wants it.  All indices are compile-time constants (no local memory), a step is one
//     reciprocal, k shuffles and k FMAs: ~1.5 us per iteration at m = 12 where the pivoting version took ~20.
// Nothing here has to be trusted: the pair is accepted only if its residual against the ORIGINAL G is at
// rounding level and Cauchy interlacing certifies it as the lowest; otherwise the caller falls back to the
// full Jacobi decomposition.
template <int MR>
__device__ __forceinline__ bool rqi_lowest(const double (*Gs)[kMaxS + 1], int m, int lane, double theta_prev,
                                           const double* y_prev, double scale, double* mu_out, double* yi_out,
                                           int* iters) {
    const int d = m - 1;
    const unsigned full = 0xffffffffu;
    double grow[MR];
#pragma unroll
    for (int j = 0; j < MR; ++j) grow[j] = (lane < m && j < m) ? Gs[lane][j] : 0.0;
    // 2 x 2 start: [theta_prev b; b a] in the basis {(y_prev, 0), e_d}
    double zi = lane < d ? y_prev[lane] : 0.0;
    const double bcoup = warp_sum(lane < d ? Gs[d][lane] * zi : 0.0);
    const double a_new = Gs[d][d];
    double mu = theta_prev;
    if (bcoup != 0.0) {
        const double half = 0.5 * (a_new - theta_prev);
        const double root = sqrt(half * half + bcoup * bcoup);
        // lower root of the 2 x 2 problem, written without cancellation
        const double shift = half >= 0.0 ? -bcoup * bcoup / (half + root) : half - root;
        mu = theta_prev + shift;
        const double t = shift / bcoup;   // eigenvector (1, t)
        const double inv = rsqrt(1.0 + t * t);
        zi = lane < d ? zi * inv : (lane == d ? t * inv : 0.0);
    }
    const double tiny = 1e-300 + 1e-18 * scale;
    bool ok = false;
    for (int it = 0; it < 8; ++it) {
        double a[MR];
#pragma unroll
        for (int j = 0; j < MR; ++j) a[j] = grow[j] - (j == lane ? mu : 0.0);
        double b = lane < m ? zi : 0.0;
        double rinv[MR];
        // elimination, pivots k = m-1 ... 1 (row 0 is last and takes the near-singular pivot)
#pragma unroll
        for (int k = MR - 1; k >= 1; --k) {
            rinv[k] = 1.0;
            if (k < m) {
                double pk = __shfl_sync(full, a[k], k);
                if (fabs(pk) < tiny) pk = pk < 0.0 ? -tiny : tiny;
                const double r = 1.0 / pk;
                rinv[k] = r;
                const double f = lane < k ? a[k] * r : 0.0;
                const double pb = __shfl_sync(full, b, k);
                b = fma(-f, pb, b);
#pragma unroll
                for (int j = 0; j < k; ++j) {
                    const double pj = __shfl_sync(full, a[j], k);
                    a[j] = fma(-f, pj, a[j]);
                }
            }
        }
        {
            double p0 = __shfl_sync(full, a[0], 0);
            if (fabs(p0) < tiny) p0 = p0 < 0.0 ? -tiny : tiny;   // mu hit the eigenvalue: that is fine
            rinv[0] = 1.0 / p0;
        }
        // forward substitution on the lower-triangular remainder, column oriented: z_k known -> every later
        // row removes its column-k term
        double znew = 0.0;
#pragma unroll
        for (int k = 0; k < MR; ++k) {
            if (k < m) {
                const double zk = __shfl_sync(full, b, k) * rinv[k];
                if (lane == k) znew = zk;
                if (lane > k) b = fma(-a[k], zk, b);
            }
        }
        // normalise (scale first: the solution is huge when mu is converged)
        double zmax = lane < m ? fabs(znew) : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) zmax = fmax(zmax, __shfl_xor_sync(full, zmax, o));
        if (!(zmax > 0.0) || !isfinite(zmax)) break;
        znew *= 1.0 / zmax;
        const double n2 = warp_sum(lane < m ? znew * znew : 0.0);
        zi = znew * rsqrt(n2);
        // Rayleigh quotient and residual against the original G: gz_i = sum_j G[i][j] z_j
        double gz = 0.0;
#pragma unroll
        for (int j = 0; j < MR; ++j) {
            const double zj = __shfl_sync(full, zi, j);
            gz = fma(grow[j], zj, gz);
        }
        const double mu_new = warp_sum(lane < m ? zi * gz : 0.0);
        const double r = lane < m ? gz - mu_new * zi : 0.0;
        const double rn = sqrt(warp_sum(r * r));
        mu = mu_new;
        *iters += 1;
        // (mu, z) is an eigenpair of G to working precision (cubic convergence: one to three iterations;
        // stagnation of mu alone is NOT convergence -- a start vector that mixes +lambda and -lambda equally
        // keeps its Rayleigh quotient for ever)
        if (rn <= 2e-14 * scale || (it == 7 && rn <= 1e-11 * scale)) {
            // interlacing: lambda_1(G_m) <= theta_prev <= lambda_2(G_m).  An eigenvalue strictly below
            // theta_prev is therefore the lowest one; one that equals theta_prev to rounding may be the
            // second (a degenerate or decoupled direction) -- the full decomposition decides
            ok = mu < theta_prev - 1.5e-14 * (fabs(theta_prev) + scale);
            break;
        }
    }
    *mu_out = mu;
    *yi_out = lane < m ? zi : 0.0;
    return ok;
}

template <int MR>
__device__ void rayleigh_ritz_body(DavState* st, const double* partials, int nblk, int m, int need_full) {
    __shared__ double A[kMaxS][kMaxS + 1];   // G - mu I (elimination) or the Jacobi work matrix
    __shared__ double J[kMaxS][kMaxS + 1];   // copy of G (Rayleigh quotients) / accumulated rotations
    __shared__ double g[kMaxS], bvec[kMaxS], zvec[kMaxS];
    __shared__ double cs_c[kMaxS / 2], cs_s[kMaxS / 2];
    __shared__ int pr_p[kMaxS / 2], pr_q[kMaxS / 2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = m - 1;
    // new Gram column (all warps), then G = [G_old g; g^T a] in the state and in shared memory
    for (int row = warp; row < m; row += blockDim.x >> 5) {
        const double v = reduce_partials(partials, row, nblk);
        if (lane == 0) g[row] = v;
    }
    __syncthreads();
    for (int idx = tid; idx < kMaxS * m; idx += blockDim.x) {
        const int i = idx / kMaxS, j = idx % kMaxS;  // compile-time divisor
        if (j >= m) continue;
        double v;
        if (i == d) v = g[j];
        else if (j == d) v = g[i];
        else v = st->G[i * kMaxS + j];
        J[i][j] = v;
        if (i == d || j == d) st->G[i * kMaxS + j] = v;
    }
    __syncthreads();
    if (warp != 0) return;
    // ---- one warp from here on ----
    bool ok = false;
    double theta = 0.0, yi = 0.0;
    const bool pair_restart = need_full && st->restart_mode == 1 && m >= 2 && m <= MR;
    if (!need_full || pair_restart) {
        if (m == 1) {
            ok = true;
            theta = J[0][0];
            yi = lane == 0 ? 1.0 : 0.0;
        } else {
            // start: previous Ritz vector padded with zero; its Rayleigh quotient is the previous theta
            const double theta_prev = st->theta;
            if (lane < d) bvec[lane] = st->y[lane];
            double scale = 0.0;
            for (int i = 0; i < m; ++i)
                if (lane < m) scale = fmax(scale, fabs(J[i][lane]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) scale = fmax(scale, __shfl_xor_sync(0xffffffffu, scale, o));
            __syncwarp();
            double mu = theta_prev;
            int iters = 0;
            if (m <= MR) ok = rqi_lowest<MR>(J, m, lane, theta_prev, bvec, scale, &mu, &yi, &iters);
            if (lane == 0) st->n_rqi_iter += iters;
            theta = mu;
        }
    }
    if (!ok) {
        // full decomposition (restart cycle, or the iteration above was rejected)
        if (lane == 0) st->n_jacobi += 1;
        for (int i = 0; i < m; ++i)
            if (lane < m) A[i][lane] = J[i][lane];
        __syncwarp();
        jacobi_full(st, A, J, m, lane, cs_c, cs_s, pr_p, pr_q);
        const int best = st->ord[0];
        theta = A[best][best];
        yi = lane < m ? J[lane][best] : 0.0;
    } else if (pair_restart) {
        // second vector of the collapsed basis: the previous Ritz vector (padded with a zero) orthogonalised
        // against the new one.  G y = theta y, so ANY unit u orthogonal to y gives a diagonal collapsed matrix
        // diag(theta, u^T G u): no decomposition needed.
        const unsigned full = 0xffffffffu;
        double u = lane < d ? bvec[lane] : 0.0;
        u -= warp_sum(u * yi) * yi;
        double nu2 = warp_sum(u * u);
        if (!(nu2 > 1e-12)) {
            // the Ritz vector did not move (converged, or the first cycles): use the newest basis vector
            const double yd = __shfl_sync(full, yi, d);
            u = (lane == d ? 1.0 : 0.0) - yd * yi;
            if (lane >= m) u = 0.0;
            nu2 = warp_sum(u * u);
        }
        u *= rsqrt(nu2);
        u -= warp_sum(u * yi) * yi;   // second pass: orthogonal to rounding after the normalisation
        u *= rsqrt(warp_sum(u * u));
        double gu = 0.0;
        for (int j = 0; j < m; ++j) {
            const double uj = __shfl_sync(full, u, j);
            if (lane < m) gu = fma(J[lane][j], uj, gu);
        }
        const double rho = warp_sum(lane < m ? u * gu : 0.0);
        if (lane < m) {
            st->Q[lane * kMaxS + 0] = yi;
            st->Q[lane * kMaxS + 1] = u;
        }
        if (lane == 0) {
            st->ord[0] = 0;
            st->ord[1] = 1;
            st->best = 0;
            st->lam[0] = theta;
            st->lam[1] = rho;
        }
    }
    publish_pair(st, yi, theta, m, lane);
}

// last node of the cycle in the device-driven loop: basis bookkeeping and the WHILE condition
__global__ void advance_kernel(DavState* st, cudaGraphConditionalHandle handle) {
    if (threadIdx.x != 0) return;
    unsigned int go = 0u;
    if (st->status == 0) {
        const int m = st->m;
        const int me = (m == st->M) ? st->q_keep : m;
        st->m = me + 1;
        st->slot = me;
        go = st->cycles < st->max_cycle ? 1u : 0u;
    }
    cudaGraphSetConditional(handle, go);
}

__device__ void convergence_body(DavState* st, const double* partials, int nblk, int m, int restart,
                                 double tol, double tol_residual) {
    __shared__ double p[kMaxS + 2];
    const int tid = threadIdx.x;
    for (int row = tid >> 5; row < kMaxS + 2; row += blockDim.x >> 5) {
        if (row < m || row >= kMaxS) {
            const double v = reduce_partials(partials, row, nblk);
            if ((tid & 31) == 0) p[row] = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        const double rnorm = sqrt(p[kMaxS]);
        st->rnorm = rnorm;
        st->tnorm2 = p[kMaxS + 1];
        st->cycles += 1;
        if (fabs(st->theta - st->theta_prev) < tol && rnorm < tol_residual) {
            st->status = 1;
        } else if (restart) {
            // the collapsed space is spanned by the `restart` lowest Ritz vectors: projections of t on
            // them, then the decomposition becomes diagonal
            double cj[kKeep], lj[kKeep];
            for (int k = 0; k < restart; ++k) {
                const int col = st->ord[k];
                double c = 0.0;
                for (int i = 0; i < m; ++i) c += (k == 0 ? st->y[i] : st->Q[i * kMaxS + col]) * p[i];
                cj[k] = c;
                lj[k] = st->lam[col];
            }
            for (int k = 0; k < restart; ++k) {
                st->c1[k] = cj[k];
                st->lam[k] = lj[k];
                st->ord[k] = k;
                for (int i = 0; i < restart; ++i) {
                    st->Q[i * kMaxS + k] = i == k ? 1.0 : 0.0;
                    st->G[i * kMaxS + k] = i == k ? lj[k] : 0.0;  // kept Ritz vectors: G is diagonal
                }
                st->y[k] = k == 0 ? 1.0 : 0.0;
            }
        } else {
            for (int i = 0; i < m; ++i) st->c1[i] = p[i];
        }
    }
}

__device__ void norm_body(DavState* st, const double* partials, int nblk, int m, double lindep) {
    __shared__ double p[kMaxS + 1];
    const int tid = threadIdx.x;
    for (int row = tid >> 5; row < kMaxS + 1; row += blockDim.x >> 5) {
        if (row < m || row == kMaxS) {
            const double v = reduce_partials(partials, row, nblk);
            if ((tid & 31) == 0) p[row] = v;
        }
    }
    __syncthreads();
    if (tid == 0) {
        double n2 = p[kMaxS];
        for (int i = 0; i < m; ++i) {
            st->c2[i] = p[i];
            n2 -= p[i] * p[i];
        }
        // lindep is relative to the size of the correction before orthogonalisation (pyscf tests the
        // normalised vector); an exactly vanishing correction also ends the iteration
        if (!(n2 > lindep * st->tnorm2) || !(n2 > 1e-300)) {
            st->status = 2;  // cannot expand the space any further: current Ritz vector is final
            st->inv_norm = 0.0;
        } else {
            st->inv_norm = 1.0 / sqrt(n2);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// helpers: init, dot, axpby, scale
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads)
dot_partial_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                   double* __restrict__ partials) {
    __shared__ double red[kRedThreads / 32];
    double acc[1] = {0.0};
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x)
        acc[0] = fma(x[j], y[j], acc[0]);
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc[0];
}

__global__ void dot_final_kernel(const double* __restrict__ partials, int nblk, double* out) {
    const double s = reduce_partials(partials, 0, nblk);  // launched with one warp
    if (threadIdx.x == 0) out[0] = s;
}

// y <- alpha * x * (1/sqrt(*norm2) if norm2 else 1)
__global__ void scale_copy_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n,
                                  const double* __restrict__ norm2) {
    const double f = norm2 ? 1.0 / sqrt(norm2[0]) : 1.0;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x)
        y[j] = x[j] * f;
}

// z <- a*x + b*y   (x, y, z may alias)
__global__ void axpby_kernel(const int* __restrict__ done, double a, const double* x, double b,
                             const double* y, double* z, int64_t n) {
    if (done && *done) return;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x)
        z[j] = a * x[j] + b * y[j];
}

__global__ void init_state_kernel(DavState* st, int M, int q_keep, int max_cycle, int restart_mode) {
    if (threadIdx.x == 0) {
        st->restart_mode = restart_mode;
        st->m = 1;
        st->slot = 0;
        st->n_jacobi = 0;
        st->n_rqi_iter = 0;
        st->M = M;
        st->q_keep = q_keep;
        st->max_cycle = max_cycle;
        st->status = 0;
        st->cycles = 0;
        st->theta = 0.0;
        st->theta_prev = INFINITY;
        st->rnorm = INFINITY;
        st->tnorm2 = 0.0;
        st->inv_norm = 0.0;
    }
    for (int i = threadIdx.x; i < kMaxS * kMaxS; i += blockDim.x) st->Q[i] = 0.0;
    for (int i = threadIdx.x; i < kMaxS; i += blockDim.x) {
        st->y[i] = st->c1[i] = st->c2[i] = st->lam[i] = 0.0;
        st->ord[i] = i;
    }
    if (threadIdx.x == 0) st->best = 0;
    if (threadIdx.x < 4) st->ticket[threadIdx.x] = 0;
}

// argmin over the (na, nb) block of hdiag (pads excluded): stage 1 per-CTA, stage 2 single thread.
__global__ void __launch_bounds__(kRedThreads)
argmin_partial_kernel(const double* __restrict__ hdiag, int na, int nb, int ldc,
                      double* __restrict__ pval, int64_t* __restrict__ pidx) {
    __shared__ double sv[kRedThreads];
    __shared__ int64_t si[kRedThreads];
    double best = INFINITY;
    int64_t bi = -1;
    const int64_t total = (int64_t)na * nb;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total;
         k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = k / nb, b = k % nb;
        const double v = hdiag[a * ldc + b];
        if (v < best || (v == best && k < bi)) {
            best = v;
            bi = k;
        }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v2 = sv[threadIdx.x + o];
            const int64_t i2 = si[threadIdx.x + o];
            if (i2 >= 0 && (v2 < sv[threadIdx.x] || si[threadIdx.x] < 0 ||
                            (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x]))) {
                sv[threadIdx.x] = v2;
                si[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        pval[blockIdx.x] = sv[0];
        pidx[blockIdx.x] = si[0];
    }
}

__global__ void init_guess_kernel(const double* __restrict__ pval, const int64_t* __restrict__ pidx,
                                  int nblk, int na, int nb, int ldc, double* __restrict__ x0) {
    // x0 was zero-filled before this launch
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double best = INFINITY;
        int64_t bi = -1;
        for (int b = 0; b < nblk; ++b) {
            if (pidx[b] >= 0 && (pval[b] < best || bi < 0 || (pval[b] == best && pidx[b] < bi))) {
                best = pval[b];
                bi = pidx[b];
            }
        }
        const int64_t a = bi / nb, b = bi % nb;
        x0[a * ldc + b] = 1.0;
        // pyscf direct_spin1._get_init_guess: ci0[0][0] += 1e-5; ci0[0][-1] -= 1e-5
        x0[0] += 1e-5;
        x0[(int64_t)(na - 1) * ldc + (nb - 1)] -= 1e-5;
    }
}

// Sign convention of the returned eigenvector: the first element of largest magnitude is positive
// (pyscf's sign is whatever LAPACK returns for the small problem).  Stage 1 per CTA, stage 2 one thread,
// stage 3 flips the vector when needed -- no host round trip.
__global__ void __launch_bounds__(kRedThreads)
absmax_partial_kernel(const double* __restrict__ x, int64_t n, double* __restrict__ pval,
                      int64_t* __restrict__ pidx) {
    __shared__ double sv[kRedThreads];
    __shared__ int64_t si[kRedThreads];
    double best = -1.0;
    int64_t bi = -1;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x) {
        const double v = fabs(x[k]);
        if (v > best) {  // k ascends per thread: ties keep the lowest index
            best = v;
            bi = k;
        }
    }
    sv[threadIdx.x] = best;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = kRedThreads / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double v2 = sv[threadIdx.x + o];
            const int64_t i2 = si[threadIdx.x + o];
            if (i2 >= 0 && (v2 > sv[threadIdx.x] || si[threadIdx.x] < 0 ||
                            (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x]))) {
                sv[threadIdx.x] = v2;
                si[threadIdx.x] = i2;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        pval[blockIdx.x] = sv[0];
        pidx[blockIdx.x] = si[0];
    }
}

__global__ void sign_flag_kernel(const double* __restrict__ pval, int64_t* __restrict__ pidx, int nblk,
                                 const double* __restrict__ x) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double best = -1.0;
        int64_t bi = -1;
        for (int b = 0; b < nblk; ++b)
            if (pidx[b] >= 0 && (pval[b] > best || (pval[b] == best && pidx[b] < bi))) {
                best = pval[b];
                bi = pidx[b];
            }
        pidx[kRedBlocks] = (bi >= 0 && x[bi] < 0.0) ? 1 : 0;  // flag slot after the partial indices
    }
}

__global__ void __launch_bounds__(kRedThreads)
flip_kernel(double* __restrict__ x, int64_t n, const int64_t* __restrict__ flag) {
    if (*flag == 0) return;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n;
         k += (int64_t)gridDim.x * blockDim.x)
        x[k] = -x[k];
}

// row / column weights of c^2 and orbital occupancies
__global__ void row_weight_kernel(const double* __restrict__ c, int na, int nb, int ldc,
                                  double* __restrict__ wa) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= na) return;
    double s = 0.0;
    for (int b = lane; b < nb; b += 32) {
        const double v = c[(int64_t)a * ldc + b];
        s = fma(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) wa[a] = s;
}

__global__ void col_weight_kernel(const double* __restrict__ c, int na, int nb, int ldc,
                                  double* __restrict__ wb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    double s = 0.0;
    for (int a = 0; a < na; ++a) {
        const double v = c[(int64_t)a * ldc + b];
        s = fma(v, v, s);
    }
    wb[b] = s;
}

// one CTA per (spin, orbital)
__global__ void __launch_bounds__(kRedThreads)
occupancy_kernel(const double* __restrict__ wa, const uint64_t* __restrict__ sa, int na,
                 const double* __restrict__ wb, const uint64_t* __restrict__ sb, int nb, int norb,
                 double* __restrict__ occ) {
    __shared__ double red[kRedThreads / 32];
    const int spin = blockIdx.x / norb, p = blockIdx.x % norb;
    const double* w = spin ? wb : wa;
    const uint64_t* s = spin ? sb : sa;
    const int n = spin ? nb : na;
    double acc[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if ((s[i] >> p) & 1ull) acc[0] += w[i];
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) occ[blockIdx.x] = acc[0];
}

static inline int red_blocks(int64_t n) {
    static const int per_thread = getenv("SQD_RED_PER_THREAD") ? atoi(getenv("SQD_RED_PER_THREAD")) : 4;
    const int64_t chunk = (int64_t)(per_thread > 0 ? per_thread : 4) * kRedThreads;
    int64_t b = (n + chunk - 1) / chunk;
    if (b > kRedBlocks) b = kRedBlocks;
    if (b < 1) b = 1;
    return (int)b;
}

template <typename F>
static int dispatch_mv(int m, F&& f) {
    if (m <= 4) return f(std::integral_constant<int, 4>());
    if (m <= 8) return f(std::integral_constant<int, 8>());
    if (m <= 12) return f(std::integral_constant<int, 12>());
    if (m <= 16) return f(std::integral_constant<int, 16>());
    if (m <= 24) return f(std::integral_constant<int, 24>());
    return f(std::integral_constant<int, 32>());
}

struct Workspace {
    double *V, *W, *T, *X, *tmp1, *tmp2, *partials;
    DavState* state;
};

static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

static int64_t workspace_bytes(int64_t n, int max_space) {
    size_t b = 0;
    b += align_up((size_t)max_space * n * sizeof(double)) * 2;  // V, W
    b += align_up((size_t)n * sizeof(double)) * 4;              // T, X, tmp1, tmp2
    b += align_up((size_t)kPartialRows * kRedBlocks * sizeof(double));
    b += align_up(sizeof(DavState));
    return (int64_t)b;
}

static void carve(void* base, int64_t n, int max_space, Workspace* ws) {
    char* p = (char*)base;
    ws->V = (double*)p;
    p += align_up((size_t)max_space * n * sizeof(double));
    ws->W = (double*)p;
    p += align_up((size_t)max_space * n * sizeof(double));
    ws->T = (double*)p;
    p += align_up((size_t)n * sizeof(double));
    ws->X = (double*)p;
    p += align_up((size_t)n * sizeof(double));
    ws->tmp1 = (double*)p;
    p += align_up((size_t)n * sizeof(double));
    ws->tmp2 = (double*)p;
    p += align_up((size_t)n * sizeof(double));
    ws->partials = (double*)p;
    p += align_up((size_t)kPartialRows * kRedBlocks * sizeof(double));
    ws->state = (DavState*)p;
}

// w <- O v (+ quadratic spin penalty)
static int apply_operator(const sqd_operator* op, const sqd_davidson_params* prm, const double* v,
                          double* w, Workspace& ws, int64_t n, cudaStream_t st) {
    const int* done = &ws.state->status;
    if (prm->nccl_comm != nullptr) {
        // sharded build: this rank owns rows [row_begin, row_end) of sigma; the other rows are zero and
        // the blocks meet in an all-reduce over NVLink.  Every rank runs the identical enqueue sequence.
        if (prm->shard_bounds != nullptr && prm->shard_world > 0 && prm->shard_world <= 64) {
            // disjoint row blocks: every rank writes its rows in place, then the blocks travel
            if (sigma_dispatch_rows(op, v, w, done, prm->row_begin, prm->row_end, st)) return -2;
            long long offs[65];
            for (int r = 0; r <= prm->shard_world; ++r) offs[r] = (long long)prm->shard_bounds[r] * op->ldc;
            if (nccl_allgather_blocks(prm->nccl_comm, w, offs, prm->shard_world, st)) return -2;
        } else {
            SQD_CUDA_OK(cudaMemsetAsync(w, 0, (size_t)n * sizeof(double), st));
            if (sigma_dispatch_rows(op, v, w, done, prm->row_begin, prm->row_end, st)) return -2;
            if (nccl_allreduce_sum_f64(prm->nccl_comm, w, n, st)) return -2;
        }
    } else if (sigma_dispatch_flag(op, v, w, done, st)) {
        return -2;
    }
    if (prm->ss_op) {
        // w += shift * (S^2 - ss)^2 v   (pyscf fix_spin_, quadratic branch)
        const int blocks = red_blocks(n);
        if (sigma_dispatch_flag(prm->ss_op, v, ws.tmp1, done, st)) return -2;
        axpby_kernel<<<blocks, kRedThreads, 0, st>>>(done, 1.0, ws.tmp1, -prm->ss_value, v, ws.tmp1, n);
        if (sigma_dispatch_flag(prm->ss_op, ws.tmp1, ws.tmp2, done, st)) return -2;
        axpby_kernel<<<blocks, kRedThreads, 0, st>>>(done, 1.0, ws.tmp2, -prm->ss_value, ws.tmp1,
                                                     ws.tmp2, n);
        axpby_kernel<<<blocks, kRedThreads, 0, st>>>(done, 1.0, w, prm->ss_shift, ws.tmp2, w, n);
        if (check_launch("axpby_kernel", 3)) return -2;
    }
    return 0;
}

using ApplyFn = std::function<int(const double*, double*, Workspace&)>;
// device-driven loop: w[slot] <- O v[slot] with the slot read from device memory (bases passed); may be empty
using ApplyCtlFn = std::function<int(const double*, double*, const int*, Workspace&, cudaStream_t)>;

// Side stream of a host thread: the full Rayleigh-Ritz decomposition of cycle k runs there, concurrently
// with the residual / orthogonalisation / next sigma build of the main stream, and is joined before the
// Rayleigh-Ritz step of cycle k+1 (or before the residual kernel of a restart cycle).
struct SideStream {
    cudaStream_t s = nullptr;
    cudaStream_t cap = nullptr;  // capture stream used when the caller's stream is a default stream
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_cap = nullptr;
    int dev = -1;
};
static thread_local SideStream g_side;

static int side_stream(SideStream** out) {
    int dev = 0;
    SQD_CUDA_OK(cudaGetDevice(&dev));
    if (g_side.s == nullptr || g_side.dev != dev) {
        if (g_side.s != nullptr) {
            cudaStreamDestroy(g_side.s);
            cudaStreamDestroy(g_side.cap);
            cudaEventDestroy(g_side.ev_fork);
            cudaEventDestroy(g_side.ev_join);
            cudaEventDestroy(g_side.ev_cap);
        }
        SQD_CUDA_OK(cudaStreamCreateWithFlags(&g_side.s, cudaStreamNonBlocking));
        SQD_CUDA_OK(cudaStreamCreateWithFlags(&g_side.cap, cudaStreamNonBlocking));
        SQD_CUDA_OK(cudaEventCreateWithFlags(&g_side.ev_fork, cudaEventDisableTiming));
        SQD_CUDA_OK(cudaEventCreateWithFlags(&g_side.ev_join, cudaEventDisableTiming));
        SQD_CUDA_OK(cudaEventCreateWithFlags(&g_side.ev_cap, cudaEventDisableTiming));
        g_side.dev = dev;
    }
    *out = &g_side;
    return 0;
}

// The whole Davidson loop as ONE graph launch: a WHILE conditional node whose body is one cycle.  Every
// kernel of the body reads the cycle's control variables (basis size, vector slot, restart) from the
// device-side state, so the body is captured once and the device replays it until the convergence flag
// (or the cycle limit) clears the condition -- the host is out of the loop.  Body:
//   gram (+ lowest Ritz pair; whole Rayleigh-Ritz on a restart cycle)
//     |- side branch: full decomposition for the next cycle
//     |- residual -> ortho1 -> ortho2 -> advance (m, slot, condition) -> sigma build on the new vector
// The first sigma build runs before the graph (eagerly).
// while_node = false: the graph is ONE cycle and the host replays it (a graph launch instead of seven kernel
// launches per cycle: the host cost of a cycle drops from ~45 us to ~10 us, which is what limits the ranks of
// a box that share few host cores); kernels after convergence return at once (status flag).
static int davidson_graph_loop(int64_t n, const ApplyCtlFn& apply_ctl, const double* d_hdiag, Workspace& ws,
                               int M, int blocks, const sqd_davidson_params* prm, SideStream* side,
                               cudaStream_t st, cudaGraph_t* g_out, cudaGraphExec_t* ex_out, bool while_node) {
    cudaGraph_t g = nullptr;
    SQD_CUDA_OK(cudaGraphCreate(&g, 0));
    *g_out = g;
    cudaGraphConditionalHandle handle = 0;
    cudaGraph_t body = g;
    if (while_node) {
        SQD_CUDA_OK(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams cp = {cudaGraphNodeTypeConditional};
        cp.conditional.handle = handle;
        cp.conditional.type = cudaGraphCondTypeWhile;
        cp.conditional.size = 1;
        cudaGraphNode_t node;
        SQD_CUDA_OK(cudaGraphAddNode(&node, g, nullptr, 0, &cp));
        body = cp.conditional.phGraph_out[0];
    }
    SQD_CUDA_OK(cudaStreamBeginCaptureToGraph(st, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
    int rc = dispatch_mv(M, [&](auto mv) {
        constexpr int MV = decltype(mv)::value;
        gram_kernel<MV><<<blocks, kRedThreads, 0, st>>>(ws.state, ws.V, ws.W, n, -1, ws.partials, 2);
        residual_kernel<MV><<<blocks, kRedThreads, 0, st>>>(ws.state, ws.V, ws.W, d_hdiag, n, -1, 0,
                                                            prm->level_shift, ws.X, ws.T, ws.partials,
                                                            prm->tol, prm->tol_residual);
        ortho1_kernel<MV><<<blocks, kRedThreads, 0, st>>>(ws.state, ws.V, n, -1, ws.T, ws.partials,
                                                          prm->lindep);
        ortho2_kernel<MV><<<blocks, kRedThreads, 0, st>>>(ws.state, ws.V, n, -1, ws.T, ws.V);
        if (while_node) advance_kernel<<<1, 32, 0, st>>>(ws.state, handle);
        else advance_plain_kernel<<<1, 32, 0, st>>>(ws.state);
        if (check_launch("davidson cycle (graph)", 5)) return -2;
        if (apply_ctl(ws.V, ws.W, &ws.state->slot, ws, st)) return -2;
        return 0;
    });
    cudaGraph_t captured = nullptr;
    const cudaError_t e = cudaStreamEndCapture(st, &captured);
    if (rc != 0 || e != cudaSuccess) {
        if (rc == 0) set_error("davidson graph capture failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return -2;
    }
    SQD_CUDA_OK(cudaGraphInstantiate(ex_out, g, 0));
    if (while_node) SQD_CUDA_OK(cudaGraphLaunch(*ex_out, st));
    return 0;
}

static int davidson_core(int64_t n, const ApplyFn& apply, const double* d_hdiag, const double* d_x0,
                         double* d_x, void* d_workspace, int64_t ws_bytes,
                         const sqd_davidson_params* prm, sqd_davidson_info* h_info, cudaStream_t st,
                         const ApplyCtlFn& apply_ctl = ApplyCtlFn()) {
    const int M = prm->max_space;
    SQD_REQUIRE(M >= 2 && M <= kMaxS, "sqd_davidson: max_space must be in [2, %d] (got %d)", kMaxS, M);
    SQD_REQUIRE(ws_bytes >= workspace_bytes(n, M), "sqd_davidson: workspace too small");
    SQD_REQUIRE(prm->max_cycle >= 1, "sqd_davidson: max_cycle must be >= 1");
    SQD_REQUIRE(n % 2 == 0, "sqd_davidson: the vector length must be even (16-byte loads); pad the rows");
    Workspace ws;
    carve(d_workspace, n, M, &ws);
    const int blocks = red_blocks(n);
    const int check_every = prm->check_every > 0 ? prm->check_every : 4;
    const int q_keep_all = restart_keep(M);
    init_state_kernel<<<1, 256, 0, st>>>(ws.state, M, q_keep_all, prm->max_cycle, restart_mode_knob());
    // V_0 = x0 / |x0|
    dot_partial_kernel<<<blocks, kRedThreads, 0, st>>>(d_x0, d_x0, n, ws.partials);
    dot_final_kernel<<<1, 32, 0, st>>>(ws.partials, blocks, ws.partials + kRedBlocks);
    scale_copy_kernel<<<blocks, kRedThreads, 0, st>>>(d_x0, ws.V, n, ws.partials + kRedBlocks);
    if (check_launch("davidson init", 4)) return -2;

    std::vector<cudaEvent_t> ev;  // profile mode only: (before, after) per operator application
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    if (prm->profile) {
        SQD_CUDA_OK(cudaEventCreate(&ev_begin));
        SQD_CUDA_OK(cudaEventCreate(&ev_end));
        SQD_CUDA_OK(cudaEventRecord(ev_begin, st));
    }
    SideStream* side = nullptr;
    if (side_stream(&side)) return -2;
    int m = 1, slot = 0, status = 0, cycle = 0, sigma_builds = 0;
    static const int knob_graph = env_knob("SQD_DAVIDSON_GRAPH", 0);
    const bool use_graph = knob_graph != 0 && apply_ctl && !prm->profile && prm->nccl_comm == nullptr &&
                           prm->ss_op == nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    if (use_graph) {
        // first sigma build eagerly (it also configures the kernels' attributes outside the capture)
        if (apply(ws.V, ws.W, ws)) return -2;
        // a default stream (legacy or per-thread) cannot be captured: the loop then runs on a private
        // stream of this host thread, ordered after / before the caller's stream with events
        const bool dflt = st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread;
        cudaStream_t gst = dflt ? side->cap : st;
        if (dflt) {
            SQD_CUDA_OK(cudaEventRecord(side->ev_cap, st));
            SQD_CUDA_OK(cudaStreamWaitEvent(gst, side->ev_cap, 0));
        }
        const bool while_node = knob_graph == 1;
        const long long launches_before = thread_launches();
        int grc = davidson_graph_loop(n, apply_ctl, d_hdiag, ws, M, blocks, prm, side, gst, &graph, &graph_exec,
                                      while_node);
        if (grc == 0 && !while_node) {
            // host-replayed cycle graph: the launches are counted like the kernels they stand for
            const long long per_cycle = thread_launches() - launches_before;   // kernels captured in the graph
            add_launches(-per_cycle);                                           // (the capture launched nothing)
            for (cycle = 0; cycle < prm->max_cycle; ++cycle) {
                add_launches(per_cycle);
                if (cudaGraphLaunch(graph_exec, gst) != cudaSuccess) {
                    set_error("cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
                    grc = -2;
                    break;
                }
                if ((cycle + 1) % check_every == 0 || cycle + 1 == prm->max_cycle) {
                    if (read_back(&status, &ws.state->status, sizeof(int), gst)) {
                        grc = -2;
                        break;
                    }
                    if (status != 0) break;
                }
            }
        }
        if (grc != 0) {
            if (graph_exec) cudaGraphExecDestroy(graph_exec);
            if (graph) cudaGraphDestroy(graph);
            return grc;
        }
        if (dflt) {
            SQD_CUDA_OK(cudaEventRecord(side->ev_cap, gst));
            SQD_CUDA_OK(cudaStreamWaitEvent(st, side->ev_cap, 0));
        }
        sigma_builds = -1;  // filled in from the device-side cycle count below
    }
    for (; !use_graph && cycle < prm->max_cycle; ++cycle) {
        if (prm->profile) {
            cudaEvent_t e0, e1;
            SQD_CUDA_OK(cudaEventCreate(&e0));
            SQD_CUDA_OK(cudaEventCreate(&e1));
            ev.push_back(e0);
            ev.push_back(e1);
            SQD_CUDA_OK(cudaEventRecord(e0, st));
        }
        if (apply(ws.V + (int64_t)slot * n, ws.W + (int64_t)slot * n, ws)) return -2;
        if (prm->profile) SQD_CUDA_OK(cudaEventRecord(ev.back(), st));
        ++sigma_builds;
        // restart when the space is full (see restart_keep)
        const int q_keep = restart_keep(M);
        const int restart = (m == M) ? q_keep : 0;
        static const int knob_skip = env_knob("SQD_DAV_SKIP", 0);  // timing experiments: skip bit 0 residual, 1 ortho, 2 gram
        int rc = dispatch_mv(m, [&](auto mv) {
            constexpr int MV = decltype(mv)::value;
            if (!(knob_skip & 4)) gram_kernel<MV><<<blocks, kRedThreads, 0, st>>>(ws.state, ws.V, ws.W + (int64_t)slot * n, n, m,
                                                            ws.partials, restart ? 1 : 0);
            if (!(knob_skip & 1)) residual_kernel<MV><<<blocks, kRedThreads, 0, st>>>(ws.state, ws.V, ws.W, d_hdiag, n, m,
                                                                restart, prm->level_shift, ws.X, ws.T,
                                                                ws.partials, prm->tol, prm->tol_residual);
            return check_launch("davidson cycle (1)", 2);
        });
        if (rc) return -2;
        const int me = restart ? restart : m;
        rc = dispatch_mv(me, [&](auto mv) {
            constexpr int MV = decltype(mv)::value;
            if (!(knob_skip & 2)) {
                ortho1_kernel<MV><<<blocks, kRedThreads, 0, st>>>(ws.state, ws.V, n, me, ws.T, ws.partials,
                                                                  prm->lindep);
                ortho2_kernel<MV><<<blocks, kRedThreads, 0, st>>>(ws.state, ws.V, n, me, ws.T,
                                                                  ws.V + (int64_t)me * n);
            }
            return check_launch("davidson cycle (2)", 2);
        });
        if (rc) return -2;
        m = me + 1;
        slot = me;
        if ((cycle + 1) % check_every == 0 || cycle + 1 == prm->max_cycle) {
            if (read_back(&status, &ws.state->status, sizeof(int), st)) return -2;
            if (status != 0) {
                ++cycle;
                break;
            }
        }
    }
    DavState hs;
    SQD_CUDA_OK(cudaMemcpyAsync(d_x, ws.X, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (prm->profile) SQD_CUDA_OK(cudaEventRecord(ev_end, st));
    if (read_back(&hs, ws.state, sizeof(DavState), st)) return -2;
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (graph) cudaGraphDestroy(graph);
    if (sigma_builds < 0) sigma_builds = hs.cycles + (hs.status == 0 ? 1 : 0);
    static const int knob_debug = env_knob("SQD_DEBUG_RITZ", 0);
    if (knob_debug)
        fprintf(stderr, "[sqd] davidson: cycles %d, full decompositions %d, rqi iterations %d, status %d\n",
                hs.cycles, hs.n_jacobi, hs.n_rqi_iter, hs.status);
    double sigma_ms = 0.0, total_ms = 0.0;
    if (prm->profile) {
        // only cycles that ran before the device-side stop flag was raised did real work
        const int real = hs.cycles < (int)(ev.size() / 2) ? hs.cycles : (int)(ev.size() / 2);
        for (int i = 0; i < (int)(ev.size() / 2); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
            if (i < real) sigma_ms += ms;
            cudaEventDestroy(ev[2 * i]);
            cudaEventDestroy(ev[2 * i + 1]);
        }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev_begin, ev_end);
        total_ms = ms;
        cudaEventDestroy(ev_begin);
        cudaEventDestroy(ev_end);
    }
    if (h_info) {
        h_info->sigma_ms = sigma_ms;
        h_info->total_ms = total_ms;
        h_info->converged = hs.status == 1 ? 1 : (hs.status == 2 ? 2 : 0);
        h_info->cycles = hs.cycles;
        h_info->sigma_builds = sigma_builds;
        h_info->theta = hs.theta;
        h_info->residual = hs.rnorm;
    }
    return 0;
}


}  // namespace sqd

using namespace sqd;

extern "C" {

int64_t sqd_davidson_workspace_bytes(int na, int ldc, int max_space) {
    if (na <= 0 || ldc <= 0 || max_space < 2 || max_space > kMaxS) return -1;
    return workspace_bytes((int64_t)na * ldc, max_space);
}

int sqd_dot(const double* d_x, const double* d_y, int64_t n, double* d_out, double* d_scratch,
            void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = red_blocks(n);
    dot_partial_kernel<<<blocks, kRedThreads, 0, st>>>(d_x, d_y, n, d_scratch);
    dot_final_kernel<<<1, 32, 0, st>>>(d_scratch, blocks, d_out);
    return check_launch("dot kernels", 2);
}

int sqd_init_guess(const double* d_hdiag, int na, int nb, int ldc, double* d_x0, void* d_scratch,
                   void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(na > 0 && nb > 0 && ldc >= nb, "sqd_init_guess: bad shape");
    const int64_t n = (int64_t)na * ldc;
    const int blocks = red_blocks((int64_t)na * nb);
    double* pval = (double*)d_scratch;
    int64_t* pidx = (int64_t*)(pval + kRedBlocks);
    SQD_CUDA_OK(cudaMemsetAsync(d_x0, 0, n * sizeof(double), st));
    argmin_partial_kernel<<<blocks, kRedThreads, 0, st>>>(d_hdiag, na, nb, ldc, pval, pidx);
    init_guess_kernel<<<1, 32, 0, st>>>(pval, pidx, blocks, na, nb, ldc, d_x0);
    return check_launch("init_guess kernels", 2);
}

int sqd_fix_sign(double* d_x, int64_t n, void* d_scratch, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(n > 0, "sqd_fix_sign: empty vector");
    const int blocks = red_blocks(n);
    double* pval = (double*)d_scratch;
    int64_t* pidx = (int64_t*)(pval + kRedBlocks);
    absmax_partial_kernel<<<blocks, kRedThreads, 0, st>>>(d_x, n, pval, pidx);
    sign_flag_kernel<<<1, 32, 0, st>>>(pval, pidx, blocks, d_x);
    flip_kernel<<<blocks, kRedThreads, 0, st>>>(d_x, n, pidx + kRedBlocks);
    return check_launch("fix_sign kernels", 3);
}

int sqd_read_back(void* h_dst, const void* d_src, int64_t bytes, void* stream) {
    SQD_REQUIRE(bytes >= 0, "sqd_read_back: negative size");
    return read_back(h_dst, d_src, (size_t)bytes, (cudaStream_t)stream);
}

int sqd_occupancies(const double* d_c, const uint64_t* d_strs_a, int na, const uint64_t* d_strs_b,
                    int nb, int ldc, int norb, double* d_occ, double* d_scratch, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    double* wa = d_scratch;
    double* wb = d_scratch + na;
    row_weight_kernel<<<(na + 7) / 8, 256, 0, st>>>(d_c, na, nb, ldc, wa);
    col_weight_kernel<<<(nb + 127) / 128, 128, 0, st>>>(d_c, na, nb, ldc, wb);
    occupancy_kernel<<<2 * norb, kRedThreads, 0, st>>>(wa, d_strs_a, na, wb, d_strs_b, nb, norb, d_occ);
    return check_launch("occupancy kernels", 3);
}

int sqd_davidson(const sqd_operator* op, const double* d_hdiag, const double* d_x0, double* d_x,
                 void* d_workspace, int64_t ws_bytes, const sqd_davidson_params* prm,
                 sqd_davidson_info* h_info, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = (int64_t)op->a.n * op->ldc;
    auto apply = [&](const double* v, double* w, Workspace& ws) {
        return apply_operator(op, prm, v, w, ws, n, st);
    };
    ApplyCtlFn apply_ctl = [&](const double* vbase, double* wbase, const int* slot, Workspace& ws,
                               cudaStream_t cst) {
        return sigma_dispatch_ctl(op, vbase, wbase, &ws.state->status, slot, n, 1, cst);
    };
    return davidson_core(n, apply, d_hdiag, d_x0, d_x, d_workspace, ws_bytes, prm, h_info, st, apply_ctl);
}

int64_t sqd_csr_davidson_workspace_bytes(int64_t d, int k, int max_space) {
    if (d <= 0 || k != 1 || max_space < 2 || max_space > kMaxS) return -1;
    // Davidson workspace on the (re, im) embedding + embedded diagonal + start vector
    return workspace_bytes(2 * d, max_space) + 2 * (int64_t)align_up((size_t)2 * d * sizeof(double));
}

int sqd_csr_davidson(int64_t d, const int32_t* d_row_ptr, const int32_t* d_col, const double* d_val,
                     int k, int max_space, int max_cycle, double tol, const double* d_start,
                     double* d_evecs, double* h_evals, int* h_cycles, double* h_residual,
                     void* d_workspace, int64_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(k == 1, "sqd_csr_davidson: only the lowest eigenpair (k=1) runs natively (got k=%d)", k);
    SQD_REQUIRE(d > 0, "sqd_csr_davidson: empty matrix");
    SQD_REQUIRE(ws_bytes >= sqd_csr_davidson_workspace_bytes(d, k, max_space),
                "sqd_csr_davidson: workspace too small");
    // A complex Hermitian matrix acting on interleaved (re, im) storage is a real symmetric operator
    // on 2d doubles with the real inner product; every eigenvalue appears twice (z and i z), which
    // is harmless for the lowest pair.
    const int64_t n = 2 * d;
    char* p = (char*)d_workspace;
    double* hdiag = (double*)p;
    p += align_up((size_t)n * sizeof(double));
    double* x0 = (double*)p;
    p += align_up((size_t)n * sizeof(double));
    if (csr_diag_embed(d, d_row_ptr, d_col, d_val, hdiag, st)) return -2;
    // start vector: the caller's (complex128[d]) or the unit vector at argmin(diag).  A unit vector never
    // leaves its own connected component of a block-diagonal operator; the host driver (qubit.py) uses
    // Gershgorin bounds to decide which other components still have to be searched.
    if (d_start != nullptr) {
        SQD_CUDA_OK(cudaMemcpyAsync(x0, d_start, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    } else {
        // scratch for the argmin lives in the (not yet used) Davidson workspace
        if (sqd_init_guess(hdiag, 1, (int)n, (int)n, x0, p, stream)) return -2;
    }
    sqd_davidson_params prm;
    memset(&prm, 0, sizeof(prm));
    prm.max_space = max_space;
    prm.max_cycle = max_cycle;
    prm.tol = tol;
    prm.tol_residual = sqrt(tol);
    prm.lindep = 1e-14;
    prm.level_shift = 1e-4;
    prm.check_every = 4;
    sqd_davidson_info info;
    auto apply = [&](const double* v, double* w, Workspace& ws) {
        return csr_matvec_flag(&ws.state->status, d, d_row_ptr, d_col, d_val, v, w, st);
    };
    const int rc = davidson_core(n, apply, hdiag, x0, d_evecs, p,
                                 ws_bytes - 2 * (int64_t)align_up((size_t)n * sizeof(double)), &prm,
                                 &info, st);
    if (rc) return rc;
    h_evals[0] = info.theta;
    if (h_cycles) *h_cycles = info.cycles;
    if (h_residual) *h_residual = info.residual;
    return 0;
}

}  // extern "C"

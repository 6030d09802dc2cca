// Spin-resolved one-particle reduced density matrices from the in-set excitation tables.
// Replaces pyscf selected_ci.make_rdm1s (FCImake_rdm1a / FCImake_rdm1b through link tables), reached
// from qiskit_addon_sqd/fermion.py:117-121 (SCIState.rdm) and :725-729, :821-825.
//
//   dm1a[p,q] = <c| a+_p a_q |c> = sum_{(a <- a', pq, sgn) in S_a} sgn * <c[a,:], c[a',:]>     (p != q)
//   dm1a[p,p] = sum_{a: p in a} |c[a,:]|^2
// and the same for beta on the transposed matrix.  Deterministic: one warp computes each row-row dot
// product with a fixed tree, one CTA per (p,q) adds the products of its excitations in table order.
#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

__global__ void transpose_kernel(const double* __restrict__ c, int na, int nb, int ldc,
                                 double* __restrict__ ct, int ldt) {
    __shared__ double tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int a = by + j, b = bx + threadIdx.x;
        tile[j][threadIdx.x] = (a < na && b < nb) ? c[(size_t)a * ldc + b] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int b = bx + j, a = by + threadIdx.x;
        if (b < nb && a < ldt) ct[(size_t)b * ldt + a] = a < na ? tile[threadIdx.x][j] : 0.0;
    }
}

// dots[e] = <x[row(e),:], x[col(e),:]> for every single-excitation entry e; diag_w[i] = |x[i,:]|^2
__global__ void link_dots_kernel(const sqd_spin_table T, const double* __restrict__ x, int ncols, int ldx,
                                 double* __restrict__ dots, double* __restrict__ roww) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= T.n) return;
    const double* xi = x + (size_t)i * ldx;
    double w = 0.0;
    for (int b = lane; b < ncols; b += 32) w = fma(xi[b], xi[b], w);
    w = warp_sum(w);
    if (lane == 0) roww[i] = w;
    const int beg = T.row_ptr[i], ns = T.n_single[i];
    for (int e = beg; e < beg + ns; ++e) {
        const double* xj = x + (size_t)T.col[e] * ldx;
        double d = 0.0;
        for (int b = lane; b < ncols; b += 32) d = fma(xi[b], xj[b], d);
        d = warp_sum(d);
        if (lane == 0) dots[e] = d;
    }
}

// one CTA per (p,q): scan the table in order and add the signed dots of the excitations with this pq
__global__ void __launch_bounds__(256)
rdm1_collect_kernel(const sqd_spin_table T, int norb, const double* __restrict__ dots,
                    const double* __restrict__ roww, double* __restrict__ dm1) {
    __shared__ double red[8];
    const int p = blockIdx.x / norb, q = blockIdx.x % norb;
    double acc[1] = {0.0};
    if (p == q) {
        for (int i = threadIdx.x; i < T.n; i += blockDim.x)
            if ((T.strs[i] >> p) & 1ull) acc[0] += roww[i];
    } else {
        const uint32_t want = (uint32_t)(p * norb + q);
        for (int i = threadIdx.x; i < T.n; i += blockDim.x) {
            const int beg = T.row_ptr[i], ns = T.n_single[i];
            for (int e = beg; e < beg + ns; ++e) {
                const uint32_t m = T.meta[e];
                if ((m & 0x7fffffffu) == want) acc[0] += (m >> 31) ? -dots[e] : dots[e];
            }
        }
    }
    block_sum<1>(acc, red);
    if (threadIdx.x == 0) dm1[p * norb + q] = acc[0];
}


// ============================================================================================
// Two-particle reduced density matrices, pyscf convention  dm2[p,q,r,s] = <c| p+ r+ s q |c>
// (selected_ci.make_rdm2s / make_rdm2, reached from fermion.py:117-128 and :728-729, :825-826).
//
// pyscf resolves the same-spin blocks through (N-2)-electron intermediates (SCIrdm2_aaaa) and the
// opposite-spin block through t1a^T t1b dgemms (FCItdm12kern_ab).  Here every block comes straight from
// the in-set excitation tables:
//   same spin   : every ordered pair (t, s) of strings that differ by 0, 1 or 2 orbitals contributes
//                 <x[t,:], x[s,:]> to a fixed, disjoint family of cells --
//                   identical strings  -> [p,p,r,r] and [p,r,r,p]          (rdm2_diag_kernel)
//                   single excitation  -> [p,q,j,j] [j,j,p,q] [p,j,j,q] [j,q,p,j], j a spectator
//                                                                            (rdm2_singles_kernel)
//                   double excitation  -> the four antisymmetric images of [a1,i1,a2,i2]
//                                                                            (rdm2_doubles_kernel)
//   opposite    : dm2ab[pq,rs] = sum_{(a<-a',pq)} sum_{(b<-b',rs)} sgn_a sgn_b c[a,b] c[a',b'] with the
//                 single-excitation lists (diagonal p=q included) grouped by orbital pair.
// No atomics: every cell is accumulated by one thread in table order, so the result is bit-reproducible.
// ============================================================================================

constexpr uint32_t kNotDouble = 0xffffffffu;

// erow[e] = row of table entry e (one warp per row)
__global__ void entry_rows_kernel(const sqd_spin_table T, int* __restrict__ erow) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= T.n) return;
    for (int e = T.row_ptr[i] + lane; e < T.row_ptr[i + 1]; e += 32) erow[e] = i;
}

__global__ void row_norms_kernel(const double* __restrict__ x, int nrows, int ncols, int ldx,
                                 double* __restrict__ roww) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= nrows) return;
    const double* xi = x + (size_t)i * ldx;
    double w = 0.0;
    for (int b = lane; b < ncols; b += 32) w = fma(xi[b], xi[b], w);
    w = warp_sum(w);
    if (lane == 0) roww[i] = w;
}

// One warp per table entry e (grid-stride): dots[e] = <x[row(e),:], x[col(e),:]> (fixed shuffle tree);
// dinfo[e] = a1 | a2<<6 | i1<<12 | i2<<18 | parity<<31 for doubles (target = (-1)^parity a1+ a2+ i2 i1
// source, i1<i2 holes of the source, a1<a2 particles of the target), kNotDouble for singles
__global__ void __launch_bounds__(256)
pair_dots_kernel(const sqd_spin_table T, const int* __restrict__ erow, int nnz, const double* __restrict__ x,
                 int ncols, int ldx, double* __restrict__ dots, uint32_t* __restrict__ dinfo) {
    const int lane = threadIdx.x & 31;
    const int nwarp = gridDim.x * (blockDim.x >> 5);
    for (int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); e < nnz; e += nwarp) {
        const int i = erow[e], j = (int)T.col[e];
        const double* xi = x + (size_t)i * ldx;
        const double* xj = x + (size_t)j * ldx;
        double d = 0.0;
        for (int b = lane; b < ncols; b += 32) d = fma(xi[b], xj[b], d);
        d = warp_sum(d);
        if (lane == 0) {
            dots[e] = d;
            uint32_t info = kNotDouble;
            if (e >= T.row_ptr[i] + T.n_single[i]) {
                const uint64_t s = T.strs[j], t = T.strs[i];
                const uint64_t xo = s ^ t;
                uint64_t holes = xo & s, parts = xo & t;
                const int i1 = lowbit64(holes);
                holes &= holes - 1;
                const int i2 = lowbit64(holes);
                const int a1 = lowbit64(parts);
                parts &= parts - 1;
                const int a2 = lowbit64(parts);
                int par = popc64(s & below_mask(i1));
                uint64_t u = s ^ (1ull << i1);
                par += popc64(u & below_mask(i2));
                u ^= (1ull << i2);
                par += popc64(u & below_mask(a2));
                u |= (1ull << a2);
                par += popc64(u & below_mask(a1));
                info = (uint32_t)a1 | ((uint32_t)a2 << 6) | ((uint32_t)i1 << 12) | ((uint32_t)i2 << 18) |
                       ((uint32_t)(par & 1) << 31);
            }
            dinfo[e] = info;
        }
    }
}

constexpr int kD2Threads = 256;
constexpr int kD2Cap = 2048;   // entries per segment = capacity of the match list: a segment never overflows it

// One CTA per row [P,Q,:,:] of the same-spin dm2.  Writes the WHOLE row (zeros included), so it runs
// before the singles / diagonal kernels, which then overwrite their own (disjoint) cells.
// The table is scanned in segments of kD2Cap entries: threads append their matches to a shared list without
// any barrier (arrival order), then the list is put in entry order by rank counting and every cell's owner
// thread adds its contributions in that order -- every cell sees its contributions in ascending table order,
// whatever the arrival order was (bit-reproducible).  Four barriers per segment; the first version compacted
// with a ballot scan and three barriers per 256 entries, which was most of its 130 us.
__global__ void __launch_bounds__(kD2Threads)
rdm2_doubles_kernel(const uint32_t* __restrict__ dinfo, const double* __restrict__ dots, int nnz, int norb,
                    double* __restrict__ dm2) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n2 = norb * norb;
    double* acc = reinterpret_cast<double*>(smem_raw);   // [n2]
    double* lval = acc + n2;                              // [kD2Cap]
    int* lcell = reinterpret_cast<int*>(lval + kD2Cap);   // [kD2Cap]
    int* lent = lcell + kD2Cap;                           // [kD2Cap] table entry of the match
    int* order = lent + kD2Cap;                           // [kD2Cap] list position of the r-th match in entry order
    __shared__ int list_n;
    const int tid = threadIdx.x;
    const int P = blockIdx.x / norb, Q = blockIdx.x % norb;
    for (int c = tid; c < n2; c += kD2Threads) acc[c] = 0.0;
    if (tid == 0) list_n = 0;
    __syncthreads();
    if (P != Q) {
        for (int seg = 0; seg < nnz; seg += kD2Cap) {
            const int seg_end = seg + kD2Cap < nnz ? seg + kD2Cap : nnz;
            for (int e = seg + tid; e < seg_end; e += kD2Threads) {
                const uint32_t info = __ldg(dinfo + e);
                if (info == kNotDouble) continue;
                const int a1 = info & 63, a2 = (info >> 6) & 63, i1 = (info >> 12) & 63, i2 = (info >> 18) & 63;
                const int pa = P == a1 ? 0 : (P == a2 ? 1 : -1);
                const int qi = Q == i1 ? 0 : (Q == i2 ? 1 : -1);
                if (pa < 0 || qi < 0) continue;
                // [a1,i1,a2,i2] carries the table phase; swapping the creators or the annihilators flips
                // the sign
                const int R = pa ? a1 : a2, S = qi ? i1 : i2;
                const int neg = (int)(info >> 31) ^ pa ^ qi;
                const double d = __ldg(dots + e);
                const int o = atomicAdd(&list_n, 1);
                lent[o] = e;
                lcell[o] = R * norb + S;
                lval[o] = neg ? -d : d;
            }
            __syncthreads();
            const int ln = list_n;
            for (int k = tid; k < ln; k += kD2Threads) {
                const int ek = lent[k];
                int r = 0;
                for (int l = 0; l < ln; ++l) r += lent[l] < ek;
                order[r] = k;
            }
            __syncthreads();
            for (int r = 0; r < ln; ++r) {
                const int k = order[r];
                const int c = lcell[k];
                if ((c & (kD2Threads - 1)) == tid) acc[c] += lval[k];  // cell owner: fixed thread, entry order
            }
            __syncthreads();
            if (tid == 0) list_n = 0;
            __syncthreads();
        }
    }
    double* row = dm2 + (size_t)blockIdx.x * n2;
    for (int c = tid; c < n2; c += kD2Threads) row[c] = acc[c];
}

// One warp per (p,q), p != q; lane j (and j+32) owns the spectator orbital j.  The single excitations
// with this orbital pair come from the grouped list (ent / ent_e), in table order.
__global__ void __launch_bounds__(128)
rdm2_singles_kernel(const sqd_spin_table T, const int* __restrict__ gptr, const int2* __restrict__ ent,
                    const int* __restrict__ ent_e, const double* __restrict__ dots, int norb,
                    double* __restrict__ dm2) {
    const int lane = threadIdx.x & 31;
    const int pq = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n2 = norb * norb;
    if (pq >= n2) return;
    const int p = pq / norb, q = pq % norb;
    if (p == q) return;
    double g0 = 0.0, g1 = 0.0;
    for (int k = gptr[pq]; k < gptr[pq + 1]; ++k) {
        const int2 E = ent[k];
        // target strs[E.x] = sgn * p+ q source; spectators = occupied orbitals of the source but q
        const uint64_t src = T.strs[E.y & 0x7fffffff] & ~(1ull << q);
        const double d0 = dots[ent_e[k]];
        const double d = E.y < 0 ? -d0 : d0;
        if ((src >> lane) & 1ull) g0 += d;
        if ((src >> (lane + 32)) & 1ull) g1 += d;
    }
    const size_t n1 = norb, n3 = (size_t)n2 * norb;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int j = lane + 32 * half;
        const double g = half ? g1 : g0;
        if (j < norb && j != p && j != q) {
            dm2[p * n3 + q * (size_t)n2 + j * n1 + j] = g;    // <p+ j+ j q>
            dm2[j * n3 + j * (size_t)n2 + p * n1 + q] = g;    // <j+ p+ q j>
            dm2[p * n3 + j * (size_t)n2 + j * n1 + q] = -g;   // <p+ j+ q j>
            dm2[j * n3 + q * (size_t)n2 + p * n1 + j] = -g;   // <j+ p+ j q>
        }
    }
}

// dm2[p,p,r,r] = sum_{s: p,r in s} |x[s,:]|^2 = -dm2[p,r,r,p]   (p != r); one CTA per p, thread r
__global__ void __launch_bounds__(64)
rdm2_diag_kernel(const sqd_spin_table T, const double* __restrict__ roww, int norb,
                 double* __restrict__ dm2) {
    const int p = blockIdx.x, r = threadIdx.x;
    if (r >= norb || r == p) return;
    const uint64_t need = (1ull << p) | (1ull << r);
    double acc = 0.0;
    for (int i = 0; i < T.n; ++i)
        if ((T.strs[i] & need) == need) acc += roww[i];
    const size_t n1 = norb, n2 = n1 * n1, n3 = n2 * n1;
    dm2[p * n3 + p * n2 + r * n1 + r] = acc;
    dm2[p * n3 + r * n2 + r * n1 + p] = -acc;
}

// ---- single-excitation lists grouped by orbital pair (diagonal p == q included) ---------------
// entry = {target index, source index | sign << 31}
__device__ __forceinline__ bool group_probe(const sqd_spin_table& T, int i, int p, int q, int pq,
                                            uint32_t* colsign, int* entry) {
    *entry = -1;
    if (p == q) {
        *colsign = (uint32_t)i;
        return (T.strs[i] >> p) & 1ull;
    }
    const int beg = T.row_ptr[i], ns = T.n_single[i];
    for (int e = beg; e < beg + ns; ++e) {
        const uint32_t m = __ldg(T.meta + e);
        if ((int)(m & 0x7fffffffu) == pq) {
            *colsign = T.col[e] | (m & 0x80000000u);
            *entry = e;
            return true;
        }
    }
    return false;
}

__global__ void __launch_bounds__(256)
group_count_kernel(const sqd_spin_table T, int norb, int* __restrict__ cnt) {
    __shared__ int red[8];
    const int pq = blockIdx.x, p = pq / norb, q = pq % norb;
    int c = 0;
    uint32_t dummy;
    int dummy_e;
    for (int i = threadIdx.x; i < T.n; i += blockDim.x)
        c += group_probe(T, i, p, q, pq, &dummy, &dummy_e) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += red[w];
        cnt[pq] = s;
    }
}

__global__ void __launch_bounds__(256)
group_fill_kernel(const sqd_spin_table T, int norb, const int* __restrict__ ptr, int2* __restrict__ ent,
                  int* __restrict__ ent_e) {
    __shared__ int wcnt[8];
    const int pq = blockIdx.x, p = pq / norb, q = pq % norb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int base = ptr[pq];
    for (int i0 = 0; i0 < T.n; i0 += 256) {
        const int i = i0 + threadIdx.x;
        uint32_t cs = 0;
        int en = -1;
        const bool hit = i < T.n && group_probe(T, i, p, q, pq, &cs, &en);
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            if (w < warp) woff += wcnt[w];
            total += wcnt[w];
        }
        if (hit) {
            const int o = base + woff + __popc(bal & ((1u << lane) - 1u));
            ent[o] = make_int2(i, (int)cs);
            ent_e[o] = en;
        }
        base += total;
        __syncthreads();
    }
}

// acc[rs] += sa * sum_{(b <- b', rs)} sgn_b X[b] Y[b'] for every beta pair rs: off-diagonal groups (a few
// entries each) by one thread per rs, the long diagonal groups rs = rr by one warp each.  Every acc[rs]
// has exactly one writer.
__device__ __forceinline__ void ab_accumulate(const double* X, const double* Y, double sa, int norb,
                                              const int* __restrict__ ptr_b, const int2* __restrict__ ent_b,
                                              double* acc, bool overwrite) {
    const int n2 = norb * norb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
    for (int rs = tid; rs < n2; rs += blockDim.x) {
        if (rs / norb == rs % norb) continue;
        double s = 0.0;
        for (int eb = ptr_b[rs]; eb < ptr_b[rs + 1]; ++eb) {
            const int2 B = ent_b[eb];
            const double t = X[B.x] * Y[B.y & 0x7fffffff];
            s += B.y < 0 ? -t : t;
        }
        acc[rs] = overwrite ? sa * s : fma(sa, s, acc[rs]);
    }
    for (int r = warp; r < norb; r += nwarp) {
        const int rs = r * norb + r;
        double s = 0.0;
        for (int eb = ptr_b[rs] + lane; eb < ptr_b[rs + 1]; eb += 32) {
            const int bb = ent_b[eb].x;
            s = fma(X[bb], Y[bb], s);
        }
        s = warp_sum(s);
        if (lane == 0) acc[rs] = overwrite ? sa * s : fma(sa, s, acc[rs]);
    }
}

// Diagonal alpha pairs: dm2ab[pp, rs] = sum_{a: p in a} T[a][rs] with the per-row beta transition
// T[a][rs] = sum_{(b <- b', rs)} sgn_b c[a,b] c[a,b'].  One CTA per alpha string builds its T row ...
__global__ void __launch_bounds__(256)
rdm2_ab_rows_kernel(const double* __restrict__ c, int nb, int ldc, int norb, const int* __restrict__ ptr_b,
                    const int2* __restrict__ ent_b, double* __restrict__ Trows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* X = reinterpret_cast<double*>(smem_raw);
    const int a = blockIdx.x;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) X[b] = c[(size_t)a * ldc + b];
    __syncthreads();
    ab_accumulate(X, X, 1.0, norb, ptr_b, ent_b, Trows + (size_t)a * norb * norb, true);
}

// ... and the rows are added in string order (one thread per rs, fixed order -> reproducible)
__global__ void __launch_bounds__(256)
rdm2_ab_diag_kernel(const uint64_t* __restrict__ strs_a, int na, int norb, const double* __restrict__ Trows,
                    double* __restrict__ dm2ab) {
    const int n2 = norb * norb;
    const int p = blockIdx.x;
    const int rs = blockIdx.y * blockDim.x + threadIdx.x;
    if (rs >= n2) return;
    double acc = 0.0;
    for (int a = 0; a < na; ++a)
        if ((strs_a[a] >> p) & 1ull) acc += Trows[(size_t)a * n2 + rs];
    dm2ab[(size_t)(p * norb + p) * n2 + rs] = acc;
}

// Off-diagonal alpha pairs: one CTA per row [p,q,:,:] of dm2ab (p != q).  For every alpha excitation
// (a <- a') of the pair the rows c[a,:] and c[a',:] are staged in shared memory.
__global__ void __launch_bounds__(256)
rdm2_ab_kernel(const double* __restrict__ c, int nb, int ldc, int norb, const int* __restrict__ ptr_a,
               const int2* __restrict__ ent_a, const int* __restrict__ ptr_b,
               const int2* __restrict__ ent_b, double* __restrict__ dm2ab) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n2 = norb * norb;
    const int pq = blockIdx.x;
    if (pq / norb == pq % norb) return;  // written by rdm2_ab_diag_kernel
    double* X = reinterpret_cast<double*>(smem_raw);  // c[a ,:]  (bra)
    double* Y = X + ldc;                                // c[a',:]  (ket)
    double* acc = Y + ldc;                              // [n2]
    const int tid = threadIdx.x;
    for (int k = tid; k < n2; k += blockDim.x) acc[k] = 0.0;
    const int ea_beg = ptr_a[pq], ea_end = ptr_a[pq + 1];
    for (int ea = ea_beg; ea < ea_end; ++ea) {
        const int2 A = ent_a[ea];
        const int a = A.x, a1 = A.y & 0x7fffffff;
        const double sa = A.y < 0 ? -1.0 : 1.0;
        __syncthreads();
        for (int b = tid; b < nb; b += blockDim.x) {
            X[b] = c[(size_t)a * ldc + b];
            Y[b] = c[(size_t)a1 * ldc + b];
        }
        __syncthreads();
        ab_accumulate(X, Y, sa, norb, ptr_b, ent_b, acc, false);
    }
    __syncthreads();
    double* row = dm2ab + (size_t)pq * n2;
    for (int k = tid; k < n2; k += blockDim.x) row[k] = acc[k];
}

}  // namespace sqd

using namespace sqd;

extern "C" {

int64_t sqd_rdm1s_workspace_bytes(const sqd_operator* op) {
    const int64_t na = op->a.n, nb = op->b.n;
    const int64_t ldt = (na + 1) / 2 * 2;
    int64_t nnz = 0;  // the caller knows nnz; bound the dots array by the table sizes passed separately
    (void)nnz;
    return (nb * ldt + na + nb) * (int64_t)sizeof(double);
}

int sqd_rdm1s(const sqd_operator* op, const double* d_c, int64_t nnz_a, int64_t nnz_b, double* d_dm1,
              double* d_workspace, double* d_dots, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int na = op->a.n, nb = op->b.n, ldc = op->ldc, norb = op->norb;
    const int ldt = (na + 1) / 2 * 2;
    SQD_REQUIRE(nnz_a >= 0 && nnz_b >= 0 && d_dots != nullptr, "sqd_rdm1s: dots scratch missing");
    double* ct = d_workspace;               // [nb][ldt]
    double* rw_a = ct + (size_t)nb * ldt;   // [na]
    double* rw_b = rw_a + na;               // [nb]
    // alpha: rows of c
    link_dots_kernel<<<(na + 7) / 8, 256, 0, st>>>(op->a, d_c, nb, ldc, d_dots, rw_a);
    rdm1_collect_kernel<<<norb * norb, 256, 0, st>>>(op->a, norb, d_dots, rw_a, d_dm1);
    // beta: rows of c^T
    dim3 tb(32, 8), tg((nb + 31) / 32, (na + 31) / 32);
    transpose_kernel<<<tg, tb, 0, st>>>(d_c, na, nb, ldc, ct, ldt);
    link_dots_kernel<<<(nb + 7) / 8, 256, 0, st>>>(op->b, ct, na, ldt, d_dots, rw_b);
    rdm1_collect_kernel<<<norb * norb, 256, 0, st>>>(op->b, norb, d_dots, rw_b, d_dm1 + norb * norb);
    return check_launch("rdm1s kernels", 5);
}

}  // extern "C"

static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

extern "C" int64_t sqd_rdm2s_workspace_bytes(const sqd_operator* op, int64_t nnz_a, int64_t nnz_b) {
    const int64_t na = op->a.n, nb = op->b.n, norb = op->norb, n2 = norb * norb;
    const int64_t ldt = (na + 1) / 2 * 2;
    const int64_t nnz = nnz_a > nnz_b ? nnz_a : nnz_b;
    size_t b = 0;
    b += al256((size_t)nb * ldt * sizeof(double));          // c^T
    b += al256((size_t)(na + nb) * sizeof(double));         // row weights
    b += al256((size_t)(nnz + 1) * sizeof(double));         // dots
    b += al256((size_t)(nnz + 1) * sizeof(uint32_t));       // dinfo
    b += al256((size_t)(nnz + 1) * sizeof(int));            // entry -> row
    b += 4 * al256((size_t)(n2 + 1) * sizeof(int));         // cnt/ptr for both spins
    b += al256((size_t)(nnz_a + na * norb + 1) * sizeof(int2)) + al256((size_t)(nnz_a + na * norb + 1) * sizeof(int));
    b += al256((size_t)(nnz_b + nb * norb + 1) * sizeof(int2)) + al256((size_t)(nnz_b + nb * norb + 1) * sizeof(int));
    b += al256((size_t)na * n2 * sizeof(double));           // per-row beta transitions
    return (int64_t)b;
}

template <typename K>
static int big_smem(K kern, size_t smem, bool* cfg) {
    if (smem > 48 * 1024) {
        int dev = 0;
        SQD_CUDA_OK(cudaGetDevice(&dev));
        if (dev < 0 || dev >= 64 || !cfg[dev]) {
            // (a kernel's static shared memory counts against the same 227 KB: leave it 2 KB)
            SQD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(225 * 1024)));
            if (dev >= 0 && dev < 64) cfg[dev] = true;
        }
    }
    return 0;
}

// same-spin block of dm2 (+ the spin's dm1 when asked) for table T over the rows of x
static int rdm2_same_spin(const sqd_spin_table& T, int64_t nnz, const double* x, int ncols, int ldx,
                          int norb, const int* gptr, const int2* ent, const int* ent_e, int* erow,
                          double* dots, double* roww, uint32_t* dinfo, double* dm2, double* dm1,
                          cudaStream_t st) {
    const int n2 = norb * norb;
    row_norms_kernel<<<(T.n + 7) / 8, 256, 0, st>>>(x, T.n, ncols, ldx, roww);
    if (nnz > 0) {
        entry_rows_kernel<<<(T.n + 7) / 8, 256, 0, st>>>(T, erow);
        const int blocks = (int)((nnz + 7) / 8 < kNumSMs * 8 ? (nnz + 7) / 8 : kNumSMs * 8);
        pair_dots_kernel<<<blocks, 256, 0, st>>>(T, erow, (int)nnz, x, ncols, ldx, dots, dinfo);
    }
    const size_t smem = (size_t)n2 * sizeof(double) + kD2Cap * (sizeof(double) + 3 * sizeof(int));
    static bool cfg[64] = {false};
    if (big_smem(rdm2_doubles_kernel, smem, cfg)) return -2;
    rdm2_doubles_kernel<<<n2, kD2Threads, smem, st>>>(dinfo, dots, (int)nnz, norb, dm2);
    rdm2_singles_kernel<<<(n2 + 3) / 4, 128, 0, st>>>(T, gptr, ent, ent_e, dots, norb, dm2);
    rdm2_diag_kernel<<<norb, 64, 0, st>>>(T, roww, norb, dm2);
    if (dm1 != nullptr) rdm1_collect_kernel<<<n2, 256, 0, st>>>(T, norb, dots, roww, dm1);
    return check_launch("rdm2 same-spin kernels", 7);
}

extern "C" int sqd_rdm2s(const sqd_operator* op, const double* d_c, int64_t nnz_a, int64_t nnz_b,
                         double* d_dm2aa, double* d_dm2ab, double* d_dm2bb, double* d_dm1,
                         void* d_workspace, int64_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int na = op->a.n, nb = op->b.n, ldc = op->ldc, norb = op->norb, n2 = norb * norb;
    const int ldt = (na + 1) / 2 * 2;
    SQD_REQUIRE(norb >= 1 && norb <= 64, "sqd_rdm2s: norb must be in [1, 64]");
    SQD_REQUIRE(ws_bytes >= sqd_rdm2s_workspace_bytes(op, nnz_a, nnz_b), "sqd_rdm2s: workspace too small");
    const int64_t nnz = nnz_a > nnz_b ? nnz_a : nnz_b;
    char* p = (char*)d_workspace;
    double* ct = (double*)p;        p += al256((size_t)nb * ldt * sizeof(double));
    double* roww = (double*)p;      p += al256((size_t)(na + nb) * sizeof(double));
    double* dots = (double*)p;      p += al256((size_t)(nnz + 1) * sizeof(double));
    uint32_t* dinfo = (uint32_t*)p; p += al256((size_t)(nnz + 1) * sizeof(uint32_t));
    int* erow = (int*)p;            p += al256((size_t)(nnz + 1) * sizeof(int));
    int* cnt_a = (int*)p;           p += al256((size_t)(n2 + 1) * sizeof(int));
    int* ptr_a = (int*)p;           p += al256((size_t)(n2 + 1) * sizeof(int));
    int* cnt_b = (int*)p;           p += al256((size_t)(n2 + 1) * sizeof(int));
    int* ptr_b = (int*)p;           p += al256((size_t)(n2 + 1) * sizeof(int));
    int2* ent_a = (int2*)p;         p += al256((size_t)(nnz_a + (int64_t)na * norb + 1) * sizeof(int2));
    int* ente_a = (int*)p;          p += al256((size_t)(nnz_a + (int64_t)na * norb + 1) * sizeof(int));
    int2* ent_b = (int2*)p;         p += al256((size_t)(nnz_b + (int64_t)nb * norb + 1) * sizeof(int2));
    int* ente_b = (int*)p;          p += al256((size_t)(nnz_b + (int64_t)nb * norb + 1) * sizeof(int));
    double* Trows = (double*)p;
    // single-excitation lists grouped by orbital pair, both spins
    group_count_kernel<<<n2, 256, 0, st>>>(op->a, norb, cnt_a);
    group_count_kernel<<<n2, 256, 0, st>>>(op->b, norb, cnt_b);
    if (check_launch("group_count_kernel", 2)) return -2;
    if (sqd_exclusive_scan(cnt_a, ptr_a, n2, nullptr, stream)) return -2;
    if (sqd_exclusive_scan(cnt_b, ptr_b, n2, nullptr, stream)) return -2;
    group_fill_kernel<<<n2, 256, 0, st>>>(op->a, norb, ptr_a, ent_a, ente_a);
    group_fill_kernel<<<n2, 256, 0, st>>>(op->b, norb, ptr_b, ent_b, ente_b);
    if (check_launch("group_fill_kernel", 2)) return -2;
    // same spin, alpha: rows of c
    if (rdm2_same_spin(op->a, nnz_a, d_c, nb, ldc, norb, ptr_a, ent_a, ente_a, erow, dots, roww, dinfo,
                       d_dm2aa, d_dm1, st))
        return -2;
    // same spin, beta: rows of c^T
    dim3 tb(32, 8), tg((nb + 31) / 32, (na + 31) / 32);
    transpose_kernel<<<tg, tb, 0, st>>>(d_c, na, nb, ldc, ct, ldt);
    if (check_launch("transpose_kernel")) return -2;
    if (rdm2_same_spin(op->b, nnz_b, ct, na, ldt, norb, ptr_b, ent_b, ente_b, erow, dots, roww + na, dinfo,
                       d_dm2bb, d_dm1 ? d_dm1 + n2 : nullptr, st))
        return -2;
    // opposite spin
    const size_t smem_rows = (size_t)ldc * sizeof(double);
    const size_t smem_ab = (size_t)(2 * ldc + n2) * sizeof(double);
    SQD_REQUIRE(smem_ab <= 225 * 1024, "sqd_rdm2s: nb=%d, norb=%d do not fit the shared-memory row staging",
                nb, norb);
    static bool cfg_rows[64] = {false}, cfg_ab[64] = {false};
    if (big_smem(rdm2_ab_rows_kernel, smem_rows, cfg_rows)) return -2;
    if (big_smem(rdm2_ab_kernel, smem_ab, cfg_ab)) return -2;
    rdm2_ab_rows_kernel<<<na, 256, smem_rows, st>>>(d_c, nb, ldc, norb, ptr_b, ent_b, Trows);
    rdm2_ab_diag_kernel<<<dim3(norb, (n2 + 255) / 256), 256, 0, st>>>(op->a.strs, na, norb, Trows, d_dm2ab);
    rdm2_ab_kernel<<<n2, 256, smem_ab, st>>>(d_c, nb, ldc, norb, ptr_a, ent_a, ptr_b, ent_b, d_dm2ab);
    return check_launch("rdm2 opposite-spin kernels", 3);
}

// Determinant-string excitation tables, opposite-spin contraction tables and operator diagonals.
//
// Replaces what the reference delegates to pyscf (recalled; not under /root/reference):
//   selected_ci._all_linkstr_index -> SCIcre_des_linkstr / SCIdes_des_linkstr, make_hdiag,
// reached from qiskit_addon_sqd/fermion.py:721-723 and :810-818.  B200 design: strings are
// bit-packed uint64 loaded coalesced; partners are found by all-pairs xor + popcount with
// warp-ballot compaction (no hash / no binary search), matrix elements follow Slater-Condon rules.
#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

// --------------------------------------------------------------------------------------------
// bit packing:  bool (n, nbits) big-endian  ->  (left, right) uint64 halves
// --------------------------------------------------------------------------------------------
__global__ void pack_bitstrings_kernel(const uint8_t* __restrict__ bits, int64_t n, int nbits,
                                       uint64_t* __restrict__ left, uint64_t* __restrict__ right) {
    // one warp per row: lanes read consecutive bytes (coalesced), ballot assembles the word
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const int half = nbits >> 1;
    const uint8_t* src = bits + row * nbits;
    for (int side = 0; side < 2; ++side) {
        uint64_t word = 0;
        // column j of the half has weight 2^(half-1-j)
        for (int j0 = 0; j0 < half; j0 += 32) {
            const int j = j0 + lane;
            const bool bit = (j < half) && src[side * half + j] != 0;
            const uint32_t m = __ballot_sync(0xffffffffu, bit);  // bit l of m <-> column j0+l
            // reverse so that column j0 becomes the most significant of this chunk
            const uint32_t rev = __brev(m);
            const int chunk = min(32, half - j0);
            const uint64_t v = (uint64_t)(rev >> (32 - chunk));
            word |= v << (half - j0 - chunk);
        }
        if (lane == 0) (side == 0 ? left : right)[row] = word;
    }
}

// --------------------------------------------------------------------------------------------
// pass 1: count partners.  One warp per string; lanes stride the partner index.
// --------------------------------------------------------------------------------------------
__global__ void excitation_count_kernel(const uint64_t* __restrict__ strs, int n,
                                        int* __restrict__ n_single, int* __restrict__ n_total) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint64_t si = strs[i];
    int c1 = 0, c2 = 0;
    for (int j = lane; j < n; j += 32) {
        const int pc = popc64(si ^ strs[j]);
        c1 += (pc == 2);
        c2 += (pc == 4);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
    }
    if (lane == 0) {
        n_single[i] = c1;
        n_total[i] = c1 + c2;
    }
}

// --------------------------------------------------------------------------------------------
// exclusive scan (single CTA, n up to a few million: this is setup, not the hot loop)
// --------------------------------------------------------------------------------------------
__global__ void exclusive_scan_kernel(const int* __restrict__ in, int* __restrict__ out, int n) {
    // the running total is kept in 64 bits: a total beyond INT_MAX is reported as out[n] = -1 (the callers
    // size 32-bit indexed tables from it and must refuse) instead of wrapping around silently
    __shared__ int warp_tot[32];
    __shared__ long long carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int idx = base + threadIdx.x;
        const int v = idx < n ? in[idx] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_tot[warp] = s;
        __syncthreads();
        if (warp == 0) {
            int w = lane < nwarp ? warp_tot[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;  // inclusive over warps (one tile of <= 1024 rows of < 2^20 entries: no overflow)
        }
        __syncthreads();
        const long long carry = carry_s;
        const int woff = warp == 0 ? 0 : warp_tot[warp - 1];
        if (idx < n) out[idx] = (int)(carry + woff + s - v);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + warp_tot[nwarp - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s > 2147483647ll ? -1 : (int)carry_s;
}

// --------------------------------------------------------------------------------------------
// Slater-Condon elements
// --------------------------------------------------------------------------------------------
// <t| H_same-spin |s> for t = a+_p a_q s (p != q), sign included.
__device__ __forceinline__ double single_element(uint64_t s, int p, int q, int norb,
                                                 const double* __restrict__ h,
                                                 const double* __restrict__ g, int* sign_out) {
    const int sign = (popc64(s & between_mask(p, q)) & 1) ? -1 : 1;
    const int64_t n1 = norb, n2 = n1 * n1, n3 = n2 * n1;
    double v = h[p * n1 + q];
    uint64_t occ = s;
    while (occ) {
        const int k = lowbit64(occ);
        occ &= occ - 1;
        v += g[p * n3 + q * n2 + k * n1 + k] - g[p * n3 + k * n2 + k * n1 + q];
    }
    *sign_out = sign;
    return sign * v;
}

// <t| H |s> for a double excitation: holes i1<i2 (in s), particles a1<a2 (in t)
__device__ __forceinline__ double double_element(uint64_t s, uint64_t t, int norb,
                                                 const double* __restrict__ g) {
    const uint64_t x = s ^ t;
    uint64_t holes = x & s, parts = x & t;
    const int i1 = lowbit64(holes);
    holes &= holes - 1;
    const int i2 = lowbit64(holes);
    const int a1 = lowbit64(parts);
    parts &= parts - 1;
    const int a2 = lowbit64(parts);
    // phase of a+_{a1} a+_{a2} a_{i2} a_{i1} |s>
    int par = popc64(s & below_mask(i1));
    uint64_t u = s ^ (1ull << i1);
    par += popc64(u & below_mask(i2));
    u ^= (1ull << i2);
    par += popc64(u & below_mask(a2));
    u |= (1ull << a2);
    par += popc64(u & below_mask(a1));
    const int64_t n1 = norb, n2 = n1 * n1, n3 = n2 * n1;
    const double v = g[a1 * n3 + i1 * n2 + a2 * n1 + i2] - g[a1 * n3 + i2 * n2 + a2 * n1 + i1];
    return (par & 1) ? -v : v;
}

__device__ __forceinline__ double diagonal_element(uint64_t s, int norb, const double* __restrict__ h,
                                                   const double* __restrict__ g) {
    const int64_t n1 = norb, n2 = n1 * n1, n3 = n2 * n1;
    double e = 0.0;
    uint64_t oi = s;
    while (oi) {
        const int i = lowbit64(oi);
        oi &= oi - 1;
        e += h[i * n1 + i];
        uint64_t oj = s;
        while (oj) {
            const int j = lowbit64(oj);
            oj &= oj - 1;
            e += 0.5 * (g[i * n3 + i * n2 + j * n1 + j] - g[i * n3 + j * n2 + j * n1 + i]);
        }
    }
    return e;
}

// --------------------------------------------------------------------------------------------
// pass 2: fill.  Structure first -- one warp per row (target t = strs[i]), ballot/popcount compaction keeps
// the "singles first, ascending" then "doubles, ascending" order -- then the values: one THREAD per table entry
// (and per diagonal element).  The Slater-Condon evaluation of an entry is a dependent chain of ~2 n_elec
// integral loads; with one warp per row (round 1) only ~300 warps shared that latency and the kernel took
// 110 us at 316 strings; per entry, 24 000 threads do.
// --------------------------------------------------------------------------------------------
__global__ void excitation_fill_kernel(const uint64_t* __restrict__ strs, int n, int norb,
                                       const int* __restrict__ row_ptr,
                                       const int* __restrict__ n_single, uint32_t* __restrict__ col,
                                       uint32_t* __restrict__ meta, uint32_t* __restrict__ pack) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const uint64_t t = strs[i];
    int pos1 = row_ptr[i];
    int pos2 = pos1 + n_single[i];
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        const uint64_t s = j < n ? strs[j] : t;
        const int pc = popc64(s ^ t);
        const uint32_t m1 = __ballot_sync(0xffffffffu, pc == 2);
        const uint32_t m2 = __ballot_sync(0xffffffffu, pc == 4);
        const uint32_t lt = (1u << lane) - 1u;
        if (pc == 2) {
            const int q = lowbit64((s ^ t) & s);  // annihilated in the source
            const int p = lowbit64((s ^ t) & t);  // created in the target
            const int sign = (popc64(s & between_mask(p, q)) & 1) ? -1 : 1;
            const int o = pos1 + __popc(m1 & lt);
            col[o] = (uint32_t)j;
            meta[o] = (uint32_t)(p * norb + q) | (sign < 0 ? 0x80000000u : 0u);
            pack[o] = (uint32_t)j | ((uint32_t)(p * norb + q) << 19) | (sign < 0 ? 0x80000000u : 0u);
        } else if (pc == 4) {
            const int o = pos2 + __popc(m2 & lt);
            col[o] = (uint32_t)j;
            meta[o] = 0u;
            pack[o] = (uint32_t)j;
        }
        pos1 += __popc(m1);
        pos2 += __popc(m2);
    }
}

__global__ void excitation_values_kernel(const uint64_t* __restrict__ strs, int n, int norb,
                                         const double* __restrict__ h, const double* __restrict__ g,
                                         const int* __restrict__ row_ptr, const int* __restrict__ n_single,
                                         const uint32_t* __restrict__ col, const uint32_t* __restrict__ meta,
                                         double* __restrict__ val, double* __restrict__ diag) {
    const bool have_ints = (h != nullptr) && (g != nullptr);  // structure-only tables when NULL
    const int nnz = row_ptr[n];
    (void)diag;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += gridDim.x * blockDim.x) {
        if (!have_ints) {
            val[e] = 0.0;
            continue;
        }
        int lo = 0, hi = n;   // row of entry e: the last i with row_ptr[i] <= e
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (row_ptr[mid] <= e) lo = mid; else hi = mid;
        }
        const int i = lo;
        const uint64_t t = strs[i], s = strs[col[e]];
        if (e < row_ptr[i] + n_single[i]) {
            const int pq = (int)(meta[e] & 0x7fffffffu);
            int sign;
            val[e] = single_element(s, pq / norb, pq % norb, norb, h, g, &sign);
        } else {
            val[e] = double_element(s, t, norb, g);
        }
    }
}

// Same-spin diagonal elements, one WARP per string: the n_elec^2 integral loads of a string are spread over the
// lanes, the additions are made in the order of diagonal_element() (every lane repeats them on shuffled terms),
// so the values are the sequential ones bit for bit.  (One thread per string -- 2 n_elec^2 dependent loads --
// was the long pole of the table set-up: 60 us at 15 electrons.)
__global__ void excitation_diag_kernel(const uint64_t* __restrict__ strs, int n, int norb,
                                       const double* __restrict__ h, const double* __restrict__ g,
                                       double* __restrict__ diag) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    if (h == nullptr || g == nullptr) {
        if (lane == 0) diag[i] = 0.0;
        return;
    }
    const uint64_t s = strs[i];
    const int ne = popc64(s);
    // lane l holds the l-th and the (l+32)-th occupied orbital (ascending)
    int o_lo = 0, o_hi = 0;
    {
        uint64_t occ = s;
        for (int k = 0; occ; ++k) {
            const int p = lowbit64(occ);
            occ &= occ - 1;
            if (k == lane) o_lo = p;
            if (k == lane + 32) o_hi = p;
        }
    }
    const int64_t n1 = norb, n2 = n1 * n1, n3 = n2 * n1;
    double e = 0.0;
    for (int ii = 0; ii < ne; ++ii) {
        const int p = __shfl_sync(0xffffffffu, ii < 32 ? o_lo : o_hi, ii & 31);
        double t_lo = 0.0, t_hi = 0.0;
        if (lane < ne) t_lo = 0.5 * (g[p * n3 + p * n2 + o_lo * n1 + o_lo] - g[p * n3 + o_lo * n2 + o_lo * n1 + p]);
        if (lane + 32 < ne) t_hi = 0.5 * (g[p * n3 + p * n2 + o_hi * n1 + o_hi] - g[p * n3 + o_hi * n2 + o_hi * n1 + p]);
        e += h[p * n1 + p];
        for (int k = 0; k < ne; ++k) e += __shfl_sync(0xffffffffu, k < 32 ? t_lo : t_hi, k & 31);
    }
    if (lane == 0) diag[i] = e;
}

// --------------------------------------------------------------------------------------------
// opposite-spin tensor and its contractions
// --------------------------------------------------------------------------------------------
__global__ void make_gab_kernel(const double* __restrict__ g, int norb, double shift, int mode,
                                double* __restrict__ gab, int ldg) {
    const int n2 = norb * norb;
    const int64_t total = (int64_t)n2 * ldg;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int pq = (int)(idx / ldg), rs = (int)(idx % ldg);
        double v = 0.0;
        if (rs < n2) {
            const int p = pq / norb, q = pq % norb, r = rs / norb, s = rs % norb;
            const double d = (p == s && q == r) ? 1.0 : 0.0;
            v = (mode == 0 ? g[(int64_t)pq * n2 + rs] : 0.0) - (mode == 0 ? shift : 1.0) * d;
        }
        gab[idx] = v;
    }
}

// Wa[a*ldg + rs] = sum_{p in a} gab[pp*ldg + rs]
__global__ void wa_kernel(const uint64_t* __restrict__ strs_a, int na, int norb,
                          const double* __restrict__ gab, int ldg, double* __restrict__ Wa) {
    const int a = blockIdx.x;
    const uint64_t s = strs_a[a];
    for (int rs = threadIdx.x; rs < ldg; rs += blockDim.x) {
        double v = 0.0;
        uint64_t occ = s;
        while (occ) {
            const int p = lowbit64(occ);
            occ &= occ - 1;
            v += gab[(int64_t)(p * norb + p) * ldg + rs];
        }
        Wa[(int64_t)a * ldg + rs] = v;
    }
}

// Wb[pq*ldc + b] = sum_{r in b} gab[pq*ldg + rr]
__global__ void wb_kernel(const uint64_t* __restrict__ strs_b, int nb, int norb,
                          const double* __restrict__ gab, int ldg, double* __restrict__ Wb, int ldc) {
    const int pq = blockIdx.x;
    const double* grow = gab + (int64_t)pq * ldg;
    for (int b = threadIdx.x; b < ldc; b += blockDim.x) {
        double v = 0.0;
        if (b < nb) {
            uint64_t occ = strs_b[b];
            while (occ) {
                const int r = lowbit64(occ);
                occ &= occ - 1;
                v += grow[r * norb + r];
            }
        }
        Wb[(int64_t)pq * ldc + b] = v;
    }
}

// diag[a*ldc+b] = da[a] + db[b] + sum_{p in a} Wb[pp*ldc+b] + c0 ; pads <- pad_value
__global__ void op_diag_kernel(const uint64_t* __restrict__ strs_a, int na, int nb, int norb,
                               const double* __restrict__ da, const double* __restrict__ db,
                               const double* __restrict__ Wb, int ldc, double c0, double pad_value,
                               double* __restrict__ diag) {
    const int a = blockIdx.x;
    const uint64_t s = strs_a[a];
    const double ea = da ? da[a] : 0.0;
    for (int b = threadIdx.x; b < ldc; b += blockDim.x) {
        double v = pad_value;
        if (b < nb) {
            v = ea + (db ? db[b] : 0.0) + c0;
            uint64_t occ = s;
            while (occ) {
                const int p = lowbit64(occ);
                occ &= occ - 1;
                v += Wb[(int64_t)(p * norb + p) * ldc + b];
            }
        }
        diag[(int64_t)a * ldc + b] = v;
    }
}

}  // namespace sqd

using namespace sqd;

extern "C" {

int sqd_pack_bitstrings(const uint8_t* d_bits, int64_t n, int nbits, uint64_t* d_left,
                        uint64_t* d_right, void* stream) {
    SQD_REQUIRE(nbits % 2 == 0 && nbits / 2 <= 64 && nbits > 0,
                "sqd_pack_bitstrings: nbits=%d must be even with nbits/2 <= 64", nbits);
    if (n == 0) return 0;
    const int wpb = 8;
    pack_bitstrings_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        d_bits, n, nbits, d_left, d_right);
    return check_launch("pack_bitstrings_kernel");
}

int sqd_excitation_count(const uint64_t* d_strs, int n, int* d_n_single, int* d_n_total,
                         void* stream) {
    SQD_REQUIRE(n > 0, "sqd_excitation_count: empty string list");
    const int wpb = 8;
    excitation_count_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(
        d_strs, n, d_n_single, d_n_total);
    return check_launch("excitation_count_kernel");
}

int sqd_exclusive_scan(const int* d_in, int* d_out, int n, int* h_total, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(n >= 0, "sqd_exclusive_scan: negative length");
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(d_in, d_out, n);
    if (check_launch("exclusive_scan_kernel")) return -2;
    if (h_total) {
        if (read_back(h_total, d_out + n, sizeof(int), st)) return -2;
    }
    return 0;
}

int sqd_excitation_fill(const uint64_t* d_strs, int n, int norb, const double* d_h,
                        const double* d_g, const int* d_row_ptr, const int* d_n_single,
                        uint32_t* d_col, double* d_val, uint32_t* d_meta, uint32_t* d_pack,
                        double* d_diag, void* stream) {
    SQD_REQUIRE(n > 0 && norb > 0 && norb <= 64, "sqd_excitation_fill: need 0 < norb <= 64 (got %d)",
                norb);
    SQD_REQUIRE(n <= (1 << 19), "sqd_excitation_fill: at most 2^19 strings per spin (got %d)", n);
    const int wpb = 8;
    cudaStream_t st = (cudaStream_t)stream;
    excitation_fill_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, st>>>(d_strs, n, norb, d_row_ptr, d_n_single, d_col,
                                                                   d_meta, d_pack);
    // ~75-150 entries per string at the BASELINE shapes: a grid that covers them in one or two strides
    const int64_t guess = (int64_t)n * 128;
    const int blocks = (int)min((int64_t)kNumSMs * 8, (guess + 255) / 256);
    excitation_values_kernel<<<blocks, 256, 0, st>>>(d_strs, n, norb, d_h, d_g, d_row_ptr, d_n_single, d_col, d_meta,
                                                    d_val, d_diag);
    excitation_diag_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, st>>>(d_strs, n, norb, d_h, d_g, d_diag);
    return check_launch("excitation fill/values/diag kernels", 3);
}

int sqd_make_gab(const double* d_g, int norb, double shift, int mode, double* d_gab, int ldg,
                 void* stream) {
    SQD_REQUIRE(norb > 0 && norb <= 64 && ldg >= norb * norb && ldg % 2 == 0,
                "sqd_make_gab: bad norb/ldg (%d, %d)", norb, ldg);
    const int64_t total = (int64_t)norb * norb * ldg;
    const int blocks = (int)min((int64_t)kNumSMs * 8, (total + 255) / 256);
    make_gab_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_g, norb, shift, mode, d_gab, ldg);
    return check_launch("make_gab_kernel");
}

int sqd_opposite_spin_tables(const uint64_t* d_strs_a, int na, const uint64_t* d_strs_b, int nb,
                             int norb, const double* d_gab, int ldg, const double* d_da,
                             const double* d_db, double diag_const, double pad_value, double* d_Wa,
                             double* d_Wb, double* d_diag, int ldc, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(na > 0 && nb > 0 && ldc >= nb && ldc % 2 == 0, "sqd_opposite_spin_tables: bad shape");
    if (d_Wa) {
        wa_kernel<<<na, 256, 0, st>>>(d_strs_a, na, norb, d_gab, ldg, d_Wa);
        if (check_launch("wa_kernel")) return -2;
    }
    SQD_REQUIRE(d_Wb != nullptr, "sqd_opposite_spin_tables: Wb scratch is required for the diagonal");
    wb_kernel<<<norb * norb, 256, 0, st>>>(d_strs_b, nb, norb, d_gab, ldg, d_Wb, ldc);
    if (check_launch("wb_kernel")) return -2;
    op_diag_kernel<<<na, 256, 0, st>>>(d_strs_a, na, nb, norb, d_da, d_db, d_Wb, ldc, diag_const,
                                       pad_value, d_diag);
    return check_launch("op_diag_kernel");
}

}  // extern "C"

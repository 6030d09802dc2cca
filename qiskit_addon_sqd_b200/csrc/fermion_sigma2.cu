// CI sigma-vector build, second generation (v2).  See the comment of sqd_sigma_v2 in include/sqd_b200.h.
//
// Replaces pyscf selected_ci.contract_2e (SCIcontract_2e_bbaa + SCIcontract_2e_aaaa, recalled) reached from
// qiskit_addon_sqd/fermion.py:721-723, 810-818 -- same mathematics as fermion_sigma.cu, different mapping:
//
//   sigma[a,b] = diag[a,b] c[a,b]
//              + sum_{a'} Ha[a,a'] c[a',b] + sum_{b'} Hb[b,b'] c[a,b']              dense FP64 tiles   (K2)
//              + sum_{a' in S_a(a)} sgn_a Wb[pq,b] c[a',b]                           epilogue           (K3)
//              + sum over the P rows of a (self item + one per a' in S_a(a)), segments of b   epilogue (K3)
//   P[self(a')][b]     =         sum_{j in S_b(b)} sgn_j Wa[a', rs_j]   c[a', b'_j]                      (K1)
//   P[row(a <- a')][b] = sgn_a   sum_{j in S_b(b)} sgn_j g_ab[pq, rs_j] c[a', b'_j]                      (K1)
//
// Why: in the v1 kernel every FMA of the opposite-spin part costs three shared-memory reads (link word,
// c[a',b'_j], g_ab[pq,rs_j]) and all warps of a CTA march in step.  Grouping the work by SOURCE string a'
// makes x_j = sgn_j c[a',b'_j] a constant of the thread for all excitations out of a', so it lives in
// registers together with the byte offsets of rs_j: one gather and one FMA per link, and a warp needs
// nothing from the other warps of its CTA.  The price is that the result rows travel through memory (P, L2
// resident) because several source strings feed one row of sigma.  The same-spin part of a HF-centred
// sample set is 15-25 % dense, where dense FP64 tiles beat gathers by a wide margin.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

constexpr int kV2MaxGroups = 64;
constexpr int kV2MaxStages = 8;
// virtual columns per K1 group aimed at / hard limit (K1 runs vc_pad + 32 threads).  With 8 links per column
// the kernel fits 80 registers and two CTAs of <= 416 threads share an SM (the K1 kernels of two concurrent
// solves then overlap instead of taking turns on the whole GPU); 16 links per column: one CTA of <= 512.
__host__ __device__ constexpr int v2_group_target(int lmax) { return lmax == 8 ? 352 : 416; }
__host__ __device__ constexpr int v2_group_max(int lmax) { return lmax == 8 ? 384 : 480; }
constexpr int kV2GroupMax = 480;
constexpr uint32_t kSelfItem = 0x40000000u;
enum { C_NITEMS = 0, C_NCHUNKS, C_NGROUPS, C_VCPAD, C_SINGLES_A, C_SINGLES_B, C_ERR, C_NVC, C_NQ, C_NHEAVY };
constexpr int kV2HeavyRow = 64;  // alpha strings with more single excitations get a CTA per column block in the epilogue

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// device-memory layout of the plan (byte offsets); every array is sized by upper bounds that are known
// on the host before the plan kernels run
struct V2Layout {
    size_t single_ptr, item_ptr, chunk_rec, item_tgt, item_gsel, item_pslot;
    size_t col_grp, col_full, col_nfull, col_rem, col_seg;
    size_t vc_src, vc_off, vc_len, vc_q, heavy_rows, counts, total;
    int maxch, capw;
};

static V2Layout v2_layout(int na, int nb, int64_t nnz_a, int64_t nnz_b, int lmax, int ipc) {
    V2Layout L{};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += al256(bytes);
        return at;
    };
    const size_t items = (size_t)na + (size_t)(nnz_a > 0 ? nnz_a : 1);
    L.maxch = (int)(na + (na + nnz_a) / ipc + 1);
    // virtual columns: sum_b ceil(len_b / lmax) <= nb + singles_b / lmax; groups add padding
    L.capw = (int)(2 * ((int64_t)nb + nnz_b / lmax) + kV2MaxGroups * 64 + 64);
    L.single_ptr = take((size_t)(na + 1) * 4);
    L.item_ptr = take((size_t)(na + 1) * 4);
    L.chunk_rec = take((size_t)L.maxch * 16);
    L.item_tgt = take(items * 4);
    L.item_gsel = take(items * 4);
    L.item_pslot = take(items * 4);
    L.col_grp = take((size_t)nb * 4);
    L.col_full = take((size_t)nb * 4);
    L.col_nfull = take((size_t)nb * 4);
    L.col_rem = take((size_t)nb * 4);
    L.col_seg = take((size_t)(nb + 1) * 4);
    L.vc_src = take((size_t)L.capw * lmax * 4);
    L.vc_off = take((size_t)L.capw * lmax * 4);
    L.vc_len = take((size_t)L.capw * 4);
    L.vc_q = take((size_t)L.capw * 4);
    L.heavy_rows = take((size_t)na * 4);
    L.counts = take((size_t)SQD_V2_COUNTS * 4);
    L.total = o;
    return L;
}

// ---------------------------------------------------------------------------------------------------
// plan kernels (set-up time)
// ---------------------------------------------------------------------------------------------------
// exclusive prefix of v over the CTA (blockDim.x <= 1024, a multiple of 32); *total = sum over the CTA
__device__ __forceinline__ int block_excl_scan(int v, int* warp_tot, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    __syncthreads();  // warp_tot may still be read from the previous call
    if (lane == 31) warp_tot[warp] = s;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarp ? warp_tot[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warp_tot[lane] = w;
    }
    __syncthreads();
    *total = warp_tot[nwarp - 1];
    return (warp == 0 ? 0 : warp_tot[warp - 1]) + s - v;
}

// alpha side: item numbering and the chunk list, sorted by descending size (the CTAs of K1 take chunks
// round robin, so every CTA gets one chunk of each size rank)
__global__ void __launch_bounds__(1024)
v2_alpha_plan_kernel(const sqd_spin_table A, int ipc, int maxch, int* __restrict__ single_ptr,
                     int* __restrict__ item_ptr, int4* __restrict__ chunk_rec, int* __restrict__ counts) {
    __shared__ int warp_tot[32];
    const int na = A.n;
    int carry = 0, tot = 0;
    for (int base = 0; base < na; base += blockDim.x) {
        const int a = base + threadIdx.x;
        const int v = a < na ? A.n_single[a] : 0;
        const int ex = block_excl_scan(v, warp_tot, &tot);
        if (a < na) {
            single_ptr[a] = carry + ex;
            item_ptr[a] = carry + ex + a;
        }
        carry += tot;
    }
    if (threadIdx.x == 0) {
        single_ptr[na] = carry;
        item_ptr[na] = carry + na;
        counts[C_SINGLES_A] = carry;
        counts[C_NITEMS] = carry + na;
    }
    // chunks by size class, ipc first
    int pos = 0;
    for (int cls = ipc; cls >= 1; --cls) {
        for (int base = 0; base < na; base += blockDim.x) {
            const int a = base + threadIdx.x;
            int cnt = 0, nch = 0, last = 0, ip = 0;
            if (a < na) {
                const int T = 1 + A.n_single[a];
                ip = item_ptr[a];  // written by this very thread above
                nch = (T + ipc - 1) / ipc;
                last = T - (nch - 1) * ipc;
                cnt = (cls == ipc ? nch - 1 : 0) + (last == cls ? 1 : 0);
            }
            const int ex = block_excl_scan(cnt, warp_tot, &tot);
            if (a < na && cnt > 0) {
                int at = pos + ex;
                if (cls == ipc) {
                    for (int j = 0; j < nch - 1; ++j, ++at)
                        if (at < maxch) chunk_rec[at] = make_int4(a, ip + j * ipc, ipc, 0);
                }
                if (last == cls && at < maxch) chunk_rec[at] = make_int4(a, ip + (nch - 1) * ipc, last, 0);
            }
            pos += tot;
        }
    }
    if (threadIdx.x == 0) {
        counts[C_NCHUNKS] = pos;
        if (pos > maxch) counts[C_ERR] = 2;
    }
}

// Per-item arrays in SOURCE order (what K1's producer reads).  Entry k of row a points at partner a'; seen
// from the source a that is the excitation a -> a' (row `a` stores E_pq a' = a, so a -> a' is E_qp), and its
// result belongs to the P row of TARGET a' that corresponds to a'-s own entry for partner a:
// item_ptr[a'] + 1 + k'.  One warp per row, lanes over its single excitations.
__global__ void v2_item_kernel(const sqd_spin_table A, int norb, const int* __restrict__ item_ptr,
                               int* __restrict__ item_tgt, uint32_t* __restrict__ item_gsel,
                               int* __restrict__ item_pslot, int* __restrict__ heavy_rows,
                               int* __restrict__ counts) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= A.n) return;
    const int beg = A.row_ptr[a], ns = A.n_single[a], ip = item_ptr[a];
    if (lane == 0) {
        item_tgt[ip] = a;
        item_gsel[ip] = kSelfItem;
        item_pslot[ip] = ip;
        // list of strings with many excitations (any order: every string is summed on its own)
        if (ns > kV2HeavyRow) heavy_rows[atomicAdd(&counts[C_NHEAVY], 1)] = a;
    }
    for (int k = lane; k < ns; k += 32) {
        const int ap = (int)A.col[beg + k];
        const uint32_t m = A.meta[beg + k];
        const int pq = (int)(m & 0x7fffffffu), p = pq / norb, q = pq - p * norb;
        const int pb = A.row_ptr[ap];
        int lo = 0, hi = A.n_single[ap] - 1, found = 0;
        while (lo <= hi) {  // singles of a row are sorted by partner index
            const int mid = (lo + hi) >> 1;
            const int v = (int)A.col[pb + mid];
            if (v == a) {
                found = mid;
                break;
            }
            if (v < a) lo = mid + 1; else hi = mid - 1;
        }
        item_tgt[ip + 1 + k] = ap;
        item_gsel[ip + 1 + k] = (uint32_t)(q * norb + p) | (m & 0x80000000u);
        item_pslot[ip + 1 + k] = item_ptr[ap] + 1 + found;
    }
}

// beta side: cut every string's single-excitation list into virtual columns of at most lmax links, deal
// the strings to groups (round robin over the length-sorted order), and number the virtual columns of a
// group: full ones first (strings in rank order), then the remainders by descending length, so that the
// 32 threads of a warp run (almost) the same trip count.  The P column of a virtual column is
// col_seg[b] + its segment number: segments of a string are adjacent and strings keep their natural order,
// which is what lets the epilogue read P with (nearly) coalesced loads.
// Single CTA; O(nb^2 / 1024) compares per thread.
__global__ void __launch_bounds__(1024)
v2_beta_plan_kernel(const sqd_spin_table B, int lmax, int capw, int* __restrict__ col_grp,
                    int* __restrict__ col_full, int* __restrict__ col_nfull, int* __restrict__ col_rem,
                    int* __restrict__ col_seg, int* __restrict__ counts) {
    extern __shared__ int bp_smem[];
    __shared__ int warp_tot[32];
    __shared__ int g_nfull[kV2MaxGroups], g_nrem[kV2MaxGroups];
    __shared__ int s_G, s_ok;
    const int nb = B.n;
    int* L = bp_smem;          // [nb] list length
    int* rk = L + nb;          // [nb] rank of the string (length descending, index ascending)
    int* by_rank = rk + nb;    // [nb]
    int nvc = 0, nz = 0, sb = 0, carry = 0, tot = 0;
    for (int base = 0; base < nb; base += blockDim.x) {
        const int b = base + threadIdx.x;
        const int l = b < nb ? B.n_single[b] : 0;
        const int nseg = (l + lmax - 1) / lmax;
        if (b < nb) L[b] = l;
        nvc += nseg;
        nz += l > 0;
        sb += l;
        const int ex = block_excl_scan(nseg, warp_tot, &tot);
        if (b < nb) col_seg[b] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) col_seg[nb] = carry;
    __syncthreads();
    // the epilogue keeps the P columns of 32 consecutive beta strings in registers: at most 128 of them
    int wide = 0;
    for (int b0 = 32 * threadIdx.x; b0 < nb; b0 += 32 * blockDim.x)
        wide |= col_seg[min(b0 + 32, nb)] - col_seg[b0] > 128;
    block_excl_scan(wide, warp_tot, &tot);
    const int too_wide = tot;
    block_excl_scan(nvc, warp_tot, &tot);
    nvc = tot;
    block_excl_scan(nz, warp_tot, &tot);
    nz = tot;
    block_excl_scan(sb, warp_tot, &tot);
    sb = tot;
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int lb = L[b];
        int r = 0;
        for (int o = 0; o < nb; ++o) {
            const int lo = L[o];
            r += (lo > lb) || (lo == lb && o < b);
        }
        rk[b] = r;
        by_rank[r] = b;
    }
    __syncthreads();
    // number of groups: smallest G >= nvc / target whose largest group fits the CTA
    if (threadIdx.x == 0) {
        s_G = nvc > 0 ? (nvc + v2_group_target(lmax) - 1) / v2_group_target(lmax) : 1;
        // a group also owns nb / G natural columns for the Wb terms, one per thread
        if (s_G < (nb + v2_group_max(lmax) - 1) / v2_group_max(lmax)) s_G = (nb + v2_group_max(lmax) - 1) / v2_group_max(lmax);
        if (s_G > kV2MaxGroups) s_G = kV2MaxGroups;
        s_ok = 0;
    }
    __syncthreads();
    for (int attempt = 0; attempt < kV2MaxGroups; ++attempt) {
        const int G = s_G;
        if (threadIdx.x < G) {
            const int g = threadIdx.x;
            int nfull = 0, nrem = 0;
            for (int r = g; r < nz; r += G) {
                const int l = L[by_rank[r]];
                nfull += l / lmax;
                nrem += (l % lmax) != 0;
            }
            g_nfull[g] = nfull;
            g_nrem[g] = nrem;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int mx = 0;
            for (int g = 0; g < G; ++g) mx = max(mx, g_nfull[g] + g_nrem[g]);
            if (mx <= v2_group_max(lmax)) s_ok = 1;
            else if (G < kV2MaxGroups) s_G = G + 1;
            else s_ok = -1;
        }
        __syncthreads();
        if (s_ok != 0) break;
    }
    const int G = s_G;
    // full virtual columns: running count over the strings of the group in rank order
    if (threadIdx.x < G) {
        const int g = threadIdx.x;
        int run = 0;
        for (int r = g; r < nz; r += G) {
            const int b = by_rank[r];
            col_full[b] = run;
            run += L[b] / lmax;
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int l = L[b], r = rk[b];
        if (l == 0) {
            col_grp[b] = -1;
            col_full[b] = 0;
            col_nfull[b] = 0;
            col_rem[b] = -1;
            continue;
        }
        const int g = r % G, rem = l % lmax;
        // remainder rank inside the group: longer remainders first, ties in rank order
        int rr = 0;
        for (int r2 = g; r2 < nz; r2 += G) {
            const int rem2 = L[by_rank[r2]] % lmax;
            rr += (rem2 > rem) || (rem2 == rem && r2 < r);
        }
        col_grp[b] = g;
        col_nfull[b] = l / lmax;
        col_rem[b] = rem ? g_nfull[g] + rr : -1;
    }
    if (threadIdx.x == 0) {
        int mx = 0;
        for (int g = 0; g < G; ++g) mx = max(mx, g_nfull[g] + g_nrem[g]);
        const int wcols = (nb + G - 1) / G;
        const int vc_pad = max(32, (max(mx, wcols) + 31) / 32 * 32);
        counts[C_NGROUPS] = G;
        counts[C_VCPAD] = vc_pad;
        counts[C_SINGLES_B] = sb;
        counts[C_NVC] = nvc;
        counts[C_NQ] = nvc;  // one P column per virtual column
        if (s_ok != 1 || (long long)G * vc_pad > capw || too_wide != 0 || vc_pad > v2_group_max(lmax)) counts[C_ERR] = 1;
    }
}

__global__ void v2_vc_init_kernel(int lmax, int capw, uint32_t zero_off, uint32_t* __restrict__ vc_src,
                                  uint32_t* __restrict__ vc_off, int* __restrict__ vc_len,
                                  int* __restrict__ vc_q) {
    const int64_t n = (int64_t)capw * lmax;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        vc_src[i] = 0u;
        vc_off[i] = zero_off;
        if (i < capw) {
            vc_len[i] = 0;
            vc_q[i] = -1;
        }
    }
}

// one warp per beta string: scatter its links into the virtual-column arrays of its group
__global__ void v2_vc_fill_kernel(const sqd_spin_table B, const int* __restrict__ counts, int lmax,
                                  const int* __restrict__ col_grp, const int* __restrict__ col_full,
                                  const int* __restrict__ col_nfull, const int* __restrict__ col_rem,
                                  const int* __restrict__ col_seg, uint32_t* __restrict__ vc_src,
                                  uint32_t* __restrict__ vc_off, int* __restrict__ vc_len,
                                  int* __restrict__ vc_q) {
    if (counts[C_ERR] != 0) return;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B.n) return;
    const int g = col_grp[b];
    if (g < 0) return;
    const int vc_pad = counts[C_VCPAD];
    const int beg = B.row_ptr[b], l = B.n_single[b];
    const int full0 = col_full[b], nfull = col_nfull[b], rem = col_rem[b], q0 = col_seg[b];
    for (int k = lane; k < l; k += 32) {
        const uint32_t pk = B.pack[beg + k];
        const int seg = k / lmax, j = k - seg * lmax;
        const int vc = seg < nfull ? full0 + seg : rem;
        const size_t at = ((size_t)g * lmax + j) * vc_pad + vc;
        vc_src[at] = (pk & 0x7ffffu) | (pk & 0x80000000u);
        vc_off[at] = ((pk >> 19) & 0xfffu) * 8u;
    }
    if (lane == 0) {
        for (int s = 0; s < nfull; ++s) {
            vc_len[g * vc_pad + full0 + s] = lmax;
            vc_q[g * vc_pad + full0 + s] = q0 + s;
        }
        if (rem >= 0) {
            vc_len[g * vc_pad + rem] = l - nfull * lmax;
            vc_q[g * vc_pad + rem] = q0 + nfull;
        }
    }
}

// Bank-aware order of the links inside a virtual column.  In K1 the 16 lanes of a half warp gather one
// 8-byte word each from the staged integral row at step j; words whose index is equal mod 16 share a bank
// pair and serialise (ncu: half of K1's shared-memory wavefronts were conflicts with the links in table
// order).  The order of a thread's links is free, so lane l aims at bank (l + 3 j) mod 16 at step j -- the 16
// lanes of a half warp then aim at 16 different banks -- and takes, among its unused links, the one whose
// bank is closest to the target.  One thread per virtual column, no communication; O(len^2), set-up time.
__global__ void v2_vc_order_kernel(const int* __restrict__ counts, int lmax, uint32_t* __restrict__ vc_src,
                                   uint32_t* __restrict__ vc_off, const int* __restrict__ vc_len) {
    if (counts[C_ERR] != 0) return;
    const int vc_pad = counts[C_VCPAD], G = counts[C_NGROUPS];
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= G * vc_pad) return;
    const int g = idx / vc_pad, t = idx - g * vc_pad;
    const int len = vc_len[idx];
    if (len <= 1) return;
    uint32_t off[16], src[16];
    const size_t base = (size_t)g * lmax * vc_pad + t;
    for (int j = 0; j < len; ++j) {
        off[j] = vc_off[base + (size_t)j * vc_pad];
        src[j] = vc_src[base + (size_t)j * vc_pad];
    }
    unsigned used = 0u;
    const int l16 = t & 15;
    for (int j = 0; j < len; ++j) {
        const int target = (l16 + 3 * j) & 15;
        int best = -1, best_d = 99;
        for (int k = 0; k < len; ++k) {
            if ((used >> k) & 1u) continue;
            const int bank = (int)(off[k] >> 3) & 15;
            int d = (bank - target) & 15;
            if (d > 8) d = 16 - d;
            if (d < best_d) {
                best_d = d;
                best = k;
            }
        }
        used |= 1u << best;
        vc_off[base + (size_t)j * vc_pad] = off[best];
        vc_src[base + (size_t)j * vc_pad] = src[best];
    }
}

// HDT[col*ld + row] = val(row, col): transposed so that the tile loader of K2 reads, for a fixed source
// string k, the elements <a|H|k> of consecutive target strings a with coalesced 16-byte loads -- and uses
// exactly the table values of row a (the v1 kernels and the oracle use those; <a|H|k> and <k|H|a> may
// differ in the last bit).  One warp per row.
__global__ void v2_dense_scatter_kernel(const sqd_spin_table T, double* __restrict__ HDT, int ld) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= T.n) return;
    for (int e = T.row_ptr[a] + lane; e < T.row_ptr[a + 1]; e += 32)
        HDT[(size_t)T.col[e] * ld + a] = T.val[e];
}

// ---------------------------------------------------------------------------------------------------
// K1: opposite-spin part by source string
// ---------------------------------------------------------------------------------------------------
struct V2Args {
    sqd_operator op;
    const double* c;
    double* sigma;
    const int* done;
    int row_begin, row_end;
    int nsplit;
    // device-driven Davidson loop (CUDA graph replay): c and sigma are BASE pointers, the vector slot is
    // read from device memory
    const int* slot_ptr;
    long long vec_stride;
};
__device__ __forceinline__ size_t v2_slot_offset(const V2Args& P) {
    return P.slot_ptr != nullptr ? (size_t)(*P.slot_ptr) * (size_t)P.vec_stride : 0;
}

__device__ __forceinline__ bool v2_try_wait(uint64_t* bar, uint32_t parity) {
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
__device__ __forceinline__ void v2_wait(uint64_t* bar, uint32_t parity) {
    while (!v2_try_wait(bar, parity)) __nanosleep(32);
}
__device__ __forceinline__ void v2_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// dot product of the first N links: the gathers are issued in batches of BATCH before their FMAs (16 in
// flight cost 16 more registers than 8 and push the kernel over the budget that lets the dense tile kernel
// share the SM: the LEAN instance uses 8), four interleaved partial sums keep the FP64 FMA latency off the
// critical path
template <int N, int LMAX, int BATCH>
__device__ __forceinline__ double v2_dot(const double (&x)[LMAX], const uint32_t (&off2)[LMAX / 2],
                                         const char* G) {
    // off2[j/2] holds the byte offsets of links j (low half) and j+1 (high half): 8 registers instead of 16
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
    for (int j0 = 0; j0 < N; j0 += BATCH) {
        double gv[BATCH];
#pragma unroll
        for (int j = 0; j < BATCH; ++j) {
            if (j0 + j < N) {
                const uint32_t w = off2[(j0 + j) >> 1];
                const uint32_t o = ((j0 + j) & 1) ? (w >> 16) : (w & 0xffffu);
                gv[j] = *reinterpret_cast<const double*>(G + o);
            }
        }
#pragma unroll
        for (int j = 0; j < BATCH; j += 4) {
            if (j0 + j < N) {
                a0 = fma(x[j0 + j], gv[j], a0);
                a1 = fma(x[j0 + j + 1], gv[j + 1], a1);
                a2 = fma(x[j0 + j + 2], gv[j + 2], a2);
                a3 = fma(x[j0 + j + 3], gv[j + 3], a3);
            }
        }
    }
    return (a0 + a1) + (a2 + a3);
}

// stage header: x = type (0 exit, 1 new source string, 2 item, 3 self item), y = P row, z = sign,
// w = source string.  A type-1 stage carries the row c[a',:] (bulk copy) when `src_smem` is set, so that
// the gather of x_j = sgn_j c[a', b'_j] hits shared memory and its L2 latency hides behind the ring like
// the integral rows'.  A type-2 stage carries g_ab[pq,:] and, behind it, Wb[pq,:]: thread t also owns the
// natural column g*wcols + t of the group and writes sgn*Wb[pq,b]*c[a',b] into the w-part of the P row.
// LEAN: register cap 96 (launch bound 640) so that one CTA of this kernel and a CTA of the dense tile kernel
// fit on an SM together (overlapped build of a lone solve); otherwise 128 registers.
template <int LMAX, bool LEAN>
__global__ void __launch_bounds__(LMAX == 8 ? 416 : (LEAN ? 640 : 512), LMAX == 8 ? 2 : 1)
sigma2_ab_kernel(const V2Args P, const int NST, const int stage_len, const int src_smem) {
    constexpr int BATCH = 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (P.done != nullptr && *P.done != 0) return;
    const double* Pc = P.c + v2_slot_offset(P);
    const sqd_operator& op = P.op;
    const sqd_sigma_v2& V = op.v2;
    const int g = blockIdx.y;
    const int vc_pad = V.vc_pad, ldg = op.ldg, ldc = op.ldc, nb = op.b.n;
    const int tid = threadIdx.x, lane = tid & 31;
    const int ncons = blockDim.x - 32, nwarp_c = ncons >> 5;
    const bool have_wb = op.Wb != nullptr;

    double* stage = reinterpret_cast<double*>(smem_raw);                            // [NST][stage_len]
    uint64_t* full = reinterpret_cast<uint64_t*>(stage + (size_t)NST * stage_len);  // [kV2MaxStages]
    uint64_t* empty = full + kV2MaxStages;                                           // [kV2MaxStages]
    int4* hdr = reinterpret_cast<int4*>(empty + kV2MaxStages);                       // [kV2MaxStages]

    if (tid == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], nwarp_c);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // =========================== producer warp ===========================
    // Chunks are dealt round robin over the CTAs of the group (the list is sorted by descending size, so
    // the loads are balanced to within one chunk); everything a chunk needs from global memory is
    // fetched one chunk ahead, so the only waits of this warp are the ring's empty barriers.
    if (tid >= ncons) {
        int s = 0, round = 0;
        const bool self_ok = op.Wa != nullptr;
        const int stride = gridDim.x;
        const int4* recs = reinterpret_cast<const int4*>(V.chunk_rec);
        int chunk = blockIdx.x;
        int4 rec = make_int4(0, 0, 0, 0), rec_next = make_int4(0, 0, 0, 0);
        int tgt = 0, pslot = 0;
        uint32_t gs = 0;
        if (chunk < V.n_chunks) {
            rec = recs[chunk];
            if (lane < rec.z) {
                tgt = V.item_tgt[rec.y + lane];
                gs = V.item_gsel[rec.y + lane];
                pslot = V.item_pslot[rec.y + lane];
            }
        }
        if (chunk + stride < V.n_chunks) rec_next = recs[chunk + stride];
        for (; chunk < V.n_chunks; chunk += stride) {
            const int ap = rec.x, n = rec.z;
            const bool is_self = (gs & kSelfItem) != 0;
            const bool live = lane < n && tgt >= P.row_begin && tgt < P.row_end && (!is_self || self_ok);
            const int my_slot = pslot;
            const int my_sel = is_self ? -1 : (int)(gs & 0xffffu);
            const int my_sign = (gs >> 31) ? -1 : 1;
            // prefetch: lane data of the next chunk, record of the one after
            const int4 rec_n = rec_next;
            int tgt_n = 0, pslot_n = 0;
            uint32_t gs_n = 0;
            if (chunk + stride < V.n_chunks && lane < rec_n.z) {
                tgt_n = V.item_tgt[rec_n.y + lane];
                gs_n = V.item_gsel[rec_n.y + lane];
                pslot_n = V.item_pslot[rec_n.y + lane];
            }
            if (chunk + 2 * stride < V.n_chunks) rec_next = recs[chunk + 2 * stride];

            uint32_t rem = __ballot_sync(0xffffffffu, live);
            if (rem != 0) {
                if (lane == 0) {
                    if (round > 0) v2_wait(&empty[s], (uint32_t)((round - 1) & 1));
                    hdr[s] = make_int4(1, 0, 0, ap);
                    if (src_smem) {
                        mbar_expect_tx(&full[s], (uint32_t)(ldc * sizeof(double)));
                        bulk_g2s(stage + (size_t)s * stage_len, Pc + (size_t)ap * ldc,
                                 (uint32_t)(ldc * sizeof(double)), &full[s]);
                    } else {
                        v2_arrive(&full[s]);
                    }
                }
                if (++s == NST) {
                    s = 0;
                    ++round;
                }
                while (rem) {
                    const int k = __ffs(rem) - 1;
                    rem &= rem - 1;
                    const int slot_k = __shfl_sync(0xffffffffu, my_slot, k);
                    const int sign_k = __shfl_sync(0xffffffffu, my_sign, k);
                    const int sel_k = __shfl_sync(0xffffffffu, my_sel, k);
                    if (lane == 0) {
                        if (round > 0) v2_wait(&empty[s], (uint32_t)((round - 1) & 1));
                        double* dst = stage + (size_t)s * stage_len;
                        if (sel_k < 0) {
                            hdr[s] = make_int4(3, slot_k, 1, ap);
                            mbar_expect_tx(&full[s], (uint32_t)(ldg * sizeof(double)));
                            bulk_g2s(dst, op.Wa + (size_t)ap * ldg, (uint32_t)(ldg * sizeof(double)), &full[s]);
                        } else {
                            hdr[s] = make_int4(2, slot_k, sign_k, ap);
                            mbar_expect_tx(&full[s], (uint32_t)((ldg + (have_wb ? ldc : 0)) * sizeof(double)));
                            bulk_g2s(dst, op.gab + (size_t)sel_k * ldg, (uint32_t)(ldg * sizeof(double)),
                                     &full[s]);
                            if (have_wb)
                                bulk_g2s(dst + ldg, op.Wb + (size_t)sel_k * ldc,
                                         (uint32_t)(ldc * sizeof(double)), &full[s]);
                        }
                    }
                    if (++s == NST) {
                        s = 0;
                        ++round;
                    }
                }
            }
            rec = rec_n;
            tgt = tgt_n;
            gs = gs_n;
            pslot = pslot_n;
        }
        if (lane == 0) {
            if (round > 0) v2_wait(&empty[s], (uint32_t)((round - 1) & 1));
            hdr[s] = make_int4(0, 0, 0, 0);
            v2_arrive(&full[s]);
        }
        return;
    }

    // =========================== consumer warps ==========================
    // A warp depends on the ring only: it may run up to NST stages ahead of the slowest warp of its CTA.
    const int t = tid;
    const size_t gbase = (size_t)g * LMAX * vc_pad;
    const int len = V.vc_len[g * vc_pad + t];
    const int myq = V.vc_q[g * vc_pad + t];
    // natural column of this thread for the Wb terms
    const int wcols = (nb + (int)gridDim.y - 1) / (int)gridDim.y;
    const int wb_col = (t < wcols && g * wcols + t < nb) ? g * wcols + t : -1;
    int wlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wlen = max(wlen, __shfl_xor_sync(0xffffffffu, wlen, o));
    const bool warp_has_w = __any_sync(0xffffffffu, wb_col >= 0) && have_wb;
    if (wlen == 0 && !warp_has_w) {
        // a warp of padding lanes still has to release the stages
        int s = 0, round = 0;
        for (;;) {
            v2_wait(&full[s], (uint32_t)(round & 1));
            const int type = hdr[s].x;
            if (type == 0) break;
            __syncwarp();
            if (lane == 0) v2_arrive(&empty[s]);
            if (++s == NST) {
                s = 0;
                ++round;
            }
        }
        return;
    }
    uint32_t off[LMAX / 2];  // two 16-bit byte offsets per register (8 * norb^2 <= 32768)
    double x[LMAX];
#pragma unroll
    for (int j = 0; j < LMAX; j += 2) {
        off[j >> 1] = V.vc_off[gbase + (size_t)j * vc_pad + t] | (V.vc_off[gbase + (size_t)(j + 1) * vc_pad + t] << 16);
        x[j] = 0.0;
        x[j + 1] = 0.0;
    }
    const size_t ldp = (size_t)V.ldp;
    double* prow = V.P + (myq >= 0 ? myq : 0);
    double* wrow = V.P + V.ldq + (wb_col >= 0 ? wb_col : 0);
    double cmine = 0.0;
    int s = 0, round = 0;
    for (;;) {
        v2_wait(&full[s], (uint32_t)(round & 1));
        const int4 h = hdr[s];
        if (h.x == 0) break;
        if (h.x == 1) {
            // the (source column, sign) words are re-read per source string: 4 bytes per link from L1/L2,
            // against 16 more registers per thread if they were kept.  Padding entries read column 0.
            const double* crow = src_smem ? stage + (size_t)s * stage_len : Pc + (size_t)h.w * ldc;
            if (wlen > 0) {
#pragma unroll
                for (int j = 0; j < LMAX; ++j) {
                    const uint32_t sv = __ldg(V.vc_src + gbase + (size_t)j * vc_pad + t);
                    const double v = crow[sv & 0x7fffffffu];
                    x[j] = j < len ? ((sv >> 31) ? -v : v) : 0.0;
                }
            }
            cmine = wb_col >= 0 ? crow[wb_col] : 0.0;
        } else {
            const char* G = reinterpret_cast<const char*>(stage + (size_t)s * stage_len);
            // all gathers of the item are issued before the first FMA (the trip count is a warp-uniform
            // multiple of 4), four interleaved partial sums keep the FP64 FMA latency off the critical path
            double wv = 0.0;
            if (h.x == 2 && warp_has_w && wb_col >= 0) wv = stage[(size_t)s * stage_len + ldg + wb_col];
            double acc = 0.0;
            if (wlen > 12 && LMAX >= 16) acc = v2_dot<(LMAX >= 16 ? 16 : LMAX), LMAX, BATCH>(x, off, G);
            else if (wlen > 8 && LMAX >= 12) acc = v2_dot<(LMAX >= 12 ? 12 : LMAX), LMAX, BATCH>(x, off, G);
            else if (wlen > 4) acc = v2_dot<8, LMAX, 8>(x, off, G);
            else if (wlen > 0) acc = v2_dot<4, LMAX, 4>(x, off, G);
            if (myq >= 0) prow[(size_t)h.y * ldp] = h.z < 0 ? -acc : acc;
            if (h.x == 2 && warp_has_w && wb_col >= 0) {
                const double w = wv * cmine;
                wrow[(size_t)h.y * ldp] = h.z < 0 ? -w : w;
            }
        }
        __syncwarp();
        if (lane == 0) v2_arrive(&empty[s]);
        if (++s == NST) {
            s = 0;
            ++round;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K2: dense same-spin partial tiles
// ---------------------------------------------------------------------------------------------------
constexpr int kTM = 64, kTN = 64, kKT = 16, kK2Threads = 128, kSdA = 66, kSdB = 64;

__global__ void __launch_bounds__(kK2Threads, 3)
sigma2_tile_kernel(const V2Args P) {
    __shared__ __align__(16) double As[2][kKT][kSdA];
    __shared__ __align__(16) double Bs[2][kKT][kSdB];
    if (P.done != nullptr && *P.done != 0) return;
    const sqd_operator& op = P.op;
    const sqd_sigma_v2& V = op.v2;
    const int na = op.a.n, nb = op.b.n, ldc = op.ldc;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int a_lo = P.row_begin & ~1;
    const int a0 = a_lo + blockIdx.x * kTM, b0 = blockIdx.y * kTN;
    const int split = blockIdx.z, nsplit = gridDim.z;
    const double* __restrict__ c = P.c + v2_slot_offset(P);

    double acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0.0;

    // combined K range: k-tiles [0, nka) run over source alpha strings (HaDT, C), [nka, nka+nkb) over
    // source beta strings (C^T, HbDT).  The split boundaries depend on (na, nb, nsplit) only, so a build
    // restricted to a block of rows adds every element in the same order as the full build.
    const int nka = (na + kKT - 1) / kKT, nkb = (nb + kKT - 1) / kKT;
    const int ntk = nka + nkb;
    const int t0 = (int)((long long)split * ntk / nsplit), t1 = (int)((long long)(split + 1) * ntk / nsplit);
    double2 ra[4], rb[4];
    auto fetch = [&](int t) {
        if (t < nka) {
            const int k0 = t * kKT;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int q = tid + kK2Threads * r, k = q >> 5, ch = q & 31;
                // HaDT is allocated lda x lda with lda >= na + 128 (zero padded): no bounds to check
                ra[r] = *reinterpret_cast<const double2*>(V.HaDT + (size_t)(k0 + k) * V.lda + a0 + 2 * ch);
                const int col = b0 + 2 * ch;
                rb[r] = (k0 + k < na && col < ldc)
                            ? *reinterpret_cast<const double2*>(c + (size_t)(k0 + k) * ldc + col)
                            : make_double2(0.0, 0.0);
            }
        } else {
            const int k0 = (t - nka) * kKT;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int q = tid + kK2Threads * r;
                const int i = q >> 3, kc = q & 7;
                const int row = a0 + i, k = k0 + 2 * kc;
                ra[r] = (row < na && k < ldc)
                            ? *reinterpret_cast<const double2*>(c + (size_t)row * ldc + k)
                            : make_double2(0.0, 0.0);
                const int kk = q >> 5, ch = q & 31;
                rb[r] = *reinterpret_cast<const double2*>(V.HbDT + (size_t)(k0 + kk) * V.ldb + b0 + 2 * ch);
            }
        }
    };
    auto stash = [&](int buf, int t) {
        if (t < nka) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int q = tid + kK2Threads * r, k = q >> 5, ch = q & 31;
                *reinterpret_cast<double2*>(&As[buf][k][2 * ch]) = ra[r];
                *reinterpret_cast<double2*>(&Bs[buf][k][2 * ch]) = rb[r];
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int q = tid + kK2Threads * r;
                const int i = q >> 3, kc = q & 7;
                As[buf][2 * kc][i] = ra[r].x;
                As[buf][2 * kc + 1][i] = ra[r].y;
                const int kk = q >> 5, ch = q & 31;
                *reinterpret_cast<double2*>(&Bs[buf][kk][2 * ch]) = rb[r];
            }
        }
    };
    if (t0 < t1) {
        fetch(t0);
        stash(0, t0);
    }
    __syncthreads();
    for (int t = t0; t < t1; ++t) {
        const int cur = (t - t0) & 1;
        if (t + 1 < t1) fetch(t + 1);
#pragma unroll
        for (int kk = 0; kk < kKT; ++kk) {
            double av[8], bv[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const double2 v = *reinterpret_cast<const double2*>(&As[cur][kk][ty * 8 + 2 * r]);
                av[2 * r] = v.x;
                av[2 * r + 1] = v.y;
            }
            const double2 b01 = *reinterpret_cast<const double2*>(&Bs[cur][kk][2 * tx]);
            const double2 b23 = *reinterpret_cast<const double2*>(&Bs[cur][kk][32 + 2 * tx]);
            bv[0] = b01.x;
            bv[1] = b01.y;
            bv[2] = b23.x;
            bv[3] = b23.y;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[r][q] = fma(av[r], bv[q], acc[r][q]);
        }
        if (t + 1 < t1) stash(cur ^ 1, t + 1);
        __syncthreads();
    }

    // ---- partial tile -> memory; the epilogue kernel adds the splits in a fixed order ----
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int a = a0 + ty * 8 + r;
        if (a >= na) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b = b0 + 32 * h + 2 * tx;
            if (b < ldc)
                *reinterpret_cast<double2*>(V.part + ((size_t)split * na + a) * ldc + b) =
                    make_double2(acc[r][2 * h], acc[r][2 * h + 1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K3: epilogue.  sigma[a,b] = partial tiles (fixed order) + diag*c + column sums of the P rows of a.
// One WARP per (alpha string, 32 beta strings); a CTA of four warps walks kK3Rows consecutive strings
// (few, fat CTAs: with one CTA per string most of the kernel's cost was CTA launches).  The q-part columns
// of 32 consecutive beta strings are one contiguous range (at most kK3MaxQ wide, enforced by the planner)
// and their w-part columns are the 32 strings themselves, so every lane sums the same few P columns over the
// string's items with coalesced loads, kK3Unroll items in flight, whatever the lengths of the excitation
// lists.  No shared-memory traffic except the final q-column -> beta-string fold inside the warp.
// ---------------------------------------------------------------------------------------------------
constexpr int kK3Cols = 32, kK3Warps = 4, kK3Threads = kK3Cols * kK3Warps;
constexpr int kK3MaxQ = 128;

template <int QC, int UN>
__device__ __forceinline__ void k3_sum_items(const double* pcol, const double* wcol, size_t ldp, int n_items,
                                             const bool (&qok)[4], bool wok, double (&accq)[4], double& accw) {
    const double* pr = pcol;
    const double* wr = wcol;
    int k = 0;
    for (; k + UN <= n_items; k += UN) {
        double pv[UN][QC], wv[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
#pragma unroll
            for (int i = 0; i < QC; ++i) pv[u][i] = qok[i] ? pr[(size_t)u * ldp + 32 * i] : 0.0;
            wv[u] = wok ? wr[(size_t)u * ldp] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
#pragma unroll
            for (int i = 0; i < QC; ++i) accq[i] += pv[u][i];
            accw += wv[u];
        }
        pr += (size_t)UN * ldp;
        wr += (size_t)UN * ldp;
    }
    for (; k < n_items; ++k) {
#pragma unroll
        for (int i = 0; i < QC; ++i)
            if (qok[i]) accq[i] += pr[32 * i];
        if (wok) accw += wr[0];
        pr += ldp;
        wr += ldp;
    }
}

// Shared body: column sums of the items [k_beg, k_end) of string a for the column block starting at b0.
struct K3Ctx {
    int lane, b, nqb, q0, q1, qlo;
    bool live, wok;
    bool qok[4];
};
__device__ __forceinline__ K3Ctx k3_ctx(const V2Args& P, int b0, int lane) {
    const sqd_sigma_v2& V = P.op.v2;
    const int nb = P.op.b.n;
    K3Ctx c;
    c.lane = lane;
    c.b = b0 + lane;
    c.live = c.b < nb;
    c.qlo = V.col_seg[min(b0, nb)];
    c.nqb = V.col_seg[min(b0 + kK3Cols, nb)] - c.qlo;  // <= kK3MaxQ
    c.q0 = c.live ? V.col_seg[c.b] - c.qlo : 0;
    c.q1 = c.live ? V.col_seg[c.b + 1] - c.qlo : 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) c.qok[i] = lane + 32 * i < c.nqb;
    c.wok = c.live && P.op.Wb != nullptr;
    return c;
}
__device__ __forceinline__ double k3_base(const V2Args& P, const K3Ctx& c, int a) {
    const sqd_sigma_v2& V = P.op.v2;
    const int na = P.op.a.n, ldc = P.op.ldc;
    double base = 0.0;
    if (c.live) {
        double pt[4];
        for (int sp0 = 0; sp0 < P.nsplit; sp0 += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u)
                pt[u] = sp0 + u < P.nsplit ? V.part[((size_t)(sp0 + u) * na + a) * ldc + c.b] : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) base += pt[u];
        }
        const size_t ab = (size_t)a * ldc + c.b;
        base = fma(P.op.diag[ab], (P.c + v2_slot_offset(P))[ab], base);
    }
    return base;
}
__device__ __forceinline__ void k3_items(const V2Args& P, const K3Ctx& c, int ip, int k_beg, int k_end,
                                         bool with_self, double (&accq)[4], double& accw) {
    const sqd_sigma_v2& V = P.op.v2;
    const size_t ldp = (size_t)V.ldp;
    // P rows of the string: ip = self item (q-part only, present when the operator has Wa), then one row
    // per single excitation
    if (with_self && P.op.Wa != nullptr) {
        const double* prow = V.P + (size_t)ip * ldp + c.qlo + c.lane;
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (c.qok[i]) accq[i] += prow[32 * i];
    }
    const double* pcol = V.P + (size_t)(ip + 1 + k_beg) * ldp + c.qlo + c.lane;
    const double* wcol = V.P + (size_t)(ip + 1 + k_beg) * ldp + V.ldq + (c.live ? c.b : 0);
    const int n = k_end - k_beg;
    if (c.nqb <= 32) k3_sum_items<1, 4>(pcol, wcol, ldp, n, c.qok, c.wok, accq, accw);
    else if (c.nqb <= 64) k3_sum_items<2, 4>(pcol, wcol, ldp, n, c.qok, c.wok, accq, accw);
    else k3_sum_items<4, 4>(pcol, wcol, ldp, n, c.qok, c.wok, accq, accw);
}

// K3.  Light CTAs (strings with at most kV2HeavyRow excitations): one WARP per (alpha string, 32 beta strings),
// the four warps of a CTA take four neighbouring column blocks of the same string, no synchronisation between
// warps.  Heavy CTAs (the few strings with many excitations -- the Hartree-Fock string has hundreds): one CTA
// per (string, 32 beta strings), its four warps split the items into contiguous parts and meet in shared
// memory in warp order.  The heavy CTAs come first in the grid (blockIdx.y < heavy_ctas_y) so that the
// longest chains start first.
__global__ void __launch_bounds__(kK3Threads)
sigma2_epilogue_kernel(const V2Args P, const int heavy_ctas_y, const int ncb) {
    __shared__ double Sq[kK3Warps][kK3MaxQ];
    __shared__ double Sw[kK3Warps][kK3Cols];
    if (P.done != nullptr && *P.done != 0) return;
    const sqd_operator& op = P.op;
    const sqd_sigma_v2& V = op.v2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ldc = op.ldc;
    if ((int)blockIdx.y < heavy_ctas_y) {
        // ---- heavy role ----
        const int h = blockIdx.y * gridDim.x + blockIdx.x;  // (heavy string, column block) pair
        const int hi = h / ncb, cb = h - hi * ncb;
        if (hi >= V.n_heavy) return;
        const int a = V.heavy_rows[hi];
        if (a < P.row_begin || a >= P.row_end) return;
        const int ns = op.a.n_single[a];
        const K3Ctx c = k3_ctx(P, cb * kK3Cols, lane);
        const double base = warp == 0 ? k3_base(P, c, a) : 0.0;
        const int per = (ns + kK3Warps - 1) / kK3Warps;
        const int k_beg = min(ns, warp * per), k_end = min(ns, k_beg + per);
        double accq[4] = {0.0, 0.0, 0.0, 0.0}, accw = 0.0;
        k3_items(P, c, V.item_ptr[a], k_beg, k_end, warp == 0, accq, accw);
#pragma unroll
        for (int i = 0; i < 4; ++i) Sq[warp][lane + 32 * i] = accq[i];
        Sw[warp][lane] = accw;
        __syncthreads();
        if (warp != 0 || c.b >= ldc) return;
        double v = 0.0;
        if (c.live) {
            v = base;
#pragma unroll
            for (int w = 0; w < kK3Warps; ++w) v += Sw[w][lane];
            for (int q = c.q0; q < c.q1; ++q)
#pragma unroll
                for (int w = 0; w < kK3Warps; ++w) v += Sq[w][q];
        }
        (P.sigma + v2_slot_offset(P))[(size_t)a * ldc + c.b] = v;  // pad column: 0
        return;
    }
    // ---- light role ----
    const int b0 = (blockIdx.x * kK3Warps + warp) * kK3Cols;
    if (b0 >= ldc) return;
    const int a = P.row_begin + (int)blockIdx.y - heavy_ctas_y;
    const int ns = op.a.n_single[a];
    if (ns > kV2HeavyRow) return;  // a heavy CTA owns this string
    const K3Ctx c = k3_ctx(P, b0, lane);
    const double base = k3_base(P, c, a);
    double accq[4] = {0.0, 0.0, 0.0, 0.0}, accw = 0.0;
    k3_items(P, c, V.item_ptr[a], 0, ns, true, accq, accw);
    // fold the q columns onto the beta strings (segments of a string are adjacent)
#pragma unroll
    for (int i = 0; i < 4; ++i) Sq[warp][lane + 32 * i] = accq[i];
    __syncwarp();
    if (c.b < ldc) {
        double v = 0.0;
        if (c.live) {
            v = base + accw;
            for (int q = c.q0; q < c.q1; ++q) v += Sq[warp][q];
        }
        (P.sigma + v2_slot_offset(P))[(size_t)a * ldc + c.b] = v;  // pad column: 0
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int v2_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

// the row c[a',:] rides the ring (type-1 stages) while it is not much longer than an integral row
static int v2_src_in_smem(const sqd_operator* op) { return 1; }
// a stage holds an integral row followed by a Wb row (or, type 1, the source row of c)
static int v2_stage_len(const sqd_operator* op) { return op->ldg + op->ldc; }
static size_t v2_k1_smem(const sqd_operator* op, int nst) {
    return (size_t)nst * v2_stage_len(op) * sizeof(double) + (size_t)2 * kV2MaxStages * sizeof(uint64_t) +
           (size_t)kV2MaxStages * sizeof(int4);
}

static int v2_pick_stages(const sqd_operator* op) {
    static const int knob = v2_env("SQD_V2_STAGES", 0);
    if (knob >= 2 && knob <= kV2MaxStages) return knob;
    // deep enough to let the warps of a CTA drift apart and to cover the L2 latency of a row copy, small
    // enough for three CTAs per SM
    static const int knob_kb = v2_env("SQD_V2_RING_KB", 100);
    int nst = 8;
    while (nst > 2 && v2_k1_smem(op, nst) > (size_t)knob_kb * 1024) --nst;
    return nst;
}

int64_t sigma2_smem_bytes(const sqd_operator* op) {
    const size_t b = v2_k1_smem(op, v2_pick_stages(op));
    return b <= 220 * 1024 ? (int64_t)b : -1;
}

// K splits of the dense tiles: a function of the operator's shape only (never of the row range of a
// sharded build), see sigma2_tile_kernel
static int v2_num_split(int na, int nb, int ldc) {
    static const int knob_split = v2_env("SQD_V2_SPLIT", 0);
    const int tiles = ((na + 1 + kTM - 1) / kTM) * ((ldc + kTN - 1) / kTN);
    const int ntk = (na + kKT - 1) / kKT + (nb + kKT - 1) / kKT;
    // one CTA per SM: more splits shorten a lone build by a microsecond but cost partial-tile traffic (K2 writes,
    // K3 reads ns x n_det x 8 bytes) and SM time when several solves share the GPU (bench step 11.7 -> 11.1 ms)
    int ns = knob_split > 0 ? knob_split : kNumSMs / (tiles > 0 ? tiles : 1);
    if (ns > ntk / 2) ns = ntk / 2;
    if (ns > 32) ns = 32;
    return ns < 1 ? 1 : ns;
}

// side stream of a host thread (one per device): the dense tile kernel runs there beside K1
struct V2Side {
    cudaStream_t s = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int dev = -1;
};
static thread_local V2Side g_v2_side;
static int v2_side(V2Side** out) {
    int dev = 0;
    SQD_CUDA_OK(cudaGetDevice(&dev));
    if (g_v2_side.s == nullptr || g_v2_side.dev != dev) {
        if (g_v2_side.s != nullptr) {
            cudaStreamDestroy(g_v2_side.s);
            cudaEventDestroy(g_v2_side.ev_fork);
            cudaEventDestroy(g_v2_side.ev_join);
        }
        SQD_CUDA_OK(cudaStreamCreateWithFlags(&g_v2_side.s, cudaStreamNonBlocking));
        SQD_CUDA_OK(cudaEventCreateWithFlags(&g_v2_side.ev_fork, cudaEventDisableTiming));
        SQD_CUDA_OK(cudaEventCreateWithFlags(&g_v2_side.ev_join, cudaEventDisableTiming));
        g_v2_side.dev = dev;
    }
    *out = &g_v2_side;
    return 0;
}

int sigma2_dispatch_rows(const sqd_operator* op, const double* d_c, double* d_sigma, const int* d_done,
                         int row_begin, int row_end, const int* d_slot, long long stride, int in_graph,
                         cudaStream_t st) {
    const sqd_sigma_v2& V = op->v2;
    SQD_REQUIRE(V.enabled && V.P != nullptr, "sqd_sigma: the operator has no v2 tables");
    SQD_REQUIRE(op->ldc % 2 == 0 && op->ldg % 2 == 0 && op->ldc >= op->b.n && op->ldg > op->norb * op->norb,
                "sqd_sigma: ldc/ldg must be even, ldc >= nb and ldg > norb^2");
    SQD_REQUIRE(!op->use_same_spin || (V.HaDT != nullptr && V.HbDT != nullptr),
                "sqd_sigma: the v2 tables were built without the dense same-spin blocks");
    SQD_REQUIRE(row_begin >= 0 && row_end <= op->a.n && row_begin <= row_end, "sqd_sigma: bad row range");
    if (row_end == row_begin) return 0;
    V2Args args{*op, d_c, d_sigma, d_done, row_begin, row_end, 0, d_slot, stride};
    static const int knob_skip = v2_env("SQD_V2_SKIP", 0);  // timing experiments only: bit 0 K1, 1 K2, 2 K3
    // K2 (FP64-pipe bound) and K1 (shared-memory gather bound) are independent: K2 can go to a side stream
    // of this host thread and run beside K1, both joining before K3.  Measured on a lone (30e,30o) 316 x 316
    // build: 43.0 us forked against 43.4 us in one stream -- the two event dependencies cost what the
    // overlap saves -- so the fork is used only inside a captured graph (no host cost there) or on request
    // (SQD_V2_OVERLAP=1).
    static const int knob_overlap = v2_env("SQD_V2_OVERLAP", -1);
    const bool want_fork = knob_overlap >= 0 ? knob_overlap != 0 : in_graph != 0;
    V2Side* side = nullptr;
    const bool fork = want_fork && op->use_same_spin != 0 && V.n_chunks > 0;
    // created on first use by this host thread -- also when this build does not fork, so that a later
    // build inside a stream capture never has to create a stream
    if (op->use_same_spin != 0 && v2_side(&side)) return -2;
    cudaStream_t st2 = fork ? side->s : st;
    int nsplit = 0;
    if (op->use_same_spin) {
        if (fork) {
            SQD_CUDA_OK(cudaEventRecord(side->ev_fork, st));
            SQD_CUDA_OK(cudaStreamWaitEvent(st2, side->ev_fork, 0));
        }
        const int a_lo = row_begin & ~1;
        const int tiles_x = (row_end - a_lo + kTM - 1) / kTM, tiles_y = (op->ldc + kTN - 1) / kTN;
        nsplit = v2_num_split(op->a.n, op->b.n, op->ldc);
        if (nsplit > V.max_split) nsplit = V.max_split;
        if (!(knob_skip & 2)) sigma2_tile_kernel<<<dim3(tiles_x, tiles_y, nsplit), kK2Threads, 0, st2>>>(args);
        if (check_launch("sigma2_tile_kernel")) return -2;
        if (fork) SQD_CUDA_OK(cudaEventRecord(side->ev_join, st2));
    }
    // K1
    const int nst = v2_pick_stages(op);
    const int src_smem = v2_src_in_smem(op);
    const int stage_len = v2_stage_len(op);
    const size_t smem = v2_k1_smem(op, nst);
    SQD_REQUIRE(smem <= 220 * 1024, "sqd_sigma: norb=%d does not fit the integral-row ring", op->norb);
    static const int knob_lean = v2_env("SQD_V2_LEAN", -1);
    const bool lean = knob_lean >= 0 ? knob_lean != 0 : fork;
    auto k1 = V.lmax == 8 ? (lean ? sigma2_ab_kernel<8, true> : sigma2_ab_kernel<8, false>)
                          : (lean ? sigma2_ab_kernel<16, true> : sigma2_ab_kernel<16, false>);
    static bool cfg[64][4] = {};
    int dev = 0;
    SQD_CUDA_OK(cudaGetDevice(&dev));
    const int li = (V.lmax == 8 ? 0 : 1) + (lean ? 2 : 0);
    if (dev >= 0 && dev < 64 && !cfg[dev][li]) {
        static const int knob_carve = v2_env("SQD_V2_CARVEOUT", 100);     // experiments: -1 leaves the default
        static const int knob_cache = v2_env("SQD_CACHE_CONFIG", -1);     // experiments: cudaFuncCache value
        SQD_CUDA_OK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        if (knob_carve >= 0)
            SQD_CUDA_OK(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout, knob_carve));
        if (knob_cache >= 0) SQD_CUDA_OK(cudaDeviceSetCacheConfig((cudaFuncCache)knob_cache));
        cfg[dev][li] = true;
    }
    // lean CTAs (<= 31 K registers, <= 100 KB of ring) share an SM two at a time
    static const int knob_ctas = v2_env("SQD_V2_CTAS_PER_SM", 0);
    const int ctas_per_sm = knob_ctas > 0 ? knob_ctas : (lean && !fork ? 2 : 1);
    int gx = (ctas_per_sm * kNumSMs + V.n_groups - 1) / V.n_groups;
    static const int knob_grid = v2_env("SQD_V2_K1_GRID", 0);  // experiments: absolute CTA count per group
    if (knob_grid > 0 && op->throughput_mode != 0) gx = knob_grid;
    if (gx < 1) gx = 1;
    if (gx > V.n_chunks) gx = V.n_chunks;
    if (V.n_chunks > 0 && !(knob_skip & 1)) {
        k1<<<dim3(gx, V.n_groups), V.vc_pad + 32, smem, st>>>(args, nst, stage_len, src_smem);
        if (check_launch("sigma2_ab_kernel")) return -2;
    }
    if (fork) SQD_CUDA_OK(cudaStreamWaitEvent(st, side->ev_join, 0));
    // K3
    args.nsplit = nsplit;
    if (!(knob_skip & 4)) {
        const int ncb = (op->ldc + kK3Cols - 1) / kK3Cols;
        const int gx3 = (ncb + kK3Warps - 1) / kK3Warps;
        const int heavy_y = (V.n_heavy * ncb + gx3 - 1) / gx3;
        sigma2_epilogue_kernel<<<dim3(gx3, heavy_y + row_end - row_begin), kK3Threads, 0, st>>>(args, heavy_y,
                                                                                                ncb);
    }
    return check_launch("sigma2_epilogue_kernel");
}

}  // namespace sqd

using namespace sqd;

extern "C" {

int sqd_sigma_v2_recommended(int na, int nb, int64_t nnz_a, int64_t nnz_b) {
    static const int knob = v2_env("SQD_SIGMA_V2", -1);  // 0 / 1 force, -1 automatic
    if (knob >= 0) return knob != 0;
    if (na <= 0 || nb <= 0 || nb > 8192 || na > 1 << 19) return 0;
    const double dense = (double)na * nb * ((double)na + nb);
    const double sparse = (double)nnz_a * nb + (double)nnz_b * na;
    return dense <= 12.0 * sparse ? 1 : 0;
}

int64_t sqd_sigma_v2_plan_bytes(int na, int nb, int64_t nnz_a, int64_t nnz_b, int lmax, int items_per_chunk) {
    if (na <= 0 || nb <= 0 || (lmax != 8 && lmax != 16) || items_per_chunk < 1 || items_per_chunk > 32)
        return -1;
    return (int64_t)v2_layout(na, nb, nnz_a, nnz_b, lmax, items_per_chunk).total;
}

const int* sqd_sigma_v2_counts_ptr(void* d_plan, int na, int nb, int64_t nnz_a, int64_t nnz_b, int lmax,
                                   int items_per_chunk) {
    const V2Layout L = v2_layout(na, nb, nnz_a, nnz_b, lmax, items_per_chunk);
    return reinterpret_cast<const int*>((char*)d_plan + L.counts);
}

int sqd_sigma_v2_plan(const sqd_spin_table* a, const sqd_spin_table* b, int norb, int64_t nnz_a,
                      int64_t nnz_b, int lmax, int ipc, void* d_plan, int64_t plan_bytes, int* h_counts,
                      void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(lmax == 8 || lmax == 16, "sqd_sigma_v2_plan: lmax must be 8 or 16");
    SQD_REQUIRE(ipc >= 1 && ipc <= 32, "sqd_sigma_v2_plan: items_per_chunk must be in [1, 32]");
    SQD_REQUIRE(a->n > 0 && b->n > 0, "sqd_sigma_v2_plan: empty string list");
    const V2Layout L = v2_layout(a->n, b->n, nnz_a, nnz_b, lmax, ipc);
    SQD_REQUIRE(plan_bytes >= (int64_t)L.total, "sqd_sigma_v2_plan: plan buffer too small");
    char* base = (char*)d_plan;
    auto I = [&](size_t off) { return reinterpret_cast<int*>(base + off); };
    auto U = [&](size_t off) { return reinterpret_cast<uint32_t*>(base + off); };
    int* counts = I(L.counts);
    SQD_CUDA_OK(cudaMemsetAsync(counts, 0, SQD_V2_COUNTS * sizeof(int), st));
    const size_t bp_smem = (size_t)3 * b->n * sizeof(int);
    if (bp_smem > 200 * 1024 || b->n > 8192) {
        // the single-CTA planner keeps three int arrays of nb entries in shared memory
        int one = 1;
        SQD_CUDA_OK(cudaMemcpyAsync(counts + C_ERR, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    } else {
        v2_alpha_plan_kernel<<<1, 1024, 0, st>>>(*a, ipc, L.maxch, I(L.single_ptr), I(L.item_ptr),
                                                 reinterpret_cast<int4*>(base + L.chunk_rec), counts);
        v2_item_kernel<<<(a->n + 7) / 8, 256, 0, st>>>(*a, norb, I(L.item_ptr), I(L.item_tgt),
                                                       U(L.item_gsel), I(L.item_pslot), I(L.heavy_rows), counts);
        static bool cfg[64] = {};
        int dev = 0;
        SQD_CUDA_OK(cudaGetDevice(&dev));
        if (bp_smem > 40 * 1024 && dev >= 0 && dev < 64 && !cfg[dev]) {
            SQD_CUDA_OK(cudaFuncSetAttribute(v2_beta_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(200 * 1024)));
            cfg[dev] = true;
        }
        v2_beta_plan_kernel<<<1, 1024, bp_smem, st>>>(*b, lmax, L.capw, I(L.col_grp), I(L.col_full),
                                                      I(L.col_nfull), I(L.col_rem), I(L.col_seg), counts);
        const uint32_t zero_off = (uint32_t)(norb * norb * 8);
        v2_vc_init_kernel<<<kNumSMs, 256, 0, st>>>(lmax, L.capw, zero_off, U(L.vc_src), U(L.vc_off),
                                                   I(L.vc_len), I(L.vc_q));
        v2_vc_fill_kernel<<<(b->n + 7) / 8, 256, 0, st>>>(*b, counts, lmax, I(L.col_grp), I(L.col_full),
                                                          I(L.col_nfull), I(L.col_rem), I(L.col_seg),
                                                          U(L.vc_src), U(L.vc_off), I(L.vc_len), I(L.vc_q));
        static const int knob_order = v2_env("SQD_V2_BANK_ORDER", 1);
        if (knob_order)
            v2_vc_order_kernel<<<(L.capw + 127) / 128, 128, 0, st>>>(counts, lmax, U(L.vc_src), U(L.vc_off),
                                                                   I(L.vc_len));
        if (check_launch("sigma v2 plan kernels", 6)) return -2;
    }
    if (h_counts == nullptr) return 0;
    return read_back(h_counts, counts, SQD_V2_COUNTS * sizeof(int), st);
}

static int v2_ld_dense(int n) { return (n + 63) / 64 * 64 + 128; }

struct V2Scratch {
    size_t P, part, HaDT, HbDT, total;
    int ldp, ldq, max_split;
};
static V2Scratch v2_scratch(const int* hc, int na, int nb, int ldc, int dense, int same_tables) {
    V2Scratch S{};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += al256(bytes);
        return at;
    };
    S.ldq = (hc[C_NQ] + 31) / 32 * 32;
    if (S.ldq < 32) S.ldq = 32;
    S.ldp = S.ldq + (nb + 31) / 32 * 32;
    S.P = take((size_t)hc[C_NITEMS] * S.ldp * sizeof(double));
    S.max_split = dense ? v2_num_split(na, nb, ldc) : 1;
    S.part = take(dense ? (size_t)S.max_split * na * ldc * sizeof(double) : 256);
    if (dense) {
        const size_t lda = v2_ld_dense(na), ldb = v2_ld_dense(nb);
        S.HaDT = take(lda * lda * sizeof(double));
        S.HbDT = same_tables ? S.HaDT : take(ldb * ldb * sizeof(double));
    }
    S.total = o;
    return S;
}

int64_t sqd_sigma_v2_scratch_bytes(const int* h_counts, int na, int nb, int ldc, int dense, int same_tables) {
    if (h_counts == nullptr || h_counts[C_ERR] != 0) return -1;
    return (int64_t)v2_scratch(h_counts, na, nb, ldc, dense, same_tables).total;
}

int sqd_sigma_v2_finish(const sqd_spin_table* a, const sqd_spin_table* b, int ldc, int64_t nnz_a,
                        int64_t nnz_b, int lmax, int ipc, const int* hc, void* d_plan, void* d_scratch,
                        int64_t scratch_bytes, int dense, sqd_sigma_v2* out, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(hc != nullptr && hc[C_ERR] == 0, "sqd_sigma_v2_finish: the plan reported an unsupported shape");
    const int na = a->n, nb = b->n;
    const int same_tables = (a->col == b->col && na == nb) ? 1 : 0;
    const V2Layout L = v2_layout(na, nb, nnz_a, nnz_b, lmax, ipc);
    const V2Scratch S = v2_scratch(hc, na, nb, ldc, dense, same_tables);
    SQD_REQUIRE(scratch_bytes >= (int64_t)S.total, "sqd_sigma_v2_finish: scratch buffer too small");
    char* pb = (char*)d_plan;
    char* sb = (char*)d_scratch;
    sqd_sigma_v2 V{};
    V.enabled = 1;
    V.lmax = lmax;
    V.n_groups = hc[C_NGROUPS];
    V.vc_pad = hc[C_VCPAD];
    V.n_items = hc[C_NITEMS];
    V.n_chunks = hc[C_NCHUNKS];
    V.max_split = S.max_split;
    V.ldp = S.ldp;
    V.ldq = S.ldq;
    V.vc_src = (const uint32_t*)(pb + L.vc_src);
    V.vc_off = (const uint32_t*)(pb + L.vc_off);
    V.vc_len = (const int*)(pb + L.vc_len);
    V.vc_q = (const int*)(pb + L.vc_q);
    V.col_seg = (const int*)(pb + L.col_seg);
    V.single_ptr = (const int*)(pb + L.single_ptr);
    V.item_ptr = (const int*)(pb + L.item_ptr);
    V.chunk_rec = (const int*)(pb + L.chunk_rec);
    V.item_tgt = (const int*)(pb + L.item_tgt);
    V.item_gsel = (const uint32_t*)(pb + L.item_gsel);
    V.item_pslot = (const int*)(pb + L.item_pslot);
    V.heavy_rows = (const int*)(pb + L.heavy_rows);
    V.n_heavy = hc[C_NHEAVY];
    V.P = (double*)(sb + S.P);
    V.part = (double*)(sb + S.part);
    SQD_REQUIRE(V.vc_pad >= 32 && V.vc_pad <= v2_group_max(V.lmax) && V.n_groups >= 1 && V.n_groups <= kV2MaxGroups,
                "sqd_sigma_v2_finish: inconsistent plan counts");
    // every P entry the epilogue reads is rewritten by each build; the pad columns are never read
    if (dense) {
        V.lda = v2_ld_dense(na);
        V.ldb = v2_ld_dense(nb);
        double* ha = (double*)(sb + S.HaDT);
        double* hb = (double*)(sb + S.HbDT);
        SQD_CUDA_OK(cudaMemsetAsync(ha, 0, (size_t)V.lda * V.lda * sizeof(double), st));
        v2_dense_scatter_kernel<<<(na + 7) / 8, 256, 0, st>>>(*a, ha, V.lda);
        if (!same_tables) {
            SQD_CUDA_OK(cudaMemsetAsync(hb, 0, (size_t)V.ldb * V.ldb * sizeof(double), st));
            v2_dense_scatter_kernel<<<(nb + 7) / 8, 256, 0, st>>>(*b, hb, V.ldb);
        }
        if (check_launch("v2_dense_scatter_kernel", same_tables ? 1 : 2)) return -2;
        V.HaDT = ha;
        V.HbDT = hb;
    }
    *out = V;
    return 0;
}

}  // extern "C"

// CI sigma-vector build, second generation (v2).  See the comment of sqd_sigma_v2 in include/sqd_b200.h.
//
// Replaces pyscf selected_ci.contract_2e (SCIcontract_2e_bbaa + SCIcontract_2e_aaaa, recalled) reached from
// qiskit_addon_sqd/fermion.py:721-723, 810-818 -- same mathematics as fermion_sigma.cu, different mapping:
//
//   sigma[a,b] = diag[a,b] c[a,b]
//              + sum_{a'} Ha[a,a'] c[a',b] + sum_{b'} Hb[b,b'] c[a,b']              dense FP64 tiles   (K2)
//              + sum_{a' in S_a(a)} sgn_a Wb[pq,b] c[a',b]                           epilogue           (K2)
//              + P[self(a)][b] + sum_{a' in S_a(a)} P[link(a' -> a)][b]              epilogue           (K2)
//   P[self(a')][b]     =         sum_{j in S_b(b)} sgn_j Wa[a', rs_j]   c[a', b'_j]                      (K1)
//   P[link(a'->a)][b]  = sgn_a   sum_{j in S_b(b)} sgn_j g_ab[pq, rs_j] c[a', b'_j]                      (K1)
//
// Why: in the v1 kernel every FMA of the opposite-spin part costs three shared-memory reads (link word,
// c[a',b'_j], g_ab[pq,rs_j]).  Grouping the work by SOURCE string a' makes x_j = sgn_j c[a',b'_j] a constant
// of the thread for all excitations out of a', so it lives in registers together with the byte offsets of
// rs_j: one gather and one FMA per link.  The price is that the result rows have to travel through memory
// (P, L2 resident) because several source strings feed one row of sigma.  The same-spin part of a
// HF-centred sample set is 15-25 % dense, where dense FP64 tiles beat gathers by a wide margin.
#include <stdlib.h>

#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

constexpr int kV2MaxGroups = 64;
constexpr int kV2MaxStages = 8;
constexpr int kV2GroupTarget = 288;   // virtual columns per group aimed at
constexpr int kV2GroupMax = 352;      // hard limit (K1 runs vc_pad + 32 threads, launch bound 384)
enum { C_NITEMS = 0, C_NCHUNKS, C_NGROUPS, C_VCPAD, C_SINGLES_A, C_SINGLES_B, C_ERR, C_NVC, C_NCOLMAX };

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// device-memory layout of the plan (byte offsets); every array is sized by upper bounds that are known
// on the host before the plan kernels run
struct V2Layout {
    size_t single_ptr, item_ptr, chunk_row, chunk_first, chunk_n, rev_slot;
    size_t col_grp, col_u, col_full, col_nfull, col_rem;
    size_t grp_ncol, vc_src, vc_off, vc_len, gcol, gcol_full, gcol_nfull, gcol_rem;
    size_t counter, counts, total;
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
    L.maxch = (int)(na + (na + nnz_a) / ipc + 1);
    // virtual columns: sum_b ceil(len_b / lmax) <= nb + singles_b / lmax; groups add padding
    L.capw = (int)(2 * ((int64_t)nb + nnz_b / lmax) + kV2MaxGroups * 64 + 64);
    L.single_ptr = take((size_t)(na + 1) * 4);
    L.item_ptr = take((size_t)(na + 1) * 4);
    L.chunk_row = take((size_t)L.maxch * 4);
    L.chunk_first = take((size_t)L.maxch * 4);
    L.chunk_n = take((size_t)L.maxch * 4);
    L.rev_slot = take((size_t)(nnz_a > 0 ? nnz_a : 1) * 4);
    L.col_grp = take((size_t)nb * 4);
    L.col_u = take((size_t)nb * 4);
    L.col_full = take((size_t)nb * 4);
    L.col_nfull = take((size_t)nb * 4);
    L.col_rem = take((size_t)nb * 4);
    L.grp_ncol = take((size_t)kV2MaxGroups * 4);
    L.vc_src = take((size_t)L.capw * lmax * 4);
    L.vc_off = take((size_t)L.capw * lmax * 4);
    L.vc_len = take((size_t)L.capw * 4);
    L.gcol = take((size_t)L.capw * 4);
    L.gcol_full = take((size_t)L.capw * 4);
    L.gcol_nfull = take((size_t)L.capw * 4);
    L.gcol_rem = take((size_t)L.capw * 4);
    L.counter = take((size_t)2 * kV2MaxGroups * 4);
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

// alpha side: item numbering and the chunk list, sorted by descending size (heaviest work units first:
// the CTAs of K1 pull chunks from a counter)
__global__ void __launch_bounds__(1024)
v2_alpha_plan_kernel(const sqd_spin_table A, int ipc, int maxch, int* __restrict__ single_ptr,
                     int* __restrict__ item_ptr, int* __restrict__ chunk_row,
                     int* __restrict__ chunk_first, int* __restrict__ chunk_n, int* __restrict__ counts) {
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
            int cnt = 0, nch = 0, last = 0;
            if (a < na) {
                const int T = 1 + A.n_single[a];
                nch = (T + ipc - 1) / ipc;
                last = T - (nch - 1) * ipc;
                cnt = (cls == ipc ? nch - 1 : 0) + (last == cls ? 1 : 0);
            }
            const int ex = block_excl_scan(cnt, warp_tot, &tot);
            if (a < na && cnt > 0) {
                int at = pos + ex;
                if (cls == ipc) {
                    for (int j = 0; j < nch - 1; ++j, ++at) {
                        if (at < maxch) {
                            chunk_row[at] = a;
                            chunk_first[at] = j * ipc;
                            chunk_n[at] = ipc;
                        }
                    }
                }
                if (last == cls && at < maxch) {
                    chunk_row[at] = a;
                    chunk_first[at] = (nch - 1) * ipc;
                    chunk_n[at] = last;
                }
            }
            pos += tot;
        }
    }
    if (threadIdx.x == 0) {
        counts[C_NCHUNKS] = pos;
        if (pos > maxch) counts[C_ERR] = 2;
    }
}

// P row of the reverse link: entry k of row a points at a'; the link a -> a' (source a, target a') is
// entry k' of row a' with col == a, so the opposite-spin result of source a' for target a sits in row
// item_ptr[a'] + 1 + k'.  One warp per row, lanes over its single excitations.
__global__ void v2_rev_slot_kernel(const sqd_spin_table A, const int* __restrict__ single_ptr,
                                   const int* __restrict__ item_ptr, int* __restrict__ rev_slot) {
    const int lane = threadIdx.x & 31;
    const int a = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (a >= A.n) return;
    const int beg = A.row_ptr[a], ns = A.n_single[a], sp = single_ptr[a];
    for (int k = lane; k < ns; k += 32) {
        const int ap = (int)A.col[beg + k];
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
        rev_slot[sp + k] = item_ptr[ap] + 1 + found;
    }
}

// beta side: cut every string's single-excitation list into virtual columns of at most lmax links, deal
// the strings to groups (round robin over the length-sorted order), and number the virtual columns of a
// group: full ones first (strings in rank order), then the remainders by descending length, so that the
// 32 threads of a warp run (almost) the same trip count.  Single CTA; O(nb^2 / 1024) compares per thread.
__global__ void __launch_bounds__(1024)
v2_beta_plan_kernel(const sqd_spin_table B, int lmax, int capw, int* __restrict__ col_grp,
                    int* __restrict__ col_u, int* __restrict__ col_full, int* __restrict__ col_nfull,
                    int* __restrict__ col_rem, int* __restrict__ grp_ncol, int* __restrict__ counts) {
    extern __shared__ int bp_smem[];
    __shared__ int warp_tot[32];
    __shared__ int g_nfull[kV2MaxGroups], g_nrem[kV2MaxGroups], g_ncol[kV2MaxGroups];
    __shared__ int s_G, s_ok;
    const int nb = B.n;
    int* L = bp_smem;          // [nb] list length
    int* rk = L + nb;          // [nb] rank of the string (length descending, index ascending)
    int* by_rank = rk + nb;    // [nb]
    int nvc = 0, nz = 0, sb = 0;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        const int l = B.n_single[b];
        L[b] = l;
        nvc += (l + lmax - 1) / lmax;
        nz += l > 0;
        sb += l;
    }
    int tot;
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
        s_G = nvc > 0 ? (nvc + kV2GroupTarget - 1) / kV2GroupTarget : 1;
        if (s_G > kV2MaxGroups) s_G = kV2MaxGroups;
        s_ok = 0;
    }
    __syncthreads();
    for (int attempt = 0; attempt < kV2MaxGroups; ++attempt) {
        const int G = s_G;
        if (threadIdx.x < G) {
            const int g = threadIdx.x;
            int nfull = 0, nrem = 0, ncol = 0;
            for (int r = g; r < nz; r += G) {
                const int l = L[by_rank[r]];
                nfull += l / lmax;
                nrem += (l % lmax) != 0;
                ++ncol;
            }
            g_nfull[g] = nfull;
            g_nrem[g] = nrem;
            g_ncol[g] = ncol;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int mx = 0;
            for (int g = 0; g < G; ++g) mx = max(mx, g_nfull[g] + g_nrem[g]);
            if (mx <= kV2GroupMax) s_ok = 1;
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
            col_u[b] = 0;
            col_full[b] = 0;
            col_nfull[b] = 0;
            col_rem[b] = -1;
            continue;
        }
        const int g = r % G, rem = l % lmax;
        // remainder rank inside the group: longer remainders first, ties in rank order
        int rr = 0, u = 0;
        for (int r2 = g; r2 < nz; r2 += G) {
            const int b2 = by_rank[r2];
            const int rem2 = L[b2] % lmax;
            rr += (rem2 > rem) || (rem2 == rem && r2 < r);
            u += b2 < b;
        }
        col_grp[b] = g;
        col_u[b] = u;
        col_nfull[b] = l / lmax;
        col_rem[b] = rem ? g_nfull[g] + rr : -1;
    }
    if (threadIdx.x < kV2MaxGroups) grp_ncol[threadIdx.x] = threadIdx.x < G ? g_ncol[threadIdx.x] : 0;
    if (threadIdx.x == 0) {
        int mx = 0, mc = 0;
        for (int g = 0; g < G; ++g) {
            mx = max(mx, g_nfull[g] + g_nrem[g]);
            mc = max(mc, g_ncol[g]);
        }
        const int vc_pad = max(32, (mx + 31) / 32 * 32);
        counts[C_NGROUPS] = G;
        counts[C_VCPAD] = vc_pad;
        counts[C_SINGLES_B] = sb;
        counts[C_NVC] = nvc;
        counts[C_NCOLMAX] = mc;
        if (s_ok != 1 || (long long)G * vc_pad > capw) counts[C_ERR] = 1;
    }
}

__global__ void v2_vc_init_kernel(const int* __restrict__ counts, int lmax, int capw, uint32_t zero_off,
                                  uint32_t* __restrict__ vc_src, uint32_t* __restrict__ vc_off,
                                  int* __restrict__ vc_len, int* __restrict__ gcol,
                                  int* __restrict__ gcol_full, int* __restrict__ gcol_nfull,
                                  int* __restrict__ gcol_rem) {
    const int64_t n = (int64_t)capw * lmax;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        vc_src[i] = 0u;
        vc_off[i] = zero_off;
        if (i < capw) {
            vc_len[i] = 0;
            gcol[i] = -1;
            gcol_full[i] = 0;
            gcol_nfull[i] = 0;
            gcol_rem[i] = -1;
        }
    }
}

// one warp per beta string: scatter its links into the virtual-column arrays of its group
__global__ void v2_vc_fill_kernel(const sqd_spin_table B, const int* __restrict__ counts, int lmax,
                                  const int* __restrict__ col_grp, const int* __restrict__ col_u,
                                  const int* __restrict__ col_full, const int* __restrict__ col_nfull,
                                  const int* __restrict__ col_rem, uint32_t* __restrict__ vc_src,
                                  uint32_t* __restrict__ vc_off, int* __restrict__ vc_len,
                                  int* __restrict__ gcol, int* __restrict__ gcol_full,
                                  int* __restrict__ gcol_nfull, int* __restrict__ gcol_rem) {
    if (counts[C_ERR] != 0) return;
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= B.n) return;
    const int g = col_grp[b];
    if (g < 0) return;
    const int vc_pad = counts[C_VCPAD];
    const int beg = B.row_ptr[b], l = B.n_single[b];
    const int full0 = col_full[b], nfull = col_nfull[b], rem = col_rem[b];
    for (int k = lane; k < l; k += 32) {
        const uint32_t pk = B.pack[beg + k];
        const int seg = k / lmax, j = k - seg * lmax;
        const int vc = seg < nfull ? full0 + seg : rem;
        const size_t at = ((size_t)g * lmax + j) * vc_pad + vc;
        vc_src[at] = (pk & 0x7ffffu) | (pk & 0x80000000u);
        vc_off[at] = ((pk >> 19) & 0xfffu) * 8u;
    }
    if (lane == 0) {
        for (int s = 0; s < nfull; ++s) vc_len[g * vc_pad + full0 + s] = lmax;
        if (rem >= 0) vc_len[g * vc_pad + rem] = l - nfull * lmax;
        const int u = g * vc_pad + col_u[b];
        gcol[u] = b;
        gcol_full[u] = full0;
        gcol_nfull[u] = nfull;
        gcol_rem[u] = rem;
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
};

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

// stage header: x = type (0 exit, 1 new source string, 2 item), y = P row, z = sign, w = source string
template <int LMAX>
__global__ void __launch_bounds__(384, 2)
sigma2_ab_kernel(const V2Args P, const int NST) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    if (P.done != nullptr && *P.done != 0) return;
    const sqd_operator& op = P.op;
    const sqd_sigma_v2& V = op.v2;
    const int g = blockIdx.y;
    const int vc_pad = V.vc_pad, ldg = op.ldg, ldc = op.ldc;
    const int tid = threadIdx.x, lane = tid & 31;
    const int ncons = blockDim.x - 32, nwarp_c = ncons >> 5;

    double* stage = reinterpret_cast<double*>(smem_raw);                 // [NST][ldg]
    double* accv = stage + (size_t)NST * ldg;                             // [2][vc_pad]
    uint64_t* full = reinterpret_cast<uint64_t*>(accv + 2 * vc_pad);     // [kV2MaxStages]
    uint64_t* empty = full + kV2MaxStages;                                // [kV2MaxStages]
    int4* hdr = reinterpret_cast<int4*>(empty + kV2MaxStages);            // [kV2MaxStages]

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
        int s = 0, round = 0;
        const bool self_ok = op.Wa != nullptr;
        const int norb = op.norb;
        for (;;) {
            int chunk = 0;
            if (lane == 0) chunk = atomicAdd(&V.counter[g], 1);
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
            if (chunk >= V.n_chunks) break;
            const int ap = V.chunk_row[chunk], first = V.chunk_first[chunk], n = V.chunk_n[chunk];
            // lane k describes item first + k of the chunk
            const int it = first + lane;
            bool live = false;
            int slot = 0, sign = 1;
            const double* grow = nullptr;
            if (lane < n) {
                slot = V.item_ptr[ap] + it;
                if (it == 0) {
                    live = self_ok && ap >= P.row_begin && ap < P.row_end;
                    grow = op.Wa + (size_t)ap * ldg;
                } else {
                    const int e = op.a.row_ptr[ap] + it - 1;
                    const int tgt = (int)op.a.col[e];
                    const uint32_t m = op.a.meta[e];
                    live = tgt >= P.row_begin && tgt < P.row_end;
                    // row `ap` stores E_pq tgt = ap; the excitation ap -> tgt is E_qp
                    const int pq = (int)(m & 0x7fffffffu), p = pq / norb, q = pq - p * norb;
                    grow = op.gab + (size_t)(q * norb + p) * ldg;
                    sign = (m >> 31) ? -1 : 1;
                }
            }
            uint32_t rem = __ballot_sync(0xffffffffu, live);
            if (rem == 0) continue;
            if (lane == 0) {
                if (round > 0) v2_wait(&empty[s], (uint32_t)((round - 1) & 1));
                hdr[s] = make_int4(1, 0, 0, ap);
                v2_arrive(&full[s]);
            }
            if (++s == NST) {
                s = 0;
                ++round;
            }
            while (rem) {
                const int k = __ffs(rem) - 1;
                rem &= rem - 1;
                const int slot_k = __shfl_sync(0xffffffffu, slot, k);
                const int sign_k = __shfl_sync(0xffffffffu, sign, k);
                const unsigned long long gp =
                    __shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)grow, k);
                if (lane == 0) {
                    if (round > 0) v2_wait(&empty[s], (uint32_t)((round - 1) & 1));
                    hdr[s] = make_int4(2, slot_k, sign_k, ap);
                    mbar_expect_tx(&full[s], (uint32_t)(ldg * sizeof(double)));
                    bulk_g2s(stage + (size_t)s * ldg, reinterpret_cast<const double*>((uintptr_t)gp),
                             (uint32_t)(ldg * sizeof(double)), &full[s]);
                }
                if (++s == NST) {
                    s = 0;
                    ++round;
                }
            }
        }
        if (lane == 0) {
            if (round > 0) v2_wait(&empty[s], (uint32_t)((round - 1) & 1));
            hdr[s] = make_int4(0, 0, 0, 0);
            v2_arrive(&full[s]);
            // the last CTA of the group to run out of work re-arms the counter for the next build
            __threadfence();
            const int t = atomicAdd(&V.counter[kV2MaxGroups + g], 1);
            if (t == (int)gridDim.x - 1) {
                V.counter[g] = 0;
                V.counter[kV2MaxGroups + g] = 0;
            }
        }
        return;
    }

    // =========================== consumer warps ==========================
    const int t = tid;
    const size_t gbase = (size_t)g * LMAX * vc_pad;
    const int len = V.vc_len[g * vc_pad + t];
    int wlen = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wlen = max(wlen, __shfl_xor_sync(0xffffffffu, wlen, o));
    uint32_t off[LMAX];
    double x[LMAX];
#pragma unroll
    for (int j = 0; j < LMAX; ++j) {
        off[j] = V.vc_off[gbase + (size_t)j * vc_pad + t];
        x[j] = 0.0;
    }
    // second role: thread u < ncol adds up the virtual columns of the u-th beta string of the group
    const int ncol = V.grp_ncol[g];
    int cb = -1, cfull = 0, cnf = 0, crem = -1;
    if (t < ncol) {
        cb = V.gcol[g * vc_pad + t];
        cfull = V.gcol_full[g * vc_pad + t];
        cnf = V.gcol_nfull[g * vc_pad + t];
        crem = V.gcol_rem[g * vc_pad + t];
    }
    int s = 0, round = 0, buf = 0;
    for (;;) {
        v2_wait(&full[s], (uint32_t)(round & 1));
        const int4 h = hdr[s];
        if (h.x == 0) break;
        if (h.x == 1) {
            const double* crow = P.c + (size_t)h.w * ldc;
#pragma unroll
            for (int j = 0; j < LMAX; ++j) {
                if (j < wlen) {
                    const uint32_t sv = __ldg(V.vc_src + gbase + (size_t)j * vc_pad + t);
                    const double v = __ldg(crow + (sv & 0x7fffffffu));
                    x[j] = j < len ? ((sv >> 31) ? -v : v) : 0.0;
                }
            }
        } else {
            const char* G = reinterpret_cast<const char*>(stage + (size_t)s * ldg);
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < LMAX; ++j)
                if (j < wlen) acc = fma(x[j], *reinterpret_cast<const double*>(G + off[j]), acc);
            accv[buf * vc_pad + t] = h.z < 0 ? -acc : acc;
        }
        __syncwarp();
        if (lane == 0) v2_arrive(&empty[s]);
        if (++s == NST) {
            s = 0;
            ++round;
        }
        if (h.x == 2) {
            asm volatile("bar.sync 1, %0;" ::"r"(ncons) : "memory");
            if (cb >= 0) {
                const double* av = accv + buf * vc_pad;
                double v = 0.0;
                for (int k = 0; k < cnf; ++k) v += av[cfull + k];
                if (crem >= 0) v += av[crem];
                V.P[(size_t)h.y * ldc + cb] = v;
            }
            buf ^= 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K2: dense same-spin tiles + epilogue
// ---------------------------------------------------------------------------------------------------
constexpr int kTM = 64, kTN = 64, kKT = 16, kK2Threads = 128, kSdA = 66, kSdB = 64;

__global__ void __launch_bounds__(kK2Threads, 3)
sigma2_tile_kernel(const V2Args P) {
    __shared__ __align__(16) double As[2][kKT][kSdA];
    __shared__ __align__(16) double Bs[2][kKT][kSdB];
    __shared__ int is_last_s;
    if (P.done != nullptr && *P.done != 0) return;
    const sqd_operator& op = P.op;
    const sqd_sigma_v2& V = op.v2;
    const int na = op.a.n, nb = op.b.n, ldc = op.ldc;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int a_lo = P.row_begin & ~1;
    const int a0 = a_lo + blockIdx.x * kTM, b0 = blockIdx.y * kTN;
    const int split = blockIdx.z, nsplit = gridDim.z;
    const bool ham = op.use_same_spin != 0;
    const double* __restrict__ c = P.c;

    double acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[r][q] = 0.0;

    // combined K range: k-tiles [0, nka) run over source alpha strings (HaDT, C), [nka, nka+nkb) over
    // source beta strings (C^T, HbDT)
    const int nka = ham ? (na + kKT - 1) / kKT : 0, nkb = ham ? (nb + kKT - 1) / kKT : 0;
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

    // ---- split-K: partial tiles meet in memory, the last CTA of the tile adds them in split order ----
    if (nsplit > 1) {
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
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            int* ticket = V.tile_ticket + blockIdx.y * gridDim.x + blockIdx.x;
            const int tk = atomicAdd(ticket, 1);
            is_last_s = tk == nsplit - 1;
            if (is_last_s) *ticket = 0;
        }
        __syncthreads();
        if (!is_last_s) return;
        __threadfence();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int a = a0 + ty * 8 + r;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int b = b0 + 32 * h + 2 * tx;
                double2 v = make_double2(0.0, 0.0);
                if (a < na && b < ldc) {
                    for (int sp = 0; sp < nsplit; ++sp) {
                        const double2 w = __ldcg(
                            reinterpret_cast<const double2*>(V.part + ((size_t)sp * na + a) * ldc + b));
                        v.x += w.x;
                        v.y += w.y;
                    }
                }
                acc[r][2 * h] = v.x;
                acc[r][2 * h + 1] = v.y;
            }
        }
    }

    // ---- epilogue -------------------------------------------------------------------------------
    const bool have_self = op.Wa != nullptr;
    const bool have_wb = op.Wb != nullptr;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int a = a0 + ty * 8 + r;
        if (a < P.row_begin || a >= P.row_end) continue;
        const int rp = op.a.row_ptr[a], ns = op.a.n_single[a], sp = V.single_ptr[a];
        const int self_row = V.item_ptr[a];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int b = b0 + 32 * h + 2 * tx;
            if (b >= ldc) continue;
            const size_t ab = (size_t)a * ldc + b;
            const double2 dg = *reinterpret_cast<const double2*>(op.diag + ab);
            const double2 cc = *reinterpret_cast<const double2*>(c + ab);
            double vx = acc[r][2 * h], vy = acc[r][2 * h + 1];
            vx = fma(dg.x, cc.x, vx);
            vy = fma(dg.y, cc.y, vy);
            if (have_self) {
                const double2 ps = *reinterpret_cast<const double2*>(V.P + (size_t)self_row * ldc + b);
                vx += ps.x;
                vy += ps.y;
            }
            for (int k = 0; k < ns; ++k) {
                const int slot = V.rev_slot[sp + k];
                const double2 pv = *reinterpret_cast<const double2*>(V.P + (size_t)slot * ldc + b);
                vx += pv.x;
                vy += pv.y;
                if (have_wb) {
                    const uint32_t m = op.a.meta[rp + k];
                    const int ap = (int)op.a.col[rp + k];
                    const double2 wb =
                        *reinterpret_cast<const double2*>(op.Wb + (size_t)(m & 0x7fffffffu) * ldc + b);
                    const double2 cs = *reinterpret_cast<const double2*>(c + (size_t)ap * ldc + b);
                    const double sg = (m >> 31) ? -1.0 : 1.0;
                    vx = fma(sg * wb.x, cs.x, vx);
                    vy = fma(sg * wb.y, cs.y, vy);
                }
            }
            if (b >= nb) vx = 0.0;       // pad column of sigma
            if (b + 1 >= nb) vy = 0.0;
            *reinterpret_cast<double2*>(P.sigma + ab) = make_double2(vx, vy);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int v2_env(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

static size_t v2_k1_smem(const sqd_operator* op, int nst) {
    return (size_t)nst * op->ldg * sizeof(double) + (size_t)2 * op->v2.vc_pad * sizeof(double) +
           (size_t)2 * kV2MaxStages * sizeof(uint64_t) + (size_t)kV2MaxStages * sizeof(int4);
}

static int v2_pick_stages(const sqd_operator* op) {
    static const int knob = v2_env("SQD_V2_STAGES", 0);
    if (knob >= 2 && knob <= kV2MaxStages) return knob;
    // deep enough to cover the L2 latency of a row copy behind ~0.2 us items, small enough for 2 CTAs/SM
    int nst = 4;
    while (nst > 2 && v2_k1_smem(op, nst) > 100 * 1024) --nst;
    return nst;
}

int64_t sigma2_smem_bytes(const sqd_operator* op) {
    const size_t b = v2_k1_smem(op, v2_pick_stages(op));
    return b <= 220 * 1024 ? (int64_t)b : -1;
}

int sigma2_dispatch_rows(const sqd_operator* op, const double* d_c, double* d_sigma, const int* d_done,
                         int row_begin, int row_end, cudaStream_t st) {
    const sqd_sigma_v2& V = op->v2;
    SQD_REQUIRE(V.enabled && V.P != nullptr, "sqd_sigma: the operator has no v2 tables");
    SQD_REQUIRE(op->ldc % 2 == 0 && op->ldg % 2 == 0 && op->ldc >= op->b.n && op->ldg > op->norb * op->norb,
                "sqd_sigma: ldc/ldg must be even, ldc >= nb and ldg > norb^2");
    SQD_REQUIRE(!op->use_same_spin || (V.HaDT != nullptr && V.HbDT != nullptr),
                "sqd_sigma: the v2 tables were built without the dense same-spin blocks");
    SQD_REQUIRE(row_begin >= 0 && row_end <= op->a.n && row_begin <= row_end, "sqd_sigma: bad row range");
    if (row_end == row_begin) return 0;
    V2Args args{*op, d_c, d_sigma, d_done, row_begin, row_end, 1};
    // K1
    const int nst = v2_pick_stages(op);
    const size_t smem = v2_k1_smem(op, nst);
    SQD_REQUIRE(smem <= 220 * 1024, "sqd_sigma: norb=%d does not fit the integral-row ring", op->norb);
    auto k1 = V.lmax == 8 ? sigma2_ab_kernel<8> : sigma2_ab_kernel<16>;
    static bool cfg[64][2] = {};
    int dev = 0;
    SQD_CUDA_OK(cudaGetDevice(&dev));
    const int li = V.lmax == 8 ? 0 : 1;
    if (dev >= 0 && dev < 64 && !cfg[dev][li]) {
        SQD_CUDA_OK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SQD_CUDA_OK(cudaFuncSetAttribute(k1, cudaFuncAttributePreferredSharedMemoryCarveout,
                                         (int)cudaSharedmemCarveoutMaxShared));
        cfg[dev][li] = true;
    }
    static const int knob_ctas = v2_env("SQD_V2_CTAS_PER_SM", 2);
    int gx = (knob_ctas * kNumSMs) / V.n_groups;
    if (gx < 1) gx = 1;
    if (gx > V.n_chunks) gx = V.n_chunks;
    if (V.n_chunks > 0) {
        k1<<<dim3(gx, V.n_groups), V.vc_pad + 32, smem, st>>>(args, nst);
        if (check_launch("sigma2_ab_kernel")) return -2;
    }
    // K2
    const int a_lo = row_begin & ~1;
    const int tiles_x = (row_end - a_lo + kTM - 1) / kTM, tiles_y = (op->ldc + kTN - 1) / kTN;
    int nsplit = 1;
    if (op->use_same_spin) {
        static const int knob_split = v2_env("SQD_V2_SPLIT", 0);
        const int ntk = (op->a.n + kKT - 1) / kKT + (op->b.n + kKT - 1) / kKT;
        nsplit = knob_split > 0 ? knob_split : (3 * kNumSMs) / (tiles_x * tiles_y);
        if (nsplit > ntk / 2) nsplit = ntk / 2;
        if (nsplit > V.max_split) nsplit = V.max_split;
        if (nsplit < 1) nsplit = 1;
    }
    args.nsplit = nsplit;
    sigma2_tile_kernel<<<dim3(tiles_x, tiles_y, nsplit), kK2Threads, 0, st>>>(args);
    return check_launch("sigma2_tile_kernel");
}

}  // namespace sqd

using namespace sqd;

extern "C" {

int sqd_sigma_v2_recommended(int na, int nb, int64_t nnz_a, int64_t nnz_b) {
    static const int knob = v2_env("SQD_SIGMA_V2", -1);  // 0 / 1 force, -1 automatic
    if (knob >= 0) return knob != 0;
    if (na <= 0 || nb <= 0 || nb > 16384 || na > 1 << 19) return 0;
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
    SQD_CUDA_OK(cudaMemsetAsync(I(L.counter), 0, 2 * kV2MaxGroups * sizeof(int), st));
    const size_t bp_smem = (size_t)3 * b->n * sizeof(int);
    if (bp_smem > 200 * 1024) {
        // the single-CTA planner keeps three int arrays of nb entries in shared memory
        int one = 1;
        SQD_CUDA_OK(cudaMemcpyAsync(counts + C_ERR, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    } else {
        v2_alpha_plan_kernel<<<1, 1024, 0, st>>>(*a, ipc, L.maxch, I(L.single_ptr), I(L.item_ptr),
                                                 I(L.chunk_row), I(L.chunk_first), I(L.chunk_n), counts);
        v2_rev_slot_kernel<<<(a->n + 7) / 8, 256, 0, st>>>(*a, I(L.single_ptr), I(L.item_ptr), I(L.rev_slot));
        static bool cfg[64] = {};
        int dev = 0;
        SQD_CUDA_OK(cudaGetDevice(&dev));
        if (bp_smem > 40 * 1024 && dev >= 0 && dev < 64 && !cfg[dev]) {
            SQD_CUDA_OK(cudaFuncSetAttribute(v2_beta_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)(200 * 1024)));
            cfg[dev] = true;
        }
        v2_beta_plan_kernel<<<1, 1024, bp_smem, st>>>(*b, lmax, L.capw, I(L.col_grp), I(L.col_u),
                                                      I(L.col_full), I(L.col_nfull), I(L.col_rem),
                                                      I(L.grp_ncol), counts);
        const uint32_t zero_off = (uint32_t)(norb * norb * 8);
        v2_vc_init_kernel<<<kNumSMs, 256, 0, st>>>(counts, lmax, L.capw, zero_off, U(L.vc_src), U(L.vc_off),
                                                   I(L.vc_len), I(L.gcol), I(L.gcol_full), I(L.gcol_nfull),
                                                   I(L.gcol_rem));
        v2_vc_fill_kernel<<<(b->n + 7) / 8, 256, 0, st>>>(*b, counts, lmax, I(L.col_grp), I(L.col_u),
                                                          I(L.col_full), I(L.col_nfull), I(L.col_rem),
                                                          U(L.vc_src), U(L.vc_off), I(L.vc_len), I(L.gcol),
                                                          I(L.gcol_full), I(L.gcol_nfull), I(L.gcol_rem));
        if (check_launch("sigma v2 plan kernels", 5)) return -2;
    }
    if (h_counts == nullptr) return 0;
    return read_back(h_counts, counts, SQD_V2_COUNTS * sizeof(int), st);
}

static int v2_max_split(int na, int nb, int ldc) {
    const int tiles = ((na + 1 + kTM - 1) / kTM) * ((ldc + kTN - 1) / kTN);
    const int ntk = (na + kKT - 1) / kKT + (nb + kKT - 1) / kKT;
    int ns = (3 * kNumSMs) / (tiles > 0 ? tiles : 1);
    if (ns > ntk / 2) ns = ntk / 2;
    if (ns > 32) ns = 32;
    return ns < 1 ? 1 : ns;
}

static int v2_ld_dense(int n) { return (n + 63) / 64 * 64 + 128; }

struct V2Scratch {
    size_t P, part, ticket, HaDT, HbDT, total;
};
static V2Scratch v2_scratch(const int* hc, int na, int nb, int ldc, int dense, int same_tables) {
    V2Scratch S{};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += al256(bytes);
        return at;
    };
    S.P = take((size_t)hc[C_NITEMS] * ldc * sizeof(double));
    const int ms = dense ? v2_max_split(na, nb, ldc) : 1;
    S.part = take(ms > 1 ? (size_t)ms * na * ldc * sizeof(double) : 256);
    S.ticket = take((size_t)(((na + 1 + kTM - 1) / kTM + 1) * ((ldc + kTN - 1) / kTN)) * sizeof(int));
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
    V.max_split = dense ? v2_max_split(na, nb, ldc) : 1;
    V.vc_src = (const uint32_t*)(pb + L.vc_src);
    V.vc_off = (const uint32_t*)(pb + L.vc_off);
    V.vc_len = (const int*)(pb + L.vc_len);
    V.grp_ncol = (const int*)(pb + L.grp_ncol);
    V.gcol = (const int*)(pb + L.gcol);
    V.gcol_full = (const int*)(pb + L.gcol_full);
    V.gcol_nfull = (const int*)(pb + L.gcol_nfull);
    V.gcol_rem = (const int*)(pb + L.gcol_rem);
    V.single_ptr = (const int*)(pb + L.single_ptr);
    V.item_ptr = (const int*)(pb + L.item_ptr);
    V.chunk_row = (const int*)(pb + L.chunk_row);
    V.chunk_first = (const int*)(pb + L.chunk_first);
    V.chunk_n = (const int*)(pb + L.chunk_n);
    V.rev_slot = (const int*)(pb + L.rev_slot);
    V.counter = (int*)(pb + L.counter);
    V.tile_ticket = (int*)(sb + S.ticket);
    V.P = (double*)(sb + S.P);
    V.part = (double*)(sb + S.part);
    SQD_REQUIRE(V.vc_pad >= 32 && V.vc_pad <= kV2GroupMax && V.n_groups >= 1 && V.n_groups <= kV2MaxGroups,
                "sqd_sigma_v2_finish: inconsistent plan counts");
    // P must be zero where K1 never writes (beta strings without single excitations)
    SQD_CUDA_OK(cudaMemsetAsync(V.P, 0, (size_t)V.n_items * ldc * sizeof(double), st));
    SQD_CUDA_OK(cudaMemsetAsync(sb + S.ticket, 0,
                                (size_t)(((na + 1 + kTM - 1) / kTM + 1) * ((ldc + kTN - 1) / kTN)) * sizeof(int),
                                st));
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

// Pauli-operator projection onto a sampled computational-basis subspace, and the CSR matvec used by
// the eigensolver.  Reference: qiskit_addon_sqd/qubit.py:78-300 (jax vmap + numpy isin/searchsorted +
// scipy coo additions, one Python iteration per Pauli term).
//
// B200 design: bitstrings are int64 keys (sorted, unique).  A Pauli term is (xmask, zmask, #Y):
//   P |key> = i^{#Y} (-1)^{popc(key & zmask)} |key ^ xmask>
// Terms are grouped by xmask on the host; one warp owns a row (source key) and its lanes stride the
// groups: binary search of key^xmask in the key list, then the group's terms are accumulated in their
// original order (same floating-point sum order as the reference's term-by-term `operator += ...`).
// Hits are ballot-compacted into the row's CSR segment and rank-sorted by column.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

__global__ void bits_to_keys_kernel(const uint8_t* __restrict__ bits, int64_t n, int nbits,
                                    int64_t* __restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const uint8_t* src = bits + row * nbits;
    uint64_t word = 0;
    for (int j0 = 0; j0 < nbits; j0 += 32) {
        const int j = j0 + lane;
        const bool bit = (j < nbits) && src[j] != 0;
        const uint32_t m = __ballot_sync(0xffffffffu, bit);
        const int chunk = min(32, nbits - j0);
        const uint64_t v = (uint64_t)(__brev(m) >> (32 - chunk));
        word |= v << (nbits - j0 - chunk);
    }
    if (lane == 0) keys[row] = (int64_t)word;
}

// index of `target` in sorted keys[0..d) or -1
__device__ __forceinline__ int64_t find_key(const int64_t* __restrict__ keys, int64_t d,
                                            int64_t target) {
    int64_t lo = 0, hi = d;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < target)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (lo < d && __ldg(keys + lo) == target) ? lo : -1;
}

// ---- key -> row hash table (open addressing, linear probing, load <= 1/2) ------------------------------
// The projection asks "is key ^ xmask in the subspace?" n_groups times per row and the answer is almost
// always no; a probe of a half-empty table settles that in ~1.5 L2 accesses where the binary search over the
// sorted keys needs log2(d) dependent ones.
constexpr int64_t kEmptyKey = (int64_t)0x8000000000000000ull;   // keys are < 2^63

__device__ __forceinline__ uint64_t hash_key(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}

struct KeyTable {
    const int64_t* slot_key;
    const int32_t* slot_row;
    uint64_t mask;   // capacity - 1 (capacity is a power of two)
};

__global__ void key_table_clear_kernel(int64_t* __restrict__ slot_key, int64_t cap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) slot_key[i] = kEmptyKey;
}

__global__ void key_table_build_kernel(const int64_t* __restrict__ keys, int64_t d, int64_t* slot_key,
                                       int32_t* __restrict__ slot_row, uint64_t mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    const int64_t k = keys[i];
    uint64_t s = hash_key((uint64_t)k) & mask;
    for (;;) {
        const unsigned long long old = atomicCAS((unsigned long long*)(slot_key + s), (unsigned long long)kEmptyKey,
                                                 (unsigned long long)k);
        if (old == (unsigned long long)kEmptyKey || old == (unsigned long long)k) {
            // duplicates (the caller promises there are none) keep the smallest row
            atomicMin(slot_row + s, (int32_t)i);
            return;
        }
        s = (s + 1) & mask;
    }
}

__device__ __forceinline__ int64_t table_find(const KeyTable& t, int64_t target) {
    uint64_t s = hash_key((uint64_t)target) & t.mask;
    for (;;) {
        const int64_t k = __ldg(t.slot_key + s);
        if (k == target) return __ldg(t.slot_row + s);
        if (k == kEmptyKey) return -1;
        s = (s + 1) & t.mask;
    }
}

__global__ void pauli_connect_kernel(const int64_t* __restrict__ keys, int64_t d, KeyTable table, int use_table,
                                     uint64_t xmask, uint64_t zmask, int32_t* __restrict__ col,
                                     uint8_t* __restrict__ par) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    const int64_t key = keys[i];
    // a diagonal Pauli (no X or Y) connects every configuration to itself: no search
    col[i] = xmask == 0 ? (int32_t)i
                        : (int32_t)(use_table ? table_find(table, key ^ (int64_t)xmask)
                                              : find_key(keys, d, key ^ (int64_t)xmask));
    par[i] = (uint8_t)(popc64((uint64_t)key & zmask) & 1);
}

// same, with the outputs in the reference's own types: complex128 amplitude i^ny (-1)^popc(key & zmask) and
// int64 column (-1: the image is not in the subspace); *n_missing counts those rows
__global__ void pauli_elements_kernel(const int64_t* __restrict__ keys, int64_t d, KeyTable table, int use_table,
                                      uint64_t xmask, uint64_t zmask, int ny, int64_t* __restrict__ col,
                                      double2* __restrict__ amp, int32_t* __restrict__ n_missing) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t j = 0;
    if (i < d) {
        const int64_t key = keys[i];
        j = xmask == 0 ? i
                       : (use_table ? table_find(table, key ^ (int64_t)xmask) : find_key(keys, d, key ^ (int64_t)xmask));
        const double sgn = (popc64((uint64_t)key & zmask) & 1) ? -1.0 : 1.0;
        double2 a;
        switch (ny & 3) {
            case 0: a = make_double2(sgn, 0.0); break;
            case 1: a = make_double2(0.0, sgn); break;
            case 2: a = make_double2(-sgn, 0.0); break;
            default: a = make_double2(0.0, -sgn); break;
        }
        if (col) col[i] = j;
        amp[i] = a;
    }
    const unsigned miss = __ballot_sync(0xffffffffu, j < 0);
    if ((threadIdx.x & 31) == 0 && miss) atomicAdd(n_missing, __popc(miss));
}

// coefficient * i^ny * (+-1)
__device__ __forceinline__ void accumulate_term(double& re, double& im, double cr, double ci, int ny,
                                                bool neg) {
    double tr, ti;
    switch (ny & 3) {
        case 0: tr = cr; ti = ci; break;
        case 1: tr = -ci; ti = cr; break;   // * i
        case 2: tr = -cr; ti = -ci; break;  // * -1
        default: tr = ci; ti = -cr; break;  // * -i
    }
    if (neg) {
        tr = -tr;
        ti = -ti;
    }
    re += tr;
    im += ti;
}

// signed value of one term for a source key: coefficient * i^ny * (+-1)
__device__ __forceinline__ void term_value(double cr, double ci, int ny, bool neg, double& tr, double& ti) {
    switch (ny & 3) {
        case 0: tr = cr; ti = ci; break;
        case 1: tr = -ci; ti = cr; break;   // * i
        case 2: tr = -cr; ti = -ci; break;  // * -1
        default: tr = ci; ti = -cr; break;  // * -i
    }
    if (neg) {
        tr = -tr;
        ti = -ti;
    }
}

// USE_TABLE: hash probes (4 independent probes per lane in flight); otherwise binary search of the key list.
// (A shared-memory membership filter in front of the table was tried and measured slower: the probes are not
// what the kernel spends its time on once the large term groups are evaluated cooperatively.)
template <bool FILL, bool USE_TABLE>
__global__ void pauli_project_kernel(const int64_t* __restrict__ keys, int64_t d, KeyTable table,
                                     const uint64_t* __restrict__ grp_xmask,
                                     const int32_t* __restrict__ grp_ptr, int32_t n_groups,
                                     const uint64_t* __restrict__ zmask, const int32_t* __restrict__ ny,
                                     const double* __restrict__ coeff,
                                     const int32_t* __restrict__ row_ptr, int32_t* __restrict__ row_nnz,
                                     int32_t* __restrict__ col, double* __restrict__ val, int32_t diag_group,
                                     const double* __restrict__ diag_val) {
    constexpr int U = USE_TABLE ? 4 : 1;
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= d) return;
    const int64_t key = keys[i];
    int count = 0;
    const int base = FILL ? row_ptr[i] : 0;
    for (int32_t k0 = 0; k0 < n_groups; k0 += 32 * U) {
        int64_t j[U];
        if (USE_TABLE) {
            // first probe of every group issued before any is examined
            uint64_t slot[U];
            int64_t tgt[U], got[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int32_t k = k0 + u * 32 + lane;
                tgt[u] = key ^ (int64_t)(k < n_groups ? grp_xmask[k] : 0ull);
                slot[u] = hash_key((uint64_t)tgt[u]) & table.mask;
                got[u] = k < n_groups ? __ldg(table.slot_key + slot[u]) : kEmptyKey;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                int64_t g = got[u];
                uint64_t sl = slot[u];
                while (g != tgt[u] && g != kEmptyKey) {
                    sl = (sl + 1) & table.mask;
                    g = __ldg(table.slot_key + sl);
                }
                j[u] = (g == tgt[u] && k0 + u * 32 + lane < n_groups) ? __ldg(table.slot_row + sl) : -1;
            }
        } else {
            const int32_t k = k0 + lane;
            j[0] = k < n_groups ? find_key(keys, d, key ^ (int64_t)grp_xmask[k]) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int32_t k = k0 + u * 32 + lane;
            double re = 0.0, im = 0.0;
            int32_t g0 = 0, g1 = 0;
            if (j[u] >= 0) {
                g0 = grp_ptr[k];
                g1 = grp_ptr[k + 1];
            }
            // A group with many terms (the diagonal group of a Hamiltonian easily holds thousands of Z strings)
            // is evaluated by the whole warp: 32 terms at a time, every lane the signed value of one term, then
            // the values are added IN TERM ORDER (shuffles) -- the same floating-point sums as the owning
            // lane's own loop, at a tenth of the instructions.
            const bool pre = j[u] >= 0 && k == diag_group;   // summed beforehand by pauli_diag_group_kernel
            if (pre) {
                re = diag_val[2 * i];
                im = diag_val[2 * i + 1];
            }
            const bool big = !pre && g1 - g0 >= 8;
            if (j[u] >= 0 && !big && !pre) {
                for (int32_t t = g0; t < g1; ++t)
                    accumulate_term(re, im, coeff[2 * t], coeff[2 * t + 1], ny[t],
                                    popc64((uint64_t)key & zmask[t]) & 1);
            }
            uint32_t mb = __ballot_sync(0xffffffffu, big);
            while (mb) {
                const int owner = __ffs(mb) - 1;
                mb &= mb - 1;
                const int32_t b0 = __shfl_sync(0xffffffffu, g0, owner), b1 = __shfl_sync(0xffffffffu, g1, owner);
                double sre = 0.0, sim = 0.0;
                for (int32_t t0 = b0; t0 < b1; t0 += 32) {
                    const int32_t t = t0 + lane;
                    double tr = 0.0, ti = 0.0;
                    if (t < b1)
                        term_value(coeff[2 * t], coeff[2 * t + 1], ny[t], popc64((uint64_t)key & zmask[t]) & 1, tr, ti);
                    const int cnt = b1 - t0 < 32 ? b1 - t0 : 32;
                    for (int l = 0; l < cnt; ++l) {
                        sre += __shfl_sync(0xffffffffu, tr, l);
                        sim += __shfl_sync(0xffffffffu, ti, l);
                    }
                }
                if (lane == owner) {
                    re = sre;
                    im = sim;
                }
            }
            if (j[u] >= 0 && re == 0.0 && im == 0.0) j[u] = -1;  // scipy drops exact zeros when summing
            const uint32_t m = __ballot_sync(0xffffffffu, j[u] >= 0);
            if (FILL && j[u] >= 0) {
                const int o = base + count + __popc(m & ((1u << lane) - 1u));
                col[o] = (int32_t)j[u];
                val[2 * o] = re;
                val[2 * o + 1] = im;
            }
            count += __popc(m);
        }
    }
    if (!FILL && lane == 0) row_nnz[i] = count;
}

// The diagonal group (X mask 0: every row is its own image) of a Hamiltonian easily holds thousands of Z strings
// and EVERY row needs its sum.  One thread per row, the terms in order: all threads of a warp read the same
// term (uniform loads) and add its signed value to their own row's running sum -- the sequential summation
// order of the reference, ~8 instructions per (32 rows, term).
__global__ void pauli_diag_group_kernel(const int64_t* __restrict__ keys, int64_t d,
                                        const uint64_t* __restrict__ zmask, const int32_t* __restrict__ ny,
                                        const double* __restrict__ coeff, int32_t t0, int32_t t1,
                                        double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    const uint64_t key = (uint64_t)keys[i];
    double re = 0.0, im = 0.0;
    for (int32_t t = t0; t < t1; ++t)
        accumulate_term(re, im, __ldg(coeff + 2 * t), __ldg(coeff + 2 * t + 1), __ldg(ny + t),
                        popc64(key & __ldg(zmask + t)) & 1);
    out[2 * i] = re;
    out[2 * i + 1] = im;
}

// rank sort of each CSR row by column (columns are unique inside a row)
__global__ void csr_row_sort_kernel(int64_t d, const int32_t* __restrict__ row_ptr,
                                    const int32_t* __restrict__ col_in, const double* __restrict__ val_in,
                                    int32_t* __restrict__ col_out, double* __restrict__ val_out) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= d) return;
    const int beg = row_ptr[i], end = row_ptr[i + 1];
    for (int e = beg + lane; e < end; e += 32) {
        const int32_t c = col_in[e];
        int rank = 0;
        for (int f = beg; f < end; ++f) rank += (col_in[f] < c);
        col_out[beg + rank] = c;
        val_out[2 * (beg + rank)] = val_in[2 * e];
        val_out[2 * (beg + rank) + 1] = val_in[2 * e + 1];
    }
}

// y = A x, complex128 CSR, one warp per row; optional done flag (Davidson no-op)
__global__ void csr_matvec_c128_kernel(const int* __restrict__ done, int64_t d,
                                       const int32_t* __restrict__ row_ptr,
                                       const int32_t* __restrict__ col, const double2* __restrict__ val,
                                       const double2* __restrict__ x, double2* __restrict__ y) {
    if (done && *done) return;
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= d) return;
    double re = 0.0, im = 0.0;
    for (int e = row_ptr[i] + lane; e < row_ptr[i + 1]; e += 32) {
        const double2 a = val[e];
        const double2 b = x[col[e]];
        re += a.x * b.x - a.y * b.y;
        im += a.x * b.y + a.y * b.x;
    }
    re = warp_sum(re);
    im = warp_sum(im);
    if (lane == 0) y[i] = make_double2(re, im);
}

// diagonal of a CSR matrix, duplicated for the (re, im) embedding:  out[2i] = out[2i+1] = Re A_ii
__global__ void csr_diag_embed_kernel(int64_t d, const int32_t* __restrict__ row_ptr,
                                      const int32_t* __restrict__ col, const double2* __restrict__ val,
                                      double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    double v = 0.0;
    for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e)
        if (col[e] == i) v = val[e].x;
    out[2 * i] = v;
    out[2 * i + 1] = v;
}

// Gershgorin data of a Hermitian CSR matrix: diag[i] = Re A_ii, lower[i] = diag[i] - sum_{j != i} |A_ij|
__global__ void csr_gershgorin_kernel(int64_t d, const int32_t* __restrict__ row_ptr,
                                      const int32_t* __restrict__ col, const double2* __restrict__ val,
                                      double* __restrict__ diag, double* __restrict__ lower) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= d) return;
    double dv = 0.0, r = 0.0;
    for (int e = row_ptr[i] + lane; e < row_ptr[i + 1]; e += 32) {
        const double2 a = val[e];
        if (col[e] == i) dv += a.x; else r += hypot(a.x, a.y);
    }
    dv = warp_sum(dv);
    r = warp_sum(r);
    if (lane == 0) {
        diag[i] = dv;
        lower[i] = dv - r;
    }
}

// ---- connected components of the projected operator (block structure of the subspace) -------------------
// label[i] = smallest row index of the component of row i.  Min-label hooking over the stored entries
// (the matrix is Hermitian, so the structure is symmetric) alternated with full pointer jumping; the host
// loop stops after the first hooking pass that changes nothing.
__global__ void cc_init_kernel(int64_t d, int32_t* __restrict__ label) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < d) label[i] = (int32_t)i;
}

__global__ void cc_hook_kernel(int64_t d, const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                               int32_t* label, int32_t* __restrict__ changed) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= d) return;
    const int32_t li = label[i];
    int32_t m = li;
    for (int e = row_ptr[i] + lane; e < row_ptr[i + 1]; e += 32) m = min(m, label[col[e]]);
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m < li) {
        atomicMin(label + li, m);   // hook the old representative too: whole trees move at once
        atomicMin(label + i, m);
        *changed = 1;
    }
}

__global__ void cc_jump_kernel(int64_t d, int32_t* label) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    int32_t l = label[i];
    int32_t ll = label[l];
    while (ll != l) {   // label[x] <= x: the chain is strictly decreasing and ends at a fixed point
        l = ll;
        ll = label[l];
    }
    label[i] = l;
}

// order-preserving map double -> uint64 (for atomicMin)
__device__ __forceinline__ unsigned long long enc_f64(double x) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dec_f64(unsigned long long e) {
    const unsigned long long u = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    return __longlong_as_double((long long)u);
}

// per component (stored at its representative row): size, min Gershgorin bound, min diagonal
__global__ void cc_stats_kernel(int64_t d, const int32_t* __restrict__ label, const double* __restrict__ diag,
                                const double* __restrict__ lower, int32_t* __restrict__ size,
                                unsigned long long* __restrict__ min_lower,
                                unsigned long long* __restrict__ min_diag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    const int32_t r = label[i];
    atomicAdd(size + r, 1);
    atomicMin(min_lower + r, enc_f64(lower[i]));
    atomicMin(min_diag + r, enc_f64(diag[i]));
}

// first row (smallest index) attaining the component's minimum diagonal
__global__ void cc_argmin_kernel(int64_t d, const int32_t* __restrict__ label, const double* __restrict__ diag,
                                 const unsigned long long* __restrict__ min_diag, int32_t* __restrict__ arg_row) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    const int32_t r = label[i];
    if (enc_f64(diag[i]) == min_diag[r]) atomicMin(arg_row + r, (int32_t)i);
}

// head[0] = number of multi-row components, head[1] = number of single-row components,
// head[2] = row of the lowest single-row component (or -1); best[0] = its (encoded) diagonal.
// Multi-row components are appended to `rec` (order not deterministic: the host sorts them).
__global__ void cc_summary_kernel(int64_t d, const int32_t* __restrict__ label, const int32_t* __restrict__ size,
                                  const unsigned long long* __restrict__ min_lower,
                                  const unsigned long long* __restrict__ min_diag,
                                  const int32_t* __restrict__ arg_row, int32_t* __restrict__ head,
                                  unsigned long long* __restrict__ best, sqd_component* __restrict__ rec,
                                  int64_t cap) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d || label[i] != (int32_t)i) return;
    if (size[i] == 1) {
        atomicAdd(head + 1, 1);
        atomicMin(best, min_diag[i]);
    } else {
        const int32_t slot = atomicAdd(head + 0, 1);
        if (slot < cap) {
            sqd_component c;
            c.root = (int32_t)i;
            c.size = size[i];
            c.row_min_diag = arg_row[i];
            c.pad = 0;
            c.lower = dec_f64(min_lower[i]);
            c.diag = dec_f64(min_diag[i]);
            rec[slot] = c;
        }
    }
}

__global__ void cc_best_single_kernel(int64_t d, const int32_t* __restrict__ label, const int32_t* __restrict__ size,
                                      const unsigned long long* __restrict__ min_diag,
                                      const unsigned long long* __restrict__ best, int32_t* __restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d || label[i] != (int32_t)i || size[i] != 1) return;
    if (min_diag[i] == *best) atomicMin(head + 2, (int32_t)i);
}

// start vector of a Davidson run confined to one component: e_row + scale * noise on the component's rows,
// noise_i = 2 frac((i + 1) / golden ratio) - 1 (deterministic, no generator state)
__global__ void cc_start_kernel(int64_t d, const int32_t* __restrict__ label, int32_t root, int32_t row,
                                double scale, double2* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d) return;
    double v = 0.0;
    if (label[i] == root) {
        const double t = (double)(i + 1) * 0.6180339887498949;
        v = scale * (2.0 * (t - floor(t)) - 1.0);
    }
    if (i == row) v = 1.0;
    out[i] = make_double2(v, 0.0);
}

int csr_matvec_flag(const int* d_done, int64_t d, const int32_t* row_ptr, const int32_t* col,
                    const double* val, const double* x, double* y, cudaStream_t st) {
    const int wpb = 8;
    csr_matvec_c128_kernel<<<(unsigned)((d + wpb - 1) / wpb), wpb * 32, 0, st>>>(
        d_done, d, row_ptr, col, (const double2*)val, (const double2*)x, (double2*)y);
    return check_launch("csr_matvec_c128_kernel");
}

int csr_diag_embed(int64_t d, const int32_t* row_ptr, const int32_t* col, const double* val,
                   double* out, cudaStream_t st) {
    csr_diag_embed_kernel<<<(unsigned)((d + 255) / 256), 256, 0, st>>>(d, row_ptr, col,
                                                                      (const double2*)val, out);
    return check_launch("csr_diag_embed_kernel");
}

}  // namespace sqd

using namespace sqd;

extern "C" {

int sqd_bits_to_keys(const uint8_t* d_bits, int64_t n, int nbits, int64_t* d_keys, void* stream) {
    SQD_REQUIRE(nbits > 0 && nbits <= 63, "sqd_bits_to_keys: nbits=%d must be in [1, 63]", nbits);
    if (n == 0) return 0;
    const int wpb = 8;
    bits_to_keys_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        d_bits, n, nbits, d_keys);
    return check_launch("bits_to_keys_kernel");
}

int64_t sqd_key_table_capacity(int64_t d) {
    if (d < 0) return -1;
    int64_t cap = 1024;
    const int64_t want = d <= (1LL << 24) ? 4 * d : 2 * d;   // load <= 1/4 (<= 1/2 for very large subspaces)
    while (cap < want) cap <<= 1;
    return cap;
}

int64_t sqd_key_table_bytes(int64_t d) {
    const int64_t cap = sqd_key_table_capacity(d);
    return cap < 0 ? -1 : cap * 12;
}

int sqd_key_table_build(const int64_t* d_keys, int64_t d, void* d_table, int64_t table_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t cap = sqd_key_table_capacity(d);
    SQD_REQUIRE(d >= 0 && d < 2147483647LL, "sqd_key_table_build: d=%lld out of range", (long long)d);
    SQD_REQUIRE(table_bytes >= cap * 12, "sqd_key_table_build: table too small");
    int64_t* slot_key = (int64_t*)d_table;
    int32_t* slot_row = (int32_t*)(slot_key + cap);
    SQD_CUDA_OK(cudaMemsetAsync(slot_row, 0x7f, (size_t)cap * 4, st));
    key_table_clear_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(slot_key, cap);
    if (d > 0)
        key_table_build_kernel<<<(unsigned)((d + 255) / 256), 256, 0, st>>>(d_keys, d, slot_key, slot_row,
                                                                            (uint64_t)cap - 1);
    return check_launch("key_table kernels", d > 0 ? 2 : 1);
}

static KeyTable make_table(const void* d_table, int64_t d) {
    KeyTable t;
    const int64_t cap = sqd_key_table_capacity(d);
    t.slot_key = (const int64_t*)d_table;
    t.slot_row = (const int32_t*)(t.slot_key + cap);
    t.mask = (uint64_t)cap - 1;
    return t;
}

int sqd_pauli_connect(const int64_t* d_keys, int64_t d, const void* d_table, uint64_t xmask, uint64_t zmask,
                      int32_t* d_col, uint8_t* d_par, void* stream) {
    if (d == 0) return 0;
    pauli_connect_kernel<<<(unsigned)((d + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        d_keys, d, d_table ? make_table(d_table, d) : KeyTable{nullptr, nullptr, 0}, d_table != nullptr, xmask,
        zmask, d_col, d_par);
    return check_launch("pauli_connect_kernel");
}

int sqd_pauli_elements(const int64_t* d_keys, int64_t d, const void* d_table, uint64_t xmask, uint64_t zmask,
                       int ny, int64_t* d_col, double* d_amp, int32_t* d_n_missing, void* stream) {
    if (d == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SQD_CUDA_OK(cudaMemsetAsync(d_n_missing, 0, sizeof(int32_t), st));
    pauli_elements_kernel<<<(unsigned)((d + 255) / 256), 256, 0, st>>>(
        d_keys, d, d_table ? make_table(d_table, d) : KeyTable{nullptr, nullptr, 0}, d_table != nullptr, xmask,
        zmask, ny, d_col, (double2*)d_amp, d_n_missing);
    return check_launch("pauli_elements_kernel");
}

int sqd_pauli_diag_group(const int64_t* d_keys, int64_t d, const uint64_t* d_zmask, const int32_t* d_ny,
                         const double* d_coeff, int32_t t0, int32_t t1, double* d_out, void* stream) {
    if (d == 0) return 0;
    SQD_REQUIRE(t0 >= 0 && t1 >= t0, "sqd_pauli_diag_group: bad term range");
    pauli_diag_group_kernel<<<(unsigned)((d + 127) / 128), 128, 0, (cudaStream_t)stream>>>(d_keys, d, d_zmask, d_ny,
                                                                                         d_coeff, t0, t1, d_out);
    return check_launch("pauli_diag_group_kernel");
}

int sqd_pauli_project_count(const int64_t* d_keys, int64_t d, const void* d_table, const uint64_t* d_grp_xmask,
                            const int32_t* d_grp_ptr, int32_t n_groups, const uint64_t* d_zmask,
                            const int32_t* d_ny, const double* d_coeff, int32_t diag_group,
                            const double* d_diag_val, int32_t* d_row_nnz, void* stream) {
    if (d == 0) return 0;
    const int wpb = 8;
    const unsigned nb = (unsigned)((d + wpb - 1) / wpb);
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t dg = d_diag_val ? diag_group : -1;
    if (d_table)
        pauli_project_kernel<false, true><<<nb, wpb * 32, 0, st>>>(
            d_keys, d, make_table(d_table, d), d_grp_xmask, d_grp_ptr, n_groups, d_zmask, d_ny, d_coeff, nullptr,
            d_row_nnz, nullptr, nullptr, dg, d_diag_val);
    else
        pauli_project_kernel<false, false><<<nb, wpb * 32, 0, st>>>(
            d_keys, d, KeyTable{nullptr, nullptr, 0}, d_grp_xmask, d_grp_ptr, n_groups, d_zmask, d_ny, d_coeff,
            nullptr, d_row_nnz, nullptr, nullptr, dg, d_diag_val);
    return check_launch("pauli_project_kernel<count>");
}

int sqd_pauli_project_fill(const int64_t* d_keys, int64_t d, const void* d_table, const uint64_t* d_grp_xmask,
                           const int32_t* d_grp_ptr, int32_t n_groups, const uint64_t* d_zmask,
                           const int32_t* d_ny, const double* d_coeff, int32_t diag_group,
                           const double* d_diag_val, const int32_t* d_row_ptr,
                           int32_t* d_col_tmp, double* d_val_tmp, int32_t* d_col, double* d_val,
                           void* stream) {
    if (d == 0) return 0;
    const int wpb = 8;
    const unsigned nb = (unsigned)((d + wpb - 1) / wpb);
    cudaStream_t st = (cudaStream_t)stream;
    const int32_t dg = d_diag_val ? diag_group : -1;
    if (d_table)
        pauli_project_kernel<true, true><<<nb, wpb * 32, 0, st>>>(
            d_keys, d, make_table(d_table, d), d_grp_xmask, d_grp_ptr, n_groups, d_zmask, d_ny, d_coeff, d_row_ptr,
            nullptr, d_col_tmp, d_val_tmp, dg, d_diag_val);
    else
        pauli_project_kernel<true, false><<<nb, wpb * 32, 0, st>>>(
            d_keys, d, KeyTable{nullptr, nullptr, 0}, d_grp_xmask, d_grp_ptr, n_groups, d_zmask, d_ny, d_coeff,
            d_row_ptr, nullptr, d_col_tmp, d_val_tmp, dg, d_diag_val);
    if (check_launch("pauli_project_kernel<fill>")) return -2;
    csr_row_sort_kernel<<<nb, wpb * 32, 0, st>>>(d, d_row_ptr, d_col_tmp, d_val_tmp, d_col, d_val);
    return check_launch("csr_row_sort_kernel");
}

int sqd_csr_gershgorin(int64_t d, const int32_t* d_row_ptr, const int32_t* d_col, const double* d_val,
                       double* d_diag, double* d_lower, void* stream) {
    if (d == 0) return 0;
    csr_gershgorin_kernel<<<(unsigned)((d + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        d, d_row_ptr, d_col, (const double2*)d_val, d_diag, d_lower);
    return check_launch("csr_gershgorin_kernel");
}

int64_t sqd_csr_components_workspace_bytes(int64_t d) {
    if (d <= 0) return -1;
    // size (i32) + arg_row (i32) + min_lower (u64) + min_diag (u64) per row, + changed flag, head[4], best
    return (int64_t)(d * (4 + 4 + 8 + 8) + 256);
}

int sqd_csr_components(int64_t d, const int32_t* d_row_ptr, const int32_t* d_col, const double* d_diag,
                       const double* d_lower, int32_t* d_label, sqd_component* d_rec, int64_t rec_cap,
                       int32_t* h_head, double* h_best_single, void* d_workspace, int64_t ws_bytes,
                       void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(d > 0 && d < 2147483647LL, "sqd_csr_components: d=%lld out of range", (long long)d);
    SQD_REQUIRE(ws_bytes >= sqd_csr_components_workspace_bytes(d), "sqd_csr_components: workspace too small");
    char* p = (char*)d_workspace;
    unsigned long long* min_lower = (unsigned long long*)p;  p += 8 * d;
    unsigned long long* min_diag = (unsigned long long*)p;   p += 8 * d;
    unsigned long long* best = (unsigned long long*)p;       p += 64;
    int32_t* size = (int32_t*)p;                             p += 4 * d;
    int32_t* arg_row = (int32_t*)p;                          p += 4 * d;
    int32_t* head = (int32_t*)p;                             p += 64;
    int32_t* changed = (int32_t*)p;
    const unsigned tb = 256, nb1 = (unsigned)((d + tb - 1) / tb), nbw = (unsigned)((d + 7) / 8);
    cc_init_kernel<<<nb1, tb, 0, st>>>(d, d_label);
    if (check_launch("cc_init_kernel")) return -2;
    for (int it = 0;; ++it) {
        SQD_REQUIRE(it < 100000, "sqd_csr_components: label propagation did not settle");
        SQD_CUDA_OK(cudaMemsetAsync(changed, 0, sizeof(int32_t), st));
        cc_hook_kernel<<<nbw, 256, 0, st>>>(d, d_row_ptr, d_col, d_label, changed);
        cc_jump_kernel<<<nb1, tb, 0, st>>>(d, d_label);
        if (check_launch("cc_hook/jump kernels", 2)) return -2;
        int32_t h_changed = 0;
        if (read_back(&h_changed, changed, sizeof(int32_t), st)) return -2;
        if (!h_changed) break;
    }
    SQD_CUDA_OK(cudaMemsetAsync(min_lower, 0xff, 16 * (size_t)d + 64, st));      // min_lower, min_diag, best
    SQD_CUDA_OK(cudaMemsetAsync(size, 0, 4 * (size_t)d, st));
    SQD_CUDA_OK(cudaMemsetAsync(arg_row, 0x7f, 4 * (size_t)d, st));
    SQD_CUDA_OK(cudaMemsetAsync(head, 0, 64, st));
    SQD_CUDA_OK(cudaMemsetAsync(head + 2, 0x7f, 4, st));
    cc_stats_kernel<<<nb1, tb, 0, st>>>(d, d_label, d_diag, d_lower, size, min_lower, min_diag);
    cc_argmin_kernel<<<nb1, tb, 0, st>>>(d, d_label, d_diag, min_diag, arg_row);
    cc_summary_kernel<<<nb1, tb, 0, st>>>(d, d_label, size, min_lower, min_diag, arg_row, head, best, d_rec,
                                         rec_cap);
    cc_best_single_kernel<<<nb1, tb, 0, st>>>(d, d_label, size, min_diag, best, head);
    if (check_launch("cc summary kernels", 4)) return -2;
    struct { int32_t head[16]; } hh;
    unsigned long long h_best = 0;
    if (read_back(&hh, head, 64, st)) return -2;
    if (read_back(&h_best, best, 8, st)) return -2;
    h_head[0] = hh.head[0];
    h_head[1] = hh.head[1];
    h_head[2] = hh.head[1] > 0 ? hh.head[2] : -1;
    const unsigned long long u = (h_best >> 63) ? (h_best & 0x7fffffffffffffffull) : ~h_best;
    double bd;
    memcpy(&bd, &u, 8);
    *h_best_single = hh.head[1] > 0 ? bd : 0.0;
    return 0;
}

int sqd_csr_component_start(int64_t d, const int32_t* d_label, int32_t root, int32_t row, double scale,
                            double* d_start, void* stream) {
    SQD_REQUIRE(d > 0 && row >= 0 && row < d, "sqd_csr_component_start: bad row");
    cc_start_kernel<<<(unsigned)((d + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d, d_label, root, row, scale,
                                                                                 (double2*)d_start);
    return check_launch("cc_start_kernel");
}

int sqd_csr_matvec_c128(int64_t d, const int32_t* d_row_ptr, const int32_t* d_col,
                        const double* d_val, const double* d_x, double* d_y, void* stream) {
    if (d == 0) return 0;
    return csr_matvec_flag(nullptr, d, d_row_ptr, d_col, d_val, d_x, d_y, (cudaStream_t)stream);
}

}  // extern "C"

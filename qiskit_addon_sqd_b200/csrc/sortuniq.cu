// Sorted distinct keys with multiplicities, and the bit-level format conversions on either side of it.
//
// Replaces the numpy calls of the reference's sample-format functions:
//   counts.py:57-60    np.unpackbits(bit_array.array)[..., -num_bits:] -> np.unique(axis=0, return_counts=True)
//   qubit.py:147-164   sort_and_remove_duplicates (np.unique over the rows' integer values), and the same
//                      np.unique inside solve_qubit (qubit.py:66)
// A bitstring of up to 128 bits is one (hi, lo) pair of 64-bit words (column 0 = most significant bit), so the
// lexicographic row order numpy sorts by is the numeric order of the pair.  Sort = bitonic network: the steps
// whose partner distance fits a 2048-element tile run in shared memory (one launch per merge size), the wider
// ones are one pass over global memory each; then neighbours are compared, the heads of the runs of equal keys
// are ranked by a scan and their distances are the multiplicities.  Integer work only: bit-exact.
#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

constexpr int kSortTile = 2048;     // elements per shared-memory tile (two per thread)
constexpr int kSortThreads = kSortTile / 2;

template <bool WIDE>
__device__ __forceinline__ bool key_less(uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl) {
    if (WIDE) return ah < bh || (ah == bh && al < bl);
    return al < bl;
}

// all steps j = j_start, j_start/2, ..., 1 of merge size k inside tiles of kSortTile elements
template <bool WIDE>
__global__ void __launch_bounds__(kSortThreads)
bitonic_tile_kernel(uint64_t* __restrict__ hi, uint64_t* __restrict__ lo, int64_t k, int j_start, bool full_sort) {
    __shared__ uint64_t s_lo[kSortTile];
    __shared__ uint64_t s_hi[WIDE ? kSortTile : 1];
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
    const int t = threadIdx.x;
    for (int u = t; u < kSortTile; u += kSortThreads) {
        s_lo[u] = lo[base + u];
        if (WIDE) s_hi[u] = hi[base + u];
    }
    __syncthreads();
    // full_sort: every merge size 2 .. kSortTile in one launch (the first phase of the network)
    for (int64_t kk = full_sort ? 2 : k; kk <= k; kk <<= 1) {
        for (int j = full_sort ? (int)(kk >> 1) : j_start; j > 0; j >>= 1) {
            const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));   // lower index of this thread's pair
            const int l = i | j;
            const bool up = ((base + i) & kk) == 0;
            const uint64_t al = s_lo[i], bl = s_lo[l];
            const uint64_t ah = WIDE ? s_hi[i] : 0, bh = WIDE ? s_hi[l] : 0;
            if (key_less<WIDE>(bh, bl, ah, al) == up) {
                s_lo[i] = bl;
                s_lo[l] = al;
                if (WIDE) {
                    s_hi[i] = bh;
                    s_hi[l] = ah;
                }
            }
            __syncthreads();
        }
    }
    for (int u = t; u < kSortTile; u += kSortThreads) {
        lo[base + u] = s_lo[u];
        if (WIDE) hi[base + u] = s_hi[u];
    }
}

// one step (merge size k, partner distance j >= kSortTile) over global memory; one thread per pair
template <bool WIDE>
__global__ void bitonic_global_kernel(uint64_t* __restrict__ hi, uint64_t* __restrict__ lo, int64_t n_pairs,
                                      int64_t k, int64_t j) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_pairs) return;
    const int64_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
    const int64_t l = i | j;
    const bool up = (i & k) == 0;
    const uint64_t al = lo[i], bl = lo[l];
    const uint64_t ah = WIDE ? hi[i] : 0, bh = WIDE ? hi[l] : 0;
    if (key_less<WIDE>(bh, bl, ah, al) == up) {
        lo[i] = bl;
        lo[l] = al;
        if (WIDE) {
            hi[i] = bh;
            hi[l] = ah;
        }
    }
}

__global__ void sort_load_kernel(const uint64_t* __restrict__ in_hi, const uint64_t* __restrict__ in_lo, int64_t n,
                                 int64_t n_pad, uint64_t* __restrict__ hi, uint64_t* __restrict__ lo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    // pads are the largest key: after the sort the first n positions hold the input multiset
    lo[i] = i < n ? in_lo[i] : ~0ull;
    if (hi != nullptr) hi[i] = i < n ? in_hi[i] : ~0ull;
}

__global__ void run_head_kernel(const uint64_t* __restrict__ hi, const uint64_t* __restrict__ lo, int64_t n,
                                int* __restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || lo[i] != lo[i - 1] || (hi != nullptr && hi[i] != hi[i - 1])) ? 1 : 0;
}

__global__ void run_emit_kernel(const uint64_t* __restrict__ hi, const uint64_t* __restrict__ lo, int64_t n,
                                const int* __restrict__ head, const int* __restrict__ rank,
                                uint64_t* __restrict__ out_hi, uint64_t* __restrict__ out_lo,
                                int* __restrict__ first) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    const int r = rank[i];
    out_lo[r] = lo[i];
    if (out_hi != nullptr) out_hi[r] = hi != nullptr ? hi[i] : 0ull;
    first[r] = (int)i;
}

__global__ void run_count_kernel(const int* __restrict__ first, int n_unique, int64_t n, int32_t* __restrict__ count) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_unique) return;
    count[r] = (r + 1 < n_unique ? first[r + 1] : (int)n) - first[r];
}

// bit array rows (big-endian bytes, left-padded) -> (hi, lo); bits beyond num_bits are not part of the sample
__global__ void bit_array_pack_kernel(const uint8_t* __restrict__ bytes, int64_t n, int row_bytes, int num_bits,
                                      uint64_t* __restrict__ hi, uint64_t* __restrict__ lo) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    const uint8_t* src = bytes + row * row_bytes;
    uint64_t h = 0, l = 0;
    const int used = (num_bits + 7) / 8;           // trailing bytes that carry sample bits
    for (int k = row_bytes - used; k < row_bytes; ++k) {
        h = (h << 8) | (l >> 56);
        l = (l << 8) | src[k];
    }
    if (num_bits < 64) {
        l &= (1ull << num_bits) - 1ull;
        h = 0;
    } else if (num_bits < 128) {
        h &= num_bits == 64 ? 0ull : (1ull << (num_bits - 64)) - 1ull;
    }
    lo[row] = l;
    if (hi != nullptr) hi[row] = h;
}

// (hi, lo) -> bool rows of num_bits columns, column 0 = most significant bit
__global__ void keys_to_bits_kernel(const uint64_t* __restrict__ hi, const uint64_t* __restrict__ lo, int64_t n,
                                    int num_bits, uint8_t* __restrict__ bits) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * num_bits) return;
    const int64_t row = i / num_bits;
    const int pos = num_bits - 1 - (int)(i - row * num_bits);   // bit position, 0 = least significant
    const uint64_t w = pos >= 64 ? (hi != nullptr ? hi[row] : 0ull) : lo[row];
    bits[i] = (uint8_t)((w >> (pos & 63)) & 1ull);
}

// ---- carry-over of the SQD loop (fermion.py:607-631) ---------------------------------------------------------
// Rows / columns of the amplitude matrix that hold an entry with |c| >= threshold, and their marginal weights
// sum |c|^2 in numpy's summation order (np.sum over the contiguous axis = pairwise_sum: blocks of <= 128 with
// eight interleaved accumulators, halves split at a multiple of 8 above that), without FMA contraction -- the
// ranking by weight that follows on the host then sees the reference's weights bit for bit.
__global__ void carry_flag_kernel(const double* __restrict__ x, int na, int nb, int ldc, double thr,
                                  int* __restrict__ row_flag, int* __restrict__ col_flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)na * ldc) return;
    const int a = (int)(i / ldc), b = (int)(i - (int64_t)a * ldc);
    if (b < nb && fabs(x[i]) >= thr) {
        row_flag[a] = 1;   // benign race: every writer stores 1
        col_flag[b] = 1;
    }
}

__device__ __forceinline__ double sq_rn(double v) { return __dmul_rn(v, v); }

__device__ double numpy_sum_sq_leaf(const double* __restrict__ a, int64_t stride, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, sq_rn(a[i * stride]));
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = sq_rn(a[j * stride]);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], sq_rn(a[(i + j) * stride]));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, sq_rn(a[i * stride]));
    return res;
}

// numpy's recursive pairwise_sum of a[i]^2, i < n, with an explicit stack (depth <= log2(n / 128) + 2)
__device__ double numpy_sum_sq(const double* __restrict__ a, int64_t stride, int64_t n) {
    struct Frame {
        int64_t off, len;
        double left;
        int stage;
    };
    Frame st[32];
    int sp = 1;
    st[0] = Frame{0, n, 0.0, 0};
    double ret = 0.0;
    while (sp > 0) {
        Frame& f = st[sp - 1];
        if (f.len <= 128) {
            ret = numpy_sum_sq_leaf(a + f.off * stride, stride, (int)f.len);
            --sp;
            continue;
        }
        int64_t n2 = f.len / 2;
        n2 -= n2 % 8;
        if (f.stage == 0) {
            f.stage = 1;
            st[sp++] = Frame{f.off, n2, 0.0, 0};
        } else if (f.stage == 1) {
            f.left = ret;
            f.stage = 2;
            st[sp++] = Frame{f.off + n2, f.len - n2, 0.0, 0};
        } else {
            ret = __dadd_rn(f.left, ret);
            --sp;
        }
    }
    return ret;
}

__global__ void carry_weight_kernel(const double* __restrict__ x, int na, int nb, int ldc,
                                    const int* __restrict__ row_flag, const int* __restrict__ col_flag,
                                    double* __restrict__ row_w, double* __restrict__ col_w) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < na) {
        row_w[t] = row_flag[t] ? numpy_sum_sq(x + (size_t)t * ldc, 1, nb) : 0.0;
    } else if (t < na + nb) {
        const int b = t - na;
        col_w[b] = col_flag[b] ? numpy_sum_sq(x + b, ldc, na) : 0.0;
    }
}

static int64_t padded_size(int64_t n) {
    int64_t p = kSortTile;
    while (p < n) p <<= 1;
    return p;
}

template <bool WIDE>
static int bitonic_sort(uint64_t* hi, uint64_t* lo, int64_t n_pad, cudaStream_t st) {
    const unsigned tiles = (unsigned)(n_pad / kSortTile);
    bitonic_tile_kernel<WIDE><<<tiles, kSortThreads, 0, st>>>(hi, lo, kSortTile, kSortTile / 2, true);
    if (check_launch("bitonic_tile_kernel")) return -2;
    const int64_t n_pairs = n_pad / 2;
    for (int64_t k = 2 * (int64_t)kSortTile; k <= n_pad; k <<= 1) {
        for (int64_t j = k >> 1; j >= kSortTile; j >>= 1) {
            bitonic_global_kernel<WIDE><<<(unsigned)((n_pairs + 255) / 256), 256, 0, st>>>(hi, lo, n_pairs, k, j);
            if (check_launch("bitonic_global_kernel")) return -2;
        }
        bitonic_tile_kernel<WIDE><<<tiles, kSortThreads, 0, st>>>(hi, lo, k, kSortTile / 2, false);
        if (check_launch("bitonic_tile_kernel")) return -2;
    }
    return 0;
}

}  // namespace sqd

using namespace sqd;

extern "C" {

int64_t sqd_sort_unique_workspace_bytes(int64_t n) {
    if (n < 0 || n > (1ll << 30)) return -1;
    const int64_t n_pad = padded_size(n);
    // hi, lo (padded) | head, rank(+1), first(+1)
    return 2 * n_pad * 8 + (3 * (n + 2)) * 4 + 256;
}

int sqd_sort_unique(const uint64_t* d_hi, const uint64_t* d_lo, int64_t n, uint64_t* d_out_hi, uint64_t* d_out_lo,
                    int32_t* d_out_count, int64_t* h_n_unique, void* d_workspace, int64_t workspace_bytes,
                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(h_n_unique != nullptr, "sqd_sort_unique: h_n_unique is required");
    *h_n_unique = 0;
    const int64_t need = sqd_sort_unique_workspace_bytes(n);
    SQD_REQUIRE(need >= 0 && workspace_bytes >= need, "sqd_sort_unique: bad size (n=%lld) or workspace too small",
                (long long)n);
    if (n == 0) return 0;
    SQD_REQUIRE(d_lo != nullptr && d_out_lo != nullptr, "sqd_sort_unique: missing key buffers");
    const bool wide = d_hi != nullptr;
    const int64_t n_pad = padded_size(n);
    uint64_t* lo = (uint64_t*)d_workspace;
    uint64_t* hi = lo + n_pad;
    int* head = (int*)(hi + n_pad);
    int* rank = head + (n + 2);
    int* first = rank + (n + 2);
    sort_load_kernel<<<(unsigned)((n_pad + 255) / 256), 256, 0, st>>>(d_hi, d_lo, n, n_pad, wide ? hi : nullptr, lo);
    if (check_launch("sort_load_kernel")) return -2;
    if (wide ? bitonic_sort<true>(hi, lo, n_pad, st) : bitonic_sort<false>(nullptr, lo, n_pad, st)) return -2;
    const unsigned nblk = (unsigned)((n + 255) / 256);
    run_head_kernel<<<nblk, 256, 0, st>>>(wide ? hi : nullptr, lo, n, head);
    if (check_launch("run_head_kernel")) return -2;
    int n_unique = 0;
    if (sqd_exclusive_scan(head, rank, (int)n, &n_unique, stream)) return -2;
    SQD_REQUIRE(n_unique > 0 && n_unique <= n, "sqd_sort_unique: inconsistent run count %d", n_unique);
    run_emit_kernel<<<nblk, 256, 0, st>>>(wide ? hi : nullptr, lo, n, head, rank, d_out_hi, d_out_lo, first);
    if (check_launch("run_emit_kernel")) return -2;
    if (d_out_count != nullptr) {
        run_count_kernel<<<(unsigned)((n_unique + 255) / 256), 256, 0, st>>>(first, n_unique, n, d_out_count);
        if (check_launch("run_count_kernel")) return -2;
    }
    *h_n_unique = n_unique;
    return 0;
}

int sqd_bit_array_pack(const uint8_t* d_bytes, int64_t n, int row_bytes, int num_bits, uint64_t* d_hi,
                       uint64_t* d_lo, void* stream) {
    SQD_REQUIRE(n >= 0 && num_bits >= 1 && num_bits <= 128 && row_bytes >= (num_bits + 7) / 8,
                "sqd_bit_array_pack: need 1 <= num_bits <= 128 and rows of at least ceil(num_bits/8) bytes "
                "(got num_bits=%d, row_bytes=%d)", num_bits, row_bytes);
    SQD_REQUIRE(d_lo != nullptr && (num_bits <= 64 || d_hi != nullptr), "sqd_bit_array_pack: missing output");
    if (n == 0) return 0;
    bit_array_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_bytes, n, row_bytes,
                                                                                          num_bits, d_hi, d_lo);
    return check_launch("bit_array_pack_kernel");
}

int sqd_keys_to_bits(const uint64_t* d_hi, const uint64_t* d_lo, int64_t n, int num_bits, uint8_t* d_bits,
                     void* stream) {
    SQD_REQUIRE(n >= 0 && num_bits >= 1 && num_bits <= 128 && (num_bits <= 64 || d_hi != nullptr),
                "sqd_keys_to_bits: need 1 <= num_bits <= 128 (got %d)", num_bits);
    if (n == 0) return 0;
    const int64_t total = n * num_bits;
    keys_to_bits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_hi, d_lo, n, num_bits,
                                                                                            d_bits);
    return check_launch("keys_to_bits_kernel");
}

int sqd_carryover(const double* d_x, int na, int nb, int ldc, double threshold, int* d_row_flag, int* d_col_flag,
                  double* d_row_weight, double* d_col_weight, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    SQD_REQUIRE(na > 0 && nb > 0 && ldc >= nb, "sqd_carryover: bad shape (na=%d, nb=%d, ldc=%d)", na, nb, ldc);
    SQD_REQUIRE(d_x != nullptr && d_row_flag != nullptr && d_col_flag != nullptr && d_row_weight != nullptr &&
                    d_col_weight != nullptr, "sqd_carryover: missing buffers");
    SQD_CUDA_OK(cudaMemsetAsync(d_row_flag, 0, (size_t)na * sizeof(int), st));
    SQD_CUDA_OK(cudaMemsetAsync(d_col_flag, 0, (size_t)nb * sizeof(int), st));
    const int64_t n = (int64_t)na * ldc;
    carry_flag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_x, na, nb, ldc, threshold, d_row_flag,
                                                                    d_col_flag);
    if (check_launch("carry_flag_kernel")) return -2;
    carry_weight_kernel<<<(unsigned)((na + nb + 63) / 64), 64, 0, st>>>(d_x, na, nb, ldc, d_row_flag, d_col_flag,
                                                                        d_row_weight, d_col_weight);
    return check_launch("carry_weight_kernel");
}

}  // extern "C"

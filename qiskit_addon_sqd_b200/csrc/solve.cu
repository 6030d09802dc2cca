// One-call ground state of the Hamiltonian projected on a product subspace A x B.
//
// Replaces the whole body of qiskit_addon_sqd/fermion.py:solve_sci (:711-740) below the Python
// signature: kernel_fixed_space (link tables, hdiag, initial guess, Davidson), make_rdm1s diagonals,
// the energy of the bare Hamiltonian, spin_square and -- on request -- make_rdm1 / make_rdm2.
//
// Why one call: a batch of K subspaces is solved by K host threads (one CUDA stream each).  Driving the
// ~40 set-up steps of one solve from Python serialises the K threads on the interpreter lock and leaves
// every stream idle most of the time; here the host thread stays inside C from the first kernel to the
// last read-back.  Scratch memory comes from the device's stream-ordered pool (cudaMallocAsync) and is
// returned before the call ends; results are written to caller-owned buffers.
#include <stdlib.h>

#include <vector>

#include "common.cuh"
#include "../../include/sqd_b200.h"

namespace sqd {

// Scratch of one solve: slabs from the device's stream-ordered pool, carved with a bump pointer.  A solve
// makes ~50 requests; one cudaMallocAsync each costs ~15 us of host time when eight solver threads
// contend for the allocator, a slab brings that down to a handful of calls.
struct Pool {
    static constexpr size_t kSlab = (size_t)16 << 20;
    cudaStream_t st;
    std::vector<void*> ptrs;
    char* cur = nullptr;
    size_t left = 0;
    bool failed = false;
    explicit Pool(cudaStream_t s) : st(s) {}
    template <typename T>
    T* get(size_t n) {
        const size_t bytes = (((n > 0 ? n : 1) * sizeof(T)) + 255) & ~(size_t)255;
        if (bytes > left) {
            const size_t slab = bytes > kSlab / 4 ? bytes : kSlab;  // large requests get their own block
            void* p = nullptr;
            if (cudaMallocAsync(&p, slab, st) != cudaSuccess) {
                set_error("sqd_solve_subspace: cudaMallocAsync of %zu bytes failed: %s", slab,
                          cudaGetErrorString(cudaGetLastError()));
                failed = true;
                return nullptr;
            }
            ptrs.push_back(p);
            if (slab == bytes && bytes > kSlab / 4) return (T*)p;  // dedicated block: keep the current slab
            cur = (char*)p;
            left = slab;
        }
        T* out = (T*)cur;
        cur += bytes;
        left -= bytes;
        return out;
    }
    ~Pool() {
        for (void* p : ptrs) cudaFreeAsync(p, st);
    }
};

// keep freed blocks in the pool instead of returning them to the driver at every synchronisation
static int configure_pool() {
    static bool done[64] = {};
    int dev = 0;
    SQD_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !done[dev]) {
        cudaMemPool_t pool;
        SQD_CUDA_OK(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long thr = ~0ull;
        SQD_CUDA_OK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        done[dev] = true;
    }
    return 0;
}

static int read_back_large(void* h_dst, const void* d_src, size_t bytes, cudaStream_t st) {
    const size_t step = 8 * 1024;
    for (size_t o = 0; o < bytes; o += step) {
        const size_t k = bytes - o < step ? bytes - o : step;
        if (read_back((char*)h_dst + o, (const char*)d_src + o, k, st)) return -2;
    }
    return 0;
}

struct TableBufs {
    int* n_single;
    int* n_total;
    int* row_ptr;
    uint32_t *col, *meta, *pack;
    double *val, *diag;
    int nnz;
};

static int table_count(Pool& P, const uint64_t* strs, int n, TableBufs* T) {
    T->n_single = P.get<int>(n);
    T->n_total = P.get<int>(n);
    T->row_ptr = P.get<int>(n + 1);
    if (P.failed) return -2;
    if (sqd_excitation_count(strs, n, T->n_single, T->n_total, P.st)) return -2;
    return sqd_exclusive_scan(T->n_total, T->row_ptr, n, nullptr, P.st);
}

static int table_fill(Pool& P, const uint64_t* strs, int n, int norb, const double* h, const double* g,
                      TableBufs* T, sqd_spin_table* out) {
    const size_t m = T->nnz > 0 ? T->nnz : 1;
    T->col = P.get<uint32_t>(m);
    T->meta = P.get<uint32_t>(m);
    T->pack = P.get<uint32_t>(m);
    T->val = P.get<double>(m);
    T->diag = P.get<double>(n);
    if (P.failed) return -2;
    if (sqd_excitation_fill(strs, n, norb, h, g, T->row_ptr, T->n_single, T->col, T->val, T->meta, T->pack,
                            T->diag, P.st))
        return -2;
    *out = sqd_spin_table{n, strs, T->row_ptr, T->n_single, T->col, T->val, T->meta, T->pack};
    return 0;
}

struct SellBufs {
    int *perm, *len, *slice_ptr;
    uint32_t* pack;
    double* val;
    int cap;
};

static int sell_launch(Pool& P, const sqd_spin_table& T, int64_t nnz, int mode, const int* long_idx,
                       SellBufs* S) {
    const int n = T.n;
    S->cap = (int)(nnz + 36 * (int64_t)n + 256);
    S->perm = P.get<int>(n);
    S->len = P.get<int>(n);
    S->slice_ptr = P.get<int>((n + 31) / 32 + 1);
    S->pack = P.get<uint32_t>(S->cap);
    S->val = mode == 1 ? P.get<double>(S->cap) : nullptr;
    if (P.failed) return -2;
    return sqd_sell_build(&T, mode, long_idx, S->cap, S->perm, S->len, S->slice_ptr, S->pack, S->val, P.st);
}

// gab, Wa, Wb and diag of one operator over tables / plan / SELL copies that are already built
static int operator_build(Pool& P, const sqd_solve_params* prm, const sqd_operator& base, int mode,
                          double shift, double diag_const, bool same_spin, bool with_w, const double* da,
                          const double* db, sqd_operator* out) {
    const int norb = base.norb, na = base.a.n, nb = base.b.n, ldc = base.ldc, ldg = base.ldg;
    const int n2 = norb * norb;
    double* gab = P.get<double>((size_t)n2 * ldg);
    double* Wa = with_w ? P.get<double>((size_t)na * ldg) : nullptr;
    double* Wb = P.get<double>((size_t)n2 * ldc);
    double* diag = P.get<double>((size_t)na * ldc);
    if (P.failed) return -2;
    if (sqd_make_gab(mode == 0 ? prm->d_g : nullptr, norb, shift, mode, gab, ldg, P.st)) return -2;
    if (sqd_opposite_spin_tables(base.a.strs, na, base.b.strs, nb, norb, gab, ldg, same_spin ? da : nullptr,
                                 same_spin ? db : nullptr, diag_const, same_spin ? 1e300 : 0.0, Wa, Wb, diag,
                                 ldc, P.st))
        return -2;
    *out = base;
    out->diag = diag;
    out->gab = gab;
    out->Wa = Wa;
    out->Wb = with_w ? Wb : nullptr;
    out->use_same_spin = same_spin ? 1 : 0;
    return 0;
}

__global__ void rdm2_spin_sum_kernel(double* __restrict__ aa, const double* __restrict__ ab,
                                     const double* __restrict__ bb, int norb) {
    // pyscf make_rdm2: dm2aa + dm2bb + dm2ab + dm2ab.transpose(2,3,0,1), accumulated into the aa block
    const int64_t n2 = (int64_t)norb * norb, n4 = n2 * n2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pq = i / n2, rs = i % n2;
        aa[i] = aa[i] + bb[i] + ab[i] + ab[rs * n2 + pq];
    }
}

__global__ void rdm1_spin_sum_kernel(const double* __restrict__ dm1, int norb, double* __restrict__ out) {
    // sqd_rdm1s stores <p+ q>; pyscf's make_rdm1 is dm1[p,q] = <q+ p>: transpose while adding the spins
    const int n2 = norb * norb;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        const int p = i / norb, q = i % norb;
        out[i] = dm1[q * norb + p] + dm1[n2 + q * norb + p];
    }
}

}  // namespace sqd

using namespace sqd;

extern "C" {

int sqd_solve_subspace(const sqd_solve_params* prm, double* d_x, double* d_rdm1, double* d_rdm2,
                       sqd_solve_result* h_res, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int norb = prm->norb, na = prm->na, nb = prm->nb;
    SQD_REQUIRE(norb >= 1 && norb <= 64, "sqd_solve_subspace: norb must be in [1, 64] (got %d)", norb);
    SQD_REQUIRE(na > 0 && nb > 0, "sqd_solve_subspace: empty string list");
    SQD_REQUIRE(prm->d_h != nullptr && prm->d_g != nullptr && d_x != nullptr && h_res != nullptr,
                "sqd_solve_subspace: missing integrals or output buffers");
    if (configure_pool()) return -2;
    const int ldc = (nb + 1) / 2 * 2;
    const int ldg = (norb * norb + 2) / 2 * 2;
    const int64_t n = (int64_t)na * ldc;
    const bool same = prm->d_strs_b == prm->d_strs_a && na == nb;
    Pool P(st);

    // ---- excitation tables (one synchronisation: the two entry counts) ----
    TableBufs Ta{}, Tb{};
    if (table_count(P, prm->d_strs_a, na, &Ta)) return -2;
    if (!same && table_count(P, prm->d_strs_b, nb, &Tb)) return -2;
    int* pair = P.get<int>(2);
    if (P.failed) return -2;
    SQD_CUDA_OK(cudaMemcpyAsync(pair, Ta.row_ptr + na, sizeof(int), cudaMemcpyDeviceToDevice, st));
    SQD_CUDA_OK(cudaMemcpyAsync(pair + 1, (same ? Ta.row_ptr + na : Tb.row_ptr + nb), sizeof(int),
                                cudaMemcpyDeviceToDevice, st));
    int h_pair[2];
    if (read_back(h_pair, pair, sizeof(h_pair), st)) return -2;
    Ta.nnz = h_pair[0];
    Tb.nnz = h_pair[1];
    // the tables are indexed with 32-bit integers (the scan reports a total beyond INT_MAX as -1), and the
    // SELL copies of the v1 kernels need nnz + 36 n + 256 entries
    SQD_REQUIRE(Ta.nnz >= 0 && Tb.nnz >= 0 && (int64_t)Ta.nnz + 36ll * na + 256 < 2147483647ll &&
                    (int64_t)Tb.nnz + 36ll * nb + 256 < 2147483647ll,
                "sqd_solve_subspace: the excitation tables of this subspace have more than 2^31 entries "
                "(na=%d, nb=%d): beyond the 32-bit table index of this library", na, nb);
    sqd_spin_table ta{}, tb{};
    if (table_fill(P, prm->d_strs_a, na, norb, prm->d_h, prm->d_g, &Ta, &ta)) return -2;
    if (same) {
        tb = ta;
        Tb = Ta;
    } else if (table_fill(P, prm->d_strs_b, nb, norb, prm->d_h, prm->d_g, &Tb, &tb)) {
        return -2;
    }

    // ---- work plan of the sigma build (one synchronisation: its sizes) ----
    // v2 kernels (fermion_sigma2.cu) when the tables are dense enough, else the v1 plan + SELL copies
    bool use_v2 = prm->sigma_path == 2 ||
                  (prm->sigma_path == 0 && sqd_sigma_v2_recommended(na, nb, Ta.nnz, Tb.nnz) != 0);
    const int v2_lmax = prm->v2_lmax == 8 ? 8 : 16;
    const int v2_ipc = prm->v2_items_per_chunk > 0 && prm->v2_items_per_chunk <= 32 ? prm->v2_items_per_chunk : 8;
    const int nsl = (nb + 31) / 32;
    int* counts = P.get<int>(8 + SQD_V2_COUNTS);
    if (P.failed) return -2;
    int* chunk[4] = {};
    int* split[3] = {};
    int *long_idx = nullptr, *long_cols = nullptr;
    SellBufs S0{}, S1{};
    auto build_v1_plan = [&]() -> int {
        const int cost = prm->cost_per_chunk > 0 ? prm->cost_per_chunk : 256;
        const int long_thr = prm->long_threshold > 0 ? prm->long_threshold : 64;
        const int max_chunks = 2 * na + (int)((16 * (int64_t)(Ta.nnz > 0 ? Ta.nnz : 1)) / cost) + 1;
        for (auto& c : chunk) c = P.get<int>(max_chunks);
        for (auto& c : split) c = P.get<int>(na);
        long_idx = P.get<int>(nb);
        long_cols = P.get<int>(SQD_MAX_LONG_COLUMNS);
        if (P.failed) return -2;
        if (sqd_sigma_plan_build(&ta, &tb, cost, long_thr, max_chunks, chunk[0], chunk[1], chunk[2], chunk[3],
                                 split[0], split[1], split[2], long_idx, long_cols, counts, nullptr, st))
            return -2;
        if (sell_launch(P, tb, Tb.nnz, 0, long_idx, &S0)) return -2;
        if (sell_launch(P, tb, Tb.nnz, 1, nullptr, &S1)) return -2;
        SQD_CUDA_OK(cudaMemcpyAsync(counts + 5, S0.slice_ptr + nsl, sizeof(int), cudaMemcpyDeviceToDevice, st));
        SQD_CUDA_OK(cudaMemcpyAsync(counts + 6, S1.slice_ptr + nsl, sizeof(int), cudaMemcpyDeviceToDevice, st));
        return 0;
    };
    void* v2_plan = nullptr;
    int64_t v2_plan_bytes = 0;
    if (use_v2) {
        v2_plan_bytes = sqd_sigma_v2_plan_bytes(na, nb, Ta.nnz, Tb.nnz, v2_lmax, v2_ipc);
        v2_plan = v2_plan_bytes > 0 ? P.get<char>((size_t)v2_plan_bytes) : nullptr;
        if (P.failed) return -2;
        if (v2_plan == nullptr) {
            use_v2 = false;
        } else {
            if (sqd_sigma_v2_plan(&ta, &tb, norb, Ta.nnz, Tb.nnz, v2_lmax, v2_ipc, v2_plan, v2_plan_bytes,
                                  nullptr, st))
                return -2;
            SQD_CUDA_OK(cudaMemcpyAsync(
                counts + 8, sqd_sigma_v2_counts_ptr(v2_plan, na, nb, Ta.nnz, Tb.nnz, v2_lmax, v2_ipc),
                SQD_V2_COUNTS * sizeof(int), cudaMemcpyDeviceToDevice, st));
        }
    }
    // neither staged path takes rows this long (or the caller asked for it): the wide kernel, no plan
    bool use_wide = prm->sigma_path == 3 || !sqd_sigma_v1_supported(ldc, ldg);
    if (!use_v2 && !use_wide && build_v1_plan()) return -2;

    // ---- operators: integrals, W tables and diagonals do not depend on the plan, so their kernels (and
    // the start vector) are enqueued BEFORE the host waits for the plan / SELL sizes ----
    sqd_operator base{};
    base.a = ta;
    base.b = tb;
    base.norb = norb;
    base.ldc = ldc;
    base.ldg = ldg;
    base.throughput_mode = prm->throughput_mode;
    const int n_alpha = prm->n_alpha, n_beta = prm->n_beta;
    const double sz = 0.5 * (n_alpha > n_beta ? n_alpha - n_beta : n_beta - n_alpha);
    const double szz = 0.5 * (n_alpha - n_beta);
    const bool have_ss = prm->penalty != 0;
    const bool linear = have_ss && prm->spin_sq < sz * (sz + 1.0) + 0.1;  // pyscf fix_spin_ branches
    const bool quad = have_ss && !linear;
    const double lin_shift = linear ? prm->shift : 0.0;
    sqd_operator ham{}, s2op{};
    if (operator_build(P, prm, base, 0, lin_shift,
                       lin_shift * (szz * (szz + 1.0) + n_beta - prm->spin_sq), true, true, Ta.diag, Tb.diag,
                       &ham))
        return -2;
    const bool need_s2 = have_ss || prm->want_spin;
    if (need_s2 && operator_build(P, prm, base, 1, 0.0, szz * (szz + 1.0) + n_beta, false, false, nullptr,
                                  nullptr, &s2op))
        return -2;
    const int M = prm->max_space < 2 ? 2 : (prm->max_space > SQD_MAX_SPACE ? SQD_MAX_SPACE : prm->max_space);
    const int64_t ws_bytes = sqd_davidson_workspace_bytes(na, ldc, M);
    void* ws = P.get<char>((size_t)ws_bytes);
    double* x0 = P.get<double>(n);
    double* scratch = P.get<double>(4096);
    if (P.failed) return -2;
    if (prm->d_ci0 != nullptr) {
        SQD_CUDA_OK(cudaMemsetAsync(x0, 0, n * sizeof(double), st));
        SQD_CUDA_OK(cudaMemcpy2DAsync(x0, (size_t)ldc * sizeof(double), prm->d_ci0, (size_t)nb * sizeof(double),
                                      (size_t)nb * sizeof(double), na, cudaMemcpyDeviceToDevice, st));
    } else if (sqd_init_guess(ham.diag, na, nb, ldc, x0, scratch, st)) {
        return -2;
    }

    int hc[8 + SQD_V2_COUNTS];
    if (read_back(hc, counts, sizeof(hc), st)) return -2;  // hc[4], hc[7] unused
    if (use_v2 && hc[8 + 6] != 0) {
        // the v2 planner refused the shape (a beta string with thousands of links, or too many strings
        // for its single-CTA passes): the v1 kernels take over
        use_v2 = false;
        if (!use_wide) {
            if (build_v1_plan()) return -2;
            if (read_back(hc, counts, 8 * sizeof(int), st)) return -2;
        }
    }
    if (use_v2) {
        const int same_tables = same ? 1 : 0;
        const int64_t sb = sqd_sigma_v2_scratch_bytes(hc + 8, na, nb, ldc, 1, same_tables);
        SQD_REQUIRE(sb > 0, "sqd_solve_subspace: bad v2 scratch size");
        void* scratch2 = P.get<char>((size_t)sb);
        if (P.failed) return -2;
        sqd_sigma_v2 v2{};
        if (sqd_sigma_v2_finish(&ta, &tb, ldc, Ta.nnz, Tb.nnz, v2_lmax, v2_ipc, hc + 8, v2_plan, scratch2, sb,
                                1, &v2, st))
            return -2;
        for (sqd_operator* o : {&base, &ham, &s2op}) o->v2 = v2;
    } else if (use_wide) {
        for (sqd_operator* o : {&base, &ham, &s2op}) o->wide = 1;
    } else {
        double* part = P.get<double>((size_t)(hc[1] > 0 ? hc[1] : 1) * ldc);
        if (P.failed) return -2;
        const sqd_sigma_plan plan{hc[0], hc[1], hc[2], hc[3], chunk[0], chunk[1], chunk[2], chunk[3],
                                  split[0], split[1], split[2], long_idx, long_cols, part};
        const sqd_sell bd{nsl, hc[5], S0.perm, S0.len, S0.slice_ptr, S0.pack, nullptr};
        const sqd_sell bb{nsl, hc[6], S1.perm, S1.len, S1.slice_ptr, S1.pack, S1.val};
        for (sqd_operator* o : {&base, &ham, &s2op}) {
            o->plan = plan;
            o->bd = bd;
            o->bb = bb;
        }
    }
    h_res->sigma_path = use_v2 ? 2 : use_wide ? 3 : 1;

    // ---- Davidson ----
    sqd_davidson_params dp{};
    dp.max_space = M;
    dp.max_cycle = prm->max_cycle;
    dp.tol = prm->tol;
    dp.tol_residual = prm->tol_residual;
    dp.lindep = prm->lindep;
    dp.level_shift = prm->level_shift;
    dp.check_every = prm->check_every > 0 ? prm->check_every : 4;
    dp.profile = prm->profile;
    dp.single_stream_ritz = prm->throughput_mode;
    if (quad) {
        dp.ss_op = &s2op;
        dp.ss_shift = prm->shift;
        dp.ss_value = prm->spin_sq;
    }
    dp.nccl_comm = prm->nccl_comm;
    dp.row_begin = prm->row_begin;
    dp.row_end = prm->row_end;
    std::vector<int> h_ns, h_ptr, shard_bounds;
    if (prm->profile || (prm->nccl_comm != nullptr && prm->row_begin < 0)) {
        h_ns.resize(na);
        h_ptr.resize(na + 1);
        if (read_back_large(h_ns.data(), Ta.n_single, (size_t)na * sizeof(int), st)) return -2;
        if (read_back_large(h_ptr.data(), Ta.row_ptr, (size_t)(na + 1) * sizeof(int), st)) return -2;
    }
    if (prm->nccl_comm != nullptr && prm->row_begin < 0) {
        // per-row cost: alpha singles (gather loops), alpha doubles (row streaming), beta part (fixed);
        // the Hartree-Fock end of the string list is much heavier than the tail
        SQD_REQUIRE(prm->shard_world >= 1 && prm->shard_rank >= 0 && prm->shard_rank < prm->shard_world,
                    "sqd_solve_subspace: bad shard rank/world");
        std::vector<long long> cum(na + 1, 0);
        const long long beta = Tb.nnz / (nb > 0 ? nb : 1) > 1 ? Tb.nnz / nb : 1;
        for (int a = 0; a < na; ++a) {
            const long long ns = h_ns[a], nd = (long long)(h_ptr[a + 1] - h_ptr[a]) - ns;
            cum[a + 1] = cum[a] + 16 * (ns + 1) + nd + beta;
        }
        auto bound = [&](int r) {
            if (r <= 0) return 0;
            if (r >= prm->shard_world) return na;
            const double target = (double)cum[na] * r / prm->shard_world;
            int lo = 0;
            while (lo < na && (double)cum[lo] < target) ++lo;
            return lo;
        };
        dp.row_begin = bound(prm->shard_rank);
        dp.row_end = bound(prm->shard_rank + 1);
        // every rank derives the same bounds from the same tables: the blocks can be exchanged as they are
        static const int knob_gather = getenv("SQD_SHARD_GATHER") ? atoi(getenv("SQD_SHARD_GATHER")) : 1;
        if (knob_gather && prm->shard_world <= 64) {
            shard_bounds.resize(prm->shard_world + 1);
            for (int r = 0; r <= prm->shard_world; ++r) shard_bounds[r] = bound(r);
            dp.shard_bounds = shard_bounds.data();
            dp.shard_world = prm->shard_world;
        }
    }
    sqd_davidson_info info{};
    if (sqd_davidson(&ham, ham.diag, x0, d_x, ws, ws_bytes, &dp, &info, st)) return -2;

    // ---- expectation values: everything lands in one small buffer, read back once ----
    // res = [<x|x>, <x|Hx>, -, <x|S^2 x>, -,-,-,-, occ_a[norb], occ_b[norb]]
    double* res = P.get<double>(8 + 2 * norb);
    double* hx = P.get<double>(n);
    double* occ_scratch = P.get<double>(na + nb);
    if (P.failed) return -2;
    SQD_CUDA_OK(cudaMemsetAsync(res, 0, (8 + 2 * norb) * sizeof(double), st));
    if (sqd_sigma(&ham, d_x, hx, st)) return -2;
    if (sqd_dot(d_x, d_x, n, res, scratch, st)) return -2;
    if (sqd_dot(d_x, hx, n, res + 1, scratch, st)) return -2;
    if (need_s2) {
        if (sqd_sigma(&s2op, d_x, hx, st)) return -2;
        if (sqd_dot(d_x, hx, n, res + 3, scratch, st)) return -2;
    }
    if (sqd_occupancies(d_x, ta.strs, na, tb.strs, nb, ldc, norb, res + 8, occ_scratch, st)) return -2;
    if (sqd_fix_sign(d_x, n, scratch, st)) return -2;

    // ---- reduced density matrices (optional; spin-summed, pyscf conventions) ----
    if (d_rdm1 != nullptr || d_rdm2 != nullptr) {
        sqd_operator rop = base;  // the RDM kernels only read the tables
        // the 1-RDMs share the row-pair dot products of the 2-RDM pass when both are wanted
        double* dm1 = d_rdm1 != nullptr ? P.get<double>(2 * (size_t)norb * norb) : nullptr;
        if (d_rdm2 != nullptr) {
            const int64_t n4 = (int64_t)norb * norb * norb * norb;
            const int64_t wb = sqd_rdm2s_workspace_bytes(&rop, Ta.nnz, Tb.nnz);
            void* w2 = P.get<char>((size_t)wb);
            double* abbb = P.get<double>(2 * (size_t)n4);
            if (P.failed) return -2;
            if (sqd_rdm2s(&rop, d_x, Ta.nnz, Tb.nnz, d_rdm2, abbb, abbb + n4, dm1, w2, wb, st)) return -2;
            rdm2_spin_sum_kernel<<<kNumSMs * 4, 256, 0, st>>>(d_rdm2, abbb, abbb + n4, norb);
            if (check_launch("rdm2_spin_sum_kernel")) return -2;
        } else {
            double* w1 = (double*)P.get<char>((size_t)sqd_rdm1s_workspace_bytes(&rop) + 16);
            const int64_t mx = Ta.nnz > Tb.nnz ? Ta.nnz : Tb.nnz;
            double* dots = P.get<double>((size_t)(mx > 0 ? mx : 1));
            if (P.failed) return -2;
            if (sqd_rdm1s(&rop, d_x, Ta.nnz, Tb.nnz, dm1, w1, dots, st)) return -2;
        }
        if (d_rdm1 != nullptr) {
            rdm1_spin_sum_kernel<<<(norb * norb + 255) / 256, 256, 0, st>>>(dm1, norb, d_rdm1);
            if (check_launch("rdm1_spin_sum_kernel")) return -2;
        }
    }

    double hr[8 + 128];
    if (read_back(hr, res, (8 + 2 * norb) * sizeof(double), st)) return -2;
    const double xx = hr[0];
    const double e_pen = hr[1] / xx;
    const double s2 = need_s2 ? hr[3] / xx : 0.0;
    h_res->energy = linear ? e_pen - lin_shift * (s2 - prm->spin_sq) : e_pen;
    h_res->spin_square = s2;
    h_res->have_spin_square = need_s2 ? 1 : 0;
    for (int p = 0; p < norb; ++p) {
        h_res->occ_a[p] = hr[8 + p];
        h_res->occ_b[p] = hr[8 + norb + p];
    }
    h_res->info = info;
    h_res->nnz_a = Ta.nnz;
    h_res->nnz_b = Tb.nnz;
    h_res->ldc = ldc;
    // singles of each table (diagnostics for the byte model of the bench): not read back on the hot path
    h_res->singles_a = h_res->singles_b = -1;
    if (prm->profile) {
        long long sa = 0, sb = 0;
        for (int a = 0; a < na; ++a) sa += h_ns[a];
        if (same) {
            sb = sa;
        } else {
            std::vector<int> h_nb(nb);
            if (read_back_large(h_nb.data(), Tb.n_single, (size_t)nb * sizeof(int), st)) return -2;
            for (int b = 0; b < nb; ++b) sb += h_nb[b];
        }
        h_res->singles_a = sa;
        h_res->singles_b = sb;
    }
    return 0;
}

}  // extern "C"

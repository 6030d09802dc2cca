/* sqd_b200 -- C-ABI of the B200-native SQD hot path (libsqd_b200.so).
 *
 * Drop-in boundary for the three hot-path entry points of Qiskit/qiskit-addon-sqd (reference paths
 * are relative to /root/reference):
 *   - fermion.solve_fermion / solve_sci / solve_sci_batch      qiskit_addon_sqd/fermion.py:643-845
 *       (the work the reference delegates to pyscf: fci.selected_ci.kernel_fixed_space,
 *        make_rdm1s, make_rdm1/make_rdm2 energy, spin_square -- fermion.py:713-729, 803-830)
 *   - qubit.project_operator_to_subspace / matrix_elements_from_pauli / solve_qubit
 *                                                               qiskit_addon_sqd/qubit.py:29-300
 *   - configuration_recovery.recover_configurations             qiskit_addon_sqd/configuration_recovery.py:59-306
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer on the current CUDA device; h_* is a host pointer;
 *   - the library never allocates or frees memory that outlives a call: the caller (Python: torch
 *     tensors) owns all buffers and sizes them with the *_count / *_bytes queries;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value 0 = ok, < 0 = error; sqd_last_error() returns a thread-local message;
 *   - no function synchronises the stream unless its comment says so.
 */
#ifndef SQD_B200_H
#define SQD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQD_B200_VERSION 100
#define SQD_MAX_SPACE 32 /* largest Davidson subspace the device-side Rayleigh-Ritz supports */
#define SQD_MAX_LONG_COLUMNS 64 /* capacity of the long-column list (the kernels use the 4 longest) */

int sqd_version(void);
const char* sqd_last_error(void);
/* number of CUDA kernels this library has launched in this process (reset != 0: return and clear) */
long long sqd_launch_count(int reset);

/* ------------------------------------------------------------------------------------------ *
 * Determinant strings and excitation tables
 * replaces: pyscf selected_ci._all_linkstr_index (cre_des_linkstr / des_des_linkstr) reached from
 *           fermion.py:721,810; packing counts.py:186-201 (the string checks of fermion.py:1075-1097 are host
 *           code: a popcount per string on arrays the caller already holds on the host)
 * ------------------------------------------------------------------------------------------ */

/* Pack a bool bitstring matrix (n rows, nbits columns, 1 byte per bit, column 0 = most significant)
 * into two uint64 halves per row: left = columns [0, nbits/2) (beta), right = columns [nbits/2, nbits)
 * (alpha).  nbits/2 <= 64.  (counts.py:186-201, fermion.py:1026-1030) */
int sqd_pack_bitstrings(const uint8_t* d_bits, int64_t n, int nbits, uint64_t* d_left,
                        uint64_t* d_right, void* stream);

/* Pass 1: for each string i count the in-set single (xor-popcount 2) and single+double (2 or 4)
 * excitation partners.  d_n_single, d_n_total: int[n]. */
int sqd_excitation_count(const uint64_t* d_strs, int n, int* d_n_single, int* d_n_total,
                         void* stream);

/* Exclusive prefix sum of int[n] into int[n+1] (out[n] = total); h_total (pinned or pageable host
 * int) receives out[n] -- synchronises the stream when h_total != NULL. */
int sqd_exclusive_scan(const int* d_in, int* d_out, int n, int* h_total, void* stream);

/* Pass 2: fill the per-string excitation table, CSR over strings; inside each row the single
 * excitations come first (ascending partner index) followed by the doubles (ascending).
 *   d_col [nnz]  partner string index (source string s of <t|H|s>, row = target t)
 *   d_val [nnz]  same-spin Hamiltonian element <t| h + 1/2 (pq|rs) |s> (Slater-Condon rules)
 *   d_meta[nnz]  singles: (p*norb+q) | sign_bit<<31 for t = sign * a+_p a_q s ; doubles: 0
 *   d_pack[nnz]  singles: col | (p*norb+q)<<19 | sign_bit<<31 (one 4-byte load in the hot loop;
 *                needs n <= 2^19 and norb <= 64); doubles: col
 *   d_diag[n]    <s|H_same-spin|s>
 * d_h: double[norb*norb]; d_g: double[norb^4] chemist order (pq|rs), C-contiguous. */
int sqd_excitation_fill(const uint64_t* d_strs, int n, int norb, const double* d_h,
                        const double* d_g, const int* d_row_ptr, const int* d_n_single,
                        uint32_t* d_col, double* d_val, uint32_t* d_meta, uint32_t* d_pack,
                        double* d_diag, void* stream);

/* Opposite-spin tensor with pyscf fix_spin_'s linear penalty folded in:
 *   g_ab[pq*ldg + rs] = (pq|rs) - shift * delta_ps delta_qr       
 * mode 0: as above; mode 1: ignore d_g and write the S^2 tensor  -delta_ps delta_qr.
 * ldg = norb*norb rounded up to a multiple of 2 (16-byte rows for bulk copies). */
int sqd_make_gab(const double* d_g, int norb, double shift, int mode, double* d_gab, int ldg,
                 void* stream);

/* One-electron-like contractions of g_ab with the occupation of the other spin:
 *   Wa[a*ldg + rs] = sum_{p in a} g_ab[pp, rs]      (na x ldg)
 *   Wb[pq*ldc + b] = sum_{r in b} g_ab[pq, rr]      (norb^2 x ldc)
 * and the diagonal of the operator (ldc = nb rounded up to a multiple of 2; pad entries get
 * `pad_value`):
 *   diag[a*ldc + b] = da[a] + db[b] + sum_{p in a} Wb[pp, b] + diag_const
 * d_da / d_db may be NULL (treated as 0: S^2 operator). */
int sqd_opposite_spin_tables(const uint64_t* d_strs_a, int na, const uint64_t* d_strs_b, int nb,
                             int norb, const double* d_gab, int ldg, const double* d_da,
                             const double* d_db, double diag_const, double pad_value, double* d_Wa,
                             double* d_Wb, double* d_diag, int ldc, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * sigma-vector build and Davidson
 * replaces: pyscf selected_ci.contract_2e (SCIcontract_2e_aaaa / _bbaa), make_hdiag,
 *           lib.davidson1 inside kernel_fixed_space  (fermion.py:721-723, 810-818)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int n;                   /* number of strings */
    const uint64_t* strs;    /* [n] sorted ascending */
    const int* row_ptr;      /* [n+1] */
    const int* n_single;     /* [n] */
    const uint32_t* col;     /* [nnz] */
    const double* val;       /* [nnz] */
    const uint32_t* meta;    /* [nnz] */
    const uint32_t* pack;    /* [nnz] */
} sqd_spin_table;

/* Work decomposition of one sigma build (built once per subspace by sqd_sigma_plan_build).  Rows of the CI
 * matrix are cut into chunks of bounded cost (one CTA each); a row with several chunks accumulates
 * per-chunk partial vectors (`part`, n_slots x ldc doubles of caller-owned scratch) that are summed in
 * chunk order; beta strings with very long excitation lists ("long columns") are reduced by a warp. */
typedef struct {
    int n_chunks, n_slots, n_split, n_long;
    const int* chunk_row;      /* [n_chunks] alpha string index */
    const int* chunk_beg;      /* [n_chunks] first entry of the alpha table covered */
    const int* chunk_end;      /* [n_chunks] one past the last entry */
    const int* chunk_slot;     /* [n_chunks] row of `part`, or -1 when the chunk owns its whole row */
    const int* split_row;      /* [n_split] rows that have more than one chunk */
    const int* split_slot_beg; /* [n_split] */
    const int* split_n;        /* [n_split] */
    const int* long_idx;       /* [nb] index into long_cols or -1 */
    const int* long_cols;      /* [n_long] */
    double* part;              /* [n_slots * ldc] scratch, written by every sigma build */
} sqd_sigma_plan;

/* SELL-32 layout of a string list's excitation entries (built by sqd_sell_build): strings sorted by
 * list length (descending), 32 per slice, entry k of the string at sorted position 32*s+i stored at
 * slice_ptr[s] + 32*k + i -- a warp reads one coalesced line per step and runs a uniform trip count. */
typedef struct {
    int n_slices;            /* ceil(n / 32) */
    int n_entries;           /* slice_ptr[n_slices]: stored entries incl. padding */
    const int* perm;         /* [n] sorted position -> string index */
    const int* len;          /* [n] list length at each sorted position */
    const int* slice_ptr;    /* [n_slices + 1] */
    const uint32_t* pack;    /* same encoding as sqd_spin_table.pack */
    const double* val;       /* mode 1 only */
} sqd_sell;

/* ---- sigma build, second generation ("v2", csrc/fermion_sigma2.cu) ---------------------------
 * Three kernels per build, chosen when the in-set connectivity is dense enough:
 *   K1  opposite-spin part, grouped by SOURCE alpha string a': a thread owns one "virtual column" (at
 *       most `lmax` beta single excitations of one beta string), gathers x_j = sgn_j c[a', b'_j] into
 *       REGISTERS once per chunk of a' and then, for every alpha single excitation a' -> a (integral row
 *       g_ab[pq,:] staged by the bulk-copy engine through an mbarrier ring), does one shared-memory gather
 *       and one FMA per link; the result goes to P[row of (a <- a')][segment of the beta string].  Warps
 *       never synchronise with each other, only with the ring.
 *   K2  64x64 output tiles of the same-spin part as dense FP64 products  HaD C + C HbD^T, K split over
 *       CTAs into partial tiles.
 *   K3  sigma = partial tiles (fixed order) + diag*c + the P rows of the row's excitations (contiguous:
 *       P is ordered by TARGET string) + the sgn*Wb[pq,b]*c[a',b] terms.
 * Every element of sigma is produced in a fixed order (bit-reproducible).
 * All arrays are device pointers into caller-owned memory (sqd_sigma_v2_plan / sqd_sigma_v2_finish). */
typedef struct {
    int enabled;             /* 0: v1 kernels; 1: v2 */
    int lmax;                /* links per virtual column: 8 or 16 */
    int n_groups;            /* column groups (K1 grid.y) */
    int vc_pad;              /* K1 consumer threads = virtual columns per group, padded to 32 */
    int n_items;             /* rows of P: na self items + alpha single excitations */
    int n_chunks;            /* K1 work units (source string, <= items_per_chunk items) */
    int max_split;           /* capacity of `part` in K splits */
    int lda, ldb;            /* row strides of HaDT / HbDT */
    int ldp, ldq;            /* row stride of P = ldq (q-part: one column per beta-string segment, padded) +
                                w-part (one column per beta string: the sgn*Wb[pq,b]*c[a',b] term), padded */
    const uint32_t* vc_src;  /* [n_groups][lmax][vc_pad] source column b' | sign << 31 */
    const uint32_t* vc_off;  /* [n_groups][lmax][vc_pad] byte offset 8*rs into an integral row (pad: zero slot) */
    const int* vc_len;       /* [n_groups][vc_pad] */
    const int* vc_q;         /* [n_groups][vc_pad] column of P the virtual column writes, or -1 */
    const int* col_seg;      /* [nb+1] first P column of beta string b (its segments are adjacent) */
    const int* single_ptr;   /* [na+1] exclusive scan of a.n_single */
    const int* item_ptr;     /* [na+1] single_ptr[a] + a: P row of a's self item; the row fed by its k-th
                                single excitation (source a.col[k], target a) follows at +1+k */
    const int* chunk_rec;    /* [n_chunks][4] {source alpha string, first item, items, 0}; sorted by descending size */
    const int* item_tgt;     /* [n_items] (source order) target alpha string of the item */
    const uint32_t* item_gsel; /* [n_items] integral row pq of source -> target | sign << 31; bit 30: self item (row of Wa) */
    const int* item_pslot;   /* [n_items] (source order) P row the item writes */
    const int* heavy_rows;   /* [n_heavy] alpha strings with more than 64 single excitations (own epilogue launch) */
    int n_heavy;
    const double* HaDT;      /* [lda*lda] dense same-spin alpha block, transposed: HaDT[a'*lda + a] */
    const double* HbDT;      /* [ldb*ldb] */
    double* P;               /* [n_items * ldp] */
    double* part;            /* [max_split * na * ldc] */
} sqd_sigma_v2;

typedef struct {
    sqd_spin_table a, b;     /* alpha strings index rows, beta strings index columns */
    int norb;
    int ldc;                 /* leading dimension of CI matrices (>= nb, multiple of 2) */
    int ldg;                 /* row stride of gab / Wa (>= norb^2, multiple of 2) */
    const double* diag;      /* [na*ldc] operator diagonal */
    const double* gab;       /* [norb^2 * ldg] */
    const double* Wa;        /* [na*ldg] or NULL */
    const double* Wb;        /* [norb^2*ldc] or NULL */
    int use_same_spin;       /* 1: include table values (Hamiltonian); 0: opposite-spin only (S^2) */
    sqd_sigma_plan plan;
    sqd_sell bd;             /* beta single excitations (mode 0; long columns have length 0) */
    sqd_sell bb;             /* every beta entry with its same-spin value (mode 1) */
    int throughput_mode;     /* see sqd_solve_params.throughput_mode */
    sqd_sigma_v2 v2;         /* v2.enabled != 0: sqd_sigma runs the v2 kernels */
    int wide;                /* != 0 (and v2 disabled): the staging-free "wide" kernel -- any nb up to the 2^19
                                strings of the tables; needs neither plan nor SELL copies */
} sqd_operator;

/* Build the v2 structures of a subspace (operator independent: shared by the Hamiltonian and S^2).
 * Step 1, sqd_sigma_v2_plan: everything that only needs the tables.  d_plan: sqd_sigma_v2_plan_bytes(...)
 *   bytes; h_counts: int[16] -- when not NULL the stream is synchronised and the counts are returned
 *   (a caller that reads d_counts itself passes NULL): [0] n_items, [1] n_chunks, [2] n_groups, [3] vc_pad,
 *   [4] alpha singles, [5] beta singles, [6] error flag (!= 0: shape unsupported, use v1), [7] virtual
 *   columns, [8] P columns.
 * Step 2, sqd_sigma_v2_finish: dense blocks, zeroed P, the struct.  d_scratch: sqd_sigma_v2_scratch_bytes
 *   bytes.  dense != 0 builds HaDT/HbDT (needed by Hamiltonian-like operators, not by S^2 alone). */
#define SQD_V2_COUNTS 16
/* 1 when the v2 kernels are expected to beat v1 for tables of this size: the dense same-spin tiles cost
 * na*nb*(na+nb) FMAs against nnz_a*nb + nnz_b*na gathered ones, worth it above a few per cent density. */
int sqd_sigma_v2_recommended(int na, int nb, int64_t nnz_a, int64_t nnz_b);
int64_t sqd_sigma_v2_plan_bytes(int na, int nb, int64_t nnz_a, int64_t nnz_b, int lmax, int items_per_chunk);
int sqd_sigma_v2_plan(const sqd_spin_table* a, const sqd_spin_table* b, int norb, int64_t nnz_a,
                      int64_t nnz_b, int lmax, int items_per_chunk, void* d_plan, int64_t plan_bytes,
                      int* h_counts, void* stream);
/* device address of the int[SQD_V2_COUNTS] counts inside d_plan */
const int* sqd_sigma_v2_counts_ptr(void* d_plan, int na, int nb, int64_t nnz_a, int64_t nnz_b, int lmax,
                                   int items_per_chunk);
int64_t sqd_sigma_v2_scratch_bytes(const int* h_counts, int na, int nb, int ldc, int dense, int same_tables);
int sqd_sigma_v2_finish(const sqd_spin_table* a, const sqd_spin_table* b, int ldc, int64_t nnz_a,
                        int64_t nnz_b, int lmax, int items_per_chunk, const int* h_counts, void* d_plan,
                        void* d_scratch, int64_t scratch_bytes, int dense, sqd_sigma_v2* out, void* stream);

/* Build a SELL-32 copy of a table.  mode 0: single excitations only, strings with more than
 * long_threshold of them get length 0 (they are reduced cooperatively, see sqd_sigma_plan_build);
 * mode 1: all entries with values.  capacity (entries of d_pack / d_val) >= nnz + 32*n.
 * d_perm, d_len: int[n]; d_slice_ptr: int[ceil(n/32)+1]. */
int sqd_sell_build(const sqd_spin_table* t, int mode, const int* d_long_idx, int capacity, int* d_perm,
                   int* d_len, int* d_slice_ptr, uint32_t* d_pack, double* d_val, void* stream);

/* Build the work plan.  cost_per_chunk: multiple of 16 (a single excitation costs 16, a double 1);
 * long_threshold: beta strings with more single excitations than this become long columns (at most
 * SQD_MAX_LONG_COLUMNS); max_chunks: capacity of the d_chunk_* arrays, 2*na + (16*nnz_a)/cost_per_chunk + 1
 * always suffices.  d_split_*: int[na]; d_long_idx: int[nb]; d_long_cols: int[SQD_MAX_LONG_COLUMNS];
 * d_counts: int[4] device scratch; h_counts receives {n_chunks, n_slots, n_split, n_long}.
 * Synchronises the stream. */
int sqd_sigma_plan_build(const sqd_spin_table* a, const sqd_spin_table* b, int cost_per_chunk,
                   int long_threshold, int max_chunks, int* d_chunk_row, int* d_chunk_beg,
                   int* d_chunk_end, int* d_chunk_slot, int* d_split_row, int* d_split_slot_beg,
                   int* d_split_n, int* d_long_idx, int* d_long_cols, int* d_counts, int* h_counts,
                   void* stream);

/* 1 when CI rows of ldc doubles and integral rows of ldg doubles fit the shared-memory staging of the v1
 * kernels (nb <= 5760 at 30 orbitals); 0: only the wide kernel (sqd_operator.wide) applies.  pyscf's
 * contract_2e has no such limit (fermion.py:721-723 reaches it for any subspace size). */
int sqd_sigma_v1_supported(int ldc, int ldg);

/* Dynamic shared memory the sigma kernel needs for this operator, or <0 if the shape is unsupported. */
int64_t sqd_sigma_smem_bytes(const sqd_operator* op);

/* d_sigma[a*ldc+b] = sum_{a'b'} <ab|O|a'b'> d_c[a'*ldc+b'];  pads of d_sigma are written as 0. */
int sqd_sigma(const sqd_operator* op, const double* d_c, double* d_sigma, void* stream);

/* As sqd_sigma, restricted to rows [row_begin, row_end) of sigma; other rows are left untouched. */
int sqd_sigma_rows(const sqd_operator* op, const double* d_c, double* d_sigma, int row_begin,
                   int row_end, void* stream);

/* Diagnostics: one sigma build that also records, per CTA of the alpha kernel, 8 values in
 * d_prof (int64[8 * plan.n_chunks]): clock64() at start / after table staging / after the doubles /
 * after the singles / at the end, then #singles, #doubles and the SM id. */
int sqd_sigma_profile(const sqd_operator* op, const double* d_c, double* d_sigma, long long* d_prof,
                      void* stream);

typedef struct {
    int max_space;       /* <= SQD_MAX_SPACE; pyscf default 12 */
    int max_cycle;       /* pyscf default 100 */
    double tol;          /* converged when |d theta| < tol and |r| < tol_residual */
    double tol_residual; /* pyscf: sqrt(tol) */
    double lindep;       /* pyscf default 1e-14 */
    double level_shift;  /* preconditioner: r / (hdiag - theta + level_shift), pyscf 1e-4 */
    int check_every;     /* host polls the device-side convergence flag every this many cycles */
    /* quadratic spin penalty  shift*(S^2-ss)^2  (pyscf fix_spin_ when ss >= sz(sz+1)+0.1) */
    const sqd_operator* ss_op; /* NULL unless the quadratic form is requested */
    double ss_shift, ss_value;
    int profile;         /* != 0: bracket every operator application with CUDA events (measurement) */
    /* one diagonalisation sharded over GPUs: NULL, or a communicator from sqd_nccl_init; this rank then
     * builds rows [row_begin, row_end) of every sigma vector and the blocks are all-reduced (sum) */
    void* nccl_comm;
    int row_begin, row_end;
    /* sharded build, optional: row bounds of ALL ranks (shard_world + 1 ints, host memory, valid during the
     * call; bounds[r] = first row of rank r).  When given, the ranks exchange their disjoint row blocks
     * (grouped broadcasts, (W-1)/W of the vector per rank) instead of all-reducing zero-padded vectors. */
    const int* shard_bounds;
    int shard_world;
    /* 0: the lowest Ritz pair comes from the secular equation on the main stream and the full
     * Rayleigh-Ritz decomposition runs on a side stream (shortest critical path of ONE solve);
     * != 0: one Rayleigh-Ritz kernel on the main stream (fewer launches when many solves share the GPU) */
    int single_stream_ritz;
} sqd_davidson_params;

typedef struct {
    int converged;   /* 1 converged, 0 hit max_cycle, 2 stopped on linear dependency */
    int cycles;      /* Davidson cycles (= sigma builds inside the loop) */
    int sigma_builds;
    double theta;    /* last Ritz value of the (possibly spin-penalised) operator */
    double residual; /* last residual norm */
    double sigma_ms; /* profile != 0: summed device time of the operator applications, else 0 */
    double total_ms; /* profile != 0: device time of the whole Davidson loop */
} sqd_davidson_info;

/* bytes of device workspace for sqd_davidson on an operator of na x ldc */
int64_t sqd_davidson_workspace_bytes(int na, int ldc, int max_space);

/* Ground state of `op`.  d_hdiag: preconditioner diagonal [na*ldc] (bare Hamiltonian diagonal, as in
 * pyscf where fix_spin_ leaves hdiag untouched).  d_x0: start vector [na*ldc] (unit vector at
 * argmin(hdiag) + pyscf's 1e-5 noise is produced by sqd_init_guess).  d_x receives the normalised
 * Ritz vector.  Synchronises the stream before returning. */
int sqd_davidson(const sqd_operator* op, const double* d_hdiag, const double* d_x0, double* d_x,
                 void* d_workspace, int64_t workspace_bytes, const sqd_davidson_params* params,
                 sqd_davidson_info* h_info, void* stream);

/* pyscf direct_spin1._get_init_guess: e_argmin(hdiag), +1e-5 on element 0 and -1e-5 on the last. */
int sqd_init_guess(const double* d_hdiag, int na, int nb, int ldc, double* d_x0, void* d_scratch,
                   void* stream);

/* ------------------------------------------------------------------------------------------ *
 * One-call subspace solve: everything below the Python signature of fermion.py:solve_sci (:711-740)
 *   kernel_fixed_space (tables, hdiag, initial guess, Davidson) -> make_rdm1s diagonals -> energy of the
 *   bare Hamiltonian -> spin_square -> optional make_rdm1 / make_rdm2.
 * The host thread stays inside this call from the first kernel to the last read-back, so K subspaces
 * solved by K host threads (one stream each) do not serialise on the interpreter lock.  Scratch memory
 * is taken from the device's stream-ordered pool (cudaMallocAsync) and returned before the call ends.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int norb, na, nb;            /* strings sorted ascending, unique, equal popcount per list */
    int n_alpha, n_beta;         /* electrons per spin (popcount of the strings) */
    const uint64_t* d_strs_a;    /* [na] device */
    const uint64_t* d_strs_b;    /* [nb] device; the same pointer as d_strs_a shares one table */
    const double* d_h;           /* [norb^2] device */
    const double* d_g;           /* [norb^4] device, chemist order (pq|rs) */
    int penalty;                 /* 0: none; 1: pyscf fix_spin_(ss=spin_sq, shift) -- the linear form when
                                    spin_sq < sz(sz+1)+0.1, else the quadratic form */
    double spin_sq, shift;
    int want_spin;               /* also return <S^2> when no penalty is requested */
    int max_space, max_cycle;    /* Davidson, see sqd_davidson_params */
    double tol, tol_residual, lindep, level_shift;
    int check_every;
    const double* d_ci0;         /* NULL, or start vector (na, nb) row-major on the device */
    int cost_per_chunk, long_threshold; /* sigma work plan, 0 = defaults (256, 64) */
    int profile;                 /* time the operator applications with CUDA events; count singles */
    void* nccl_comm;             /* NULL, or sharded build (see sqd_davidson_params) */
    int row_begin, row_end;      /* rows of sigma this rank builds; row_begin < 0: split the rows into
                                    shard_world contiguous blocks of equal estimated cost, take block
                                    shard_rank */
    int shard_rank, shard_world;
    int throughput_mode;         /* != 0: this solve shares the GPU with others (one stream each): the sigma
                                    kernel is launched under a register cap that lets three of its CTAs
                                    share an SM -- slower alone, faster in aggregate */
    int sigma_path;              /* 0: choose by table density (sqd_sigma_v2_recommended); 1: v1 kernels;
                                    2: v2 kernels whenever their planner accepts the shape; 3: the wide
                                    kernel.  Rows too long for the staged kernels (nb > 8192 for v2, > 5760
                                    for v1) take the wide kernel whatever is asked for */
    int v2_lmax, v2_items_per_chunk; /* v2 tuning, 0 = defaults (16, 8) */
} sqd_solve_params;

typedef struct {
    double energy;               /* <x|H|x> of the bare Hamiltonian (fermion.py:730-732) */
    double spin_square;          /* <x|S^2|x> when have_spin_square */
    int have_spin_square;
    double occ_a[64], occ_b[64]; /* diagonals of the spin 1-RDMs (fermion.py:725-726) */
    sqd_davidson_info info;
    int64_t nnz_a, nnz_b;        /* entries of the two excitation tables */
    int64_t singles_a, singles_b;/* profile != 0: single excitations among them, else -1 */
    int ldc;                     /* row stride of d_x */
    int sigma_path;              /* 1 / 2 / 3: the sigma kernels that ran (v1, v2, wide) */
} sqd_solve_result;

/* d_x: double[na * ldc], ldc = nb rounded up to even, receives the normalised ground state (pad columns
 * zero; sign: first element of largest magnitude positive).  d_rdm1: NULL or double[norb^2] (pyscf
 * make_rdm1).  d_rdm2: NULL or double[norb^4] (pyscf make_rdm2, dm2[p,q,r,s] = <p+ r+ s q>). */
int sqd_solve_subspace(const sqd_solve_params* params, double* d_x, double* d_rdm1, double* d_rdm2,
                       sqd_solve_result* h_result, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Multi-GPU exchange for one sharded diagonalisation (the K-batches mode needs no collective).
 * NCCL is resolved with dlopen at first use.  Bootstrap: rank 0 calls sqd_nccl_unique_id and ships the
 * 128 bytes to the other ranks by any means (the Python host uses torch.distributed); every rank then
 * calls sqd_nccl_init with its CUDA device current.
 * ------------------------------------------------------------------------------------------ */
int sqd_nccl_unique_id(char* h_id128);
int sqd_nccl_init(const char* h_id128, int rank, int world, void** comm_out);
int sqd_nccl_destroy(void* comm);
int sqd_allreduce_sum_f64(void* comm, double* d_buf, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Expectation values   (fermion.py:821-830: make_rdm1s diagonals, energy, spin_square)
 * ------------------------------------------------------------------------------------------ */
/* d_out[0] = sum_i x_i y_i over na*ldc (deterministic two-stage reduction; scratch >= 1024 doubles) */
int sqd_dot(const double* d_x, const double* d_y, int64_t n, double* d_out, double* d_scratch,
            void* stream);

/* Sign convention of a returned eigenvector (fermion.py:837-845 hands pyscf's vector through as is; its
 * sign is LAPACK's): flip d_x in place so that its first element of largest magnitude is positive.
 * d_scratch: at least 4096 doubles. */
int sqd_fix_sign(double* d_x, int64_t n, void* d_scratch, void* stream);

/* Wait until everything enqueued on `stream` has finished.  Unlike cudaStreamSynchronize the wait can be
 * told not to spin (environment SQD_WAIT_MODE: 0 spin, 1 blocking-sync event, 2 poll + yield/sleep; default 2
 * when the process is one of more than two ranks, WORLD_SIZE > 2, else 0): with 8 ranks x 8 solver threads on
 * a 32-core host spinning waits starve the launching threads. */
int sqd_stream_wait(void* stream);
/* Device -> host copy into the caller's (pageable or pinned) array, stream-ordered, returns when the data
 * has arrived.  For bindings whose host language holds a global lock while it copies (Python): the copy
 * runs inside the call, outside that lock. */
int sqd_download(void* h_dst, const void* d_src, long long bytes, void* stream);

/* Small (<= 64 KB) device -> host read-back through a per-thread pinned staging buffer followed by a
 * stream synchronisation; never blocks the launches of other host threads. */
int sqd_read_back(void* h_dst, const void* d_src, int64_t bytes, void* stream);

/* occ_a[p] = sum_{a: p in a} sum_b c[a,b]^2, occ_b[p] likewise.  d_occ: double[2*norb] (alpha first).
 * d_scratch: double[na + nb]. */
int sqd_occupancies(const double* d_c, const uint64_t* d_strs_a, int na, const uint64_t* d_strs_b,
                    int nb, int ldc, int norb, double* d_occ, double* d_scratch, void* stream);

/* Spin-resolved 1-RDMs  dm1a[p*norb+q] = <c| a+_p a_q |c>, dm1b likewise (pyscf make_rdm1s reached from
 * fermion.py:117-121, 725-729).  d_dm1: double[2*norb*norb] (alpha block first).  d_workspace:
 * sqd_rdm1s_workspace_bytes(op) bytes; d_dots: double[max(nnz_a, nnz_b)] scratch. */
int64_t sqd_rdm1s_workspace_bytes(const sqd_operator* op);
int sqd_rdm1s(const sqd_operator* op, const double* d_c, int64_t nnz_a, int64_t nnz_b, double* d_dm1,
              double* d_workspace, double* d_dots, void* stream);

/* Spin-separated 2-RDMs in pyscf's convention  dm2[p,q,r,s] = <c| p+ r+ s q |c>  (selected_ci.make_rdm2s:
 * (alpha,alpha), (alpha,beta), (beta,beta) blocks), reached from fermion.py:117-128 (SCIState.rdm) and
 * fermion.py:728-729, 825-826 (energy from RDMs, SCIResult.rdm2).  Each output is double[norb^4],
 * C-order [p][q][r][s]; for the opposite-spin block p,q are alpha and r,s beta orbitals.  nnz_a / nnz_b =
 * entries of the excitation tables.  d_dm1: NULL, or double[2*norb^2] that receives the spin 1-RDMs in
 * the layout of sqd_rdm1s (they share the row-pair dot products).  Deterministic (no atomics). */
int64_t sqd_rdm2s_workspace_bytes(const sqd_operator* op, int64_t nnz_a, int64_t nnz_b);
int sqd_rdm2s(const sqd_operator* op, const double* d_c, int64_t nnz_a, int64_t nnz_b, double* d_dm2aa,
              double* d_dm2ab, double* d_dm2bb, double* d_dm1, void* d_workspace, int64_t ws_bytes,
              void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Qubit path   (qubit.py:78-300)
 * ------------------------------------------------------------------------------------------ */
/* big-endian bool rows -> int64 keys (qubit.py:280-300); nbits <= 63 */
int sqd_bits_to_keys(const uint8_t* d_bits, int64_t n, int nbits, int64_t* d_keys, void* stream);

/* key -> row hash table over the (unique) keys of the subspace: open addressing, linear probing, capacity a
 * power of two >= 4 d (>= 2 d above 2^24 rows), 12 bytes per slot.  The projection kernels below take it as
 * an optional argument (NULL: binary search of the sorted key list, as in round 1): "is key ^ xmask in the
 * subspace" is almost always answered no, which a probe of a sparse table settles in ~1.4 L2 accesses instead
 * of log2(d) dependent ones.  (Reference: np.isin + searchsorted per term, qubit.py:225-237.) */
int64_t sqd_key_table_capacity(int64_t d);
int64_t sqd_key_table_bytes(int64_t d);
int sqd_key_table_build(const int64_t* d_keys, int64_t d, void* d_table, int64_t table_bytes, void* stream);

/* One Pauli term (qubit.py:167-240): for every row i whose image key_i ^ xmask is in the sorted key
 * list, amplitude (-1)^popc(key_i & zmask) * i^ny, row i, col index of the image (-1: not in the subspace).
 * sqd_pauli_connect returns the parity bit; sqd_pauli_elements returns the outputs in the reference's own
 * types (complex128 amplitude, int64 column; d_col may be NULL when xmask == 0, where col[i] = i) and the
 * number of rows without an image in *d_n_missing. */
int sqd_pauli_connect(const int64_t* d_keys, int64_t d, const void* d_table /* or NULL */, uint64_t xmask,
                      uint64_t zmask, int32_t* d_col, uint8_t* d_par, void* stream);
int sqd_pauli_elements(const int64_t* d_keys, int64_t d, const void* d_table /* or NULL */, uint64_t xmask,
                       uint64_t zmask, int ny, int64_t* d_col, double* d_amp /* re,im pairs */,
                       int32_t* d_n_missing, void* stream);

/* Projection of a whole operator (qubit.py:78-144).  Terms are pre-grouped by X mask on the host:
 * group k owns terms [grp_ptr[k], grp_ptr[k+1]) (original order kept inside a group), each with a
 * zmask, a number of Y's and a complex coefficient.  Pass 1 counts, pass 2 fills a CSR matrix whose
 * row i holds A[i, col] for every group that connects row i (transpose convention of the reference:
 * A[source, image]), columns ascending, exact zeros dropped (scipy canonical format). */
/* Optional pre-pass for the diagonal group (X mask 0, terms [t0, t1) ): d_out[2i], d_out[2i+1] = the sum of
 * the group's terms for row i, in term order.  Hand the group's index and d_out to count/fill (diag_group,
 * d_diag_val; -1 / NULL: none) and they read the sums instead of re-evaluating thousands of Z strings per row. */
int sqd_pauli_diag_group(const int64_t* d_keys, int64_t d, const uint64_t* d_zmask, const int32_t* d_ny,
                         const double* d_coeff, int32_t t0, int32_t t1, double* d_out, void* stream);
int sqd_pauli_project_count(const int64_t* d_keys, int64_t d, const void* d_table /* or NULL */,
                            const uint64_t* d_grp_xmask,
                            const int32_t* d_grp_ptr, int32_t n_groups, const uint64_t* d_zmask,
                            const int32_t* d_ny, const double* d_coeff /* re,im pairs */,
                            int32_t diag_group, const double* d_diag_val,
                            int32_t* d_row_nnz, void* stream);
int sqd_pauli_project_fill(const int64_t* d_keys, int64_t d, const void* d_table /* or NULL */,
                           const uint64_t* d_grp_xmask,
                           const int32_t* d_grp_ptr, int32_t n_groups, const uint64_t* d_zmask,
                           const int32_t* d_ny, const double* d_coeff, int32_t diag_group,
                           const double* d_diag_val, const int32_t* d_row_ptr,
                           int32_t* d_col_tmp, double* d_val_tmp /* scratch, nnz entries each */,
                           int32_t* d_col, double* d_val /* re,im pairs */, void* stream);

/* y = A x for a complex128 CSR matrix (warp per row). */
int sqd_csr_matvec_c128(int64_t d, const int32_t* d_row_ptr, const int32_t* d_col,
                        const double* d_val, const double* d_x, double* d_y, void* stream);

/* Gershgorin data of a Hermitian CSR matrix: d_diag[i] = Re A_ii and the row's lower bound
 * d_lower[i] = A_ii - sum_{j != i} |A_ij|  (no eigenvalue of the connected component containing row i
 * lies below the minimum of d_lower over that component). */
int sqd_csr_gershgorin(int64_t d, const int32_t* d_row_ptr, const int32_t* d_col, const double* d_val,
                       double* d_diag, double* d_lower, void* stream);

/* Connected components of a Hermitian CSR matrix (the block structure of the projected operator: sets of
 * configurations no Pauli term connects).  d_label[i] = smallest row index of the component of row i.
 * Single-row components are exact eigenpairs (A_ii, e_i) and are reduced on the device:
 * h_head[0] = number of multi-row components (their records, in no particular order, in d_rec[0..min(.,cap)) ),
 * h_head[1] = number of single-row components, h_head[2] = row of the lowest one (-1 if none) and
 * *h_best_single = its diagonal.  d_diag / d_lower come from sqd_csr_gershgorin. */
typedef struct sqd_component {
    int32_t root;          /* label of the component */
    int32_t size;          /* rows in it (>= 2) */
    int32_t row_min_diag;  /* first row with the smallest diagonal: the Davidson start row */
    int32_t pad;
    double lower;          /* min over the rows of A_ii - sum_j |A_ij|: no eigenvalue of the block is below it */
    double diag;           /* smallest diagonal: the block's lowest eigenvalue is <= it */
} sqd_component;
int64_t sqd_csr_components_workspace_bytes(int64_t d);
int sqd_csr_components(int64_t d, const int32_t* d_row_ptr, const int32_t* d_col, const double* d_diag,
                       const double* d_lower, int32_t* d_label, sqd_component* d_rec, int64_t rec_cap,
                       int32_t* h_head /* [3] */, double* h_best_single, void* d_workspace,
                       int64_t ws_bytes, void* stream);
/* complex128[d] start vector confined to one component: e_row + scale * deterministic noise in [-1, 1) on the
 * rows with d_label == root (pyscf perturbs its unit start vector for the same reason: a ground state with a
 * node on the start row must still have a component in the start vector). */
int sqd_csr_component_start(int64_t d, const int32_t* d_label, int32_t root, int32_t row, double scale,
                            double* d_start, void* stream);

/* Lowest eigenpair (k must be 1) of the Hermitian matrix A -- the matrix the reference hands to eigsh
 * (qubit.py:73) -- by the device-resident Davidson on the (re, im) embedding, replacing ARPACK.
 * d_start: complex128[d] start vector or NULL (unit vector at argmin of the diagonal); the iteration
 * stays inside the connected component(s) the start vector touches.  d_evecs: complex128[d];
 * h_evals: double[1]; h_cycles / h_residual may be NULL.  Synchronises. */
int64_t sqd_csr_davidson_workspace_bytes(int64_t d, int k, int max_space);
int sqd_csr_davidson(int64_t d, const int32_t* d_row_ptr, const int32_t* d_col, const double* d_val,
                     int k, int max_space, int max_cycle, double tol, const double* d_start,
                     double* d_evecs, double* h_evals, int* h_cycles, double* h_residual,
                     void* d_workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Configuration recovery   (configuration_recovery.py:59-306)
 * ------------------------------------------------------------------------------------------ */
/* Per-bitstring correction.  Rows are packed halves (left = beta, right = alpha) as produced by
 * sqd_pack_bitstrings.  d_occ_left / d_occ_right: occupancy of every bit in COLUMN order of the half
 * (configuration_recovery.py:113).
 * mode 0 = exact stream: d_rng_state holds numpy's PCG64 (state_hi, state_lo, inc_hi, inc_lo); rows are
 *   processed in order and consume uniform doubles exactly as Generator.choice does
 *   (left half first, `size - n_uniq` doubles per retry round); the advanced state is written back, so
 *   the caller's Generator continues as if numpy had run.
 * mode 1 = one PCG64 substream per row derived from (seed, row): same distribution, fully parallel.
 * d_status[0] != 0: numpy would have raised "Fewer non-zero entries in p than size" at row
 * d_status[1]. */
int64_t sqd_recover_workspace_bytes(int64_t n, int norb);
int sqd_recover(const uint64_t* d_left, const uint64_t* d_right, int64_t n, int norb,
                const double* d_occ_left, const double* d_occ_right, int hamming_left,
                int hamming_right, int mode, uint64_t* d_rng_state, uint64_t seed,
                uint64_t* d_left_out, uint64_t* d_right_out, int32_t* d_status, void* d_workspace,
                int64_t workspace_bytes, void* stream);

/* Duplicate merge of the repaired rows (configuration_recovery.py:112-126): the distinct (left, right) rows in
 * first-seen order and, per distinct row, the probabilities of its occurrences added in input order starting
 * from 0.0 (the reference's row-by-row `freqs[idx] += p`; bit-identical sums).  Outputs hold *h_n_unique
 * entries.  Hash table over the rows, first occurrences ranked by a scan, members of a group sorted by row. */
int64_t sqd_merge_rows_workspace_bytes(int64_t n);
int sqd_merge_rows(const uint64_t* d_left, const uint64_t* d_right, int64_t n, const double* d_prob,
                   uint64_t* d_out_left, uint64_t* d_out_right, double* d_out_sum, int32_t* h_n_unique,
                   void* d_workspace, int64_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------ *
 * Sample formats   (counts.py:45-61 bit_array_to_arrays; qubit.py:147-164 sort_and_remove_duplicates and
 *                   the np.unique of solve_qubit, qubit.py:66)
 * A bitstring of up to 128 bits travels as a (hi, lo) pair of 64-bit words, column 0 of the bool matrix = most
 * significant bit, so numpy's lexicographic row order is the numeric order of the pair.
 * ------------------------------------------------------------------------------------------ */
/* Distinct keys in ascending order with their multiplicities (np.unique(..., return_counts=True)).
 * d_hi == NULL: 64-bit keys (d_out_hi may then be NULL too).  d_out_*: n entries; d_out_count may be NULL.
 * *h_n_unique receives the number of distinct keys.  Bitonic sort in a padded copy inside the workspace
 * (inputs are not modified).  Synchronises the stream. */
int64_t sqd_sort_unique_workspace_bytes(int64_t n);
int sqd_sort_unique(const uint64_t* d_hi, const uint64_t* d_lo, int64_t n, uint64_t* d_out_hi,
                    uint64_t* d_out_lo, int32_t* d_out_count, int64_t* h_n_unique, void* d_workspace,
                    int64_t workspace_bytes, void* stream);

/* Rows of a sampled bit array (qiskit BitArray.array: uint8, big-endian, left-padded, row_bytes per shot) ->
 * (hi, lo); bits beyond num_bits are masked off (counts.py:57).  num_bits <= 128; d_hi may be NULL up to 64. */
int sqd_bit_array_pack(const uint8_t* d_bytes, int64_t n, int row_bytes, int num_bits, uint64_t* d_hi,
                       uint64_t* d_lo, void* stream);

/* (hi, lo) -> bool matrix rows of num_bits columns (1 byte per bit, column 0 = most significant bit). */
int sqd_keys_to_bits(const uint64_t* d_hi, const uint64_t* d_lo, int64_t n, int num_bits, uint8_t* d_bits,
                     void* stream);

/* Carry-over selection of the SQD loop (fermion.py:607-631, _process_sci_results): d_row_flag[a] / d_col_flag[b]
 * = 1 when row a / column b of the amplitude matrix d_x (na x ldc, row-major) holds an entry with
 * |c| >= threshold, and for those the marginal weights sum_b |c[a,b]|^2 / sum_a |c[a,b]|^2 added in numpy's
 * pairwise summation order without FMA contraction (bit-identical to np.sum(np.abs(amps[rows])**2, axis=1) and
 * np.sum(np.abs(amps[:, cols])**2, axis=0)); weights of unflagged rows / columns are 0.  The full CI vector
 * does not travel to the host for this step, and the reference's argsort over all n_det magnitudes is gone. */
int sqd_carryover(const double* d_x, int na, int nb, int ldc, double threshold, int* d_row_flag,
                  int* d_col_flag, double* d_row_weight, double* d_col_weight, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SQD_B200_H */

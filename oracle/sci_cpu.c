/* CPU restatement (plain C + OpenMP) of the fermionic subspace diagonalisation.
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY -- never linked into or called from the product package.
 *
 * What it restates: the work behind qiskit_addon_sqd/fermion.py:713-732 (solve_sci) and :803-830
 * (solve_fermion), i.e. pyscf's fci.selected_ci.kernel_fixed_space + make_rdm1s diagonals + energy
 * expectation.  pyscf (>= 2.9, unpinned, pyproject.toml:30) is not in /root/reference nor in this
 * image, so PARITY IS UNPINNED against pyscf itself; this file is checked against
 * oracle/fermion_oracle.py (dense Slater-Condon + Jordan-Wigner) in tests/test_oracle_cpu.py.
 *
 * Two sigma-vector algorithms are provided:
 *   algo 0 "direct"  : in-set excitation tables, O(n_det * links) -- the same mathematics as the CUDA
 *                      path, written as straightforward loops (independent code).
 *   algo 1 "pyscf"   : pyscf's published selected_ci.contract_2e structure (recalled, Appendix A of
 *                      SURVEY.md): h1e absorbed into the two-electron tensor; same-spin part through
 *                      (N-2)-electron intermediate strings with a gather -> dgemm -> scatter per
 *                      intermediate; opposite-spin part gather(alpha links) -> dgemm(eri) -> scatter
 *                      (beta links).  O(n_det * npair^2) flops -- this is the cost model of the
 *                      reference's CPU path and is what `bench.py --impl reference` times.
 * Davidson: unit start vector at argmin(hdiag) (+1e-5/-1e-5 noise on first/last element), diagonal
 * preconditioner r/(hdiag - e + 1e-4), max_space with collapse, |dE| < tol and |r| < sqrt(tol).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IDX4(p, q, r, s, n) ((((size_t)(p) * (n) + (q)) * (n) + (r)) * (n) + (s))

static inline int popc(uint64_t x) { return __builtin_popcountll(x); }
static inline int lowbit(uint64_t x) { return __builtin_ctzll(x); }
static inline uint64_t between(int p, int q) {
    int lo = p < q ? p : q, hi = p < q ? q : p;
    return ((1ull << hi) - 1ull) & ~((2ull << lo) - 1ull);
}

/* ---------------- optional BLAS (scipy's bundled OpenBLAS), resolved by the Makefile ------------ */
#ifdef SCI_USE_BLAS
extern void SCI_DGEMM(const char*, const char*, const int*, const int*, const int*, const double*,
                      const double*, const int*, const double*, const int*, const double*, double*,
                      const int*);
extern void SCI_BLAS_SET_THREADS(int);
#endif
/* C(m x n) = A(m x k) * B(k x n), row-major */
static void gemm_rm(int m, int n, int k, const double* A, const double* B, double* C) {
#ifdef SCI_USE_BLAS
    /* row-major C = A B  <=>  column-major C^T = B^T A^T */
    const double one = 1.0, zero = 0.0;
    SCI_DGEMM("N", "N", &n, &m, &k, &one, B, &n, A, &k, &zero, C, &n);
#else
    memset(C, 0, sizeof(double) * (size_t)m * n);
    for (int i = 0; i < m; ++i)
        for (int l = 0; l < k; ++l) {
            const double a = A[(size_t)i * k + l];
            if (a == 0.0) continue;
            const double* b = B + (size_t)l * n;
            double* c = C + (size_t)i * n;
            for (int j = 0; j < n; ++j) c[j] += a * b[j];
        }
#endif
}

/* ------------------------------------ excitation tables ---------------------------------------- */
typedef struct {
    int n;
    int* ptr;      /* [n+1] */
    int* nsingle;  /* [n] */
    int* col;      /* partner (source) index */
    double* val;   /* same-spin matrix element */
    int* pq;       /* singles: p*norb+q */
    int* sgn;      /* singles: +-1 */
    double* diag;  /* [n] */
} table_t;

static double diag_elem(uint64_t s, int norb, const double* h, const double* g) {
    double e = 0;
    for (uint64_t oi = s; oi; oi &= oi - 1) {
        int i = lowbit(oi);
        e += h[i * norb + i];
        for (uint64_t oj = s; oj; oj &= oj - 1) {
            int j = lowbit(oj);
            e += 0.5 * (g[IDX4(i, i, j, j, norb)] - g[IDX4(i, j, j, i, norb)]);
        }
    }
    return e;
}

static void build_table(const uint64_t* strs, int n, int norb, const double* h, const double* g,
                        table_t* t) {
    t->n = n;
    t->ptr = calloc(n + 1, sizeof(int));
    t->nsingle = calloc(n, sizeof(int));
    t->diag = malloc(sizeof(double) * n);
    int* ntot = calloc(n, sizeof(int));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        int c1 = 0, c2 = 0;
        for (int j = 0; j < n; ++j) {
            int pc = popc(strs[i] ^ strs[j]);
            c1 += pc == 2;
            c2 += pc == 4;
        }
        t->nsingle[i] = c1;
        ntot[i] = c1 + c2;
        t->diag[i] = diag_elem(strs[i], norb, h, g);
    }
    for (int i = 0; i < n; ++i) t->ptr[i + 1] = t->ptr[i] + ntot[i];
    int nnz = t->ptr[n];
    t->col = malloc(sizeof(int) * (nnz + 1));
    t->val = malloc(sizeof(double) * (nnz + 1));
    t->pq = calloc(nnz + 1, sizeof(int));
    t->sgn = calloc(nnz + 1, sizeof(int));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        const uint64_t tt = strs[i];
        int o1 = t->ptr[i], o2 = t->ptr[i] + t->nsingle[i];
        for (int j = 0; j < n; ++j) {
            const uint64_t s = strs[j], x = s ^ tt;
            int pc = popc(x);
            if (pc == 2) {
                int q = lowbit(x & s), p = lowbit(x & tt);
                int sg = (popc(s & between(p, q)) & 1) ? -1 : 1;
                double v = h[p * norb + q];
                for (uint64_t o = s; o; o &= o - 1) {
                    int k = lowbit(o);
                    v += g[IDX4(p, q, k, k, norb)] - g[IDX4(p, k, k, q, norb)];
                }
                t->col[o1] = j; t->val[o1] = sg * v; t->pq[o1] = p * norb + q; t->sgn[o1] = sg;
                ++o1;
            } else if (pc == 4) {
                uint64_t holes = x & s, parts = x & tt;
                int i1 = lowbit(holes); holes &= holes - 1; int i2 = lowbit(holes);
                int a1 = lowbit(parts); parts &= parts - 1; int a2 = lowbit(parts);
                int par = popc(s & ((1ull << i1) - 1));
                uint64_t u = s ^ (1ull << i1);
                par += popc(u & ((1ull << i2) - 1)); u ^= 1ull << i2;
                par += popc(u & ((1ull << a2) - 1)); u |= 1ull << a2;
                par += popc(u & ((1ull << a1) - 1));
                double v = g[IDX4(a1, i1, a2, i2, norb)] - g[IDX4(a1, i2, a2, i1, norb)];
                t->col[o2] = j; t->val[o2] = (par & 1) ? -v : v;
                ++o2;
            }
        }
    }
    free(ntot);
}

static void free_table(table_t* t) {
    free(t->ptr); free(t->nsingle); free(t->col); free(t->val); free(t->pq); free(t->sgn); free(t->diag);
}

/* ------------------------------------ direct sigma ---------------------------------------------- */
typedef struct {
    int na, nb, norb;
    table_t ta, tb;
    int same_tables;
    double* gab;   /* norb^2 x norb^2 */
    double* Wa;    /* na x norb^2 */
    double* Wb;    /* norb^2 x nb */
    double* hdiag; /* na x nb */
} direct_t;

static void direct_setup(direct_t* d, const uint64_t* sa, int na, const uint64_t* sb, int nb, int norb,
                         const double* h, const double* g, double shift, double ss) {
    const int n2 = norb * norb;
    d->na = na; d->nb = nb; d->norb = norb;
    build_table(sa, na, norb, h, g, &d->ta);
    d->same_tables = (na == nb && memcmp(sa, sb, sizeof(uint64_t) * na) == 0);
    if (d->same_tables) d->tb = d->ta; else build_table(sb, nb, norb, h, g, &d->tb);
    d->gab = malloc(sizeof(double) * (size_t)n2 * n2);
    for (int p = 0; p < norb; ++p) for (int q = 0; q < norb; ++q)
        for (int r = 0; r < norb; ++r) for (int s = 0; s < norb; ++s)
            d->gab[(size_t)(p * norb + q) * n2 + r * norb + s] =
                g[IDX4(p, q, r, s, norb)] - ((p == s && q == r) ? shift : 0.0);
    d->Wa = calloc((size_t)na * n2, sizeof(double));
    d->Wb = calloc((size_t)n2 * nb, sizeof(double));
    d->hdiag = malloc(sizeof(double) * (size_t)na * nb);
#pragma omp parallel for schedule(static)
    for (int a = 0; a < na; ++a)
        for (uint64_t o = sa[a]; o; o &= o - 1) {
            int p = lowbit(o);
            const double* src = d->gab + (size_t)(p * norb + p) * n2;
            for (int rs = 0; rs < n2; ++rs) d->Wa[(size_t)a * n2 + rs] += src[rs];
        }
#pragma omp parallel for schedule(static)
    for (int pq = 0; pq < n2; ++pq)
        for (int b = 0; b < nb; ++b) {
            double v = 0;
            for (uint64_t o = sb[b]; o; o &= o - 1) { int r = lowbit(o); v += d->gab[(size_t)pq * n2 + r * norb + r]; }
            d->Wb[(size_t)pq * nb + b] = v;
        }
    const int nea = popc(sa[0]), neb = popc(sb[0]);
    const double sz = 0.5 * (nea - neb);
    const double c0 = shift * (sz * (sz + 1.0) + neb - ss);
#pragma omp parallel for schedule(static)
    for (int a = 0; a < na; ++a)
        for (int b = 0; b < nb; ++b) {
            double v = d->ta.diag[a] + d->tb.diag[b] + c0;
            for (uint64_t o = sa[a]; o; o &= o - 1) { int p = lowbit(o); v += d->Wb[(size_t)(p * norb + p) * nb + b]; }
            d->hdiag[(size_t)a * nb + b] = v;
        }
}

static void direct_free(direct_t* d) {
    if (!d->same_tables) free_table(&d->tb);
    free_table(&d->ta);
    free(d->gab); free(d->Wa); free(d->Wb); free(d->hdiag);
}

static void direct_sigma(const direct_t* d, const double* c, double* out) {
    const int na = d->na, nb = d->nb, n2 = d->norb * d->norb;
    const table_t *ta = &d->ta, *tb = &d->tb;
#pragma omp parallel for schedule(dynamic, 4)
    for (int a = 0; a < na; ++a) {
        double* o = out + (size_t)a * nb;
        const double* ca = c + (size_t)a * nb;
        const double* wa = d->Wa + (size_t)a * n2;
        for (int b = 0; b < nb; ++b) {
            double acc = d->hdiag[(size_t)a * nb + b] * ca[b];
            const int beg = tb->ptr[b], ns = tb->nsingle[b], end = tb->ptr[b + 1];
            for (int e = beg; e < beg + ns; ++e)
                acc += (tb->val[e] + tb->sgn[e] * wa[tb->pq[e]]) * ca[tb->col[e]];
            for (int e = beg + ns; e < end; ++e) acc += tb->val[e] * ca[tb->col[e]];
            o[b] = acc;
        }
        const int beg = ta->ptr[a], ns = ta->nsingle[a], end = ta->ptr[a + 1];
        for (int e = beg + ns; e < end; ++e) {
            const double v = ta->val[e];
            const double* cp = c + (size_t)ta->col[e] * nb;
            for (int b = 0; b < nb; ++b) o[b] += v * cp[b];
        }
        for (int e = beg; e < beg + ns; ++e) {
            const double* cp = c + (size_t)ta->col[e] * nb;
            const double* gr = d->gab + (size_t)ta->pq[e] * n2;
            const double* wb = d->Wb + (size_t)ta->pq[e] * nb;
            const double sa_ = ta->sgn[e], va = ta->val[e];
            for (int b = 0; b < nb; ++b) {
                double s = 0;
                for (int f = tb->ptr[b]; f < tb->ptr[b] + tb->nsingle[b]; ++f)
                    s += tb->sgn[f] * gr[tb->pq[f]] * cp[tb->col[f]];
                o[b] += sa_ * s + (va + sa_ * wb[b]) * cp[b];
            }
        }
    }
}

/* ------------------------------------ pyscf-style sigma ----------------------------------------- */
/* link tables as pyscf: cd = single excitations incl. diagonal (p>=q folded to tril pair index),
 * dd = (N-2)-electron intermediates with the annihilation pairs reaching them. */
typedef struct { int pair, addr, sign; } link_t;
typedef struct {
    int n, nlink; link_t* cd;      /* [n][nlink] creation-destruction links (sign 0 = padding) */
    int nint, mlink; link_t* dd;   /* [nint][mlink] des-des links into the string list */
} plinks_t;

static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : x > y;
}
static int find_str(const uint64_t* strs, int n, uint64_t key) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (strs[mid] < key) lo = mid + 1; else hi = mid; }
    return (lo < n && strs[lo] == key) ? lo : -1;
}

static void build_plinks(const uint64_t* strs, int n, int norb, int nelec, plinks_t* L) {
    const int nvir = norb - nelec;
    L->n = n; L->nlink = nelec * (nvir + 1);
    L->cd = calloc((size_t)n * L->nlink, sizeof(link_t));
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
        link_t* row = L->cd + (size_t)i * L->nlink;
        int k = 0;
        for (uint64_t o = strs[i]; o; o &= o - 1) {
            int q = lowbit(o);
            row[k].pair = q * norb + q; row[k].addr = i; row[k].sign = 1; ++k;  /* E_qq */
            for (int p = 0; p < norb; ++p) {
                if ((strs[i] >> p) & 1ull) continue;
                uint64_t t = (strs[i] ^ (1ull << q)) | (1ull << p);
                int j = find_str(strs, n, t);
                row[k].pair = p * norb + q; row[k].addr = j < 0 ? 0 : j;
                row[k].sign = j < 0 ? 0 : ((popc(strs[i] & between(p, q)) & 1) ? -1 : 1);
                ++k;
            }
        }
    }
    L->nint = 0; L->mlink = 0; L->dd = NULL;
    if (nelec < 2) return;
    const int npr = nelec * (nelec - 1) / 2;
    uint64_t* inter = malloc(sizeof(uint64_t) * (size_t)n * npr);
    size_t m = 0;
    for (int i = 0; i < n; ++i)
        for (uint64_t o1 = strs[i]; o1; o1 &= o1 - 1)
            for (uint64_t o2 = o1 & (o1 - 1); o2; o2 &= o2 - 1)
                inter[m++] = strs[i] ^ (1ull << lowbit(o1)) ^ (1ull << lowbit(o2));
    qsort(inter, m, sizeof(uint64_t), cmp_u64);
    size_t u = 0;
    for (size_t k = 0; k < m; ++k) if (k == 0 || inter[k] != inter[k - 1]) inter[u++] = inter[k];
    L->nint = (int)u;
    L->mlink = (norb - nelec + 2) * (norb - nelec + 1) / 2;
    L->dd = calloc((size_t)L->nint * L->mlink, sizeof(link_t));
#pragma omp parallel for schedule(static)
    for (int k = 0; k < L->nint; ++k) {
        link_t* row = L->dd + (size_t)k * L->mlink;
        int c = 0;
        const uint64_t s2 = inter[k];
        for (int i = 1; i < norb; ++i) {
            if ((s2 >> i) & 1ull) continue;
            for (int j = 0; j < i; ++j) {
                if ((s2 >> j) & 1ull) continue;
                uint64_t t = s2 | (1ull << i) | (1ull << j);
                int addr = find_str(strs, n, t);
                if (addr < 0) continue;
                /* a_j a_i |t> (i > j): sign */
                int par = popc(t & ((1ull << i) - 1)) + popc((t ^ (1ull << i)) & ((1ull << j) - 1));
                row[c].pair = i * (i - 1) / 2 + j; row[c].addr = addr; row[c].sign = (par & 1) ? -1 : 1;
                ++c;
            }
        }
    }
    free(inter);
}

typedef struct {
    int na, nb, norb;
    plinks_t la, lb;
    double* eri_aa;   /* npair x npair antisymmetrised */
    double* eri_ab;   /* norb^2 x norb^2 with h1e absorbed */
    double* hdiag;
    int nea, neb;
} pyscf_t;

static void pyscf_setup(pyscf_t* P, const uint64_t* sa, int na, const uint64_t* sb, int nb, int norb,
                        const double* h, const double* g, const double* hdiag) {
    const int n2 = norb * norb, npair = norb * (norb - 1) / 2;
    P->na = na; P->nb = nb; P->norb = norb;
    P->nea = popc(sa[0]); P->neb = popc(sb[0]);
    build_plinks(sa, na, norb, P->nea, &P->la);
    build_plinks(sb, nb, norb, P->neb, &P->lb);
    /* absorb_h1e(h, eri, norb, nelec, 0.5): h2e = eri with (h - 1/2 sum_j (pj|jq)) / N spread on the
     * diagonals; then contract_2e uses  eri1 = h2e*2 + h_ps corrections.  Net effect restated:
     *   sigma = sum_{pq,rs} 1/2 G[pq,rs] E_pq E_rs c  with
     *   G[pq,rs] = (pq|rs) + (delta_rs f_pq + delta_pq f_rs) / N ,  f = h - 1/2 sum_j (pj|jq)  */
    const int N = P->nea + P->neb;
    double* f = malloc(sizeof(double) * n2);
    for (int p = 0; p < norb; ++p) for (int q = 0; q < norb; ++q) {
        double v = h[p * norb + q];
        for (int j = 0; j < norb; ++j) v -= 0.5 * g[IDX4(p, j, j, q, norb)];
        f[p * norb + q] = v;
    }
    double* G = malloc(sizeof(double) * (size_t)n2 * n2);
    for (int p = 0; p < norb; ++p) for (int q = 0; q < norb; ++q)
        for (int r = 0; r < norb; ++r) for (int s = 0; s < norb; ++s) {
            double v = g[IDX4(p, q, r, s, norb)];
            if (r == s) v += f[p * norb + q] / N;
            if (p == q) v += f[r * norb + s] / N;
            G[(size_t)(p * norb + q) * n2 + r * norb + s] = v;
        }
    /* same-spin part: 1/2 sum G[pq,rs] E_pq E_rs = 1/2 sum G[pq,rs] a+_p a+_r a_s a_q + sum_ps K_ps E_ps,
     * K_ps = 1/2 sum_q G[pq,qs].  pyscf contracts the first term through (N-2)-electron intermediates
     * with the antisymmetrised tensor on strictly-lower-triangular pairs,
     *   W[(p>r),(q>s)] = G[pq,rs] - G[ps,rq],
     * and folds the one-body remainder K into the opposite-spin tensor with the number operator of the
     * OTHER spin (h_ps/nelec in selected_ci.contract_2e):
     *   Gab[pq,rs] = G[pq,rs] + delta_rs K_pq / N_beta + delta_pq K_rs / N_alpha.                  */
    P->eri_aa = malloc(sizeof(double) * (size_t)(npair ? npair : 1) * (npair ? npair : 1));
    for (int p = 1; p < norb; ++p) for (int r = 0; r < p; ++r)
        for (int q = 1; q < norb; ++q) for (int s = 0; s < q; ++s)
            P->eri_aa[(size_t)(p * (p - 1) / 2 + r) * npair + q * (q - 1) / 2 + s] =
                G[(size_t)(p * norb + q) * n2 + r * norb + s] - G[(size_t)(p * norb + s) * n2 + r * norb + q];
    double* K = calloc(n2, sizeof(double));
    for (int p = 0; p < norb; ++p) for (int s = 0; s < norb; ++s) {
        double v = 0;
        for (int q = 0; q < norb; ++q) v += 0.5 * G[(size_t)(p * norb + q) * n2 + q * norb + s];
        K[p * norb + s] = v;
    }
    double* Gab = malloc(sizeof(double) * (size_t)n2 * n2);
    for (int pq = 0; pq < n2; ++pq) for (int rs = 0; rs < n2; ++rs) {
        double v = G[(size_t)pq * n2 + rs];
        if (rs / norb == rs % norb) v += K[pq] / P->neb;
        if (pq / norb == pq % norb) v += K[rs] / P->nea;
        Gab[(size_t)rs * n2 + pq] = v;  /* stored transposed: row = beta pair rs, column = alpha pair pq */
    }
    P->eri_ab = Gab;
    free(K); free(G);
    free(f);
    P->hdiag = malloc(sizeof(double) * (size_t)na * nb);
    memcpy(P->hdiag, hdiag, sizeof(double) * (size_t)na * nb);
}

static void pyscf_free(pyscf_t* P) {
    free(P->la.cd); free(P->la.dd); free(P->lb.cd); free(P->lb.dd);
    free(P->eri_aa); free(P->eri_ab); free(P->hdiag);
}

/* same-spin contraction through N-2 intermediates on the ROW index of c (n x m, row-major) */
static void contract_same_spin(const plinks_t* L, int norb, const double* eri_aa, const double* c,
                               double* out, int n, int m) {
    const int npair = norb * (norb - 1) / 2;
    if (L->nint == 0) return;
#pragma omp parallel
    {
        double* t1 = malloc(sizeof(double) * (size_t)npair * m);
        double* vt = malloc(sizeof(double) * (size_t)npair * m);
        double* loc = calloc((size_t)n * m, sizeof(double));
#pragma omp for schedule(dynamic, 8)
        for (int k = 0; k < L->nint; ++k) {
            const link_t* row = L->dd + (size_t)k * L->mlink;
            memset(t1, 0, sizeof(double) * (size_t)npair * m);
            int used = 0;
            for (int l = 0; l < L->mlink && row[l].sign; ++l, ++used) {
                const double* src = c + (size_t)row[l].addr * m;
                double* dst = t1 + (size_t)row[l].pair * m;
                const double sg = row[l].sign;
                for (int b = 0; b < m; ++b) dst[b] += sg * src[b];
            }
            if (!used) continue;
            gemm_rm(npair, m, npair, eri_aa, t1, vt);
            for (int l = 0; l < used; ++l) {
                const double* src = vt + (size_t)row[l].pair * m;
                double* dst = loc + (size_t)row[l].addr * m;
                const double sg = row[l].sign;
                for (int b = 0; b < m; ++b) dst[b] += sg * src[b];
            }
        }
#pragma omp critical
        for (size_t i = 0; i < (size_t)n * m; ++i) out[i] += loc[i];
        free(t1); free(vt); free(loc);
    }
}

static void transpose(const double* a, double* at, int n, int m) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) for (int j = 0; j < m; ++j) at[(size_t)j * n + i] = a[(size_t)i * m + j];
}

static void pyscf_sigma(const pyscf_t* P, const double* c, double* out) {
    const int na = P->na, nb = P->nb, norb = P->norb, n2 = norb * norb;
#ifdef SCI_USE_BLAS
    SCI_BLAS_SET_THREADS(1); /* dgemm is called from inside OpenMP loops, as pyscf does */
#endif
    memset(out, 0, sizeof(double) * (size_t)na * nb);
    /* (aa|aa) */
    contract_same_spin(&P->la, norb, P->eri_aa, c, out, na, nb);
    /* (bb|bb) on the transpose */
    double* ct = malloc(sizeof(double) * (size_t)na * nb);
    double* ot = calloc((size_t)na * nb, sizeof(double));
    transpose(c, ct, na, nb);
    contract_same_spin(&P->lb, norb, P->eri_aa, ct, ot, nb, na);
    transpose(ot, ct, nb, na);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < (size_t)na * nb; ++i) out[i] += ct[i];
    free(ct); free(ot);
    /* (bb|aa): for each alpha string, gather t1[rs, a'] ... pyscf loops over beta blocks; restated per
     * alpha target row: t[pq, b] = sum_{alpha links a<-a'} sign c[a', b];  then v = G^T-contract over
     * beta links.  Here: for each alpha string a:  t1[pq, :] += sign * c[a', :]  (gather),
     * vt = eri_ab (rs,pq) x t1 (pq, nb)  (dgemm), scatter through beta links:
     * out[a, b] += sum_{beta links b<-b', rs} sign * vt[rs, b']. */
#pragma omp parallel
    {
        double* t1 = malloc(sizeof(double) * (size_t)n2 * nb);
        double* vt = malloc(sizeof(double) * (size_t)n2 * nb);
#pragma omp for schedule(dynamic, 4)
        for (int a = 0; a < na; ++a) {
            memset(t1, 0, sizeof(double) * (size_t)n2 * nb);
            const link_t* ra = P->la.cd + (size_t)a * P->la.nlink;
            for (int l = 0; l < P->la.nlink; ++l) {
                if (!ra[l].sign) continue;
                /* link stored on the SOURCE side: E_pq |a> = sign |addr>; we need rows reaching a:
                 * by hermiticity use E_qp: <a|E_qp|addr> = sign  -> contributes with pair (q,p) */
                const int p = ra[l].pair / norb, q = ra[l].pair % norb;
                const double* src = c + (size_t)ra[l].addr * nb;
                double* dst = t1 + (size_t)(q * norb + p) * nb;
                const double sg = ra[l].sign;
                for (int b = 0; b < nb; ++b) dst[b] += sg * src[b];
            }
            /* vt[rs, b'] = sum_pq Gab[pq, rs] t1[pq, b']  (eri_ab holds Gab^T; Gab is not symmetric when
             * N_alpha != N_beta because of the K/N_beta and K/N_alpha terms) */
            gemm_rm(n2, nb, n2, P->eri_ab, t1, vt);
            double* o = out + (size_t)a * nb;
            for (int b = 0; b < nb; ++b) {
                const link_t* rb = P->lb.cd + (size_t)b * P->lb.nlink;
                double acc = 0;
                for (int l = 0; l < P->lb.nlink; ++l) {
                    if (!rb[l].sign) continue;
                    const int r = rb[l].pair / norb, s = rb[l].pair % norb;
                    acc += rb[l].sign * vt[(size_t)(s * norb + r) * nb + rb[l].addr];
                }
                o[b] += acc;
            }
        }
        free(t1); free(vt);
    }
}

/* ------------------------------------ Davidson ------------------------------------------------- */
static void jacobi_lowest(double* A, int m, int lda, double* y, double* theta) {
    double Q[64 * 64];
    for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) Q[i * m + j] = i == j;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, dia = 0;
        for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) {
            double v = A[i * lda + j] * A[i * lda + j];
            if (i == j) dia += v; else off += v;
        }
        if (off <= 1e-32 * dia) break;
        for (int p = 0; p < m - 1; ++p) for (int q = p + 1; q < m; ++q) {
            double apq = A[p * lda + q];
            if (fabs(apq) < 1e-300) continue;
            double tau = (A[q * lda + q] - A[p * lda + p]) / (2 * apq);
            double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1 + tau * tau));
            double c = 1 / sqrt(1 + t * t), s = t * c;
            for (int k = 0; k < m; ++k) {
                double xp = A[k * lda + p], xq = A[k * lda + q];
                A[k * lda + p] = c * xp - s * xq; A[k * lda + q] = s * xp + c * xq;
                double qp = Q[k * m + p], qq = Q[k * m + q];
                Q[k * m + p] = c * qp - s * qq; Q[k * m + q] = s * qp + c * qq;
            }
            for (int k = 0; k < m; ++k) {
                double xp = A[p * lda + k], xq = A[q * lda + k];
                A[p * lda + k] = c * xp - s * xq; A[q * lda + k] = s * xp + c * xq;
            }
        }
    }
    int best = 0;
    for (int i = 1; i < m; ++i) if (A[i * lda + i] < A[best * lda + best]) best = i;
    *theta = A[best * lda + best];
    for (int i = 0; i < m; ++i) y[i] = Q[i * m + best];
}

static double dotp(const double* x, const double* y, size_t n) {
    double s = 0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (size_t i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

typedef void (*sigma_fn)(const void*, const double*, double*);

static int davidson(sigma_fn apply, const void* ctx, const double* hdiag, size_t n, double tol,
                    int max_cycle, int max_space, double* x, double* theta_out, int* cycles_out) {
    double* V = malloc(sizeof(double) * n * max_space);
    double* W = malloc(sizeof(double) * n * max_space);
    double* t = malloc(sizeof(double) * n);
    double* hx = malloc(sizeof(double) * n);
    double G[64 * 64], A[64 * 64], y[64];
    memset(V, 0, sizeof(double) * n);
    size_t k0 = 0;
    for (size_t i = 1; i < n; ++i) if (hdiag[i] < hdiag[k0]) k0 = i;
    V[k0] = 1.0; V[0] += 1e-5; V[n - 1] -= 1e-5;
    double nrm = sqrt(dotp(V, V, n));
    for (size_t i = 0; i < n; ++i) V[i] /= nrm;
    int m = 1, conv = 0, cyc = 0;
    double theta = 0, theta_prev = INFINITY;
    apply(ctx, V, W);
    for (cyc = 0; cyc < max_cycle; ++cyc) {
        const int slot = m - 1;
        for (int i = 0; i < m; ++i) G[i * 64 + slot] = G[slot * 64 + i] = dotp(V + (size_t)i * n, W + (size_t)slot * n, n);
        for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) A[i * 64 + j] = G[i * 64 + j];
        jacobi_lowest(A, m, 64, y, &theta);
        double rr = 0;
#pragma omp parallel for reduction(+ : rr) schedule(static)
        for (size_t j = 0; j < n; ++j) {
            double xv = 0, hv = 0;
            for (int i = 0; i < m; ++i) { xv += y[i] * V[(size_t)i * n + j]; hv += y[i] * W[(size_t)i * n + j]; }
            x[j] = xv; hx[j] = hv;
            double r = hv - theta * xv;
            rr += r * r;
            double den = hdiag[j] - theta + 1e-4;
            if (fabs(den) < 1e-8) den = den < 0 ? -1e-8 : 1e-8;
            t[j] = r / den;
        }
        if (fabs(theta - theta_prev) < tol && sqrt(rr) < sqrt(tol)) { conv = 1; ++cyc; break; }
        theta_prev = theta;
        if (m == max_space) {
            memcpy(V, x, sizeof(double) * n); memcpy(W, hx, sizeof(double) * n);
            G[0] = theta; m = 1;
        }
        for (int pass = 0; pass < 2; ++pass)
            for (int i = 0; i < m; ++i) {
                double cf = dotp(V + (size_t)i * n, t, n);
#pragma omp parallel for schedule(static)
                for (size_t j = 0; j < n; ++j) t[j] -= cf * V[(size_t)i * n + j];
            }
        double tn = dotp(t, t, n);
        if (!(tn > 1e-14)) { conv = 2; ++cyc; break; }
        tn = 1 / sqrt(tn);
#pragma omp parallel for schedule(static)
        for (size_t j = 0; j < n; ++j) V[(size_t)m * n + j] = t[j] * tn;
        apply(ctx, V + (size_t)m * n, W + (size_t)m * n);
        ++m;
    }
    *theta_out = theta; *cycles_out = cyc;
    free(V); free(W); free(t); free(hx);
    return conv;
}

static void apply_direct(const void* ctx, const double* c, double* o) { direct_sigma((const direct_t*)ctx, c, o); }
static void apply_pyscf(const void* ctx, const double* c, double* o) { pyscf_sigma((const pyscf_t*)ctx, c, o); }

/* ------------------------------------ public entry points -------------------------------------- */
/* Ground state in A x B.  algo 0 = direct, 1 = pyscf-style.  spin penalty: shift*(S^2 - ss) when
 * use_penalty (linear form only, algo 0).  Outputs: amps[na*nb], energy (bare-H expectation),
 * occ[2*norb], info[0]=converged, info[1]=cycles.  max_cycle <= 0 -> only time `-max_cycle` sigma
 * builds (bounded CPU-baseline sample) and return. */
int sci_cpu_solve(const uint64_t* sa, int na, const uint64_t* sb, int nb, int norb, const double* h,
                  const double* g, int algo, int use_penalty, double shift, double ss, double tol,
                  int max_cycle, int max_space, int nthreads, double* amps, double* energy,
                  double* occ, double* theta_out, int* info) {
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    if (max_space > 64) max_space = 64;
    const size_t n = (size_t)na * nb;
    direct_t D;
    direct_setup(&D, sa, na, sb, nb, norb, h, g, use_penalty ? shift : 0.0, use_penalty ? ss : 0.0);
    pyscf_t P;
    if (algo == 1) pyscf_setup(&P, sa, na, sb, nb, norb, h, g, D.hdiag);
    sigma_fn fn = algo == 1 ? apply_pyscf : apply_direct;
    const void* ctx = algo == 1 ? (const void*)&P : (const void*)&D;
    int conv = 0, cyc = 0;
    double theta = 0;
    if (max_cycle <= 0) {
        /* timing sample: -max_cycle sigma builds on a normalised vector */
        double* x = malloc(sizeof(double) * n);
        double* y = malloc(sizeof(double) * n);
        for (size_t i = 0; i < n; ++i) x[i] = 1.0 / sqrt((double)n);
        for (int it = 0; it < -max_cycle; ++it) fn(ctx, x, y);
        memcpy(amps, y, sizeof(double) * n);
        free(x); free(y);
        cyc = -max_cycle;
    } else {
        conv = davidson(fn, ctx, D.hdiag, n, tol, max_cycle, max_space, amps, &theta, &cyc);
        double* hx = malloc(sizeof(double) * n);
        fn(ctx, amps, hx);
        double e = dotp(amps, hx, n) / dotp(amps, amps, n);
        if (use_penalty && algo == 0) {
            /* remove the penalty expectation: needs <S^2>; apply the S^2 operator through a second setup */
            direct_t S;
            double* zero_h = calloc((size_t)norb * norb, sizeof(double));
            double* zero_g = calloc((size_t)norb * norb * norb * norb, sizeof(double));
            direct_setup(&S, sa, na, sb, nb, norb, zero_h, zero_g, 1.0, 0.0);
            direct_sigma(&S, amps, hx);
            double s2 = dotp(amps, hx, n) / dotp(amps, amps, n);
            e -= shift * (s2 - ss);
            direct_free(&S); free(zero_h); free(zero_g);
        }
        *energy = e;
        free(hx);
        for (int p = 0; p < 2 * norb; ++p) occ[p] = 0;
        for (int a = 0; a < na; ++a) for (int b = 0; b < nb; ++b) {
            const double w = amps[(size_t)a * nb + b] * amps[(size_t)a * nb + b];
            for (uint64_t o = sa[a]; o; o &= o - 1) occ[lowbit(o)] += w;
            for (uint64_t o = sb[b]; o; o &= o - 1) occ[norb + lowbit(o)] += w;
        }
    }
    *theta_out = theta;
    info[0] = conv; info[1] = cyc;
    if (algo == 1) pyscf_free(&P);
    direct_free(&D);
    return 0;
}

/* sigma = H c for validation of both algorithms against the dense oracle */
int sci_cpu_sigma(const uint64_t* sa, int na, const uint64_t* sb, int nb, int norb, const double* h,
                  const double* g, int algo, const double* c, double* out) {
    direct_t D;
    direct_setup(&D, sa, na, sb, nb, norb, h, g, 0.0, 0.0);
    if (algo == 1) {
        pyscf_t P;
        pyscf_setup(&P, sa, na, sb, nb, norb, h, g, D.hdiag);
        pyscf_sigma(&P, c, out);
        pyscf_free(&P);
    } else {
        direct_sigma(&D, c, out);
    }
    direct_free(&D);
    return 0;
}

"""CPU oracle for the fermionic subspace projection + diagonalisation.  TEST INFRASTRUCTURE ONLY.

PARITY STATUS: **parity unpinned against pyscf**.  The arithmetic of the reference's fermion path
lives in a third-party dependency that is absent from ``/root/reference`` and from this image:
``pyscf >= 2.9`` (``pyproject.toml:30``, no upper pin, no lock file), reached through
``fci.selected_ci.kernel_fixed_space`` / ``make_rdm1s`` / ``make_rdm1`` / ``make_rdm2`` /
``spin_square`` (``qiskit_addon_sqd/fermion.py:713-729, 803-830``).  The reference's own tests for
this path (``test/test_fermion.py:54-342``) need pyscf to build their inputs, so no golden vector of
the reference can be produced here.  This file therefore restates the *mathematics* the reference
call sites define (projected Hamiltonian in the product space of alpha and beta strings, energy as the
bare-H expectation ``fermion.py:824-827``, occupancies as diagonals of the spin 1-RDMs
``fermion.py:821-822``, <S^2> ``fermion.py:830``) and pins it against an INDEPENDENT construction:
a Jordan-Wigner many-body matrix on 2*norb qubit modes (``jordan_wigner_hamiltonian``), plus the
physical invariants (full product space == FCI, tr(dm1) == N, E == Ritz value without spin penalty).

Conventions (``fermion.py:61-65, 234-237, 1035``):
  * bit i of a string <=> spatial orbital i occupied;
  * E_pq |s> = a+_p a_q |s> = sign * |s - q + p>, sign = (-1)^(occupied orbitals strictly between);
  * determinant |a, b>, flat index a_idx * nb + b_idx; amplitudes are an (na, nb) C-order matrix;
  * eri in chemist order (pq|rs) with the 8-fold symmetry of real orbitals.
"""

from __future__ import annotations

import itertools
from functools import reduce

import numpy as np


# --------------------------------------------------------------------------------------------
# string algebra
# --------------------------------------------------------------------------------------------
def popcount(x: int) -> int:
    return bin(int(x)).count("1")


def apply_excitation(p: int, q: int, s: int):
    """``a+_p a_q |s>`` -> ``(sign, t)`` or ``None`` when it vanishes."""
    s = int(s)
    if not (s >> q) & 1:
        return None
    t = s ^ (1 << q)
    if (t >> p) & 1:
        return None
    t |= 1 << p
    lo, hi = (p, q) if p < q else (q, p)
    between = ((1 << hi) - 1) & ~((1 << (lo + 1)) - 1)
    sign = -1 if popcount(s & between) & 1 else 1
    return sign, t


def annihilate(k: int, s: int):
    if not (s >> k) & 1:
        return None
    sign = -1 if popcount(s & ((1 << k) - 1)) & 1 else 1
    return sign, s ^ (1 << k)


def create(k: int, s: int):
    if (s >> k) & 1:
        return None
    sign = -1 if popcount(s & ((1 << k) - 1)) & 1 else 1
    return sign, s | (1 << k)


def single_excitation_links(strs, norb: int, include_diagonal: bool = True):
    """All ``<t|E_pq|s>`` with s, t in ``strs``: list of ``(t_idx, s_idx, p, q, sign)``."""
    strs = [int(s) for s in strs]
    pos = {s: i for i, s in enumerate(strs)}
    out = []
    for i, s in enumerate(strs):
        for q in range(norb):
            if not (s >> q) & 1:
                continue
            for p in range(norb):
                if p == q:
                    if include_diagonal:
                        out.append((i, i, p, q, 1))
                    continue
                r = apply_excitation(p, q, s)
                if r is not None and r[1] in pos:
                    out.append((pos[r[1]], i, p, q, r[0]))
    return out


# --------------------------------------------------------------------------------------------
# same-spin block: Slater-Condon rules (complete intermediate states)
# --------------------------------------------------------------------------------------------
def same_spin_diagonal(s: int, h: np.ndarray, g: np.ndarray, norb: int) -> float:
    occ = [i for i in range(norb) if (s >> i) & 1]
    e = sum(h[i, i] for i in occ)
    for i in occ:
        for j in occ:
            e += 0.5 * (g[i, i, j, j] - g[i, j, j, i])
    return float(e)


def same_spin_matrix(strs, h: np.ndarray, g: np.ndarray, norb: int) -> np.ndarray:
    """``P_S [sum h_pq E_pq + 1/2 sum (pq|rs)(E_pq E_rs - d_qr E_ps)] P_S`` by Slater-Condon rules."""
    strs = [int(s) for s in strs]
    n = len(strs)
    H = np.zeros((n, n))
    for j, s in enumerate(strs):
        occ = [i for i in range(norb) if (s >> i) & 1]
        for i2, t in enumerate(strs):
            x = s ^ t
            nx = popcount(x)
            if nx == 0:
                H[i2, j] = same_spin_diagonal(s, h, g, norb)
            elif nx == 2:
                # single i -> a
                i = (x & s).bit_length() - 1
                a = (x & t).bit_length() - 1
                sign, t2 = apply_excitation(a, i, s)
                assert t2 == t
                v = h[a, i]
                for k in occ:
                    v += g[a, i, k, k] - g[a, k, k, i]
                H[i2, j] = sign * v
            elif nx == 4:
                holes = [k for k in range(norb) if ((x & s) >> k) & 1]  # i < j occupied in s
                parts = [k for k in range(norb) if ((x & t) >> k) & 1]  # a < b occupied in t
                i, jj = holes
                a, b = parts
                # phase of a+_a a+_b a_j a_i applied to |s>
                sg1, u = annihilate(i, s)
                sg2, u = annihilate(jj, u)
                sg3, u = create(b, u)
                sg4, u = create(a, u)
                assert u == t
                H[i2, j] = sg1 * sg2 * sg3 * sg4 * (g[a, i, b, jj] - g[a, jj, b, i])
    return H


def same_spin_matrix_complete(strs, h: np.ndarray, g: np.ndarray, norb: int) -> np.ndarray:
    """Same operator through explicit E_pq matrices on the COMPLETE N-electron string space.

    Exponential in norb -- validation of ``same_spin_matrix`` for norb <= 8 only.
    """
    strs = [int(s) for s in strs]
    nel = popcount(strs[0])
    all_strs = sorted(sum(1 << i for i in c) for c in itertools.combinations(range(norb), nel))
    pos = {s: i for i, s in enumerate(all_strs)}
    N = len(all_strs)
    E = np.zeros((norb, norb, N, N))
    for j, s in enumerate(all_strs):
        for p in range(norb):
            for q in range(norb):
                r = apply_excitation(p, q, s) if p != q else ((1, s) if (s >> q) & 1 else None)
                if r is not None:
                    E[p, q, pos[r[1]], j] = r[0]
    Hs = np.einsum("pq,pqij->ij", h, E)
    Hs += 0.5 * np.einsum("pqrs,pqik,rskj->ij", g, E, E, optimize=True)
    Hs -= 0.5 * np.einsum("pqqs,psij->ij", g, E)
    sel = [pos[s] for s in strs]
    return Hs[np.ix_(sel, sel)]


# --------------------------------------------------------------------------------------------
# projected Hamiltonian and S^2 in the product space A x B
# --------------------------------------------------------------------------------------------
def projected_hamiltonian(strs_a, strs_b, h, g, norb: int, g_ab=None) -> np.ndarray:
    """Dense ``H_sub``.  ``g_ab`` overrides the opposite-spin tensor."""
    g_ab = g if g_ab is None else g_ab
    na, nb = len(strs_a), len(strs_b)
    Ha = same_spin_matrix(strs_a, h, g, norb)
    Hb = same_spin_matrix(strs_b, h, g, norb)
    H = np.kron(Ha, np.eye(nb)) + np.kron(np.eye(na), Hb)
    la = single_excitation_links(strs_a, norb)
    lb = single_excitation_links(strs_b, norb)
    # group by operator index to vectorise:  sum_pq Ea_pq (x) (sum_rs g[pq,rs] Eb_rs)
    Eb = {}
    for (tb, sb, r, s, sg) in lb:
        Eb.setdefault((r, s), []).append((tb, sb, sg))
    Ea = {}
    for (ta, sa, p, q, sg) in la:
        Ea.setdefault((p, q), []).append((ta, sa, sg))
    for (p, q), alist in Ea.items():
        Gb = np.zeros((nb, nb))
        for (r, s), blist in Eb.items():
            gv = g_ab[p, q, r, s]
            if gv != 0.0:
                for (tb, sb, sg) in blist:
                    Gb[tb, sb] += gv * sg
        Am = np.zeros((na, na))
        for (ta, sa, sg) in alist:
            Am[ta, sa] += sg
        H += np.kron(Am, Gb)
    return H


def spin_square_matrix(strs_a, strs_b, norb: int) -> np.ndarray:
    """``P S^2 P`` with ``S^2 = Sz(Sz+1) + N_b - sum_pq Ea_pq Eb_qp``."""
    na, nb = len(strs_a), len(strs_b)
    n_a, n_b = popcount(strs_a[0]), popcount(strs_b[0])
    sz = 0.5 * (n_a - n_b)
    S2 = (sz * (sz + 1) + n_b) * np.eye(na * nb)
    la = single_excitation_links(strs_a, norb)
    lb = single_excitation_links(strs_b, norb)
    Eb = {}
    for (tb, sb, r, s, sg) in lb:
        Eb.setdefault((r, s), []).append((tb, sb, sg))
    for (ta, sa, p, q, sga) in la:
        for (tb, sb, sgb) in Eb.get((q, p), []):
            S2[ta * nb + tb, sa * nb + sb] -= sga * sgb
    return S2


def make_hdiag(strs_a, strs_b, h, g, norb: int) -> np.ndarray:
    """Diagonal of ``H_sub`` as an (na, nb) array (pyscf ``make_hdiag`` equivalent, Appendix A)."""
    da = np.array([same_spin_diagonal(int(s), h, g, norb) for s in strs_a])
    db = np.array([same_spin_diagonal(int(s), h, g, norb) for s in strs_b])
    J = np.einsum("iijj->ij", g)
    oa = np.array([[(int(s) >> i) & 1 for i in range(norb)] for s in strs_a], dtype=float)
    ob = np.array([[(int(s) >> i) & 1 for i in range(norb)] for s in strs_b], dtype=float)
    return da[:, None] + db[None, :] + oa @ J @ ob.T


# --------------------------------------------------------------------------------------------
# independent construction: Jordan-Wigner on 2*norb modes
# --------------------------------------------------------------------------------------------
def jordan_wigner_operators(norb: int):
    n = 2 * norb
    I2 = np.eye(2)
    Z = np.diag([1.0, -1.0])
    sm = np.array([[0.0, 1.0], [0.0, 0.0]])

    def kron(ops):
        return reduce(np.kron, ops)

    a = [kron([Z] * k + [sm] + [I2] * (n - k - 1)) for k in range(n)]
    return a, [x.T for x in a]


def jordan_wigner_hamiltonian(h, g, norb: int) -> np.ndarray:
    """Full 4^norb many-body matrix; mode k = orbital k (alpha) or norb + k (beta).  norb <= 5."""
    a, ad = jordan_wigner_operators(norb)
    n = 2 * norb
    H = np.zeros((2**n, 2**n))
    for s in (0, 1):
        for p in range(norb):
            for q in range(norb):
                H += h[p, q] * ad[s * norb + p] @ a[s * norb + q]
    for s in (0, 1):
        for t in (0, 1):
            for p, q, r, u in itertools.product(range(norb), repeat=4):
                if g[p, q, r, u] != 0.0:
                    H += (
                        0.5
                        * g[p, q, r, u]
                        * ad[s * norb + p]
                        @ ad[t * norb + r]
                        @ a[t * norb + u]
                        @ a[s * norb + q]
                    )
    return H


def jordan_wigner_spin_square(norb: int) -> np.ndarray:
    a, ad = jordan_wigner_operators(norb)
    Sp = sum(ad[p] @ a[norb + p] for p in range(norb))
    Na = sum(ad[p] @ a[p] for p in range(norb))
    Nb = sum(ad[norb + p] @ a[norb + p] for p in range(norb))
    Sz = 0.5 * (Na - Nb)
    return Sp.T @ Sp + Sz @ Sz + Sz


def jordan_wigner_index(astr: int, bstr: int, norb: int) -> int:
    idx = 0
    for k in range(2 * norb):
        occ = (int(astr) >> k) & 1 if k < norb else (int(bstr) >> (k - norb)) & 1
        idx = idx * 2 + occ
    return idx


# --------------------------------------------------------------------------------------------
# the solve, mirroring the order of steps of fermion.solve_fermion (fermion.py:745-845)
# --------------------------------------------------------------------------------------------
def occupancies_from_amplitudes(c: np.ndarray, strs_a, strs_b, norb: int):
    """Diagonals of the spin-resolved 1-RDMs (``fermion.py:821-822``), orbital 0 first."""
    wa = (c * c).sum(axis=1)
    wb = (c * c).sum(axis=0)
    occ_a = np.array([sum(w for w, s in zip(wa, strs_a) if (int(s) >> p) & 1) for p in range(norb)])
    occ_b = np.array([sum(w for w, s in zip(wb, strs_b) if (int(s) >> p) & 1) for p in range(norb)])
    return occ_a, occ_b


def solve_dense(strs_a, strs_b, h, g, norb: int, spin_sq=None, shift: float = 0.1):
    """Dense-eigh ground state in A x B.

    Returns ``(energy, amplitudes(na, nb), (occ_a, occ_b), spin_square, ritz_value)`` where
    ``energy`` is the bare-H expectation (``fermion.py:806-809, 824-827``) and the spin penalty is
    pyscf ``fix_spin_``'s: ``shift*(S^2 - ss)`` when ``ss < sz(sz+1)+0.1`` else ``shift*(S^2-ss)^2``.
    """
    strs_a = [int(s) for s in strs_a]
    strs_b = [int(s) for s in strs_b]
    na, nb = len(strs_a), len(strs_b)
    H = projected_hamiltonian(strs_a, strs_b, h, g, norb)
    S2 = spin_square_matrix(strs_a, strs_b, norb)
    Hp = H
    if spin_sq is not None:
        sz = 0.5 * abs(popcount(strs_a[0]) - popcount(strs_b[0]))
        D = S2 - spin_sq * np.eye(na * nb)
        Hp = H + shift * (D if spin_sq < sz * (sz + 1) + 0.1 else D @ D)
    w, v = np.linalg.eigh(Hp)
    c = v[:, 0]
    k = int(np.argmax(np.abs(c)))
    if c[k] < 0:
        c = -c
    energy = float(c @ H @ c)
    s2 = float(c @ S2 @ c)
    cm = c.reshape(na, nb)
    return energy, cm, occupancies_from_amplitudes(cm, strs_a, strs_b, norb), s2, float(w[0])


# --------------------------------------------------------------------------------------------
# medium sizes: matrix-free operator on scipy sparse blocks (independent of the C/CUDA code)
# --------------------------------------------------------------------------------------------
def _links_arrays(strs, norb: int):
    """Vectorised in-set single-excitation links incl. diagonal: arrays (t, s, p, q, sign)."""
    strs = np.asarray(strs, dtype=np.uint64)
    n = len(strs)
    T, S, P, Q, SG = [], [], [], [], []
    idx = np.arange(n)
    one = np.uint64(1)
    for q in range(norb):
        occq = ((strs >> np.uint64(q)) & one).astype(bool)
        for p in range(norb):
            if p == q:
                T.append(idx[occq]); S.append(idx[occq])
                P.append(np.full(occq.sum(), p)); Q.append(np.full(occq.sum(), q))
                SG.append(np.ones(occq.sum()))
                continue
            ok = occq & ~((strs >> np.uint64(p)) & one).astype(bool)
            if not ok.any():
                continue
            src = idx[ok]
            t = (strs[ok] ^ (one << np.uint64(q))) | (one << np.uint64(p))
            j = np.searchsorted(strs, t)
            j[j >= n] = n - 1
            hit = strs[j] == t
            if not hit.any():
                continue
            lo, hi = min(p, q), max(p, q)
            mask = np.uint64(((1 << hi) - 1) & ~((1 << (lo + 1)) - 1))
            between = strs[ok][hit] & mask
            par = np.zeros(between.shape, dtype=np.int64)
            b = between.copy()
            while b.any():
                par += (b & one).astype(np.int64)
                b >>= one
            T.append(j[hit]); S.append(src[hit])
            P.append(np.full(hit.sum(), p)); Q.append(np.full(hit.sum(), q))
            SG.append(1.0 - 2.0 * (par & 1))
    return (np.concatenate(T), np.concatenate(S), np.concatenate(P), np.concatenate(Q),
            np.concatenate(SG))


def same_spin_sparse(strs, h, g, norb: int):
    """Sparse same-spin block by all-pairs xor/popcount (vectorised Slater-Condon)."""
    import scipy.sparse as sp

    strs = np.asarray(strs, dtype=np.uint64)
    n = len(strs)
    occ = ((strs[:, None] >> np.arange(norb, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(float)
    J = np.einsum("iijj->ij", g)
    K = np.einsum("ijji->ij", g)
    diag = occ @ np.diag(h) + 0.5 * np.einsum("si,ij,sj->s", occ, J - K, occ)
    rows, cols, vals = [np.arange(n)], [np.arange(n)], [diag]
    for j in range(n):
        s = int(strs[j])
        x = strs ^ strs[j]
        nx = np.array([popcount(v) for v in x])
        for i2 in np.nonzero((nx == 2) | (nx == 4))[0]:
            t = int(strs[i2])
            xx = s ^ t
            if nx[i2] == 2:
                i = (xx & s).bit_length() - 1
                a = (xx & t).bit_length() - 1
                sign, _ = apply_excitation(a, i, s)
                v = h[a, i] + occ[j] @ (g[a, i].diagonal() - g[a, :, :, i].diagonal())
                vals.append([sign * v])
            else:
                holes = [k for k in range(norb) if ((xx & s) >> k) & 1]
                parts = [k for k in range(norb) if ((xx & t) >> k) & 1]
                i, jj = holes
                a, b = parts
                sg1, u = annihilate(i, s)
                sg2, u = annihilate(jj, u)
                sg3, u = create(b, u)
                sg4, u = create(a, u)
                vals.append([sg1 * sg2 * sg3 * sg4 * (g[a, i, b, jj] - g[a, jj, b, i])])
            rows.append([i2]); cols.append([j])
    return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), (n, n))


class SparseProjectedHamiltonian:
    """Matrix-free ``H_sub`` for n_det up to ~1e5: ``sigma = Ha C + C Hb^T + sum_pq Ea_pq C Gb_pq^T``."""

    def __init__(self, strs_a, strs_b, h, g, norb: int, g_ab=None):
        import scipy.sparse as sp

        g_ab = g if g_ab is None else g_ab
        self.na, self.nb = len(strs_a), len(strs_b)
        self.Ha = same_spin_sparse(strs_a, h, g, norb)
        self.Hb = same_spin_sparse(strs_b, h, g, norb)
        ta, sa, pa, qa, sga = _links_arrays(strs_a, norb)
        tb, sb, pb, qb, sgb = _links_arrays(strs_b, norb)
        self.terms = []
        pq_a = pa * norb + qa
        g2 = np.asarray(g_ab).reshape(norb * norb, norb * norb)
        rs_b = pb * norb + qb
        for pq in np.unique(pq_a):
            m = pq_a == pq
            Ea = sp.csr_matrix((sga[m], (ta[m], sa[m])), (self.na, self.na))
            Gb = sp.csr_matrix((sgb * g2[pq, rs_b], (tb, sb)), (self.nb, self.nb))
            self.terms.append((Ea, Gb.T.tocsr()))

    def matvec(self, x):
        C = np.asarray(x).reshape(self.na, self.nb)
        out = self.Ha @ C + (self.Hb @ C.T).T
        for Ea, GbT in self.terms:
            out += (Ea @ C) @ GbT
        return out.reshape(-1)

    def ground_state(self, tol: float = 1e-12):
        from scipy.sparse.linalg import LinearOperator, eigsh

        n = self.na * self.nb
        op = LinearOperator((n, n), matvec=self.matvec, dtype=float)
        w, v = eigsh(op, k=1, which="SA", tol=tol)
        return float(w[0]), v[:, 0].reshape(self.na, self.nb)


def rdm1s(c: np.ndarray, strs_a, strs_b, norb: int):
    """Spin-resolved 1-RDMs ``(dm1a, dm1b)``, ``dm1[p, q] = <c| a+_q a_p |c>`` (pyscf's convention,
    ``make_rdm1s`` reached from ``fermion.py:117-121``)."""
    c = np.asarray(c)
    dma = np.zeros((norb, norb))
    dmb = np.zeros((norb, norb))
    for (t, s, p, q, sg) in single_excitation_links(strs_a, norb):
        dma[q, p] += sg * float(c[t] @ c[s])
    for (t, s, p, q, sg) in single_excitation_links(strs_b, norb):
        dmb[q, p] += sg * float(c[:, t] @ c[:, s])
    return dma, dmb


def rdm2s(c: np.ndarray, strs_a, strs_b, norb: int):
    """Spin-separated 2-RDMs ``(dm2aa, dm2ab, dm2bb)`` in pyscf's convention
    ``dm2[p, q, r, s] = <c| p^+ r^+ s q |c>`` (``selected_ci.make_rdm2s``, reached from
    ``fermion.py:117-128`` and ``:728-729``); for ``dm2ab`` p, q are alpha and r, s beta orbitals.

    Brute force on the 2*norb spin-orbital occupation strings (alpha modes first, then beta: the
    determinant ordering of SURVEY Appendix B.1), operator by operator with ``annihilate`` / ``create``
    -- independent of the excitation-table construction the CUDA kernels use.  Small cases only."""
    c = np.asarray(c, dtype=float)
    sa = [int(x) for x in strs_a]
    sb = [int(x) for x in strs_b]
    index = {(a | (b << norb)): (i, j) for i, a in enumerate(sa) for j, b in enumerate(sb)}
    out = {k: np.zeros((norb,) * 4) for k in ("aa", "ab", "bb")}
    off = {"a": 0, "b": norb}
    for det, (i, j) in index.items():
        ck = c[i, j]
        if ck == 0.0:
            continue
        for key in ("aa", "ab", "bb"):
            o1, o2 = off[key[0]], off[key[1]]   # spin of (p, q) and of (r, s)
            for q in range(norb):
                r1 = annihilate(q + o1, det)
                if r1 is None:
                    continue
                for s in range(norb):
                    r2 = annihilate(s + o2, r1[1])
                    if r2 is None:
                        continue
                    for r in range(norb):
                        r3 = create(r + o2, r2[1])
                        if r3 is None:
                            continue
                        for p in range(norb):
                            r4 = create(p + o1, r3[1])
                            if r4 is None:
                                continue
                            tgt = index.get(r4[1])
                            if tgt is None:
                                continue
                            sign = r1[0] * r2[0] * r3[0] * r4[0]
                            out[key][p, q, r, s] += sign * c[tgt] * ck
    return out["aa"], out["ab"], out["bb"]


def rdm2(c: np.ndarray, strs_a, strs_b, norb: int) -> np.ndarray:
    """Spin-summed 2-RDM (pyscf ``make_rdm2``): ``aa + bb + ab + ab.transpose(2, 3, 0, 1)``."""
    aa, ab, bb = rdm2s(c, strs_a, strs_b, norb)
    return aa + bb + ab + ab.transpose(2, 3, 0, 1)

"""Closed-form STO-3G integrals for chains of hydrogen atoms -- an EXTERNAL pin for the fermion oracle.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference pins its fermion path with pyscf-built
molecular integrals (``test/test_fermion.py:54-125``: N2 CASCI energy to two decimals); pyscf is not
available here, so the pin is rebuilt from textbook formulas that need no integral library:

* basis: one contracted s function per hydrogen, STO-3G with zeta = 1.24 (Hehre, Stewart, Pople 1969;
  Szabo & Ostlund, *Modern Quantum Chemistry*, section 3.5.1, eqs. 3.219-3.221 and Appendix A);
* integrals over s-type Gaussians in closed form (Szabo & Ostlund Appendix A, eqs. A.9, A.11, A.33, A.41)
  with the Boys function F0(t) = (1/2) sqrt(pi/t) erf(sqrt(t));
* literature values for H2 at R = 1.4 a0 in this basis (Szabo & Ostlund, Tables 3.x / section 4.1):
  E_HF = -1.1167 Ha, full CI E = -1.1373 Ha (E_corr = -0.0206 Ha), both including 1/R.

The integrals are returned in a symmetrically orthonormalised orbital basis (the full-CI energy does not
depend on the choice of orthonormal basis), in the conventions of ``solve_fermion``: ``hcore[p,q]`` and
``eri[p,q,r,s] = (pq|rs)`` (chemist order, ``fermion.py:827``).
"""

from __future__ import annotations

import math

import numpy as np

# STO-3G fit of a Slater 1s function with zeta = 1.0 (Szabo & Ostlund eq. 3.221); exponents scale with zeta^2
_ALPHA_1 = np.array([2.227660584, 0.405771156, 0.109818036])
_COEF = np.array([0.154328967, 0.535328142, 0.444634542])
ZETA_H = 1.24

#: literature values for H2, R = 1.4 a0, STO-3G (zeta = 1.24), total energies in Hartree
H2_R = 1.4
H2_E_HF = -1.1167
H2_E_FCI = -1.1373


def _f0(t: float) -> float:
    if t < 1e-12:
        return 1.0 - t / 3.0
    return 0.5 * math.sqrt(math.pi / t) * math.erf(math.sqrt(t))


def _primitives(zeta: float):
    a = _ALPHA_1 * zeta * zeta
    norm = (2.0 * a / math.pi) ** 0.75
    return a, _COEF * norm


def hydrogen_chain_ao(n_atoms: int, spacing: float, zeta: float = ZETA_H):
    """Overlap, kinetic + nuclear-attraction and two-electron integrals over the n contracted 1s
    functions of an equally spaced linear H_n chain (atomic units).  Returns (S, hcore, eri, e_nuc)."""
    pos = np.array([[0.0, 0.0, spacing * i] for i in range(n_atoms)])
    a, d = _primitives(zeta)
    n = n_atoms
    S = np.zeros((n, n))
    T = np.zeros((n, n))
    V = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            r2 = float(np.sum((pos[i] - pos[j]) ** 2))
            for p in range(3):
                for q in range(3):
                    ap, aq = a[p], a[q]
                    s = ap + aq
                    mu = ap * aq / s
                    pref = d[p] * d[q] * (math.pi / s) ** 1.5 * math.exp(-mu * r2)
                    S[i, j] += pref
                    T[i, j] += pref * mu * (3.0 - 2.0 * mu * r2)
                    rp = (ap * pos[i] + aq * pos[j]) / s
                    for c in range(n):
                        rpc2 = float(np.sum((rp - pos[c]) ** 2))
                        V[i, j] += -d[p] * d[q] * (2.0 * math.pi / s) * math.exp(-mu * r2) * _f0(s * rpc2)
    eri = np.zeros((n, n, n, n))
    for i in range(n):
        for j in range(n):
            rij2 = float(np.sum((pos[i] - pos[j]) ** 2))
            for k in range(n):
                for l in range(n):
                    rkl2 = float(np.sum((pos[k] - pos[l]) ** 2))
                    val = 0.0
                    for p in range(3):
                        for q in range(3):
                            s1 = a[p] + a[q]
                            rp = (a[p] * pos[i] + a[q] * pos[j]) / s1
                            e1 = math.exp(-a[p] * a[q] / s1 * rij2)
                            for r in range(3):
                                for t in range(3):
                                    s2 = a[r] + a[t]
                                    rq = (a[r] * pos[k] + a[t] * pos[l]) / s2
                                    e2 = math.exp(-a[r] * a[t] / s2 * rkl2)
                                    rpq2 = float(np.sum((rp - rq) ** 2))
                                    val += (d[p] * d[q] * d[r] * d[t] * 2.0 * math.pi ** 2.5
                                            / (s1 * s2 * math.sqrt(s1 + s2)) * e1 * e2
                                            * _f0(s1 * s2 / (s1 + s2) * rpq2))
                    eri[i, j, k, l] = val
    e_nuc = sum(1.0 / (spacing * abs(i - j)) for i in range(n) for j in range(i))
    return S, T + V, eri, e_nuc


def hydrogen_chain(n_atoms: int, spacing: float, zeta: float = ZETA_H):
    """(hcore, eri, e_nuc) in the symmetrically orthonormalised basis X = S^(-1/2)."""
    S, h, g, e_nuc = hydrogen_chain_ao(n_atoms, spacing, zeta)
    w, U = np.linalg.eigh(S)
    X = U @ np.diag(w ** -0.5) @ U.T
    h_o = X.T @ h @ X
    g_o = np.einsum("ap,bq,cr,ds,abcd->pqrs", X, X, X, X, g, optimize=True)
    return np.ascontiguousarray(h_o), np.ascontiguousarray(g_o), e_nuc


def h2_rhf_energy(spacing: float = H2_R, zeta: float = ZETA_H) -> float:
    """Restricted Hartree-Fock energy of H2 in the minimal basis: the occupied orbital is fixed by symmetry
    (sigma_g = (phi_1 + phi_2) / sqrt(2 + 2 S12)), so no SCF iteration is needed."""
    S, h, g, e_nuc = hydrogen_chain_ao(2, spacing, zeta)
    c = np.array([1.0, 1.0]) / math.sqrt(2.0 + 2.0 * S[0, 1])
    h11 = c @ h @ c
    j11 = np.einsum("a,b,c,d,abcd->", c, c, c, c, g)
    return 2.0 * h11 + j11 + e_nuc

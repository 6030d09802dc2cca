"""``pyscf.fci.selected_ci`` stand-in backed by the dense CPU oracle (small subspaces only).

Conventions follow pyscf (recalled): ``kernel_fixed_space`` returns the eigenvalue of
the (possibly spin-penalised) operator and a CI vector carrying ``_strs``; ``make_rdm1s`` is
``dm1[p, q] = <q^+ p>``; ``make_rdm2`` is ``dm2[p, q, r, s] = <p^+ r^+ s q>``; ``spin_square`` returns
``(<S^2>, multiplicity)``.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(
    os.path.abspath(__file__)))))))
from oracle import fermion_oracle as fo  # noqa: E402


class SCIvector(np.ndarray):
    """ndarray that remembers its string lists (pyscf ``_SCIvector``)."""

    def __array_finalize__(self, obj):
        self._strs = getattr(obj, "_strs", None)


def _as_SCIvector(civec, ci_strs):
    civec = np.asarray(civec).view(SCIvector)
    civec._strs = ci_strs
    return civec


def _norb_of(civec_strs, norb):
    return int(norb)


def make_rdm1s(civec_strs, norb, nelec, link_index=None):
    sa, sb = civec_strs._strs
    return fo.rdm1s(np.asarray(civec_strs), sa, sb, int(norb))


def make_rdm1(civec_strs, norb, nelec, link_index=None):
    a, b = make_rdm1s(civec_strs, norb, nelec)
    return a + b


def make_rdm2s(civec_strs, norb, nelec, link_index=None, **kwargs):
    sa, sb = civec_strs._strs
    return fo.rdm2s(np.asarray(civec_strs), sa, sb, int(norb))


def make_rdm2(civec_strs, norb, nelec, link_index=None, **kwargs):
    aa, ab, bb = make_rdm2s(civec_strs, norb, nelec)
    return aa + bb + ab + ab.transpose(2, 3, 0, 1)


def spin_square(civec_strs, norb, nelec):
    sa, sb = civec_strs._strs
    c = np.asarray(civec_strs).reshape(-1)
    s2m = fo.spin_square_matrix(sa, sb, int(norb))
    ss = float(c @ s2m @ c) / float(c @ c)
    return ss, float(np.sqrt(ss + 0.25) * 2.0)


class SelectedCI:
    conv_tol = 1e-9
    max_cycle = 100
    max_space = 12
    lindep = 1e-14
    nroots = 1

    def __init__(self, mol=None):
        self._spin_penalty = None

    def make_rdm1s(self, civec_strs, norb, nelec, link_index=None):
        return make_rdm1s(civec_strs, norb, nelec)

    def make_rdm1(self, civec_strs, norb, nelec, link_index=None):
        return make_rdm1(civec_strs, norb, nelec)

    def make_rdm2s(self, civec_strs, norb, nelec, link_index=None, **kwargs):
        return make_rdm2s(civec_strs, norb, nelec)

    def make_rdm2(self, civec_strs, norb, nelec, link_index=None, **kwargs):
        return make_rdm2(civec_strs, norb, nelec)

    def spin_square(self, civec_strs, norb, nelec):
        return spin_square(civec_strs, norb, nelec)


SCI = SelectedCI


def kernel_fixed_space(myci, h1e, eri, norb, nelec, ci_strs, ci0=None, tol=None, lindep=None,
                       max_cycle=None, max_space=None, nroots=None, davidson_only=None, max_memory=None,
                       verbose=None, ecore=0, **kwargs):
    sa, sb = (np.asarray(s) for s in ci_strs)
    pen = getattr(myci, "_spin_penalty", None)
    if pen is None:
        e, c, _occ, _s2, ritz = fo.solve_dense(sa, sb, np.asarray(h1e), np.asarray(eri), int(norb))
    else:
        shift, ss = pen
        e, c, _occ, _s2, ritz = fo.solve_dense(sa, sb, np.asarray(h1e), np.asarray(eri), int(norb),
                                               spin_sq=ss, shift=shift)
    return ritz + ecore, _as_SCIvector(c, (sa, sb))

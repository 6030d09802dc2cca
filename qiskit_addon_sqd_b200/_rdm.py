"""Reduced density matrices of an ``SCIState`` on the GPU (reference ``fermion.py:113-128``).

Rank 1 (``sqd_rdm1s``) and rank 2 (``sqd_rdm2s``) come straight from the in-set excitation tables; the
conventions are pyscf's (``selected_ci.make_rdm1s / make_rdm2s``):  ``dm1[p, q] = <q^+ p>``,
``dm2[p, q, r, s] = <p^+ r^+ s q>``, spin-summed ``dm2 = aa + bb + ab + ab.transpose(2, 3, 0, 1)``.
There is deliberately no CPU fallback.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def subspace_rdm1s(sub, c) -> tuple[np.ndarray, np.ndarray]:
    """``(dm1a, dm1b)`` with pyscf's convention ``dm1[p, q] = <q^+ p>`` (symmetric for real states)."""
    torch, lib = sub.torch, sub.lib
    op = sub.spin_operator()  # any operator carries the tables; the S^2 one needs no integrals
    norb = sub.norb
    dm1 = torch.empty(2 * norb * norb, dtype=torch.float64, device=sub.device)
    ws_bytes = lib.sqd_rdm1s_workspace_bytes(C.byref(op.struct))
    ws = torch.empty(ws_bytes // 8 + 1, dtype=torch.float64, device=sub.device)
    dots = torch.empty(max(sub.ta.nnz, sub.tb.nnz, 1), dtype=torch.float64, device=sub.device)
    _lib.check(lib.sqd_rdm1s(C.byref(op.struct), _lib.ptr(c), sub.ta.nnz, sub.tb.nnz, _lib.ptr(dm1),
                             _lib.ptr(ws), _lib.ptr(dots), _lib.stream_ptr(torch)), "sqd_rdm1s")
    d = _lib.download(torch, dm1).reshape(2, norb, norb)
    return d[0].T.copy(), d[1].T.copy()


def subspace_rdm2s_device(sub, c):
    """Device tensors ``(dm2aa, dm2ab, dm2bb)``, each ``(norb,)*4``, ``dm2[p,q,r,s] = <p^+ r^+ s q>``."""
    torch, lib = sub.torch, sub.lib
    op = sub.spin_operator()
    norb = sub.norb
    out = torch.empty(3, norb, norb, norb, norb, dtype=torch.float64, device=sub.device)
    ws_bytes = lib.sqd_rdm2s_workspace_bytes(C.byref(op.struct), sub.ta.nnz, sub.tb.nnz)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=sub.device)
    _lib.check(lib.sqd_rdm2s(C.byref(op.struct), _lib.ptr(c), sub.ta.nnz, sub.tb.nnz, _lib.ptr(out[0]),
                             _lib.ptr(out[1]), _lib.ptr(out[2]), None, _lib.ptr(ws), ws_bytes,
                             _lib.stream_ptr(torch)), "sqd_rdm2s")
    return out[0], out[1], out[2]


def subspace_rdm2s(sub, c) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    aa, ab, bb = subspace_rdm2s_device(sub, c)
    torch = sub.torch
    return _lib.download(torch, aa), _lib.download(torch, ab), _lib.download(torch, bb)


def subspace_rdm2(sub, c) -> np.ndarray:
    """Spin-summed 2-RDM (pyscf ``make_rdm2``)."""
    aa, ab, bb = subspace_rdm2s_device(sub, c)
    return _lib.download(sub.torch, aa + bb + ab + ab.permute(2, 3, 0, 1))


def subspace_rdms(sub, c):
    """Spin-summed ``(rdm1, rdm2)`` for ``SCIResult`` (reference ``fermion.py:728-729``)."""
    a, b = subspace_rdm1s(sub, c)
    return a + b, subspace_rdm2(sub, c)


def state_rdm(state, rank: int, spin_summed: bool):
    from .fermion import _Subspace

    sub = _Subspace(state.ci_strs_a, state.ci_strs_b, int(state.norb), None, None)
    c = sub.upload_amplitudes(state.amplitudes)
    if rank == 1:
        a, b = subspace_rdm1s(sub, c)
        return a + b if spin_summed else (a, b)
    if rank == 2:
        return subspace_rdm2(sub, c) if spin_summed else subspace_rdm2s(sub, c)
    raise NotImplementedError(
        f"Computing the rank {rank} reduced density matrix is currently not supported."
    )

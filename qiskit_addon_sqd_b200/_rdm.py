"""Reduced density matrices of an ``SCIState`` on the GPU (reference ``fermion.py:113-128``).

Rank 1 (spin-resolved and spin-summed) is built from the in-set excitation tables by ``sqd_rdm1s``.
Rank 2 is the next item of the scope contract and raises
``NotImplementedError`` until its kernel exists -- there is deliberately no CPU fallback.
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
    d = dm1.cpu().numpy().reshape(2, norb, norb)
    return d[0].T.copy(), d[1].T.copy()


def subspace_rdms(sub, c):
    """Spin-summed ``(rdm1, rdm2)`` for ``SCIResult``; rdm2 is not built yet (``None``)."""
    a, b = subspace_rdm1s(sub, c)
    return a + b, None


def state_rdm(state, rank: int, spin_summed: bool):
    from .fermion import _Subspace

    if rank == 2:
        raise NotImplementedError(
            "The rank-2 reduced density matrix has no CUDA kernel yet in qiskit_addon_sqd_b200 "
            "(SURVEY.md section 8f); energies and <S^2> do not need it."
        )
    sub = _Subspace(state.ci_strs_a, state.ci_strs_b, int(state.norb), None, None)
    c = sub.upload_amplitudes(state.amplitudes)
    a, b = subspace_rdm1s(sub, c)
    return a + b if spin_summed else (a, b)

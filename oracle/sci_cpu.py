"""ctypes wrapper of ``oracle/_ref/libsci_cpu.so`` (``oracle/sci_cpu.c``).  TEST INFRASTRUCTURE / CPU
BASELINE ONLY.  Built by ``make -C oracle`` (``__graft_entry__.build()`` does it)."""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libsci_cpu.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_PATH):
            raise ImportError(f"{_PATH} missing: run `make -C oracle`")
        _lib = C.CDLL(_PATH)
        _lib.sci_cpu_solve.restype = C.c_int
        _lib.sci_cpu_sigma.restype = C.c_int
    return _lib


def _u64(a):
    return np.ascontiguousarray(np.asarray(a).astype(np.uint64))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def sigma(strs_a, strs_b, h, g, c, algo: int = 0) -> np.ndarray:
    lib = load()
    sa, sb = _u64(strs_a), _u64(strs_b)
    h = np.ascontiguousarray(h, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    out = np.empty_like(c)
    lib.sci_cpu_sigma(_p(sa), C.c_int(len(sa)), _p(sb), C.c_int(len(sb)), C.c_int(h.shape[0]), _p(h),
                      _p(g), C.c_int(algo), _p(c), _p(out))
    return out


def solve(strs_a, strs_b, h, g, *, algo: int = 0, spin_sq=None, shift: float = 0.1, tol: float = 1e-12,
          max_cycle: int = 200, max_space: int = 12, nthreads: int = 0):
    """Returns ``(energy, amplitudes, (occ_a, occ_b), info)``; ``max_cycle <= 0`` times sigma builds only."""
    lib = load()
    sa, sb = _u64(strs_a), _u64(strs_b)
    h = np.ascontiguousarray(h, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    norb = h.shape[0]
    amps = np.zeros((len(sa), len(sb)))
    occ = np.zeros(2 * norb)
    e = C.c_double(0.0)
    theta = C.c_double(0.0)
    info = (C.c_int * 2)()
    lib.sci_cpu_solve(_p(sa), C.c_int(len(sa)), _p(sb), C.c_int(len(sb)), C.c_int(norb), _p(h), _p(g),
                      C.c_int(algo), C.c_int(0 if spin_sq is None else 1), C.c_double(shift),
                      C.c_double(0.0 if spin_sq is None else spin_sq), C.c_double(tol),
                      C.c_int(max_cycle), C.c_int(max_space), C.c_int(nthreads), _p(amps), C.byref(e),
                      _p(occ), C.byref(theta), info)
    return e.value, amps, (occ[:norb].copy(), occ[norb:].copy()), dict(converged=info[0], cycles=info[1],
                                                                      theta=theta.value)

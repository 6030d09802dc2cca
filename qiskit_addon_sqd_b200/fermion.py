"""Fermionic subspace projection + diagonalisation on B200.

Host-side mirror of the reference's ``qiskit_addon_sqd/fermion.py`` for the hot path only
(``solve_fermion`` :745-845, ``solve_sci`` :684-742, ``solve_sci_batch`` :643-681, ``SCIState`` :57-139,
``SCIResult`` :142-159, ``bitstring_matrix_to_ci_strs`` :1004-1035, ``_check_ci_strs`` :1075-1097):
same names, argument meaning, return layout and error messages.  Everything the reference delegates
to pyscf (``fci.selected_ci.kernel_fixed_space``, ``make_rdm1s``, the RDM energy, ``spin_square``) runs
as hand-written sm_100a kernels behind the C-ABI in ``include/sqd_b200.h``; device buffers are torch
tensors, torch is plumbing only.  There is no CPU fallback.

``solve_sci_batch`` is a drop-in ``sci_solver`` for ``diagonalize_fermionic_hamiltonian``
(``fermion.py:216-220, 432``): ``functools.partial(solve_sci_batch, spin_sq=0.0)``.

Limits of the solvers (``solve_fermion``, ``solve_sci``, ``solve_sci_batch``, ``solve_sci_sharded``); a
violation raises ``ValueError`` / ``RuntimeError`` naming the limit, nothing falls back silently:

* ``norb <= 64`` spatial orbitals (one 64-bit word per determinant string);
* at most 2^19 = 524 288 strings per spin, and fewer than 2^31 stored same-spin table entries per spin;
* beta strings per subspace: the fast sigma kernels stage whole rows of the CI matrix in shared memory
  (``nb <= 8192`` for the v2 kernels every BASELINE.json shape runs on, ``nb <= 5760`` for the v1 kernels of
  sparse subspaces); longer rows are solved by the staging-free "wide" kernel (any ``nb`` the tables take,
  several times slower per determinant) -- nothing raises.  The full 2-RDM pass stages two rows
  (``2*nb + norb^2`` doubles <= 225 KB, ``nb <= ~13 900`` at 30 orbitals): beyond, pass ``compute_rdms=False``;
* ``nroots > 1`` is not supported (``kernel_fixed_space`` default 1 is what the reference uses).
"""

from __future__ import annotations

import ctypes as C
import os
import sys
import threading
import time
import warnings
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass
from typing import Sequence, cast

import numpy as np

from . import _lib

# pyscf SelectedCI defaults (recalled) -- except ``max_space``: pyscf keeps 12 basis
# vectors and collapses onto the Ritz vector.  Here a restart keeps the Ritz vector AND the previous one
# (locally optimal restart), which makes the iteration count practically independent of the basis
# size (measured on the bench workloads: 234 cycles in total for 8 solves at max_space 12, 238 at 6, 241 at
# 5) while every basis vector costs a pass over HBM/L2 in three kernels of every cycle: 6 is ~8 % faster per
# step than 12 and halves the workspace.  ``max_space=`` passes through literally.
_PYSCF_DEFAULTS = dict(max_cycle=100, max_space=int(os.environ.get("SQD_MAX_SPACE", "6")), lindep=1e-14, level_shift=1e-4)
# pyscf's conv_tol is 1e-9 with |r| < sqrt(tol).  That leaves O(1e-9/gap) in the energy; to meet the
# 1e-8 Ha parity bar against any converged solver the default here is tighter.  Passing ``tol=`` gives
# pyscf's rule (|dE| < tol and |r| < sqrt(tol)) literally.
_DEFAULT_TOL = 1e-12
# sigma-build work plan: cost units per CTA (single excitation = 16, double = 1) and the length above
# which a beta string's single-excitation list is reduced by a whole warp
_SIGMA_COST_PER_CHUNK = int(__import__("os").environ.get("SQD_SIGMA_CHUNK_COST", "256"))
_SIGMA_LONG_THRESHOLD = int(__import__("os").environ.get("SQD_SIGMA_LONG_THRESHOLD", "32"))
# v2 sigma kernels: links per virtual column (8 or 16) and items (alpha excitations) per work unit
_SIGMA_V2_LMAX = int(__import__("os").environ.get("SQD_V2_LMAX", "16"))
_SIGMA_V2_ITEMS_PER_CHUNK = int(__import__("os").environ.get("SQD_V2_ITEMS_PER_CHUNK", "8"))
_FIX_SPIN_DEFAULT_SHIFT = 0.2  # pyscf.fci.addons.fix_spin_ default, used by solve_sci (fermion.py:715)


# ------------------------------------------------------------------------------------------------
# result types (reference fermion.py:57-159)
# ------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class SCIState:
    """The amplitudes and determinants describing a quantum state (reference ``fermion.py:57-139``)."""

    amplitudes: np.ndarray
    ci_strs_a: np.ndarray
    ci_strs_b: np.ndarray
    norb: int
    nelec: tuple[int, int]

    def __post_init__(self):
        object.__setattr__(self, "amplitudes", np.asarray(self.amplitudes))
        if self.amplitudes.shape != (len(self.ci_strs_a), len(self.ci_strs_b)):
            raise ValueError(
                f"'amplitudes' shape must be ({len(self.ci_strs_a)}, {len(self.ci_strs_b)}) "
                f"but got {self.amplitudes.shape}"
            )

    def save(self, filename):
        """Save to ``.npz`` with the reference's keys (``fermion.py:90-99``)."""
        np.savez(
            filename,
            amplitudes=self.amplitudes,
            ci_strs_a=self.ci_strs_a,
            ci_strs_b=self.ci_strs_b,
            norb=self.norb,
            nelec=self.nelec,
        )

    @classmethod
    def load(cls, filename):
        with np.load(filename) as data:
            return cls(
                data["amplitudes"],
                data["ci_strs_a"],
                data["ci_strs_b"],
                norb=data["norb"],
                nelec=tuple(data["nelec"]),
            )

    def rdm(self, rank: int = 1, spin_summed: bool = False) -> np.ndarray:
        """Reduced density matrix (reference ``fermion.py:113-128``); rank 1 and 2 on the GPU."""
        from ._rdm import state_rdm

        if rank not in (1, 2):
            raise NotImplementedError(
                f"Computing the rank {rank} reduced density matrix is currently not supported."
            )
        return state_rdm(self, rank, spin_summed)

    def spin_square(self) -> float:
        """``<S^2>`` projected in the product subspace (reference ``fermion.py:130-134``)."""
        with _Subspace(self.ci_strs_a, self.ci_strs_b, int(self.norb), None, None) as sub:
            c = sub.upload_amplitudes(self.amplitudes)
            return cast(float, sub.spin_square(c))

    def orbital_occupancies(self) -> tuple[np.ndarray, np.ndarray]:
        with _Subspace(self.ci_strs_a, self.ci_strs_b, int(self.norb), None, None) as sub:
            c = sub.upload_amplitudes(self.amplitudes)
            return sub.occupancies(c)


@dataclass(frozen=True)
class SCIResult:
    """Result of an SCI calculation (reference ``fermion.py:142-159``)."""

    energy: float
    sci_state: SCIState
    orbital_occupancies: tuple[np.ndarray, np.ndarray]
    rdm1: np.ndarray | None = None
    rdm2: np.ndarray | None = None


# ------------------------------------------------------------------------------------------------
# string helpers
# ------------------------------------------------------------------------------------------------
def _as_uint64(strs) -> np.ndarray:
    arr = np.asarray(strs)
    if arr.dtype == object:
        arr = np.array([int(x) for x in arr], dtype=np.uint64)
    return np.ascontiguousarray(arr).astype(np.uint64, copy=False)


def _validate_strings(u: np.ndarray, which: str) -> None:
    """The C-ABI builds its excitation tables from string lists that are sorted ascending, unique and of one
    Hamming weight (``include/sqd_b200.h``; the SQD loop delivers them that way, ``fermion.py:556-557``).
    Anything else would silently produce a wrong Hamiltonian, so it is rejected here -- O(n) on the host."""
    if u.size > 1 and not bool(np.all(u[1:] > u[:-1])):
        raise ValueError(f"The {which} CI string list must be sorted in ascending order without duplicates.")
    pc = np.bitwise_count(u)
    if pc.size and not bool(np.all(pc == pc[0])):
        raise ValueError(
            f"Ci string in {which} list does not match hamming weight of the first string in that list."
        )


def _popcounts(strs_u64: np.ndarray) -> np.ndarray:
    return np.bitwise_count(strs_u64).astype(np.int64)


def _check_ci_strs(ci_strs: tuple[np.ndarray, np.ndarray]) -> tuple[np.ndarray, np.ndarray]:
    """Hamming weights must be consistent (reference ``fermion.py:1075-1097``, same messages)."""
    out = []
    for name, addr in zip(("Spin-up", "Spin-down"), ci_strs):
        u = _as_uint64(addr)
        ham = _popcounts(u)
        bad = np.nonzero(ham != ham[0])[0]
        if bad.size:
            i = int(bad[0])
            raise ValueError(
                f"{name} CI string in index 0 has hamming weight {int(ham[0])}, but CI string in "
                f"index {i} has hamming weight {int(ham[i])}."
            )
        out.append(np.sort(np.unique(np.asarray(addr))))
    return out[0], out[1]


def bitstring_matrix_to_ci_strs(
    bitstring_matrix: np.ndarray, open_shell: bool = False
) -> tuple[np.ndarray, np.ndarray]:
    """Bitstring rows -> (alpha, beta) determinant lists (reference ``fermion.py:1004-1035``).

    The left half of each row (columns ``[:norb]``) is the beta string and the right half the alpha
    string, column 0 the most significant bit (``counts.py:186-201``).  Packing runs on the GPU
    (``sqd_pack_bitstrings``); unique/sort/union of the resulting integers on the host.
    """
    torch = _lib.require_cuda()
    lib = _lib.load()
    bitstring_matrix = np.asarray(bitstring_matrix)
    n, nbits = bitstring_matrix.shape
    norb = nbits // 2
    if norb > 64:
        raise ValueError("qiskit_addon_sqd_b200 supports at most 64 spatial orbitals per spin.")
    if nbits % 2:
        bitstring_matrix = bitstring_matrix[:, : 2 * norb]
    if n == 0:
        empty = np.zeros(0, dtype=np.int64)
        return empty, empty
    bits = torch.from_numpy(np.ascontiguousarray(bitstring_matrix, dtype=np.uint8)).cuda()
    left = torch.empty(n, dtype=torch.int64, device="cuda")
    right = torch.empty(n, dtype=torch.int64, device="cuda")
    _lib.check(
        lib.sqd_pack_bitstrings(
            _lib.ptr(bits), n, 2 * norb, _lib.ptr(left), _lib.ptr(right), _lib.stream_ptr(torch)
        ),
        "sqd_pack_bitstrings",
    )
    left_u = left.cpu().numpy().view(np.uint64)
    right_u = right.cpu().numpy().view(np.uint64)
    ci_left = np.unique(left_u)
    ci_right = np.unique(right_u)
    if not open_shell:
        ci_left = ci_right = np.union1d(ci_left, ci_right)
    # reference returns python ``int`` (int64) below 64 bits, object dtype at 64; closed shell returns
    # the SAME array object for both spins (fermion.py:1032-1035)
    dt = np.int64 if norb < 64 else object
    if not open_shell:
        both = ci_right.astype(dt)
        return both, both
    return ci_right.astype(dt), ci_left.astype(dt)


# ------------------------------------------------------------------------------------------------
# device-side subspace
# ------------------------------------------------------------------------------------------------
class _DeviceIntegrals:
    """hcore / eri resident on one device (shared by the K subspaces of a batch)."""

    def __init__(self, torch, hcore: np.ndarray, eri: np.ndarray, device):
        norb = hcore.shape[0]
        if hcore.shape != (norb, norb) or tuple(eri.shape) != (norb,) * 4:
            raise ValueError(
                f"hcore must be (norb, norb) and eri (norb,)*4; got {hcore.shape} and {eri.shape}"
            )
        self.norb = norb
        self.h = torch.as_tensor(np.ascontiguousarray(hcore, dtype=np.float64)).to(device)
        self.g = torch.as_tensor(np.ascontiguousarray(eri, dtype=np.float64)).to(device)
        self.h2d_bytes = self.h.numel() * 8 + self.g.numel() * 8


class _SpinTableDev:
    """Excitation table of one string list (CSR; singles first then doubles)."""

    def __init__(self, torch, lib, strs_u64: np.ndarray, ints: _DeviceIntegrals | None, norb: int,
                 device, strs_dev=None):
        n = len(strs_u64)
        self.n = n
        self._torch, self._lib, self._sell = torch, lib, {}
        st = _lib.stream_ptr(torch)
        self.strs = strs_dev if strs_dev is not None else \
            torch.from_numpy(strs_u64.view(np.int64).copy()).to(device)
        self.n_single = torch.empty(n, dtype=torch.int32, device=device)
        n_total = torch.empty(n, dtype=torch.int32, device=device)
        self.row_ptr = torch.empty(n + 1, dtype=torch.int32, device=device)
        _lib.check(lib.sqd_excitation_count(_lib.ptr(self.strs), n, _lib.ptr(self.n_single),
                                            _lib.ptr(n_total), st), "sqd_excitation_count")
        total = C.c_int(0)
        _lib.check(lib.sqd_exclusive_scan(_lib.ptr(n_total), _lib.ptr(self.row_ptr), n,
                                          C.byref(total), st), "sqd_exclusive_scan")
        self.nnz = int(total.value)
        if self.nnz < 0 or self.nnz + 36 * n + 256 >= 2**31:
            raise ValueError(f"The excitation table of {n} strings has more than 2^31 entries: beyond the "
                             "32-bit table index of qiskit_addon_sqd_b200.")
        m = max(self.nnz, 1)
        self.col = torch.empty(m, dtype=torch.int32, device=device)
        self.val = torch.empty(m, dtype=torch.float64, device=device)
        self.meta = torch.empty(m, dtype=torch.int32, device=device)
        self.pack = torch.empty(m, dtype=torch.int32, device=device)
        self.diag = torch.empty(n, dtype=torch.float64, device=device)
        # ints None: structure-only table (S^2, occupancies) -- the kernel takes NULL integrals
        h, g = (None, None) if ints is None else (ints.h, ints.g)
        _lib.check(lib.sqd_excitation_fill(_lib.ptr(self.strs), n, norb, _lib.ptr(h), _lib.ptr(g),
                                           _lib.ptr(self.row_ptr), _lib.ptr(self.n_single),
                                           _lib.ptr(self.col), _lib.ptr(self.val),
                                           _lib.ptr(self.meta), _lib.ptr(self.pack),
                                           _lib.ptr(self.diag), st),
                   "sqd_excitation_fill")

    def sell(self, mode: int, long_idx=None) -> _lib.Sell:
        """SELL-32 copy of the table (mode 0: singles only, mode 1: all entries + values); cached."""
        if mode in self._sell:
            return self._sell[mode][0]
        torch = self._torch
        dev = self.strs.device
        n = self.n
        cap = self.nnz + 32 * n + 32
        ns = (n + 31) // 32
        perm = torch.empty(n, dtype=torch.int32, device=dev)
        ln = torch.empty(n, dtype=torch.int32, device=dev)
        sptr = torch.empty(ns + 1, dtype=torch.int32, device=dev)
        pack = torch.empty(cap, dtype=torch.int32, device=dev)
        val = torch.empty(cap, dtype=torch.float64, device=dev) if mode == 1 else None
        t = self.struct()
        _lib.check(self._lib.sqd_sell_build(C.byref(t), mode, _lib.ptr(long_idx), cap,
                                            _lib.ptr(perm), _lib.ptr(ln), _lib.ptr(sptr), _lib.ptr(pack),
                                            _lib.ptr(val), _lib.stream_ptr(torch)), "sqd_sell_build")
        n_entries = int(_lib.read_back(torch, sptr[-1:])[0])
        st = _lib.Sell(ns, n_entries, _lib.ptr(perm), _lib.ptr(ln), _lib.ptr(sptr), _lib.ptr(pack),
                       _lib.ptr(val))
        self._sell[mode] = (st, (perm, ln, sptr, pack, val))
        return st

    def struct(self) -> _lib.SpinTable:
        return _lib.SpinTable(self.n, _lib.ptr(self.strs), _lib.ptr(self.row_ptr),
                              _lib.ptr(self.n_single), _lib.ptr(self.col), _lib.ptr(self.val),
                              _lib.ptr(self.meta), _lib.ptr(self.pack))


class _OperatorDev:
    """One projected operator (Hamiltonian, spin-penalised Hamiltonian, or S^2) in A x B."""

    def __init__(self, sub: "_Subspace", *, mode: int, shift: float, diag_const: float,
                 same_spin: bool, with_w: bool):
        torch, lib = sub.torch, sub.lib
        norb, na, nb, ldc, ldg = sub.norb, sub.na, sub.nb, sub.ldc, sub.ldg
        dev = sub.device
        st = _lib.stream_ptr(torch)
        self.gab = torch.empty(norb * norb * ldg, dtype=torch.float64, device=dev)
        g_ptr = _lib.ptr(sub.ints.g) if (sub.ints is not None and mode == 0) else 0
        _lib.check(lib.sqd_make_gab(g_ptr, norb, float(shift), mode, _lib.ptr(self.gab), ldg, st),
                   "sqd_make_gab")
        self.Wa = torch.empty(na * ldg, dtype=torch.float64, device=dev) if with_w else None
        self.Wb = torch.empty(norb * norb * ldc, dtype=torch.float64, device=dev)
        self.diag = torch.empty(na * ldc, dtype=torch.float64, device=dev)
        da = _lib.ptr(sub.ta.diag) if same_spin else 0
        db = _lib.ptr(sub.tb.diag) if same_spin else 0
        # pads of the diagonal get a huge value: the preconditioner then maps pad entries to ~0
        _lib.check(
            lib.sqd_opposite_spin_tables(
                _lib.ptr(sub.ta.strs), na, _lib.ptr(sub.tb.strs), nb, norb, _lib.ptr(self.gab), ldg,
                da, db, float(diag_const), 1e300 if same_spin else 0.0, _lib.ptr(self.Wa),
                _lib.ptr(self.Wb), _lib.ptr(self.diag), ldc, st),
            "sqd_opposite_spin_tables")
        if not with_w:
            self.Wb = None
        v2 = sub.sigma_v2()
        # rows too long for the staged kernels (or sigma_path="wide"): the wide kernel, no plan / SELL copies
        self.uses_wide = not v2.enabled and (
            sub.sigma_path == "wide" or not lib.sqd_sigma_v1_supported(ldc, ldg))
        if self.uses_wide:
            plan, bd, bb = _lib.SigmaPlan(), _lib.Sell(), _lib.Sell()
        else:
            plan = sub.sigma_plan()
            bd, bb = sub.tb.sell(0, sub._plan_keep[2]), sub.tb.sell(1)
        self.struct = _lib.Operator(sub.ta.struct(), sub.tb.struct(), norb, ldc, ldg,
                                    _lib.ptr(self.diag), _lib.ptr(self.gab), _lib.ptr(self.Wa),
                                    _lib.ptr(self.Wb), 1 if same_spin else 0, plan, bd, bb, 0, v2,
                                    1 if self.uses_wide else 0)
        self.uses_v2 = bool(v2.enabled)
        if lib.sqd_sigma_smem_bytes(C.byref(self.struct)) < 0:
            raise ValueError(
                f"subspace shape (na={na}, nb={nb}, norb={norb}) exceeds the shared-memory row "
                "staging of the sigma kernel (need 3*nb + 2*norb^2 doubles <= 227 KB)"
            )


class _Subspace:
    """One product subspace A x B kept alive on the current CUDA device / stream: tables, operators and
    vectors through the fine-grained C entry points.  Used by ``SCIState``'s methods (RDMs, <S^2>,
    occupancies of a stored state) and by the kernel-level tests; a ground-state solve does not go through
    this class but through the one-call ``sqd_solve_subspace`` (``_solve_on_device``)."""

    def __init__(self, strs_a, strs_b, norb: int, hcore, eri, ints: _DeviceIntegrals | None = None,
                 strs_dev=(None, None), sigma_path: str = "auto"):
        self.torch = torch = _lib.require_cuda()
        self.sigma_path = sigma_path
        self.lib = lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device())
        if norb > 64 or norb < 1:
            raise ValueError("qiskit_addon_sqd_b200 supports 1..64 spatial orbitals.")
        self.norb = norb
        if ints is None and hcore is not None:
            ints = _DeviceIntegrals(torch, np.asarray(hcore), np.asarray(eri), self.device)
        self.ints = ints
        same = strs_a is strs_b
        ua = _as_uint64(strs_a)
        ub = ua if same else _as_uint64(strs_b)
        if ua.size == 0 or ub.size == 0:
            raise ValueError("The subspace must contain at least one alpha and one beta string.")
        self.strs_a_host, self.strs_b_host = ua, ub
        self.na, self.nb = len(ua), len(ub)
        self.ldc = (self.nb + 1) // 2 * 2
        self.ldg = (norb * norb + 2) // 2 * 2  # even and > norb^2: the pad column is a zero slot for gathers
        self.n_alpha = int(np.bitwise_count(ua[0]))
        self.n_beta = int(np.bitwise_count(ub[0]))
        self.ta = _SpinTableDev(torch, lib, ua, ints, norb, self.device, strs_dev[0])
        self.tb = self.ta if (same or (ua.shape == ub.shape and np.array_equal(ua, ub))) else \
            _SpinTableDev(torch, lib, ub, ints, norb, self.device, strs_dev[1])
        self._ss_op = None
        self._plan = None
        self._v2 = None
        self._scratch = torch.empty(4096, dtype=torch.float64, device=self.device)
        self._scalar = torch.empty(8, dtype=torch.float64, device=self.device)

    def sigma_plan(self) -> _lib.SigmaPlan:
        """Work decomposition of the sigma build (chunks of rows, long columns); built once."""
        if self._plan is not None:
            return self._plan
        torch, lib = self.torch, self.lib
        dev = self.device
        cost = int(_SIGMA_COST_PER_CHUNK)
        max_chunks = 2 * self.na + (16 * max(self.ta.nnz, 1)) // cost + 1
        i32 = dict(dtype=torch.int32, device=dev)
        bufs = [torch.empty(max_chunks, **i32) for _ in range(4)]
        split = [torch.empty(self.na, **i32) for _ in range(3)]
        long_idx = torch.empty(self.nb, **i32)
        long_cols = torch.empty(_lib.MAX_LONG_COLUMNS, **i32)
        counts_d = torch.empty(4, **i32)
        counts = (C.c_int * 8)()
        ta, tb = self.ta.struct(), self.tb.struct()
        _lib.check(lib.sqd_sigma_plan_build(C.byref(ta), C.byref(tb), cost, int(_SIGMA_LONG_THRESHOLD),
                                      max_chunks, *[_lib.ptr(t) for t in bufs],
                                      *[_lib.ptr(t) for t in split], _lib.ptr(long_idx),
                                      _lib.ptr(long_cols), _lib.ptr(counts_d), counts,
                                      _lib.stream_ptr(torch)), "sqd_sigma_plan_build")
        n_chunks, n_slots, n_split, n_long = (int(v) for v in counts[:4])
        part = torch.empty(max(n_slots, 1) * self.ldc, dtype=torch.float64, device=dev)
        self._plan_keep = (bufs, split, long_idx, long_cols, part)
        self._plan = _lib.SigmaPlan(n_chunks, n_slots, n_split, n_long, *[_lib.ptr(t) for t in bufs],
                                    *[_lib.ptr(t) for t in split], _lib.ptr(long_idx),
                                    _lib.ptr(long_cols), _lib.ptr(part))
        return self._plan

    def sigma_v2(self) -> _lib.SigmaV2:
        """Tables of the v2 sigma kernels (``csrc/fermion_sigma2.cu``); built once.  ``sigma_path``:
        ``"v1"`` never, ``"v2"`` whenever the shape is supported, ``"auto"`` by the library's density rule."""
        if self._v2 is not None:
            return self._v2
        torch, lib = self.torch, self.lib
        na, nb = self.na, self.nb
        nnz_a, nnz_b = self.ta.nnz, self.tb.nnz
        want = self.sigma_path == "v2" or (
            self.sigma_path == "auto" and lib.sqd_sigma_v2_recommended(na, nb, nnz_a, nnz_b))
        self._v2 = _lib.SigmaV2()
        if not want:
            return self._v2
        lmax, ipc = int(_SIGMA_V2_LMAX), int(_SIGMA_V2_ITEMS_PER_CHUNK)
        pbytes = int(lib.sqd_sigma_v2_plan_bytes(na, nb, nnz_a, nnz_b, lmax, ipc))
        if pbytes < 0:
            return self._v2
        plan = torch.empty(pbytes, dtype=torch.uint8, device=self.device)
        counts = (C.c_int * _lib.V2_COUNTS)()
        ta, tb = self.ta.struct(), self.tb.struct()
        st = _lib.stream_ptr(torch)
        _lib.check(lib.sqd_sigma_v2_plan(C.byref(ta), C.byref(tb), self.norb, nnz_a, nnz_b, lmax, ipc,
                                         _lib.ptr(plan), pbytes, counts, st), "sqd_sigma_v2_plan")
        same = 1 if self.tb is self.ta else 0
        sbytes = int(lib.sqd_sigma_v2_scratch_bytes(counts, na, nb, self.ldc, 1, same))
        if sbytes < 0:   # shape the v2 planner does not support: v1 stays the product path
            return self._v2
        scratch = torch.empty(sbytes, dtype=torch.uint8, device=self.device)
        v2 = _lib.SigmaV2()
        _lib.check(lib.sqd_sigma_v2_finish(C.byref(ta), C.byref(tb), self.ldc, nnz_a, nnz_b, lmax, ipc,
                                           counts, _lib.ptr(plan), _lib.ptr(scratch), sbytes, 1,
                                           C.byref(v2), st), "sqd_sigma_v2_finish")
        self._v2_keep = (plan, scratch)
        self.v2_counts = [int(c) for c in counts]
        self._v2 = v2
        return v2

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    # -- operators ---------------------------------------------------------------------------
    def hamiltonian(self, penalty_shift: float = 0.0, penalty_ss: float = 0.0) -> _OperatorDev:
        sz = 0.5 * (self.n_alpha - self.n_beta)
        const = penalty_shift * (sz * (sz + 1.0) + self.n_beta - penalty_ss)
        return _OperatorDev(self, mode=0, shift=penalty_shift, diag_const=const, same_spin=True,
                            with_w=True)

    def spin_operator(self) -> _OperatorDev:
        if self._ss_op is None:
            sz = 0.5 * (self.n_alpha - self.n_beta)
            self._ss_op = _OperatorDev(self, mode=1, shift=0.0,
                                       diag_const=sz * (sz + 1.0) + self.n_beta, same_spin=False,
                                       with_w=False)
        return self._ss_op

    # -- vectors -----------------------------------------------------------------------------
    def new_vector(self):
        return self.torch.empty(self.na * self.ldc, dtype=self.torch.float64, device=self.device)

    def upload_amplitudes(self, amps: np.ndarray):
        torch = self.torch
        amps = np.ascontiguousarray(amps, dtype=np.float64)
        c = torch.zeros(self.na, self.ldc, dtype=torch.float64, device=self.device)
        c[:, : self.nb] = torch.from_numpy(amps).to(self.device)
        return c.reshape(-1)

    def download_amplitudes(self, c) -> np.ndarray:
        return _lib.download(self.torch, c.reshape(self.na, self.ldc)[:, : self.nb])

    def apply(self, op: _OperatorDev, c, out=None):
        out = self.new_vector() if out is None else out
        _lib.check(self.lib.sqd_sigma(C.byref(op.struct), _lib.ptr(c), _lib.ptr(out),
                                      _lib.stream_ptr(self.torch)), "sqd_sigma")
        return out

    def dot(self, x, y) -> float:
        _lib.check(self.lib.sqd_dot(_lib.ptr(x), _lib.ptr(y), x.numel(), _lib.ptr(self._scalar),
                                    _lib.ptr(self._scratch), _lib.stream_ptr(self.torch)), "sqd_dot")
        return float(_lib.read_back(self.torch, self._scalar[:1])[0])

    # -- observables -------------------------------------------------------------------------
    def spin_square(self, c) -> float:
        s2c = self.apply(self.spin_operator(), c)
        return self.dot(c, s2c) / self.dot(c, c)

    def occupancies(self, c) -> tuple[np.ndarray, np.ndarray]:
        torch = self.torch
        occ = torch.empty(2 * self.norb, dtype=torch.float64, device=self.device)
        scratch = torch.empty(self.na + self.nb, dtype=torch.float64, device=self.device)
        _lib.check(self.lib.sqd_occupancies(_lib.ptr(c), _lib.ptr(self.ta.strs), self.na,
                                            _lib.ptr(self.tb.strs), self.nb, self.ldc, self.norb,
                                            _lib.ptr(occ), _lib.ptr(scratch),
                                            _lib.stream_ptr(torch)), "sqd_occupancies")
        o = _lib.read_back(torch, occ)
        return o[: self.norb].copy(), o[self.norb:].copy()


# ------------------------------------------------------------------------------------------------
# the solve
# ------------------------------------------------------------------------------------------------
def _solver_options(kwargs: dict) -> dict:
    """Map pyscf ``kernel_fixed_space`` keyword arguments (``fermion.py:651,722,817``)."""
    kw = dict(kwargs)
    opts = dict(_PYSCF_DEFAULTS)
    if "tol" in kw and kw["tol"] is not None:
        tol = float(kw.pop("tol"))
        opts["tol"], opts["tol_residual"] = tol, float(np.sqrt(tol))
    else:
        kw.pop("tol", None)
        opts["tol"], opts["tol_residual"] = _DEFAULT_TOL, float(np.sqrt(_DEFAULT_TOL))
    for key in ("max_cycle", "max_space", "lindep"):
        if kw.get(key) is not None:
            opts[key] = kw.pop(key)
        else:
            kw.pop(key, None)
    opts["ci0"] = kw.pop("ci0", None)
    opts["sigma_path"] = kw.pop("sigma_path", None) or "auto"  # not a pyscf option: "auto", "v1", "v2", "wide"
    if opts["sigma_path"] not in ("auto", "v1", "v2", "wide"):
        raise ValueError("sigma_path must be one of 'auto', 'v1', 'v2', 'wide'")
    nroots = kw.pop("nroots", None)
    if nroots not in (None, 1):
        raise NotImplementedError("qiskit_addon_sqd_b200 computes the ground state only (nroots=1).")
    for ignored in ("davidson_only", "max_memory", "verbose", "ecore", "pspace_size", "orbsym"):
        kw.pop(ignored, None)  # no effect on the returned state / expectation values
    if kw:
        warnings.warn(f"Ignoring unsupported solver options: {sorted(kw)}", stacklevel=3)
    return opts


@dataclass
class SolveStats:
    """Bookkeeping of the last solves on this thread (used by bench.py; not part of the reference API)."""

    cycles: int = 0
    sigma_builds: int = 0
    converged: int = 0
    residual: float = 0.0
    theta: float = 0.0
    n_det: int = 0
    nnz_a: int = 0
    nnz_b: int = 0
    singles_a: int = 0
    singles_b: int = 0
    sigma_ms: float = 0.0   # profile mode only
    davidson_ms: float = 0.0
    na: int = 0
    nb: int = 0
    norb: int = 0
    host_ms: tuple = ()     # host wall time of (preparation, library call, downloads)
    sigma_path: int = 0     # 1: fermion_sigma.cu, 2: fermion_sigma2.cu, 3: the wide kernel (what the solve ran)


_tls = threading.local()
_pool_lock = threading.Lock()
_pool: ThreadPoolExecutor | None = None
_streams: dict[tuple[int, int], object] = {}


def _throughput_mode(n_solves: int, total_dets: int) -> bool:
    """Launch configuration of concurrent solves on one device.  ``True`` (several solves that together
    keep the GPU busy, ~3e5 determinants or more): sigma kernels under a register cap with persistent
    CTAs, one Rayleigh-Ritz kernel on the main stream -- best aggregate throughput.  ``False`` (a lone
    solve, or a handful of small ones that leave the GPU mostly idle): shortest critical path per solve --
    lowest Ritz pair from the secular equation, full decomposition on a side stream."""
    return n_solves > 1 and total_dets >= 300_000


def _worker_pool() -> ThreadPoolExecutor:
    """Host threads that drive the concurrent subspaces (created once; ctypes calls release the GIL)."""
    global _pool
    with _pool_lock:
        if _pool is None:
            _pool = ThreadPoolExecutor(max_workers=16, thread_name_prefix="sqd-b200")
        return _pool


def _stream_for(torch, device: int, slot: int):
    with _pool_lock:
        key = (device, slot)
        if key not in _streams:
            _streams[key] = torch.cuda.Stream(device=device)
        return _streams[key]


def last_solve_stats() -> list[SolveStats]:
    return list(getattr(_tls, "stats", []))


def _residual_tolerance(opts: dict, spin_sq) -> float:
    """Residual norm at which the Davidson iteration may stop.  ``sqrt(tol)`` as in pyscf -- the Ritz value is
    second order in the eigenvector error.  With the spin penalty the energy that is REPORTED is the expectation
    value of the bare Hamiltonian in an eigenvector of H + shift (S^2 - s(s+1)), which is only first order
    (the state is not an eigenvector of H unless it is spin pure): ten times tighter there, so that the
    reported energy stays within the 1e-8 Ha of the north star."""
    t = float(opts["tol_residual"])
    return t if spin_sq is None else 0.1 * t


def _solve_on_device(strs_a, strs_b, norb: int, ints: _DeviceIntegrals, spin_sq, shift, opts,
                     want_spin: bool, want_rdm: bool, *, strs_dev=(None, None), download: bool = True,
                     profile: bool = False, shard_group=None, throughput: bool = False):
    """Ground state of H projected on A x B: ONE call into the library (``sqd_solve_subspace``), so the
    host thread does not touch the interpreter between the first kernel and the last read-back.
    Returns dict of results (host arrays; with ``download=False`` the amplitudes stay on the device as a
    padded ``(na, ldc)`` tensor)."""
    t_host0 = time.perf_counter()
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    if norb > 64 or norb < 1:
        raise ValueError("qiskit_addon_sqd_b200 supports 1..64 spatial orbitals.")
    same = strs_a is strs_b
    ua = _as_uint64(strs_a)
    ub = ua if same else _as_uint64(strs_b)
    if ua.size == 0 or ub.size == 0:
        raise ValueError("The subspace must contain at least one alpha and one beta string.")
    same = same or (ua.shape == ub.shape and np.array_equal(ua, ub))
    na, nb = len(ua), len(ub)
    ldc = (nb + 1) // 2 * 2
    da = strs_dev[0] if strs_dev[0] is not None else torch.from_numpy(ua.view(np.int64).copy()).to(dev)
    db = da if same else (strs_dev[1] if strs_dev[1] is not None
                          else torch.from_numpy(ub.view(np.int64).copy()).to(dev))
    n_alpha, n_beta = int(np.bitwise_count(ua[0])), int(np.bitwise_count(ub[0]))
    x = torch.empty(na * ldc, dtype=torch.float64, device=dev)
    rdm1_d = torch.empty(norb * norb, dtype=torch.float64, device=dev) if want_rdm else None
    rdm2_d = torch.empty(norb**4, dtype=torch.float64, device=dev) if want_rdm else None
    ci0 = opts.get("ci0")
    ci0_d = None
    if ci0 is not None:
        ci0_d = torch.from_numpy(np.ascontiguousarray(ci0, dtype=np.float64).reshape(na, nb)).to(dev)
    prm = _lib.SolveParams()
    prm.norb, prm.na, prm.nb, prm.n_alpha, prm.n_beta = norb, na, nb, n_alpha, n_beta
    prm.d_strs_a, prm.d_strs_b = _lib.ptr(da), _lib.ptr(db)
    prm.d_h, prm.d_g = _lib.ptr(ints.h), _lib.ptr(ints.g)
    prm.penalty = 0 if spin_sq is None else 1
    prm.spin_sq = 0.0 if spin_sq is None else float(spin_sq)
    prm.shift = float(shift)
    prm.want_spin = 1 if want_spin else 0
    prm.max_space = int(min(max(2, opts["max_space"]), _lib.MAX_SPACE))
    prm.max_cycle = int(opts["max_cycle"])
    prm.tol, prm.tol_residual = float(opts["tol"]), _residual_tolerance(opts, spin_sq)
    prm.lindep, prm.level_shift = float(opts["lindep"]), float(opts["level_shift"])
    prm.check_every = 4
    prm.d_ci0 = _lib.ptr(ci0_d) or None
    prm.cost_per_chunk, prm.long_threshold = int(_SIGMA_COST_PER_CHUNK), int(_SIGMA_LONG_THRESHOLD)
    prm.profile = 1 if profile else 0
    prm.throughput_mode = 1 if throughput else 0
    prm.sigma_path = {"auto": 0, "v1": 1, "v2": 2, "wide": 3}[opts.get("sigma_path", "auto")]
    if shard_group is not None:
        # rows are split inside the call into blocks of equal estimated sigma-build cost
        prm.nccl_comm, prm.row_begin, prm.row_end = shard_group._comm.value, -1, -1
        prm.shard_rank, prm.shard_world = shard_group.rank, shard_group.world
    res = _lib.SolveResult()
    t_host1 = time.perf_counter()
    _lib.check(lib.sqd_solve_subspace(C.byref(prm), _lib.ptr(x), _lib.ptr(rdm1_d), _lib.ptr(rdm2_d),
                                      C.byref(res), _lib.stream_ptr(torch)), "sqd_solve_subspace")
    t_host2 = time.perf_counter()
    info = res.info
    occ = (np.array(res.occ_a[:norb]), np.array(res.occ_b[:norb]))
    s2 = float(res.spin_square) if res.have_spin_square else None
    amps = _lib.download(torch, x.reshape(na, ldc)[:, :nb]) if download else x.reshape(na, ldc)
    rdm1 = rdm2 = None
    if want_rdm:
        rdm1 = _lib.download(torch, rdm1_d.reshape(norb, norb))
        rdm2 = _lib.download(torch, rdm2_d.reshape((norb,) * 4))
    t_host3 = time.perf_counter()
    stats = SolveStats(info.cycles, info.sigma_builds, info.converged, info.residual, info.theta,
                       na * nb, int(res.nnz_a), int(res.nnz_b), max(int(res.singles_a), 0),
                       max(int(res.singles_b), 0), info.sigma_ms, info.total_ms, na, nb, norb,
                       (1e3 * (t_host1 - t_host0), 1e3 * (t_host2 - t_host1), 1e3 * (t_host3 - t_host2)),
                       int(res.sigma_path))
    if not hasattr(_tls, "stats"):
        _tls.stats = []
    _tls.stats.append(stats)
    return dict(energy=float(res.energy), amplitudes=amps, occupancies=occ, spin_square=s2,
                nelec=(n_alpha, n_beta), rdm1=rdm1, rdm2=rdm2, stats=stats)


def solve_sci(
    ci_strings: tuple[np.ndarray, np.ndarray],
    one_body_tensor: np.ndarray,
    two_body_tensor: np.ndarray,
    norb: int,
    nelec: tuple[int, int],
    *,
    spin_sq: float | None = None,
    **kwargs,
) -> SCIResult:
    """Diagonalize the Hamiltonian in the subspace defined by CI strings (reference ``fermion.py:684-742``).

    As in the reference (``fermion.py:728-740``) the result carries the spin-summed ``rdm1`` and ``rdm2``.
    Extra keyword ``compute_rdms=False`` skips them (the SQD loop never reads them, ``fermion.py:577-622``);
    the energy does not depend on it -- it is the Rayleigh quotient of the bare Hamiltonian, which equals
    the reference's RDM contraction (``fermion.py:730-732``).
    """
    return solve_sci_batch([ci_strings], one_body_tensor, two_body_tensor, norb, nelec,
                           spin_sq=spin_sq, **kwargs)[0]


def solve_sci_batch(
    ci_strings: list[tuple[np.ndarray, np.ndarray]],
    one_body_tensor: np.ndarray,
    two_body_tensor: np.ndarray,
    norb: int,
    nelec: tuple[int, int],
    *,
    spin_sq: float | None = None,
    **kwargs,
) -> list[SCIResult]:
    """Diagonalize the Hamiltonian in K subspaces (reference ``fermion.py:643-681``).

    The reference runs the K solves one after the other; here they run concurrently, one CUDA stream
    per subspace.  ``devices`` (extra keyword): ``None`` = current device only (one process per GPU,
    as under ``torchrun``); a list of device indices or ``"all"`` = shard the batch round-robin over
    those GPUs from this single process -- no collective in either case.
    """
    torch = _lib.require_cuda()
    _lib.load()
    one_body_tensor = np.asarray(one_body_tensor)
    two_body_tensor = np.asarray(two_body_tensor)
    norb, _ = one_body_tensor.shape  # as the reference: norb comes from the tensor (fermion.py:711)
    devices = kwargs.pop("devices", None)
    want_rdm = bool(kwargs.pop("compute_rdms", True))
    ints_cache = kwargs.pop("_resident_integrals", None)   # private: set by diagonalize_fermionic_hamiltonian
    shift = float(kwargs.pop("shift", _FIX_SPIN_DEFAULT_SHIFT))
    opts = _solver_options(kwargs)
    if devices is None:
        dev_list = [torch.cuda.current_device()]
    elif devices == "all":
        dev_list = list(range(torch.cuda.device_count()))
    else:
        dev_list = [int(d) for d in devices]
    _tls.stats = []
    K = len(ci_strings)
    if K == 0:
        return []

    # One upload, K library calls, one download per device.  Everything the interpreter has to do for a
    # subspace happens here, on the calling thread, before / after the solves; a worker thread only makes
    # the `sqd_solve_subspace` call (ctypes drops the interpreter lock for its whole duration), so the K
    # solves overlap instead of queueing on the lock.
    lib = _lib.load()
    if norb > 64 or norb < 1:
        raise ValueError("qiskit_addon_sqd_b200 supports 1..64 spatial orbitals.")
    jobs: list[dict] = []
    per_device: dict[int, list[int]] = {d: [] for d in dev_list}
    for k, (strs_a, strs_b) in enumerate(ci_strings):
        same = strs_a is strs_b
        ua = _as_uint64(strs_a)
        ub = ua if same else _as_uint64(strs_b)
        if ua.size == 0 or ub.size == 0:
            raise ValueError("The subspace must contain at least one alpha and one beta string.")
        _validate_strings(ua, "first")
        if not same:
            _validate_strings(ub, "second")
        same = same or (ua.shape == ub.shape and np.array_equal(ua, ub))
        d = dev_list[k % len(dev_list)]
        per_device[d].append(k)
        jobs.append(dict(ua=ua, ub=ub, same=same, na=len(ua), nb=len(ub), ldc=(len(ub) + 1) // 2 * 2,
                         device=d, slot=len(per_device[d]) - 1))
    n2, n4 = norb * norb, norb**4
    keep = {}
    for d, ks in per_device.items():
        if not ks:
            continue
        with torch.cuda.device(d):
            dev = torch.device("cuda", d)
            # inside the SQD loop the integrals are uploaded once per device and stay resident between iterations
            ints = ints_cache.get(d) if ints_cache is not None else None
            if ints is None:
                ints = _DeviceIntegrals(torch, one_body_tensor, two_body_tensor, dev)
                if ints_cache is not None:
                    ints_cache[d] = ints
            # strings of every subspace of this device in one array -> one host-to-device copy
            parts, off = [], 0
            for k in ks:
                j = jobs[k]
                j["off_a"] = off
                parts.append(j["ua"])
                off += j["na"]
                if j["same"]:
                    j["off_b"] = j["off_a"]
                else:
                    j["off_b"] = off
                    parts.append(j["ub"])
                    off += j["nb"]
            strs_d = torch.from_numpy(np.concatenate(parts).view(np.int64)).to(dev)
            x_off, tot = [], 0
            for k in ks:
                x_off.append(tot)
                tot += jobs[k]["na"] * jobs[k]["ldc"]
            x_d = torch.empty(tot, dtype=torch.float64, device=dev)
            rdm1_d = torch.empty(len(ks) * n2, dtype=torch.float64, device=dev) if want_rdm else None
            rdm2_d = torch.empty(len(ks) * n4, dtype=torch.float64, device=dev) if want_rdm else None
            ci0 = opts.get("ci0")
            for i, k in enumerate(ks):
                j = jobs[k]
                prm = _lib.SolveParams()
                prm.norb, prm.na, prm.nb = norb, j["na"], j["nb"]
                prm.n_alpha, prm.n_beta = int(np.bitwise_count(j["ua"][0])), int(np.bitwise_count(j["ub"][0]))
                prm.d_strs_a = strs_d.data_ptr() + 8 * j["off_a"]
                prm.d_strs_b = strs_d.data_ptr() + 8 * j["off_b"]
                prm.d_h, prm.d_g = _lib.ptr(ints.h), _lib.ptr(ints.g)
                prm.penalty = 0 if spin_sq is None else 1
                prm.spin_sq = 0.0 if spin_sq is None else float(spin_sq)
                prm.shift = shift
                prm.max_space = int(min(max(2, opts["max_space"]), _lib.MAX_SPACE))
                prm.max_cycle = int(opts["max_cycle"])
                prm.tol, prm.tol_residual = float(opts["tol"]), _residual_tolerance(opts, spin_sq)
                prm.lindep, prm.level_shift = float(opts["lindep"]), float(opts["level_shift"])
                prm.check_every = 4
                if ci0 is not None:
                    j["ci0_d"] = torch.from_numpy(
                        np.ascontiguousarray(ci0, dtype=np.float64).reshape(j["na"], j["nb"])).to(dev)
                    prm.d_ci0 = j["ci0_d"].data_ptr()
                prm.cost_per_chunk, prm.long_threshold = int(_SIGMA_COST_PER_CHUNK), int(_SIGMA_LONG_THRESHOLD)
                prm.throughput_mode = 1 if _throughput_mode(
                    len(ks), sum(jobs[q]["na"] * jobs[q]["nb"] for q in ks)) else 0
                prm.sigma_path = {"auto": 0, "v1": 1, "v2": 2, "wide": 3}[opts.get("sigma_path", "auto")]
                j["prm"], j["res"] = prm, _lib.SolveResult()
                j["x_ptr"] = x_d.data_ptr() + 8 * x_off[i]
                j["x_off"] = x_off[i]
                j["rdm1_ptr"] = rdm1_d.data_ptr() + 8 * i * n2 if want_rdm else None
                j["rdm2_ptr"] = rdm2_d.data_ptr() + 8 * i * n4 if want_rdm else None
                j["n_alpha"], j["n_beta"] = prm.n_alpha, prm.n_beta
            keep[d] = (ints, strs_d, x_d, rdm1_d, rdm2_d)

    def work(k: int):
        j = jobs[k]
        d = j["device"]
        with torch.cuda.device(d):
            stream = _stream_for(torch, d, j["slot"]) if K > 1 else torch.cuda.current_stream()
            if K > 1:
                stream.wait_stream(torch.cuda.current_stream())
            rc = lib.sqd_solve_subspace(C.byref(j["prm"]), j["x_ptr"], j["rdm1_ptr"], j["rdm2_ptr"],
                                        C.byref(j["res"]), stream.cuda_stream)
            if rc:
                return rc, lib.sqd_last_error()   # the error string is per host thread
            # results come back on the worker's own stream through its pinned staging buffer, so the K
            # device-to-host copies and the K host copies overlap
            _ints, _strs, x_d, rdm1_d, rdm2_d = keep[d]
            na, nb, ldc, i = j["na"], j["nb"], j["ldc"], j["slot"]
            with torch.cuda.stream(stream):
                j["amps"] = _lib.download(torch, x_d[j["x_off"]: j["x_off"] + na * ldc].view(na, ldc)[:, :nb])
                if want_rdm:
                    j["rdm1"] = _lib.download(torch, rdm1_d[i * n2: (i + 1) * n2].view(norb, norb))
                    j["rdm2"] = _lib.download(torch, rdm2_d[i * n4: (i + 1) * n4].view((norb,) * 4))
        return 0, None

    if K == 1:
        status = [work(0)]
    else:
        status = list(_worker_pool().map(work, range(K)))
    for rc, err in status:
        if rc:
            raise _lib.SqdCudaError(f"sqd_solve_subspace failed ({rc}): {err.decode() if err else ''}")

    out, stats = [], []
    # the amplitude matrices of THIS call stay on the device until the next call on this thread: the SQD loop
    # selects its carry-over strings from them without touching the host copies (_SQDRun.digest)
    resident = _tls.resident_amplitudes = {}
    for k, (strs_a, strs_b) in enumerate(ci_strings):
        j = jobs[k]
        res, info = j["res"], j["res"].info
        na, nb = j["na"], j["nb"]
        amps = j["amps"]
        occ = (np.array(res.occ_a[:norb]), np.array(res.occ_b[:norb]))
        state = SCIState(amplitudes=amps, ci_strs_a=np.asarray(strs_a), ci_strs_b=np.asarray(strs_b),
                         norb=norb, nelec=tuple(nelec))
        out.append(SCIResult(float(res.energy), state, orbital_occupancies=occ,
                             rdm1=j.get("rdm1"), rdm2=j.get("rdm2")))
        x_all = keep[j["device"]][2]
        resident[id(state)] = (state, x_all[j["x_off"]: j["x_off"] + na * j["ldc"]].view(na, j["ldc"]), j["device"])
        stats.append(SolveStats(info.cycles, info.sigma_builds, info.converged, info.residual, info.theta,
                                na * nb, int(res.nnz_a), int(res.nnz_b), 0, 0, info.sigma_ms,
                                info.total_ms, na, nb, norb, (), int(res.sigma_path)))
    _tls.stats = stats
    return out


# ------------------------------------------------------------------------------------------------
# the self-consistent SQD loop (reference fermion.py:204-640)
# ------------------------------------------------------------------------------------------------
def _first_occurrences(values: np.ndarray) -> np.ndarray:
    """Distinct values in order of first appearance (reference ``fermion.py:465-469``)."""
    _, first = np.unique(values, return_index=True)
    return values[np.sort(first)]


def _by_descending(keys: np.ndarray) -> np.ndarray:
    """Permutation used by the reference wherever it ranks strings: numpy's default argsort, reversed."""
    return np.argsort(keys)[::-1]


def _carryover_on_device(x, nb: int, device: int, threshold: float):
    """Rows / columns of the device-resident amplitude matrix ``x`` (``na x ldc``) holding an entry with
    ``|c| >= threshold`` and their marginal weights, bit-identical to the reference's numpy expressions
    (``fermion.py:607-622``): ``sqd_carryover``.  Returns ``(rows, cols, weight_a, weight_b)``."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    na, ldc = int(x.shape[0]), int(x.shape[1])
    with torch.cuda.device(device):
        flags = torch.empty(na + nb, dtype=torch.int32, device=x.device)
        weights = torch.empty(na + nb, dtype=torch.float64, device=x.device)
        _lib.check(lib.sqd_carryover(_lib.ptr(x), na, nb, ldc, float(threshold), _lib.ptr(flags),
                                     flags.data_ptr() + 4 * na, _lib.ptr(weights), weights.data_ptr() + 8 * na,
                                     _lib.stream_ptr(torch)), "sqd_carryover")
        f = _lib.read_back(torch, flags)
        w = _lib.read_back(torch, weights)
    rows, cols = np.flatnonzero(f[:na]), np.flatnonzero(f[na:])
    return rows, cols, w[:na][rows], w[na:][cols]


class _SQDRun:
    """State of one ``diagonalize_fermionic_hamiltonian`` call: the loop-invariant inputs plus what one
    configuration-recovery iteration hands to the next (occupancies, reference result, carry-over
    strings).  ``strings_for_iteration`` is the first half of an iteration (reference
    ``_prepare_ci_strings``, ``fermion.py:472-560``), ``digest`` the second (``_process_sci_results``,
    ``fermion.py:563-640``)."""

    def __init__(self, raw_bitstrings, raw_probs, norb, nelec, samples_per_batch, num_batches,
                 symmetrize_spin, include, max_dims, energy_tol, occupancies_tol, carryover_threshold,
                 rng, occupancies):
        self.bits, self.probs = raw_bitstrings, raw_probs
        self.norb = norb
        self.n_alpha, self.n_beta = nelec
        self.samples_per_batch, self.num_batches = samples_per_batch, num_batches
        self.symmetric = symmetrize_spin
        self.include_a, self.include_b = include
        self.max_dim_a, self.max_dim_b = max_dims
        self.energy_tol, self.occupancies_tol = energy_tol, occupancies_tol
        self.carryover_threshold = carryover_threshold
        self.rng = rng
        self.occupancies = occupancies
        self.reference = None            # result the next iteration is compared with
        self.best = None
        self.carry_a = np.array([], dtype=np.int64)
        self.carry_b = np.array([], dtype=np.int64)

    # -- first half: noisy samples -> K pairs of sorted string lists ------------------------------
    def strings_for_iteration(self) -> list[tuple[np.ndarray, np.ndarray]]:
        from .configuration_recovery import recover_configurations
        from .counts import bitstring_matrix_to_integers
        from .subsampling import postselect_by_hamming_right_and_left, subsample

        if self.occupancies is None:
            rows, probs = postselect_by_hamming_right_and_left(
                self.bits, self.probs, hamming_right=self.n_alpha, hamming_left=self.n_beta)
            if not rows.size:
                raise ValueError(
                    "The input bit array did not contain any valid bitstrings. "
                    "Either pass a bit array that contains at least one valid bitstring "
                    "(with the correct right and left Hamming weights), or specify a value for initial_occupancies."
                )
        else:
            rows, probs = recover_configurations(self.bits, self.probs, self.occupancies, self.n_alpha,
                                                 self.n_beta, rand_seed=self.rng)
        out = []
        for batch in subsample(rows, probs, samples_per_batch=self.samples_per_batch,
                               num_batches=self.num_batches, rand_seed=self.rng):
            halves = []
            for cols in (batch[:, self.norb:], batch[:, : self.norb]):   # alpha = right half, beta = left
                halves.append(np.unique(bitstring_matrix_to_integers(cols), return_counts=True))
            (str_a, cnt_a), (str_b, cnt_b) = halves
            if self.symmetric:
                pooled = np.concatenate((str_a, str_b))[_by_descending(np.concatenate((cnt_a, cnt_b)))]
                ranked = np.concatenate((self.include_a, self.include_b, self.carry_a, pooled))
                list_a = list_b = _first_occurrences(ranked)[: self.max_dim_a]
                list_a.sort()
            else:
                list_a = _first_occurrences(np.concatenate(
                    (self.include_a, self.carry_a, str_a[_by_descending(cnt_a)])))[: self.max_dim_a]
                list_b = _first_occurrences(np.concatenate(
                    (self.include_b, self.carry_b, str_b[_by_descending(cnt_b)])))[: self.max_dim_b]
                list_a.sort()
                list_b.sort()
            out.append((list_a, list_b))
        return out

    # -- second half: K results -> best / converged / carry-over ----------------------------------
    def digest(self, results: list[SCIResult]) -> bool:
        """Returns True when the loop has converged."""
        winner = min(results, key=lambda r: r.energy)
        if self.best is None or winner.energy < self.best.energy:
            self.best = winner
        if self.reference is not None and abs(self.reference.energy - winner.energy) < self.energy_tol:
            drift = np.ravel(self.occupancies) - np.ravel(winner.orbital_occupancies)
            if np.linalg.norm(drift, ord=np.inf) < self.occupancies_tol:
                return True
        self.reference = winner
        self.occupancies = winner.orbital_occupancies
        # strings of every determinant whose |amplitude| reaches the threshold, ranked by marginal weight
        state = winner.sci_state
        held = getattr(_tls, "resident_amplitudes", {}).get(id(state))
        if held is not None and held[0] is state:
            # result of solve_sci_batch: selection and weights on the device (sqd_carryover), from the
            # amplitude matrix the solve left there -- no argsort over all n_det magnitudes
            rows, cols, weight_a, weight_b = _carryover_on_device(held[1], state.amplitudes.shape[1], held[2],
                                                                  self.carryover_threshold)
        else:
            # result of a foreign sci_solver plugin: its amplitudes only exist on the host
            magnitude = np.abs(state.amplitudes.reshape(-1))
            order = np.argsort(magnitude)
            big = order[np.searchsorted(magnitude, self.carryover_threshold, sorter=order):]
            rows, cols = np.divmod(big, state.amplitudes.shape[1])
            rows, cols = np.unique(rows), np.unique(cols)
            weight_a = np.sum(np.abs(state.amplitudes[rows]) ** 2, axis=1)
            weight_b = np.sum(np.abs(state.amplitudes[:, cols]) ** 2, axis=0)
        keep_a, keep_b = state.ci_strs_a[rows], state.ci_strs_b[cols]
        if self.symmetric:
            pooled = np.concatenate((keep_a, keep_b))[_by_descending(np.concatenate((weight_a, weight_b)))]
            self.carry_a = self.carry_b = _first_occurrences(pooled)
        else:
            self.carry_a, self.carry_b = keep_a[_by_descending(weight_a)], keep_b[_by_descending(weight_b)]
        return False


def diagonalize_fermionic_hamiltonian(
    one_body_tensor: np.ndarray,
    two_body_tensor: np.ndarray,
    bit_array,
    samples_per_batch: int,
    norb: int,
    nelec: tuple[int, int],
    *,
    num_batches: int = 1,
    energy_tol: float = 1e-8,
    occupancies_tol: float = 1e-5,
    max_iterations: int = 100,
    sci_solver=None,
    symmetrize_spin: bool = False,
    max_dim: int | tuple[int, int] | None = None,
    include_configurations=None,
    initial_occupancies: tuple[np.ndarray, np.ndarray] | None = None,
    carryover_threshold: float = 1e-4,
    callback=None,
    seed: int | np.random.Generator | None = None,
) -> SCIResult:
    """Sample-based quantum diagonalisation: the self-consistent configuration-recovery loop (reference
    ``fermion.py:204-462``, same arguments, same error messages, same use of the random generator).

    Per iteration: configuration recovery on the GPU (``recover_configurations``; Hamming post-selection
    on the first pass without ``initial_occupancies``) -> ``subsample`` -> string lists ->
    ``sci_solver`` (default: ``solve_sci_batch``, the K subspaces solved concurrently on the GPU) ->
    best energy / convergence / carry-over.  ``bit_array`` is duck-typed (``.array``, ``.num_bits``,
    ``.num_shots``), so qiskit is not needed.  Single process: the reference's optional MPI broadcast
    (``fermion.py:429, 451``) has no counterpart here -- shard the K subspaces with
    ``solve_sci_batch(devices=...)`` or one process per GPU instead.
    """
    from .counts import bit_array_to_arrays

    if max_iterations < 1:
        raise ValueError("Maximum number of iterations must be at least 1.")
    n_alpha, n_beta = nelec
    if symmetrize_spin and n_alpha != n_beta:
        raise ValueError(
            "Spin symmetrization is only possible if the numbers of alpha and beta "
            f"electrons are equal. Instead, got {n_alpha} and {n_beta}."
        )
    dims = max_dim if isinstance(max_dim, tuple) else (max_dim, max_dim)
    if symmetrize_spin and dims[0] != dims[1]:
        raise ValueError(
            "When requesting spin symmetrization, the maximum dimension must be "
            "the same for both spin alpha and spin beta. "
            f"Instead, got {dims[0]} and {dims[1]}"
        )
    if include_configurations is None:
        include = (np.array([], dtype=int), np.array([], dtype=int))
    elif isinstance(include_configurations, tuple):
        include = include_configurations
    else:
        include = (include_configurations, include_configurations)
    include = (np.unique(include[0]), np.unique(include[1]))
    solver = solve_sci_batch if sci_solver is None else sci_solver
    if getattr(solver, "func", solver) is solve_sci_batch:
        # our own solver (plain or functools.partial of it): the Hamiltonian does not change inside this call,
        # so its device copy is made once (per device) instead of once per iteration
        import functools

        solver = functools.partial(solver, _resident_integrals={})
    raw_bitstrings, raw_probs = bit_array_to_arrays(bit_array)
    run = _SQDRun(raw_bitstrings, raw_probs, norb, (n_alpha, n_beta), samples_per_batch, num_batches,
                  symmetrize_spin, include, dims, energy_tol, occupancies_tol, carryover_threshold,
                  np.random.default_rng(seed), initial_occupancies)
    for _ in range(max_iterations):
        results = solver(run.strings_for_iteration(), one_body_tensor, two_body_tensor, norb, nelec)
        if callback is not None:
            callback(results)
        if run.digest(results):
            break
    return cast(SCIResult, run.best)


class ShardGroup:
    """NCCL communicator for ONE diagonalisation sharded over the ranks of ``torch.distributed``.

    Rank 0 creates the NCCL unique id, ``torch.distributed`` (any backend) ships its 128 bytes, every
    rank joins with its current CUDA device.  Rows of the CI matrix are split into contiguous blocks of
    equal estimated sigma-build cost (the Hartree-Fock end of the string list is much heavier).
    """

    def __init__(self):
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("ShardGroup needs an initialised torch.distributed process group")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        lib = _lib.load()
        buf = C.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(lib.sqd_nccl_unique_id(buf), "sqd_nccl_unique_id")
        box = [buf.raw if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        self._comm = C.c_void_p()
        # NCCL announces its version on stdout when a communicator is created; a caller that prints
        # machine-readable output there (bench.py's JSON line) must not see it: route fd 1 to stderr meanwhile
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            rc = lib.sqd_nccl_init(box[0], self.rank, self.world, C.byref(self._comm))
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        _lib.check(rc, "sqd_nccl_init")

    def close(self):
        if self._comm:
            _lib.load().sqd_nccl_destroy(self._comm)
            self._comm = C.c_void_p()


def solve_sci_sharded(
    ci_strings: tuple[np.ndarray, np.ndarray],
    one_body_tensor: np.ndarray,
    two_body_tensor: np.ndarray,
    norb: int,
    nelec: tuple[int, int],
    *,
    group: ShardGroup,
    spin_sq: float | None = None,
    **kwargs,
) -> SCIResult:
    """One large diagonalisation with the sigma build sharded over the ranks of ``group``.

    SPMD, like the reference's MPI mode (``fermion.py:316-324``): every rank calls this with the same
    arguments and gets the same ``SCIResult``.  Vectors are replicated, rank r builds its block of rows of
    sigma and the blocks are summed with an NCCL all-reduce over NVLink (BASELINE.json config 5).
    """
    torch = _lib.require_cuda()
    one_body_tensor = np.asarray(one_body_tensor)
    two_body_tensor = np.asarray(two_body_tensor)
    norb, _ = one_body_tensor.shape
    shift = float(kwargs.pop("shift", _FIX_SPIN_DEFAULT_SHIFT))
    opts = _solver_options(kwargs)
    _tls.stats = []
    ints = _DeviceIntegrals(torch, one_body_tensor, two_body_tensor,
                            torch.device("cuda", torch.cuda.current_device()))
    r = _solve_on_device(ci_strings[0], ci_strings[1], norb, ints, spin_sq, shift, opts, want_spin=False,
                         want_rdm=False, shard_group=group)
    state = SCIState(amplitudes=r["amplitudes"], ci_strs_a=np.asarray(ci_strings[0]),
                     ci_strs_b=np.asarray(ci_strings[1]), norb=norb, nelec=tuple(nelec))
    return SCIResult(r["energy"], state, orbital_occupancies=r["occupancies"])


def solve_fermion(
    bitstring_matrix: tuple[np.ndarray, np.ndarray] | np.ndarray,
    /,
    hcore: np.ndarray,
    eri: np.ndarray,
    *,
    open_shell: bool = False,
    spin_sq: float | None = None,
    shift: float = 0.1,
    **kwargs,
) -> tuple[float, SCIState, tuple[np.ndarray, np.ndarray], float]:
    """Approximate the ground state given integrals and configurations (reference ``fermion.py:745-845``).

    Returns ``(e_sci, SCIState, (occ_a, occ_b), spin_squared)``; ``e_sci`` is the expectation value of
    the bare Hamiltonian in the returned state (``fermion.py:806-809, 824-827``).
    """
    torch = _lib.require_cuda()
    if isinstance(bitstring_matrix, tuple):
        ci_strs = bitstring_matrix
    else:
        ci_strs = bitstring_matrix_to_ci_strs(bitstring_matrix, open_shell=open_shell)
    ci_strs = _check_ci_strs(ci_strs)
    hcore = np.asarray(hcore)
    eri = np.asarray(eri)
    norb = hcore.shape[0]
    opts = _solver_options(kwargs)
    _tls.stats = []
    ints = _DeviceIntegrals(torch, hcore, eri, torch.device("cuda", torch.cuda.current_device()))
    r = _solve_on_device(ci_strs[0], ci_strs[1], norb, ints, spin_sq, shift, opts, want_spin=True,
                         want_rdm=False)
    state = SCIState(amplitudes=r["amplitudes"], ci_strs_a=ci_strs[0], ci_strs_b=ci_strs[1], norb=norb,
                     nelec=r["nelec"])
    return r["energy"], state, r["occupancies"], r["spin_square"]

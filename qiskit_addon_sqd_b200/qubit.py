"""Qubit-operator subspace projection + diagonalisation on B200.

Host-side mirror of ``qiskit_addon_sqd/qubit.py`` (``solve_qubit`` :29-75,
``project_operator_to_subspace`` :78-144, ``sort_and_remove_duplicates`` :147-164,
``matrix_elements_from_pauli`` :167-240): same names, arguments, return layout and error message.
``hamiltonian`` is duck-typed exactly as the reference uses it: ``.paulis`` (objects with little-endian
bool arrays ``.x`` and ``.z``), ``.coeffs`` and ``.size`` -- a ``qiskit.quantum_info.SparsePauliOp``
works unchanged, qiskit itself is not required.

The reference loops over Pauli terms in Python (jax vmap + numpy isin/searchsorted + a scipy sparse add
per term).  Here all terms are projected by one pair of CUDA launches (``sqd_pauli_project_count`` /
``_fill``) and the ground state comes from the device-resident Davidson (``sqd_csr_davidson``) instead
of ARPACK; anything other than "the lowest eigenpair" still goes through ``scipy.sparse.linalg.eigsh``
-- the reference's own solver -- but with the matrix-vector product on the GPU.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
from scipy.sparse import csr_matrix, spmatrix
from scipy.sparse.linalg import LinearOperator, eigsh

from . import _lib

_LEN_ERR = "Bitstrings (rows) in bitstring_matrix must have length < 64."


def _keys_device(torch, lib, bitstring_matrix: np.ndarray):
    n, nq = bitstring_matrix.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    if bitstring_matrix.dtype in (np.bool_, np.uint8, np.int8) and bitstring_matrix.flags.c_contiguous:
        host = bitstring_matrix.view(np.uint8)      # no host-side copy of a (possibly multi-GB) bool matrix
    else:
        host = np.ascontiguousarray(bitstring_matrix != 0).view(np.uint8)
    bits = torch.from_numpy(host).to(dev)
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.check(lib.sqd_bits_to_keys(_lib.ptr(bits), n, nq, _lib.ptr(keys), _lib.stream_ptr(torch)),
               "sqd_bits_to_keys")
    return keys


def _key_table(torch, lib, keys):
    """Hash table key -> row over the unique keys of the subspace (``sqd_key_table_build``)."""
    d = int(keys.numel())
    nbytes = lib.sqd_key_table_bytes(d)
    table = torch.empty(nbytes, dtype=torch.uint8, device=keys.device)
    _lib.check(lib.sqd_key_table_build(_lib.ptr(keys), d, _lib.ptr(table), nbytes, _lib.stream_ptr(torch)),
               "sqd_key_table_build")
    return table


def _masks(x: np.ndarray, z: np.ndarray) -> tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Per-term (xmask, zmask, #Y); bit k of a mask <-> qubit k <-> column nq-1-k (``qubit.py:214-216``)."""
    x = np.atleast_2d(np.asarray(x, dtype=bool))
    z = np.atleast_2d(np.asarray(z, dtype=bool))
    w = np.uint64(1) << np.arange(x.shape[1], dtype=np.uint64)
    xm = (x.astype(np.uint64) * w[None, :]).sum(axis=1, dtype=np.uint64)
    zm = (z.astype(np.uint64) * w[None, :]).sum(axis=1, dtype=np.uint64)
    ny = np.count_nonzero(x & z, axis=1).astype(np.int32)
    return xm, zm, ny


def sort_and_remove_duplicates(bitstring_matrix: np.ndarray) -> np.ndarray:
    """Sort a bitstring matrix by integer value and drop repeated rows (reference ``qubit.py:147-164``)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    bitstring_matrix = np.asarray(bitstring_matrix)
    if bitstring_matrix.shape[0] == 0:
        return bitstring_matrix
    n_bits = bitstring_matrix.shape[1]
    # keys, bitonic sort and run heads on the device (csrc/sortuniq.cu); equal keys are equal rows, so the
    # reference's "first occurrence of every distinct key" is the key itself written back as a row
    _, keys, _ = _lib.sort_unique(torch, None, _keys_device(torch, lib, bitstring_matrix))
    d = int(keys.numel())
    bits = torch.empty((d, n_bits), dtype=torch.uint8, device=keys.device)
    _lib.check(lib.sqd_keys_to_bits(0, _lib.ptr(keys), d, n_bits, _lib.ptr(bits), _lib.stream_ptr(torch)),
               "sqd_keys_to_bits")
    rows = _lib.download(torch, bits)
    return rows.view(bool) if bitstring_matrix.dtype == np.bool_ else rows.astype(bitstring_matrix.dtype)


def matrix_elements_from_pauli(bitstring_matrix: np.ndarray, pauli):
    """Sparse matrix elements of one Pauli in the subspace (reference ``qubit.py:167-240``).

    Returns ``(amplitudes, rows, cols)`` with ``A[rows[k], cols[k]] = amplitudes[k]``; rows are the
    source configurations, cols the index of the connected configuration.
    """
    bitstring_matrix = np.asarray(bitstring_matrix)
    if bitstring_matrix.shape[1] > 63:
        raise ValueError(_LEN_ERR)
    torch = _lib.require_cuda()
    lib = _lib.load()
    d = bitstring_matrix.shape[0]
    if d == 0:
        return np.zeros(0, dtype=np.complex128), np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    keys = _keys_device(torch, lib, bitstring_matrix)
    xm, zm, ny = _masks(pauli.x, pauli.z)
    diagonal = int(xm[0]) == 0
    dev = keys.device
    table = None if diagonal else _key_table(torch, lib, keys)
    col = None if diagonal else torch.empty(d, dtype=torch.int64, device=dev)
    amp = torch.empty(2 * d, dtype=torch.float64, device=dev)
    n_missing = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.check(lib.sqd_pauli_elements(_lib.ptr(keys), d, _lib.ptr(table), int(xm[0]), int(zm[0]), int(ny[0]),
                                      _lib.ptr(col), _lib.ptr(amp), _lib.ptr(n_missing), _lib.stream_ptr(torch)),
               "sqd_pauli_elements")
    del keys
    amp_h = amp.cpu().numpy().view(np.complex128)
    rows = np.arange(d)
    if diagonal:
        return amp_h, rows, rows.copy()       # a Pauli without X or Y connects every configuration to itself
    col_h = col.cpu().numpy()
    if int(n_missing.item()) == 0:
        return amp_h, rows, col_h
    mask = col_h >= 0
    return amp_h[mask], rows[mask], col_h[mask]


class _DeviceCSR:
    def __init__(self, d, row_ptr, col, val):
        self.d, self.row_ptr, self.col, self.val = d, row_ptr, col, val

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    def to_scipy(self) -> csr_matrix:
        data = self.val.cpu().numpy().view(np.complex128).reshape(-1)
        return csr_matrix((data, self.col.cpu().numpy(), self.row_ptr.cpu().numpy()),
                          shape=(self.d, self.d))


def _project_device(torch, lib, keys, hamiltonian, timings: dict | None = None) -> _DeviceCSR:
    d = int(keys.numel())
    dev = keys.device
    st = _lib.stream_ptr(torch)
    paulis = hamiltonian.paulis
    # qiskit's PauliList exposes the whole (T, nq) x and z tables; anything else is read term by term
    px, pz = getattr(paulis, "x", None), getattr(paulis, "z", None)
    if px is None or pz is None or np.ndim(px) != 2:
        plist = list(paulis)
        px = np.array([p.x for p in plist], dtype=bool).reshape(len(plist), -1)
        pz = np.array([p.z for p in plist], dtype=bool).reshape(len(plist), -1)
    T = int(np.shape(px)[0])
    coeffs = np.asarray(hamiltonian.coeffs, dtype=np.complex128).reshape(-1)
    if T == 0 or d == 0:
        z = torch.zeros(d + 1, dtype=torch.int32, device=dev)
        return _DeviceCSR(d, z, torch.zeros(0, dtype=torch.int32, device=dev),
                          torch.zeros(0, dtype=torch.float64, device=dev))
    xm, zm, ny = _masks(px, pz)
    # group terms by X mask; groups in order of first appearance, original order inside a group
    uniq, first, inv = np.unique(xm, return_index=True, return_inverse=True)
    g_order = np.argsort(first, kind="stable")
    g_rank = np.empty_like(g_order)
    g_rank[g_order] = np.arange(len(g_order))
    grp_of_term = g_rank[inv.reshape(-1)]
    perm = np.argsort(grp_of_term, kind="stable")
    counts = np.bincount(grp_of_term, minlength=len(uniq))
    grp_ptr = np.zeros(len(uniq) + 1, dtype=np.int32)
    np.cumsum(counts, out=grp_ptr[1:])
    grp_x = uniq[g_order]

    def up(a, dt):
        return torch.from_numpy(np.ascontiguousarray(a).view(dt) if a.dtype != dt else
                                np.ascontiguousarray(a)).to(dev)

    d_gx = up(grp_x.astype(np.uint64), np.int64)
    d_gp = up(grp_ptr, np.int32)
    d_z = up(zm[perm].astype(np.uint64), np.int64)
    d_ny = up(ny[perm].astype(np.int32), np.int32)
    d_cf = torch.from_numpy(np.ascontiguousarray(coeffs[perm]).view(np.float64)).to(dev)
    row_nnz = torch.empty(d, dtype=torch.int32, device=dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timings is not None else None
    if ev:
        ev[0].record()
    table = _key_table(torch, lib, keys)
    # the diagonal group (X mask 0: every row is its own image): summed once per row by a thread-per-row pre-pass
    diag_group, diag_val = -1, None
    zero = np.flatnonzero(grp_x == 0)
    if zero.size and counts[zero[0]] >= 32:
        diag_group = int(zero[0])
        diag_val = torch.empty(2 * d, dtype=torch.float64, device=dev)
        _lib.check(lib.sqd_pauli_diag_group(_lib.ptr(keys), d, _lib.ptr(d_z), _lib.ptr(d_ny), _lib.ptr(d_cf),
                                            int(grp_ptr[diag_group]), int(grp_ptr[diag_group + 1]),
                                            _lib.ptr(diag_val), st), "sqd_pauli_diag_group")
    _lib.check(lib.sqd_pauli_project_count(_lib.ptr(keys), d, _lib.ptr(table), _lib.ptr(d_gx), _lib.ptr(d_gp),
                                           len(uniq), _lib.ptr(d_z), _lib.ptr(d_ny), _lib.ptr(d_cf),
                                           diag_group, _lib.ptr(diag_val),
                                           _lib.ptr(row_nnz), st), "sqd_pauli_project_count")
    if ev:
        ev[1].record()
    row_ptr = torch.empty(d + 1, dtype=torch.int32, device=dev)
    total = C.c_int(0)
    _lib.check(lib.sqd_exclusive_scan(_lib.ptr(row_nnz), _lib.ptr(row_ptr), d, C.byref(total), st),
               "sqd_exclusive_scan")
    nnz = int(total.value)
    col = torch.empty(nnz, dtype=torch.int32, device=dev)
    val = torch.empty(2 * nnz, dtype=torch.float64, device=dev)
    col_tmp = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
    val_tmp = torch.empty(max(2 * nnz, 1), dtype=torch.float64, device=dev)
    if ev:
        ev[2].record()
    if nnz:
        _lib.check(lib.sqd_pauli_project_fill(_lib.ptr(keys), d, _lib.ptr(table), _lib.ptr(d_gx), _lib.ptr(d_gp),
                                              len(uniq), _lib.ptr(d_z), _lib.ptr(d_ny),
                                              _lib.ptr(d_cf), diag_group, _lib.ptr(diag_val),
                                              _lib.ptr(row_ptr), _lib.ptr(col_tmp),
                                              _lib.ptr(val_tmp), _lib.ptr(col), _lib.ptr(val), st),
                   "sqd_pauli_project_fill")
    if ev:
        ev[3].record()
        torch.cuda.synchronize()
        timings["table_and_count_ms"] = ev[0].elapsed_time(ev[1])
        timings["fill_and_sort_ms"] = ev[2].elapsed_time(ev[3])
    return _DeviceCSR(d, row_ptr, col, val)


def project_operator_to_subspace(
    bitstring_matrix: np.ndarray,
    hamiltonian,
    *,
    verbose: bool = False,
) -> spmatrix:
    """Project a Pauli operator onto the subspace spanned by the rows of ``bitstring_matrix``.

    As the reference (``qubit.py:78-144``): rows must be unique and sorted ascending by integer value
    (not checked); the result is a complex128 ``scipy.sparse.csr_matrix`` ``A`` with
    ``A[source, connected] = amplitude`` -- the TRANSPOSE of the operator's matrix, canonical format,
    exact zeros dropped.
    """
    bitstring_matrix = np.asarray(bitstring_matrix)
    if bitstring_matrix.shape[1] > 63:
        raise ValueError(_LEN_ERR)
    torch = _lib.require_cuda()
    lib = _lib.load()
    d = bitstring_matrix.shape[0]
    if d == 0:
        return csr_matrix((0, 0), dtype="complex128")
    if verbose:  # pragma: no cover
        print(f"Projecting {hamiltonian.size} Pauli terms onto {d} configurations on the GPU ...")
    keys = _keys_device(torch, lib, bitstring_matrix)
    return _project_device(torch, lib, keys, hamiltonian).to_scipy()


#: numpy view of ``sqd_component`` records (include/sqd_b200.h)
_COMPONENT_DTYPE = np.dtype([("root", np.int32), ("size", np.int32), ("row", np.int32), ("pad", np.int32),
                             ("lower", np.float64), ("diag", np.float64)])
_START_NOISE = 1e-3


def _lowest_eigenpair_device(torch, lib, csr: "_DeviceCSR", kw: dict, max_rounds: int = 64):
    """Lowest eigenpair of the projected operator with the device-resident Davidson.

    The projected matrix is usually block diagonal (sets of configurations no Pauli term connects), so the
    blocks are found first (``sqd_csr_components``: label propagation over the CSR structure):

    * a single-row block is an exact eigenpair ``(A_ii, e_i)``; the lowest one is reduced on the device;
    * multi-row blocks are visited in order of their Gershgorin lower bound ``min_i (A_ii - sum_j |A_ij|)``
      and only while that bound is below the best eigenvalue found so far.  Each is solved by a Davidson run
      started from the block's lowest-diagonal basis state plus a small deterministic perturbation on the rows
      OF THAT BLOCK (pyscf perturbs its unit start vector for the same reason, ARPACK starts from a random
      vector: a ground state with a node on the start row still has a component in the start vector).  The
      perturbation never leaves the block: rows of other blocks stay exactly zero, so the run is the
      diagonalisation of that block alone.

    Exact up to the Davidson tolerance.  Returns ``(None, None)`` when a run does not converge or more than
    ``max_rounds`` blocks would have to be solved; the caller then lets ARPACK drive the GPU matvec.
    """
    d = csr.d
    dev = csr.row_ptr.device
    st = _lib.stream_ptr(torch)
    tol = float(kw.get("tol", 0) or 0)
    tol = tol * tol if tol > 0 else 1e-14  # eigsh's tol bounds the relative accuracy of the Ritz value
    # basis of at most 16 vectors: up to there the Rayleigh-Ritz step of a cycle is the register-resident
    # Rayleigh-quotient iteration (a few microseconds); beyond, every cycle pays a full Jacobi decomposition
    max_space = int(min(kw.get("ncv") or 16, 16, _lib.MAX_SPACE))
    max_cycle = int(kw.get("maxiter") or 500)
    diag = torch.empty(d, dtype=torch.float64, device=dev)
    lower = torch.empty(d, dtype=torch.float64, device=dev)
    _lib.check(lib.sqd_csr_gershgorin(d, _lib.ptr(csr.row_ptr), _lib.ptr(csr.col), _lib.ptr(csr.val),
                                      _lib.ptr(diag), _lib.ptr(lower), st), "sqd_csr_gershgorin")
    label = torch.empty(d, dtype=torch.int32, device=dev)
    rec_cap = d // 2 + 1
    rec = torch.empty(rec_cap * _COMPONENT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    cws_bytes = lib.sqd_csr_components_workspace_bytes(d)
    cws = torch.empty(cws_bytes, dtype=torch.uint8, device=dev)
    head = (C.c_int32 * 3)()
    best_single = C.c_double(0.0)
    _lib.check(lib.sqd_csr_components(d, _lib.ptr(csr.row_ptr), _lib.ptr(csr.col), _lib.ptr(diag),
                                      _lib.ptr(lower), _lib.ptr(label), _lib.ptr(rec), rec_cap, head,
                                      C.byref(best_single), _lib.ptr(cws), cws_bytes, st), "sqd_csr_components")
    n_multi, n_single, single_row = int(head[0]), int(head[1]), int(head[2])
    comps = np.zeros(0, dtype=_COMPONENT_DTYPE)
    if n_multi:
        comps = rec[: n_multi * _COMPONENT_DTYPE.itemsize].cpu().numpy().view(_COMPONENT_DTYPE)
        comps = comps[np.lexsort((comps["root"], comps["lower"]))]   # the device appends in no fixed order

    ws_bytes = lib.sqd_csr_davidson_workspace_bytes(d, 1, max_space)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    start = torch.empty(2 * d, dtype=torch.float64, device=dev)
    evec = torch.empty(2 * d, dtype=torch.float64, device=dev)
    best_e, best_vec, best_row = None, None, -1

    def run():
        evals = (C.c_double * 1)()
        cycles, resid = C.c_int(0), C.c_double(0.0)
        _lib.check(lib.sqd_csr_davidson(d, _lib.ptr(csr.row_ptr), _lib.ptr(csr.col), _lib.ptr(csr.val), 1,
                                        max_space, max_cycle, tol, _lib.ptr(start), _lib.ptr(evec), evals,
                                        C.byref(cycles), C.byref(resid), _lib.ptr(ws), ws_bytes, st),
                   "sqd_csr_davidson")
        converged = cycles.value < max_cycle or resid.value <= 10.0 * np.sqrt(tol) * max(1.0, abs(evals[0]))
        return float(evals[0]), converged

    v0 = kw.get("v0")
    if v0 is not None:
        # the caller's start vector, over whatever blocks it touches; the block search below still runs
        start.copy_(torch.from_numpy(np.ascontiguousarray(v0, dtype=np.complex128).reshape(-1).view(np.float64)))
        e, ok = run()
        if not ok:
            return None, None
        best_e, best_vec = e, evec.clone()
    if n_single and (best_e is None or best_single.value < best_e):
        best_e, best_vec, best_row = float(best_single.value), None, single_row
    rounds = 0
    for c in comps:
        if best_e is not None and c["lower"] >= best_e:
            break   # sorted by lower bound: no later block can hold a lower eigenvalue either
        rounds += 1
        if rounds > max_rounds:
            return None, None
        _lib.check(lib.sqd_csr_component_start(d, _lib.ptr(label), int(c["root"]), int(c["row"]), _START_NOISE,
                                               _lib.ptr(start), st), "sqd_csr_component_start")
        e, ok = run()
        if not ok:
            return None, None
        if best_e is None or e < best_e:
            best_e, best_vec, best_row = e, evec.clone(), -1
    if best_vec is None:
        vec = np.zeros(d, dtype=np.complex128)
        vec[best_row] = 1.0
        return best_e, vec
    return best_e, best_vec.cpu().numpy().view(np.complex128)


def _native_ok(kw: dict) -> bool:
    if kw.get("k", 6) != 1 or kw.get("which", "LM") != "SA":
        return False
    return all(kw.get(key) is None for key in ("sigma", "M", "Minv", "OPinv")) and \
        kw.get("mode", "normal") == "normal" and kw.get("return_eigenvectors", True)


def solve_qubit(
    bitstring_matrix: np.ndarray,
    hamiltonian,
    *,
    verbose: bool = False,
    **scipy_kwargs,
) -> tuple[np.ndarray, np.ndarray]:
    """Energies and eigenstates of a Pauli Hamiltonian projected into a subspace (``qubit.py:29-75``).

    ``**scipy_kwargs`` has ``scipy.sparse.linalg.eigsh``'s meaning.  ``k=1, which="SA"`` (the usual
    ground-state call) runs the device-resident Davidson; any other request is served by ``eigsh``
    itself with the sparse matrix-vector product on the GPU.
    """
    bitstring_matrix = np.asarray(bitstring_matrix)
    if bitstring_matrix.shape[1] > 63:
        raise ValueError(_LEN_ERR)
    torch = _lib.require_cuda()
    lib = _lib.load()
    keys_all = _keys_device(torch, lib, bitstring_matrix)
    _, keys, _ = _lib.sort_unique(torch, None, keys_all)  # sorted ascending, duplicates removed (qubit.py:66)
    d = int(keys.numel())
    if verbose:  # pragma: no cover
        print(f"Projecting {hamiltonian.size} Pauli terms onto {d} configurations on the GPU ...")
    csr = _project_device(torch, lib, keys, hamiltonian)
    st = _lib.stream_ptr(torch)
    if verbose:  # pragma: no cover
        print("Diagonalizing Hamiltonian in the subspace...")

    if _native_ok(scipy_kwargs) and d > 1:
        e0, vec = _lowest_eigenpair_device(torch, lib, csr, scipy_kwargs)
        if vec is not None:
            return np.array([e0]), vec.reshape(d, 1)
        # (the search did not settle within its round budget: let ARPACK drive the GPU matvec below)

    # general eigsh request: ARPACK on the host, A @ x on the device
    x_dev = torch.empty(2 * d, dtype=torch.float64, device=keys.device)
    y_dev = torch.empty(2 * d, dtype=torch.float64, device=keys.device)

    def matvec(x):
        xx = np.ascontiguousarray(x, dtype=np.complex128).reshape(-1)
        x_dev.copy_(torch.from_numpy(xx.view(np.float64)))
        _lib.check(lib.sqd_csr_matvec_c128(d, _lib.ptr(csr.row_ptr), _lib.ptr(csr.col),
                                           _lib.ptr(csr.val), _lib.ptr(x_dev), _lib.ptr(y_dev), st),
                   "sqd_csr_matvec_c128")
        return y_dev.cpu().numpy().view(np.complex128)

    if d <= 1 or scipy_kwargs.get("sigma") is not None or scipy_kwargs.get("M") is not None:
        # shift-invert / generalized problems need a factorisation: hand scipy the explicit matrix
        return eigsh(csr.to_scipy(), **scipy_kwargs)
    op = LinearOperator((d, d), matvec=matvec, dtype=np.complex128)
    return eigsh(op, **scipy_kwargs)

"""Parity of the CUDA qubit path against the reference's goldens, its known answers and the oracle."""

import os

import numpy as np
import pytest
from scipy.sparse import coo_matrix
from scipy.sparse.linalg import eigsh

from oracle import qubit_oracle as qo
from qiskit_addon_sqd_b200._synthetic import PauliSum, PauliTerm, random_pauli_operator

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "qubit_golden.npz")

BS7 = np.array([[0, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0], [0, 0, 1, 1], [0, 1, 0, 0], [1, 0, 0, 0],
                [1, 1, 0, 0]])


def test_reference_known_answers(cuda_lib):
    """Ports of test/test_qubit.py:31-164."""
    from qiskit_addon_sqd_b200 import qubit

    amps, rows, cols = qubit.matrix_elements_from_pauli(
        np.array([[0, 0], [0, 1], [1, 0], [1, 1]]), PauliTerm([False, True], [True, False]))  # "XZ"
    assert (np.array([1, -1, 1, -1]) == amps).all()
    assert (np.array([0, 1, 2, 3]) == rows).all() and (np.array([2, 3, 0, 1]) == cols).all()

    op = PauliSum.from_labels(["XZIY"])
    amps, rows, cols = qubit.matrix_elements_from_pauli(BS7, op.paulis[0])
    assert np.allclose([-1j, 1j], amps) and np.allclose([1, 5], rows) and np.allclose([5, 1], cols)

    proj = qubit.project_operator_to_subspace(BS7, PauliSum.from_labels(["XZIY"], [0.5]))
    assert proj.shape == (7, 7) and np.allclose(proj.data, [-0.5j, 0.5j])
    assert proj.format == "csr" and proj.dtype == np.complex128

    test = coo_matrix((np.array([-1j, 1j]), (np.array([1, 5]), np.array([5, 1]))), (7, 7))
    e_test, _ = eigsh(test, k=1, which="SA")
    e, v = qubit.solve_qubit(BS7, op, k=1, which="SA")
    assert np.allclose(e_test, e) and v.shape == (7, 1)

    new = qubit.sort_and_remove_duplicates(np.array([[0, 0], [1, 0], [0, 1], [0, 0], [1, 1]]))
    assert (np.array([[0, 0], [0, 1], [1, 0], [1, 1]]) == new).all()

    z64 = PauliSum.from_labels(["Z" * 64])
    for fn in (qubit.solve_qubit, qubit.project_operator_to_subspace):
        with pytest.raises(ValueError) as e_info:
            fn(np.array([[1] * 64]), z64)
        assert e_info.value.args[0] == "Bitstrings (rows) in bitstring_matrix must have length < 64."
    with pytest.raises(ValueError):
        qubit.matrix_elements_from_pauli(np.array([[1] * 64]), z64.paulis[0])


@pytest.mark.parametrize("ci", range(5))
def test_matches_reference_golden(cuda_lib, ci):
    from qiskit_addon_sqd_b200 import qubit

    g = np.load(GOLD)
    rows = g[f"c{ci}_rows_in"]
    srt = qubit.sort_and_remove_duplicates(rows)
    assert np.array_equal(srt, g[f"c{ci}_rows_sorted"])  # bit-exact ordering / de-duplication
    labels = [str(s) for s in g[f"c{ci}_labels"]]
    op = PauliSum.from_labels(labels, g[f"c{ci}_coeffs"])
    proj = qubit.project_operator_to_subspace(srt, op)
    assert proj.has_canonical_format
    assert np.array_equal(proj.indptr, g[f"c{ci}_indptr"])    # bit-exact structure, zeros dropped
    assert np.array_equal(proj.indices, g[f"c{ci}_indices"])
    assert np.allclose(proj.data, g[f"c{ci}_data"], rtol=0, atol=1e-6)  # reference has a complex64 factor
    # ... and identical to the float64-exact oracle
    ref = qo.project_operator_to_subspace(srt, op)
    ref.sort_indices()
    assert np.array_equal(proj.data, ref.data)
    for ti in range(4):
        amp, r, c = qubit.matrix_elements_from_pauli(srt, op.paulis[ti])
        assert np.array_equal(r, g[f"c{ci}_t{ti}_row"]) and np.array_equal(c, g[f"c{ci}_t{ti}_col"])
        assert np.allclose(amp, g[f"c{ci}_t{ti}_amp"], rtol=0, atol=1e-6)
    n_terms = len(labels) - 3
    hop = PauliSum.from_labels(labels[:n_terms], np.real(g[f"c{ci}_coeffs"][:n_terms]))
    e, v = qubit.solve_qubit(rows, hop, k=1, which="SA")
    assert abs(e[0] - g[f"c{ci}_e0"][0]) < 1e-7
    # eigenvector of the matrix the reference diagonalises (A = transpose of the operator)
    A = qo.project_operator_to_subspace(srt, hop)
    assert np.linalg.norm(A @ v[:, 0] - e[0] * v[:, 0]) < 1e-6
    assert abs(np.linalg.norm(v[:, 0]) - 1) < 1e-10


def test_general_eigsh_requests_use_gpu_matvec(cuda_lib):
    from qiskit_addon_sqd_b200 import qubit

    g = np.load(GOLD)
    ci = 1
    rows = g[f"c{ci}_rows_in"]
    labels = [str(s) for s in g[f"c{ci}_labels"]]
    n_terms = len(labels) - 3
    hop = PauliSum.from_labels(labels[:n_terms], np.real(g[f"c{ci}_coeffs"][:n_terms]))
    e3, v3 = qubit.solve_qubit(rows, hop, k=3, which="SA")
    e_ref, _ = qo.solve_qubit(rows, hop, k=3, which="SA")
    assert np.allclose(np.sort(e3), np.sort(e_ref), atol=1e-8) and v3.shape[1] == 3
    e1 = qubit.solve_qubit(rows, hop, k=1, which="SA")[0]
    assert abs(e1[0] - np.min(e3)) < 1e-8


def test_medium_random_operator_against_oracle(cuda_lib):
    """40-qubit style workload at a size the numpy oracle finishes in seconds."""
    from qiskit_addon_sqd_b200 import qubit

    nq, d0 = 40, 3000
    rng = np.random.default_rng(103)
    base = rng.integers(0, 2, nq).astype(bool)
    rows = np.tile(base, (d0, 1))
    for r in range(d0):
        k = rng.integers(0, 5)
        if k:
            rows[r, rng.choice(nq, k, replace=False)] ^= True
    x, z, c = random_pauli_operator(nq, 120, 4, 3, 7)
    op = PauliSum(x, z, c)
    srt = qubit.sort_and_remove_duplicates(rows)
    assert np.array_equal(srt, qo.sort_and_remove_duplicates(rows))
    proj = qubit.project_operator_to_subspace(srt, op)
    ref = qo.project_operator_to_subspace(srt, op)
    ref.sort_indices()
    assert np.array_equal(proj.indptr, ref.indptr) and np.array_equal(proj.indices, ref.indices)
    assert np.array_equal(proj.data, ref.data)  # same summation order -> bit-equal
    e, v = qubit.solve_qubit(rows, op, k=1, which="SA")
    e_ref, _ = eigsh(ref, k=1, which="SA")
    assert abs(e[0] - e_ref[0]) < 1e-8


def _device_csr(torch, A):
    from qiskit_addon_sqd_b200 import qubit

    A = A.tocsr().astype(np.complex128)
    A.sort_indices()
    dev = torch.device("cuda", torch.cuda.current_device())
    return qubit._DeviceCSR(A.shape[0], torch.from_numpy(A.indptr.astype(np.int32)).to(dev),
                            torch.from_numpy(A.indices.astype(np.int32)).to(dev),
                            torch.from_numpy(np.ascontiguousarray(A.data).view(np.float64)).to(dev))


def test_components_match_scipy(cuda_lib):
    """Block structure found on the device == scipy.sparse.csgraph.connected_components."""
    import ctypes as C

    import torch
    from scipy.sparse import random as sprandom
    from scipy.sparse.csgraph import connected_components

    from qiskit_addon_sqd_b200 import _lib, qubit

    lib = _lib.load()
    rng = np.random.default_rng(5)
    for d, density in ((1, 1.0), (50, 0.02), (3000, 0.0004), (20000, 0.00004)):
        M = sprandom(d, d, density=density, random_state=rng.integers(1 << 30), format="csr")
        A = (M + M.T + coo_matrix((rng.standard_normal(d), (np.arange(d), np.arange(d))), shape=(d, d))).tocsr()
        # a long chain inside: worst case for label propagation
        if d >= 3000:
            idx = np.arange(1000, 1999)
            A = (A + coo_matrix((np.ones(999), (idx, idx + 1)), shape=(d, d))
                 + coo_matrix((np.ones(999), (idx + 1, idx)), shape=(d, d))).tocsr()
        csr = _device_csr(torch, A)
        dev = csr.row_ptr.device
        st = _lib.stream_ptr(torch)
        diag = torch.empty(d, dtype=torch.float64, device=dev)
        lower = torch.empty(d, dtype=torch.float64, device=dev)
        _lib.check(lib.sqd_csr_gershgorin(d, _lib.ptr(csr.row_ptr), _lib.ptr(csr.col), _lib.ptr(csr.val),
                                          _lib.ptr(diag), _lib.ptr(lower), st), "gershgorin")
        label = torch.empty(d, dtype=torch.int32, device=dev)
        cap = d // 2 + 1
        rec = torch.empty(cap * qubit._COMPONENT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        wsb = lib.sqd_csr_components_workspace_bytes(d)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        head = (C.c_int32 * 3)()
        best = C.c_double(0.0)
        _lib.check(lib.sqd_csr_components(d, _lib.ptr(csr.row_ptr), _lib.ptr(csr.col), _lib.ptr(diag), _lib.ptr(lower),
                                          _lib.ptr(label), _lib.ptr(rec), cap, head, C.byref(best), _lib.ptr(ws),
                                          wsb, st), "components")
        lab = label.cpu().numpy()
        nc, ref = connected_components(A, directed=False)
        # same partition, and every label is the smallest member of its block
        first = np.full(nc, d, dtype=np.int64)
        np.minimum.at(first, ref, np.arange(d))
        assert np.array_equal(lab, first[ref])
        sizes = np.bincount(ref)
        assert head[0] == np.count_nonzero(sizes > 1) and head[1] == np.count_nonzero(sizes == 1)
        dg = A.diagonal().real
        if head[1]:
            single = np.flatnonzero(sizes[ref] == 1)
            assert best.value == dg[single].min() and head[2] == single[np.argmin(dg[single])]
        recs = rec[: head[0] * qubit._COMPONENT_DTYPE.itemsize].cpu().numpy().view(qubit._COMPONENT_DTYPE)
        off = np.asarray(abs(A).sum(axis=1)).reshape(-1) - np.abs(dg)
        for r in recs:
            members = np.flatnonzero(lab == r["root"])
            assert r["size"] == len(members) and r["diag"] == dg[members].min()
            assert r["row"] == members[np.argmin(dg[members])]
            assert abs(r["lower"] - (dg[members] - off[members]).min()) < 1e-12


def test_ground_state_with_node_on_start_row(cuda_lib):
    """The block's lowest state has zero amplitude on the block's lowest-diagonal row: a unit start
    vector never reaches it; the perturbed start does.  Plus single-row blocks around it."""
    import torch

    from qiskit_addon_sqd_b200 import _lib, qubit

    d = 40
    A = np.diag(np.linspace(1.0, 2.0, d)).astype(np.complex128)
    a, b, c = 7, 8, 9
    A[a, a] = A[c, c] = 0.1
    A[b, b] = 0.0
    A[a, b] = A[b, a] = A[b, c] = A[c, b] = 0.1
    A[a, c] = A[c, a] = 5.0                      # (a - c)/sqrt(2): energy 0.1 - 5, node on b
    A[20, 21] = A[21, 20] = 0.3j * 1j            # another small block, not the lowest
    from scipy.sparse import csr_matrix

    csr = _device_csr(torch, csr_matrix(A))
    e, vec = qubit._lowest_eigenpair_device(torch, _lib.load(), csr, {"k": 1, "which": "SA"})
    w = np.linalg.eigvalsh(A)
    assert abs(e - w[0]) < 1e-10 and abs(e + 4.9) < 1e-10
    assert abs(abs(vec[a]) - 2 ** -0.5) < 1e-7 and abs(vec[b]) < 1e-7
    # all blocks single rows: the answer is the lowest diagonal entry, no Davidson run at all
    D = np.diag(np.array([3.0, -2.5, 0.5, -2.5]))
    e, vec = qubit._lowest_eigenpair_device(torch, _lib.load(), _device_csr(torch, csr_matrix(D)), {})
    assert e == -2.5 and np.array_equal(vec, np.array([0, 1, 0, 0], dtype=np.complex128))

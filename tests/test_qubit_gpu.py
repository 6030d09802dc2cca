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

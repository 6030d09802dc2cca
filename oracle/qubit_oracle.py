"""CPU oracle for the qubit path.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Vectorised numpy restatement of ``qiskit_addon_sqd/qubit.py``:
  * ``_int_conversion_from_bts_array`` :280-300  -> ``bits_to_keys``
  * ``sort_and_remove_duplicates``     :147-164
  * ``_connected_elements_and_amplitudes_bool`` :243-277 and ``matrix_elements_from_pauli`` :167-240
  * ``project_operator_to_subspace``   :78-144  (returns csr, TRANSPOSE convention A[source, image])
  * ``solve_qubit``                    :29-75   (scipy eigsh, the reference's own dependency)

PARITY STATUS: pinned.  ``tests/golden/make_golden.py`` runs the UNMODIFIED reference module (with
import shims for qiskit/jax, ``oracle/shims``) and stores its outputs in ``tests/golden/qubit_*.npz``;
``tests/test_oracle_cpu.py`` checks this restatement against them, together with the known answers of
the reference's own ``test/test_qubit.py:107-164``.

The one deliberate deviation: the reference multiplies by ``jnp.array(1j, dtype="complex64") ** imag``
(:267), which is only single-precision exact; the oracle (and the CUDA path) use the exact value
``i**(#Y)``, so Y-type amplitudes agree with the reference to ~1e-7 (the reference's own test uses
``allclose`` there, ``test/test_qubit.py:145``) and X/Z-type amplitudes exactly.
"""

from __future__ import annotations

import numpy as np
from scipy.sparse import coo_matrix, csr_matrix
from scipy.sparse.linalg import eigsh

_ERR = "Bitstrings (rows) in bitstring_matrix must have length < 64."


def bits_to_keys(bitstring_matrix: np.ndarray) -> np.ndarray:
    bitstring_matrix = np.asarray(bitstring_matrix, dtype=bool)
    n, nq = bitstring_matrix.shape
    w = (np.int64(1) << np.arange(nq - 1, -1, -1, dtype=np.int64))
    return (bitstring_matrix.astype(np.int64) * w[None, :]).sum(axis=1).astype(np.int64)


def sort_and_remove_duplicates(bitstring_matrix: np.ndarray) -> np.ndarray:
    keys = bits_to_keys(bitstring_matrix)
    _, idx = np.unique(keys, return_index=True)
    return np.asarray(bitstring_matrix)[idx, :]


def pauli_masks(pauli) -> tuple[int, int, int]:
    """(xmask, zmask, number of Y) with bit k of a mask <-> qubit k <-> column nq-1-k (qubit.py:214-216)."""
    x = np.asarray(pauli.x, dtype=bool)
    z = np.asarray(pauli.z, dtype=bool)
    xm = sum(1 << k for k in range(len(x)) if x[k])
    zm = sum(1 << k for k in range(len(z)) if z[k])
    return xm, zm, int(np.count_nonzero(x & z))


def matrix_elements_from_pauli(bitstring_matrix: np.ndarray, pauli):
    if bitstring_matrix.shape[1] > 63:
        raise ValueError(_ERR)
    keys = bits_to_keys(bitstring_matrix)
    d = len(keys)
    xm, zm, ny = pauli_masks(pauli)
    conn = keys ^ np.int64(xm)
    par = np.bitwise_count((keys & np.int64(zm)).astype(np.uint64)) & 1
    amp = (1 - 2 * par.astype(np.float64)) * (1j ** (ny % 4))
    pos = np.searchsorted(keys, conn)
    pos[pos >= d] = d - 1 if d else 0
    mask = keys[pos] == conn if d else np.zeros(0, dtype=bool)
    return amp[mask].astype(np.complex128), np.arange(d)[mask], pos[mask]


def project_operator_to_subspace(bitstring_matrix: np.ndarray, hamiltonian) -> csr_matrix:
    if bitstring_matrix.shape[1] > 63:
        raise ValueError(_ERR)
    d = bitstring_matrix.shape[0]
    operator = coo_matrix((d, d), dtype="complex128")
    for i, pauli in enumerate(hamiltonian.paulis):
        amp, rows, cols = matrix_elements_from_pauli(bitstring_matrix, pauli)
        operator += hamiltonian.coeffs[i] * coo_matrix((amp, (rows, cols)), (d, d))
    return operator


def solve_qubit(bitstring_matrix: np.ndarray, hamiltonian, **scipy_kwargs):
    if bitstring_matrix.shape[1] > 63:
        raise ValueError(_ERR)
    bitstring_matrix = sort_and_remove_duplicates(bitstring_matrix)
    ham = project_operator_to_subspace(bitstring_matrix, hamiltonian)
    return eigsh(ham, **scipy_kwargs)

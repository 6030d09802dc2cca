"""B200-native hot path of sample-based quantum diagonalisation (SQD).

Drop-in for the subspace-projection-and-diagonalisation functions of Qiskit/qiskit-addon-sqd:
``fermion.solve_fermion / solve_sci / solve_sci_batch``, ``qubit.project_operator_to_subspace /
matrix_elements_from_pauli / solve_qubit`` and ``configuration_recovery.recover_configurations``.
Host code is Python over a C-ABI (``include/sqd_b200.h``, ``libsqd_b200.so``) of hand-written sm_100a
CUDA kernels.  Importing the package does not need a GPU; calling any solver does (no CPU fallback).
"""

__version__ = "0.1.0"

from . import _lib  # noqa: F401

__all__ = ["fermion", "qubit", "configuration_recovery"]

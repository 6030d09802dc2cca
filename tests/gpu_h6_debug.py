"""Diagnostics: convergence history of the H6 sampled-subspace solves (false-convergence hunt)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from itertools import combinations

from oracle import fermion_oracle as fo, sto3g
from qiskit_addon_sqd_b200 import fermion

n_atoms = 6
h, g, en = sto3g.hydrogen_chain(n_atoms, 1.4)
full = np.array(sorted(sum(1 << i for i in c) for c in combinations(range(n_atoms), n_atoms // 2)), dtype=np.int64)
rng = np.random.default_rng(n_atoms)
for trial in range(3):
    sa = np.sort(rng.choice(full, size=max(2, len(full) * 2 // 3), replace=False))
    sb = np.sort(rng.choice(full, size=max(2, len(full) // 2), replace=False))
    for spin_sq in (None, 0.0):
        e_ref, *_ = fo.solve_dense(sa, sb, h, g, n_atoms, spin_sq=spin_sq, shift=0.1)
        H = fo.projected_hamiltonian(sa, sb, h, g, n_atoms)
        if spin_sq is not None:
            S2 = fo.spin_square_matrix(sa, sb, n_atoms)
            H = H + 0.1 * (S2 - spin_sq * np.eye(len(H)))
        w = np.linalg.eigvalsh(H)
        e, st, oc, ss = fermion.solve_fermion((sa, sb), h, g, spin_sq=spin_sq, shift=0.1)
        cyc = fermion.last_solve_stats()[0].cycles
        print(f"trial {trial} spin {spin_sq}: e-e_ref {e - e_ref:.2e} cycles {cyc} lowest eigenvalues of the operator {w[:4]}")
        if abs(e - e_ref) > 1e-9:
            for k in range(max(1, cyc - 12), cyc + 8):
                ek, *_ = fermion.solve_fermion((sa, sb), h, g, spin_sq=spin_sq, shift=0.1, max_cycle=k)
                s = fermion.last_solve_stats()[0]
                print(f"   max_cycle {k}: cycles {s.cycles} conv {s.converged} E-Eref {ek - e_ref:.3e} resid {s.residual:.2e}")

import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, itertools
from oracle import fermion_oracle as fo
from qiskit_addon_sqd_b200 import fermion
from qiskit_addon_sqd_b200._synthetic import random_integrals
norb, nel = 6, 3
h, g = random_integrals(norb, 77)
sa = np.array(sorted(sum(1 << i for i in c) for c in itertools.combinations(range(norb), nel)))
e_ref, c_ref, _, s2_ref, w0 = fo.solve_dense(sa, sa, h, g, norb, spin_sq=2.0, shift=0.5)
print("ref", e_ref, s2_ref, w0)
ci0 = np.random.default_rng(3).standard_normal((len(sa), len(sa)))
for kw in [dict(), dict(max_space=30), dict(max_cycle=400), dict(tol=1e-9)]:
    e, state, occ, s2 = fermion.solve_fermion((sa, sa), h, g, spin_sq=2.0, shift=0.5, ci0=ci0, **kw)
    st = fermion.last_solve_stats()[-1]
    print(kw, "E", e, "s2", s2, "stats", st)
# no penalty but ci0
e, state, occ, s2 = fermion.solve_fermion((sa, sa), h, g, ci0=ci0)
print("no penalty ci0:", e, fo.solve_dense(sa, sa, h, g, norb)[0], fermion.last_solve_stats()[-1])
e, state, occ, s2 = fermion.solve_fermion((sa, sa), h, g, spin_sq=2.0, shift=0.5, ci0=c_ref)
print("start from exact:", e, s2, fermion.last_solve_stats()[-1])

"""Stand-in for the few ``pyscf`` names the reference's ``fermion.py`` imports (``fermion.py:25-33``).

TEST INFRASTRUCTURE ONLY.  pyscf is not installed in the build image; this shim lets the UNMODIFIED
reference package import and run its own SQD loop (``diagonalize_fermionic_hamiltonian``, ``solve_sci``,
``solve_fermion``) with the arithmetic of ``pyscf.fci.selected_ci`` replaced by the dense CPU oracle
(``oracle/fermion_oracle.py``).  It is used by ``tests/golden/make_golden.py`` to pin the loop logic,
string handling and result conventions of our drop-in against the reference's own orchestration code.
"""
from . import fci  # noqa: F401

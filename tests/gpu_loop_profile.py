"""cProfile of the whole SQD loop at BASELINE configs[1] (host-side hot spots; run under gpurun, not a test)."""
import cProfile
import functools
import io
import pstats
import sys
import time

sys.path.insert(0, ".")
import torch  # noqa: E402

from qiskit_addon_sqd_b200 import fermion  # noqa: E402
from qiskit_addon_sqd_b200._synthetic import noisy_samples, random_integrals  # noqa: E402

norb, nelec = 16, (5, 5)
h, g = random_integrals(norb, 102)
record = noisy_samples(norb, nelec, shots=10_000, n_strings=400, noise=0.04, seed=202)
solver = functools.partial(fermion.solve_sci_batch, spin_sq=0.0, compute_rdms=False)


def run():
    return fermion.diagonalize_fermionic_hamiltonian(
        h, g, record, samples_per_batch=300, norb=norb, nelec=nelec, num_batches=5, max_iterations=3,
        max_dim=100, sci_solver=solver, symmetrize_spin=True, seed=5)


for _ in range(3):
    run()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    run()
print("loop ms", 1e3 * (time.perf_counter() - t0) / 5)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    run()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue())

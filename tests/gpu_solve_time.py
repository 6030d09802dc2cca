"""Diagnostics: wall time of solve_sci_batch for K subspaces of a bench workload (no RDMs)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from qiskit_addon_sqd_b200 import fermion  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
extra = {}
if len(sys.argv) > 4:   # fixed number of cycles (timing experiments): max_cycle, never converge
    extra = dict(max_cycle=int(sys.argv[4]), tol=1e-30)
if os.environ.get("MAX_SPACE"):
    extra["max_space"] = int(os.environ["MAX_SPACE"])
norb, nelec, h, g, batches = bench.make_batches(wl, 0, K)
for _ in range(3):
    res = fermion.solve_sci_batch(batches, h, g, norb, nelec, compute_rdms=False, **extra)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    res = fermion.solve_sci_batch(batches, h, g, norb, nelec, compute_rdms=False, **extra)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
st = fermion.last_solve_stats()
print(f"{wl} K={K}: {1e3 * dt:.2f} ms per batch; cycles {[s.cycles for s in st]}; E0 {res[0].energy:.10f}")

"""Diagnostics: where does the host time of solve_sci_batch go? (cProfile of the calling thread + wall clock)"""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from qiskit_addon_sqd_b200 import fermion  # noqa: E402

norb, nelec, h, g, batches = bench.make_batches("c4", 0, 8)
for _ in range(3):
    fermion.solve_sci_batch(batches, h, g, norb, nelec)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    fermion.solve_sci_batch(batches, h, g, norb, nelec)
torch.cuda.synchronize()
print("e2e ms/step (8 subspaces):", (time.perf_counter() - t0) / 5 * 1e3)
t0 = time.perf_counter()
for _ in range(5):
    fermion.solve_sci_batch(batches[:1], h, g, norb, nelec)
torch.cuda.synchronize()
print("e2e ms (1 subspace):", (time.perf_counter() - t0) / 5 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    fermion.solve_sci_batch(batches[:1], h, g, norb, nelec)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

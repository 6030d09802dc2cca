"""Driver for ncu: one subspace diagonalisation of the bench workload (default c4), single stream."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from qiskit_addon_sqd_b200 import fermion  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
norb, nelec, h, g, batches = bench.make_batches(wl, 0, 1)
for _ in range(reps):
    res = fermion.solve_sci_batch(batches, h, g, norb, nelec)
torch.cuda.synchronize()
st = fermion.last_solve_stats()[0]
print("energy", res[0].energy, st)

"""Driver for ncu: projection + solve of the C3 qubit workload (bench_extras inputs)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_extras as bx
from qiskit_addon_sqd_b200 import _lib, qubit

rows, op = bx._c3_inputs()
lib = _lib.load()
_, keys, _ = _lib.sort_unique(torch, None, qubit._keys_device(torch, lib, rows))
for _ in range(2):
    csr = qubit._project_device(torch, lib, keys, op)
torch.cuda.synchronize()
if len(sys.argv) > 1:
    e, v = qubit.solve_qubit(rows, op, k=1, which="SA")
    print(e)

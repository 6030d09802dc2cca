"""Diagnostics: per-round convergence of the device Davidson on the C3 test operator."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from qiskit_addon_sqd_b200 import _lib, qubit
from qiskit_addon_sqd_b200._synthetic import PauliSum, random_pauli_operator

nq, d0 = 40, 100_000
rng = np.random.default_rng(103)
base = rng.integers(0, 2, nq).astype(bool)
rows = np.tile(base, (d0, 1))
k = np.minimum(rng.geometric(0.18, d0) + 1, 14)
for r in range(d0):
    rows[r, rng.choice(nq, k[r], replace=False)] ^= True
x, z, c = random_pauli_operator(nq, 2500, 4, 4, 103)
op = PauliSum(x, z, c)
lib = _lib.load()
keys = torch.unique(qubit._keys_device(torch, lib, rows))
d = int(keys.numel())
csr = qubit._project_device(torch, lib, keys, op)
dev = keys.device
st = _lib.stream_ptr(torch)
diag = torch.empty(d, dtype=torch.float64, device=dev)
lower = torch.empty(d, dtype=torch.float64, device=dev)
lib.sqd_csr_gershgorin(d, _lib.ptr(csr.row_ptr), _lib.ptr(csr.col), _lib.ptr(csr.val), _lib.ptr(diag), _lib.ptr(lower), st)
row = int(torch.argmin(diag).item())
print("d", d, "nnz", csr.nnz, "argmin row", row, "diag min", float(diag[row]), "lower min", float(lower.min()))
noise = torch.from_numpy(2.0 * ((np.arange(1, d + 1) * 0.6180339887498949) % 1.0) - 1.0).to(dev)
A = csr.to_scipy()
for scale in (0.0, 1e-3):
    for max_space in (12, 20):
        for max_cycle in (100, 500, 2000):
            ws_bytes = lib.sqd_csr_davidson_workspace_bytes(d, 1, max_space)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            start = torch.zeros(2 * d, dtype=torch.float64, device=dev)
            start.view(d, 2)[:, 0] = scale * noise
            start[2 * row] = 1.0
            evec = torch.empty(2 * d, dtype=torch.float64, device=dev)
            evals = (C.c_double * 1)()
            cycles, resid = C.c_int(0), C.c_double(0.0)
            rc = lib.sqd_csr_davidson(d, _lib.ptr(csr.row_ptr), _lib.ptr(csr.col), _lib.ptr(csr.val), 1, max_space,
                                      max_cycle, 1e-14, _lib.ptr(start), _lib.ptr(evec), evals, C.byref(cycles),
                                      C.byref(resid), _lib.ptr(ws), ws_bytes, st)
            v = evec.cpu().numpy().view(np.complex128)
            r = A @ v - evals[0] * v
            print(f"scale {scale} space {max_space} max_cycle {max_cycle}: rc {rc} E {evals[0]:.10f} cycles {cycles.value} "
                  f"resid {resid.value:.2e} true |r| {np.linalg.norm(r):.2e} |v| {np.linalg.norm(v):.6f}", flush=True)
from scipy.sparse.linalg import eigsh
e, _ = eigsh(A, k=2, which="SA")
print("eigsh:", e)

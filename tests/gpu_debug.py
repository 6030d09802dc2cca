import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from scipy.sparse.linalg import eigsh
from qiskit_addon_sqd_b200 import qubit, _lib
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
srt = qubit.sort_and_remove_duplicates(rows)
A = qubit.project_operator_to_subspace(srt, op)
d = A.shape[0]
print("d", d, "nnz", A.nnz, "diag range", A.diagonal().real.min(), A.diagonal().real.max())
t0 = time.time(); e_ref, v_ref = eigsh(A, k=1, which="SA"); print("eigsh", e_ref, time.time() - t0, "s")
lib = _lib.load()
keys = qubit._keys_device(torch, lib, srt)
csr = qubit._project_device(torch, lib, keys, op)
for max_space, max_cycle in [(12, 500), (20, 500), (32, 500), (32, 3000)]:
    ws_bytes = lib.sqd_csr_davidson_workspace_bytes(d, 1, max_space)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    evec = torch.empty(2 * d, dtype=torch.float64, device="cuda")
    evals = (C.c_double * 1)(); cycles = C.c_int(0)
    t0 = time.time()
    _lib.check(lib.sqd_csr_davidson(d, _lib.ptr(csr.row_ptr), _lib.ptr(csr.col), _lib.ptr(csr.val), 1, max_space,
                                    max_cycle, 1e-14, _lib.ptr(evec), evals, C.byref(cycles), _lib.ptr(ws), ws_bytes,
                                    _lib.stream_ptr(torch)))
    v = evec.cpu().numpy().view(np.complex128)
    print(max_space, max_cycle, "theta", evals[0], "cycles", cycles.value, "res", np.linalg.norm(A @ v - evals[0] * v),
          "time", time.time() - t0)

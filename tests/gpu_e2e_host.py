"""Diagnostics: host-side phases of the end-to-end plugin call (solve_sci_batch with host inputs)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from qiskit_addon_sqd_b200 import fermion  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
norb, nelec, h, g, batches = bench.make_batches(wl, 0, K)
for rdm in (True, False):
    for _ in range(3):
        fermion.solve_sci_batch(batches, h, g, norb, nelec, compute_rdms=rdm)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        out = fermion.solve_sci_batch(batches, h, g, norb, nelec, compute_rdms=rdm)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    st = fermion.last_solve_stats()
    hm = np.array([s.host_ms for s in st])
    print(f"compute_rdms={rdm}: {dt*1e3:.2f} ms per call; per-solve host ms mean (prep, library call, downloads) = "
          f"{hm.mean(axis=0).round(2).tolist()}, max = {hm.max(axis=0).round(2).tolist()}")
t0 = time.perf_counter()
for _ in range(10):
    ints = fermion._DeviceIntegrals(torch, h, g, torch.device("cuda", 0))
torch.cuda.synchronize()
print(f"integral upload: {(time.perf_counter()-t0)*100:.2f} ms per call")

# finer: where does a call with RDMs spend its wall time?
import gc  # noqa: E402

for rdm in (True,):
    out = None
    for rep in range(4):
        t0 = time.perf_counter()
        new = fermion.solve_sci_batch(batches, h, g, norb, nelec, compute_rdms=rdm)
        t1 = time.perf_counter()
        out = new          # drops the previous generation of results
        del new
        t2 = time.perf_counter()
        gc.collect()
        t3 = time.perf_counter()
        print(f"rep {rep}: call {1e3*(t1-t0):.2f} ms, drop previous results {1e3*(t2-t1):.2f} ms, gc {1e3*(t3-t2):.2f} ms")
    t0 = time.perf_counter()
    keep = [fermion.solve_sci_batch(batches, h, g, norb, nelec, compute_rdms=rdm) for _ in range(3)]
    print(f"3 calls keeping all results: {1e3*(time.perf_counter()-t0)/3:.2f} ms per call")

"""Scale-up point of SURVEY 8(d): ONE subspace of 1e4 x 1e4 = 1e8 determinants, (30e,30o) -- beyond the row
staging of the v1/v2 sigma kernels, so the wide kernel runs.  Prints one JSON line: table sizes, sigma-build time,
and the split of a few Davidson cycles into sigma and streaming vector kernels (run under gpurun; not a test)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from qiskit_addon_sqd_b200 import fermion  # noqa: E402
from qiskit_addon_sqd_b200._synthetic import hf_centred_strings, random_integrals  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
cycles = int(sys.argv[2]) if len(sys.argv) > 2 else 4
norb, ne = 30, 15
h, g = random_integrals(norb, 108)
sa = hf_centred_strings(norb, ne, n, 21)
sb = hf_centred_strings(norb, ne, n, 22)
dev = torch.device("cuda", 0)
ints = fermion._DeviceIntegrals(torch, h, g, dev)
opts = fermion._solver_options({"max_cycle": cycles})
t0 = time.perf_counter()
r = fermion._solve_on_device(sa, sb, norb, ints, None, 0.2, opts, want_spin=False, want_rdm=False,
                             download=False, profile=True)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
st = r["stats"]
n_det = st.n_det
links = st.singles_a * st.nb + st.singles_b * st.na
# gathered FMAs of one build: opposite-spin pairs + same-spin entries
fma = (st.singles_a / st.na) * (st.singles_b / st.nb) * n_det + st.nnz_a * st.nb + st.nnz_b * st.na
sigma_ms = st.sigma_ms / max(st.sigma_builds, 1)
rest_ms = (st.davidson_ms - st.sigma_ms) / max(st.cycles, 1)
# basis sizes m = 1 .. cycles: gram m+1 vectors, residual 2m+3, ortho1 / ortho2 m+2 each
vec_bytes = 8.0 * n_det * sum((m + 1) + (2 * m + 3) + 2 * (m + 2) for m in range(1, st.cycles + 1)) / max(st.cycles, 1)
print(json.dumps({
    "workload": f"s8: (30e,30o) {n}x{n} = {n_det} determinants, one subspace",
    "sigma_path": {1: "v1", 2: "v2", 3: "wide"}[st.sigma_path],
    "nnz_a": st.nnz_a, "nnz_b": st.nnz_b, "singles_a": st.singles_a, "singles_b": st.singles_b,
    "sigma_ms_per_build": sigma_ms, "sigma_builds": st.sigma_builds, "gathered_fma_per_build": fma,
    "sigma_gfma_per_s": fma / (sigma_ms * 1e-3) / 1e9,
    "vector_kernels_ms_per_cycle": rest_ms, "vector_bytes_per_cycle_model": vec_bytes,
    "vector_kernels_gbs": vec_bytes / (rest_ms * 1e-3) / 1e9,
    "cycles": st.cycles, "davidson_ms": st.davidson_ms, "wall_s_incl_tables": wall,
    "theta_after_cycles": st.theta, "links_model": links,
    "gpu_mem_gb": (torch.cuda.mem_get_info()[1] - torch.cuda.mem_get_info()[0]) / 1e9,
}))

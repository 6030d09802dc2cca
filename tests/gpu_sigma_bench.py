"""Diagnostics: CUDA-event time of one sigma build per kernel generation (v1 / v2) on a bench workload.

    python tests/gpu_sigma_bench.py [workload] [reps] [paths]

Run under ``ncu --metrics gpu__time_duration.sum`` for the per-kernel split.  Not a bench value.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from qiskit_addon_sqd_b200 import fermion  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
paths = sys.argv[3].split(",") if len(sys.argv) > 3 else ["v1", "v2"]
norb, nelec, h, g, batches = bench.make_batches(wl, 0, 1)
sa, sb = batches[0]
out = {}
for path in paths:
    sub = fermion._Subspace(sa, sb, norb, h, g, sigma_path=path)
    ham = sub.hamiltonian()
    x = sub.upload_amplitudes(np.random.default_rng(0).standard_normal((sub.na, sub.nb)))
    y = sub.new_vector()
    for _ in range(5):
        sub.apply(ham, x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sub.apply(ham, x, y)
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    out[path] = y.clone()
    extra = ""
    if path == "v2" and ham.uses_v2:
        c = sub.v2_counts
        extra = f" items {c[0]} chunks {c[1]} groups {c[2]} vc_pad {c[3]} singles {c[4]}/{c[5]} nvc {c[7]}"
    print(f"{wl} {path}: {us:.1f} us per sigma build (uses_v2={ham.uses_v2}){extra}", flush=True)
if len(out) == 2:
    a, b = out["v1"], out["v2"]
    print("max |v1 - v2| =", float((a - b).abs().max()), "max |v1| =", float(a.abs().max()))

"""Diagnostics: do K concurrent sigma builds (one stream each) overlap on the device?"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from qiskit_addon_sqd_b200 import _lib, fermion  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = 50
norb, nelec, h, g, batches = bench.make_batches(wl, 0, K)
ints = fermion._DeviceIntegrals(torch, h, g, torch.device("cuda", 0))
subs, hams, xs, ys = [], [], [], []
for sa, sb in batches:
    sub = fermion._Subspace(sa, sb, norb, None, None, ints=ints)
    subs.append(sub)
    hams.append(sub.hamiltonian())
    xs.append(sub.upload_amplitudes(np.random.default_rng(0).standard_normal((sub.na, sub.nb))))
    ys.append(sub.new_vector())
print("chunks", [hm.struct.plan.n_chunks for hm in hams], "smem", _lib.load().sqd_sigma_smem_bytes(__import__("ctypes").byref(hams[0].struct)))
streams = [torch.cuda.Stream() for _ in range(K)]
main = torch.cuda.current_stream()


def run(kk):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for s in streams[:kk]:
        s.wait_stream(main)
    for _ in range(reps):
        for k in range(kk):
            with torch.cuda.stream(streams[k]):
                subs[k].apply(hams[k], xs[k], ys[k])
    t_host = time.perf_counter() - t0
    for s in streams[:kk]:
        main.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("K=%d: %.1f us per sigma build per stream-round (%.1f us per build), host enqueue %.1f us per build"
          % (kk, ms * 1e3 / reps, ms * 1e3 / reps / kk, t_host * 1e6 / reps / kk))


for kk in (1, 1, 2, 4, K):
    run(kk)

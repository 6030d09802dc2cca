"""Diagnostics: host-side timeline of one end-to-end bench step (8 x 1e5-determinant subspaces through
solve_sci_batch with RDMs): when does each worker's library call end, how long do its downloads take?"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from qiskit_addon_sqd_b200 import _lib, fermion  # noqa: E402

norb, nelec, h, g, batches = bench.make_batches("c4", 0, 8)
lib = _lib.load()
events = []
lock = threading.Lock()
orig_solve, orig_download = lib.sqd_solve_subspace, _lib.download


class SolveProxy:
    def __call__(self, *a):
        t0 = time.perf_counter()
        rc = orig_solve(*a)
        with lock:
            events.append(("solve", threading.get_ident(), t0, time.perf_counter(), 0))
        return rc


def download(torch_, t):
    t0 = time.perf_counter()
    out = orig_download(torch_, t)
    with lock:
        events.append(("download", threading.get_ident(), t0, time.perf_counter(), out.nbytes))
    return out


class LibProxy:
    def __getattr__(self, name):
        return SolveProxy() if name == "sqd_solve_subspace" else getattr(lib, name)


for rdm in (True, False):
    for _ in range(3):
        fermion.solve_sci_batch(batches, h, g, norb, nelec, compute_rdms=rdm)
    torch.cuda.synchronize()
    _lib.download = download
    _lib._lib = LibProxy()
    events.clear()
    t_start = time.perf_counter()
    fermion.solve_sci_batch(batches, h, g, norb, nelec, compute_rdms=rdm)
    t_end = time.perf_counter()
    _lib.download = orig_download
    _lib._lib = lib
    print(f"\ncompute_rdms={rdm}: step {1e3 * (t_end - t_start):.2f} ms")
    by_thread = {}
    for kind, tid, a, b, nbytes in events:
        by_thread.setdefault(tid, []).append((kind, 1e3 * (a - t_start), 1e3 * (b - t_start), nbytes))
    for tid, evs in sorted(by_thread.items(), key=lambda kv: kv[1][0][2]):
        evs.sort(key=lambda e: e[1])
        s = evs[0]
        dl = [e for e in evs if e[0] == "download"]
        print(f"  solve {s[1]:6.2f} -> {s[2]:6.2f} ms | downloads "
              + ", ".join(f"{e[3] / 1e6:.1f} MB {e[2] - e[1]:.2f} ms" for e in dl)
              + f" | done at {max(e[2] for e in evs):6.2f} ms")

"""Diagnostics: per-CTA phase timings of the alpha sigma kernel (sqd_sigma_profile)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from qiskit_addon_sqd_b200 import _lib, fermion  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
norb, nelec, h, g, batches = bench.make_batches(wl, 0, 1)
sub = fermion._Subspace(batches[0][0], batches[0][1], norb, h, g)
ham = sub.hamiltonian()
lib = _lib.load()
x = sub.upload_amplitudes(np.random.default_rng(0).standard_normal((sub.na, sub.nb)))
y = sub.new_vector()
nch = ham.struct.plan.n_chunks
prof = torch.zeros(8 * nch, dtype=torch.int64, device="cuda")
for _ in range(3):
    sub.apply(ham, x, y)
torch.cuda.synchronize()
_lib.check(lib.sqd_sigma_profile(C.byref(ham.struct), _lib.ptr(x), _lib.ptr(y), _lib.ptr(prof),
                                 _lib.stream_ptr(torch)))
torch.cuda.synchronize()
p = prof.cpu().numpy().reshape(nch, 8)
# slots: 0 = sum(pre-wait work) 1 = sum(wait for staged rows) 4 = sum(gather loop + arrive)  [warp 0, per CTA]
#        2 = stamp after phase C, 3 = stamp after phase D, 5 = #singles, 6 = #doubles, 7 = SM id
it = p[:, 5] > 0
print("chunks", nch, "with items", int(it.sum()))
print("per item, warp 0 (cycles): pre %.0f  wait %.0f  loop %.0f  | phase D total/items %.0f" % (
    (p[it, 0] / p[it, 5]).mean(), (p[it, 1] / p[it, 5]).mean(), (p[it, 4] / p[it, 5]).mean(),
    ((p[it, 3] - p[it, 2]) / p[it, 5]).mean()))
order = np.argsort(-(p[:, 3] - p[:, 2]))[:8]
for i in order:
    print("cta %4d sm %3d items %3d dbl %4d | pre %6d wait %7d loop %6d | D %7d" % (
        i, p[i, 7], p[i, 5], p[i, 6], p[i, 0], p[i, 1], p[i, 4], p[i, 3] - p[i, 2]))
keep = sub.tb._sell[0][1]
sptr = keep[2].cpu().numpy()
print("SELL-D slice lengths:", (np.diff(sptr) // 32).tolist(), "entries", int(sptr[-1]))
print("n_long", ham.struct.plan.n_long, "long cols", sub._plan_keep[3].cpu().numpy()[:ham.struct.plan.n_long],
      "n_single of them", sub.tb.n_single.cpu().numpy()[sub._plan_keep[3].cpu().numpy()[:ham.struct.plan.n_long]])
keepb = sub.tb._sell[1][1]
print("SELL-B slice lengths:", (np.diff(keepb[2].cpu().numpy()) // 32).tolist())

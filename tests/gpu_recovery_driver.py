"""Driver for ncu: one exact-stream recover_configurations call on the bench workload."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from qiskit_addon_sqd_b200 import configuration_recovery as cr
from qiskit_addon_sqd_b200._synthetic import noisy_samples

norb, nelec, n = 30, (15, 15), 100_000
ba = noisy_samples(norb, nelec, n, 316, 0.03, 106)
bits = np.unpackbits(ba.array, axis=1)[:, -2 * norb:].astype(bool)
probs = np.full(n, 1.0 / n)
occ = (bits[:, norb:][:, ::-1].mean(axis=0), bits[:, :norb][:, ::-1].mean(axis=0))
mode = sys.argv[1] if len(sys.argv) > 1 else "exact"
for _ in range(2):
    cr.recover_configurations(bits, probs, occ, nelec[0], nelec[1], rand_seed=np.random.default_rng(7), rng_mode=mode)
torch.cuda.synchronize()

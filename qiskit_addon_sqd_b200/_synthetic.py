"""Synthetic inputs for tests and benchmarks.

All draws come from ``numpy.random.default_rng(seed)`` (PCG64) so that every process -- the CUDA arm,
the CPU oracle arm, every rank -- regenerates identical inputs from the seed alone.
"""

from __future__ import annotations

import numpy as np


def random_integrals(norb: int, seed: int) -> tuple[np.ndarray, np.ndarray]:
    """``h = diag(linspace(-1,1)) + 0.1 sym(N(0,1))``; ``g = 0.5/norb * sym8(N(0,1))``, chemist order."""
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((norb, norb))
    h = np.diag(np.linspace(-1.0, 1.0, norb)) + 0.1 * 0.5 * (a + a.T)
    g = rng.standard_normal((norb,) * 4)
    g = g + g.transpose(1, 0, 2, 3)
    g = g + g.transpose(0, 1, 3, 2)
    g = g + g.transpose(2, 3, 0, 1)
    g *= 0.5 / norb / 8.0
    return np.ascontiguousarray(h), np.ascontiguousarray(g)


def hf_centred_strings(norb: int, nelec: int, n: int, seed: int) -> np.ndarray:
    """``n`` unique fixed-popcount strings clustered around the Hartree-Fock string, sorted ascending.

    Excitation rank ``k ~ min(Geometric(0.5), nelec, norb - nelec)``, k occupied -> k virtual chosen
    uniformly; the HF string ``(1 << nelec) - 1`` is always included.
    """
    rng = np.random.default_rng(seed)
    hf = (1 << nelec) - 1
    out = {hf}
    occ0 = np.arange(nelec)
    vir0 = np.arange(nelec, norb)
    kmax = min(nelec, norb - nelec)
    import math

    if n > math.comb(norb, nelec):
        raise ValueError("more strings requested than exist")
    while len(out) < n:
        k = min(int(rng.geometric(0.5)), kmax)
        o = rng.choice(occ0, k, replace=False)
        v = rng.choice(vir0, k, replace=False)
        s = hf
        for i in o:
            s ^= 1 << int(i)
        for a in v:
            s |= 1 << int(a)
        out.add(s)
    return np.array(sorted(out), dtype=np.int64)


def uniform_strings(norb: int, nelec: int, n: int, seed: int) -> np.ndarray:
    """``n`` unique uniformly random fixed-popcount strings, sorted ascending."""
    rng = np.random.default_rng(seed)
    out: set[int] = set()
    while len(out) < n:
        out.add(sum(1 << int(i) for i in rng.choice(norb, nelec, replace=False)))
    return np.array(sorted(out), dtype=np.int64)


def strings_to_bitstring_matrix(strs_a: np.ndarray, strs_b: np.ndarray, norb: int) -> np.ndarray:
    """Pair string i of each list into one row ``[b_{N-1}..b_0, a_{N-1}..a_0]`` (``fermion.py:234-237``)."""
    n = min(len(strs_a), len(strs_b))
    out = np.zeros((n, 2 * norb), dtype=bool)
    for k in range(norb):
        out[:, norb - 1 - k] = (np.asarray(strs_b[:n]) >> k) & 1
        out[:, 2 * norb - 1 - k] = (np.asarray(strs_a[:n]) >> k) & 1
    return out


def random_pauli_operator(nq: int, n_masks: int, z_per_mask: int, max_weight: int, seed: int):
    """Hermitian sum of real-weighted Pauli strings: ``n_masks`` X-masks x ``z_per_mask`` Z-masks.

    Returns ``(x_bits, z_bits, coeffs)`` with ``x_bits``/``z_bits`` bool ``(T, nq)`` in qiskit's
    little-endian convention (index k = qubit k) and real coefficients as complex128.
    """
    rng = np.random.default_rng(seed)
    xs, zs = [], []
    for _ in range(n_masks):
        w = int(rng.integers(0, max_weight + 1))
        x = np.zeros(nq, dtype=bool)
        if w:
            x[rng.choice(nq, w, replace=False)] = True
        for _ in range(z_per_mask):
            xs.append(x.copy())
            zs.append(rng.random(nq) < 0.25)
    coeffs = rng.standard_normal(len(xs)).astype(np.complex128)
    return np.array(xs), np.array(zs), coeffs


class PauliTerm:
    """Minimal duck-type of ``qiskit.quantum_info.Pauli`` (``qubit.py:214-216`` reads only x and z)."""

    def __init__(self, x: np.ndarray, z: np.ndarray):
        self.x = np.asarray(x, dtype=bool)
        self.z = np.asarray(z, dtype=bool)

    def to_label(self) -> str:
        tab = {(False, False): "I", (True, False): "X", (False, True): "Z", (True, True): "Y"}
        return "".join(tab[(bool(a), bool(b))] for a, b in zip(self.x[::-1], self.z[::-1]))


class PauliTable(list):
    """Duck-type of ``qiskit.quantum_info.PauliList``: a sequence of terms that also exposes the whole
    ``(T, nq)`` boolean ``x`` and ``z`` tables."""

    def __init__(self, x_bits, z_bits):
        self.x = np.atleast_2d(np.asarray(x_bits, dtype=bool))
        self.z = np.atleast_2d(np.asarray(z_bits, dtype=bool))
        super().__init__(PauliTerm(x, z) for x, z in zip(self.x, self.z))


class PauliSum:
    """Minimal duck-type of ``qiskit.quantum_info.SparsePauliOp``: ``.paulis``, ``.coeffs``, ``.size``."""

    def __init__(self, x_bits, z_bits, coeffs):
        self.paulis = PauliTable(x_bits, z_bits)
        self.coeffs = np.asarray(coeffs, dtype=np.complex128)

    @classmethod
    def from_labels(cls, labels, coeffs=None):
        if isinstance(labels, str):
            labels = [labels]
        xs = [[c in "XY" for c in lab[::-1]] for lab in labels]
        zs = [[c in "ZY" for c in lab[::-1]] for lab in labels]
        if coeffs is None:
            coeffs = np.ones(len(labels))
        return cls(np.array(xs, dtype=bool), np.array(zs, dtype=bool), coeffs)

    @property
    def size(self) -> int:
        return len(self.paulis)


class PackedBitArray:
    """Minimal stand-in for ``qiskit.primitives.BitArray`` (what the SQD loop reads: ``.array`` uint8
    ``(shots, bytes)`` big-endian left-padded, ``.num_bits``, ``.num_shots``)."""

    def __init__(self, array: np.ndarray, num_bits: int):
        self.array = np.ascontiguousarray(array, dtype=np.uint8)
        self.num_bits = int(num_bits)
        self.num_shots = int(self.array.shape[0])

    @classmethod
    def from_bool_array(cls, bits: np.ndarray) -> "PackedBitArray":
        bits = np.asarray(bits, dtype=bool)
        pad = (-bits.shape[1]) % 8
        return cls(np.packbits(np.pad(bits, ((0, 0), (pad, 0))), axis=1), bits.shape[1])


def noisy_samples(norb: int, nelec: tuple[int, int], shots: int, n_strings: int, noise: float,
                  seed: int) -> PackedBitArray:
    """Synthetic measurement record for the SQD loop: alpha/beta strings drawn from HF-centred pools with
    exponentially decaying weights, columns ``[b_{N-1}..b_0, a_{N-1}..a_0]``, every bit then flipped with
    probability ``noise`` (so that configuration recovery has something to repair)."""
    import math

    rng = np.random.default_rng(seed)
    pools = [hf_centred_strings(norb, ne, min(n_strings, math.comb(norb, ne)), seed + 1 + s)
             for s, ne in enumerate(nelec)]
    bits = np.zeros((shots, 2 * norb), dtype=bool)
    for spin, pool in enumerate(pools):
        w = np.exp(-4.0 * np.arange(len(pool)) / len(pool))
        pick = pool[rng.choice(len(pool), shots, p=w / w.sum())].astype(np.uint64)
        first = norb if spin == 0 else 0   # alpha on the right
        for k in range(norb):
            bits[:, first + norb - 1 - k] = (pick >> np.uint64(k)) & np.uint64(1)
    bits ^= rng.random(bits.shape) < noise
    return PackedBitArray.from_bool_array(bits)

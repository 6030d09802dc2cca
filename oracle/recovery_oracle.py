"""CPU oracle for configuration recovery.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restatement of ``qiskit_addon_sqd/configuration_recovery.py``:
  * ``recover_configurations``            :59-128
  * ``_p_flip_0_to_1`` / ``_p_flip_1_to_0`` :131-178
  * ``_bipartite_bitstring_correcting``   :181-306
and of the third-party routine that owns the randomness, ``numpy.random.Generator.choice(a, size,
replace=False, p=p)`` on a PCG64 bit generator (numpy >= 2.0, ``pyproject.toml:29``; algorithm
restated from numpy's published ``_generator.pyx``: draw ``size - n_uniq`` uniforms, zero ``p[found]``,
``cdf = cumsum(p); cdf /= cdf[-1]``, ``searchsorted(cdf, x, side='right')``, keep first occurrences).

PARITY STATUS: pinned.  ``tests/golden/make_golden.py`` runs the UNMODIFIED reference function with
seeded generators and stores inputs, outputs and the generator's final state in
``tests/golden/recovery_*.npz``; ``tests/test_oracle_cpu.py`` checks this file against them (rows,
probabilities and PCG64 state all bit-equal) and against ``numpy.random.Generator`` directly.
"""

from __future__ import annotations

import numpy as np

_MULT = 0x2360ED051FC65DA44385DF649FCCF645
_MASK128 = (1 << 128) - 1
_MASK64 = (1 << 64) - 1


class PCG64Stream:
    """numpy's PCG64 (XSL-RR 128/64) -> uniform doubles, restated."""

    def __init__(self, state: int, inc: int):
        self.state, self.inc = int(state), int(inc)

    @classmethod
    def from_generator(cls, rng: np.random.Generator) -> "PCG64Stream":
        st = rng.bit_generator.state
        if st["bit_generator"] != "PCG64":
            raise ValueError("exact-stream recovery needs a PCG64 generator (numpy's default)")
        return cls(st["state"]["state"], st["state"]["inc"])

    def to_generator(self, rng: np.random.Generator) -> None:
        st = rng.bit_generator.state
        st["state"]["state"] = self.state
        rng.bit_generator.state = st

    def next_u64(self) -> int:
        self.state = (self.state * _MULT + self.inc) & _MASK128
        hi, lo = self.state >> 64, self.state & _MASK64
        x = hi ^ lo
        rot = self.state >> 122
        return ((x >> rot) | (x << ((-rot) & 63))) & _MASK64

    def random(self, n: int) -> np.ndarray:
        return np.array([(self.next_u64() >> 11) * (1.0 / 9007199254740992.0) for _ in range(n)])


def numpy_pairwise_sum(a) -> float:
    """The summation order of ``np.sum`` on < 128 contiguous doubles (numpy ``pairwise_sum``)."""
    a = [float(v) for v in a]
    n = len(a)
    if n < 8:
        res = -0.0
        for v in a:
            res += v
        return res
    r = a[:8]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] += a[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res += a[i]
        i += 1
    return res


def choice_without_replacement(stream: PCG64Stream, p: np.ndarray, size: int) -> np.ndarray:
    """Indices chosen by ``Generator.choice(len(p), size, replace=False, p=p)``, in numpy's order."""
    p = np.array(p, dtype=np.float64)
    if np.count_nonzero(p > 0) < size:
        raise ValueError("Fewer non-zero entries in p than size")
    found = np.zeros(size, dtype=np.int64)
    n_uniq = 0
    while n_uniq < size:
        x = stream.random(size - n_uniq)
        if n_uniq > 0:
            p[found[:n_uniq]] = 0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        new = cdf.searchsorted(x, side="right")
        _, first = np.unique(new, return_index=True)
        first.sort()
        new = new.take(first)
        found[n_uniq:n_uniq + new.size] = new
        n_uniq += new.size
    return found


def flip_weight_0_to_1(ratio: float, occ: float, eps: float = 0.01) -> float:
    if occ < ratio:
        return occ * eps / ratio
    if ratio == 1.0:
        return eps
    slope = (1 - eps) / (1 - ratio)
    return occ * slope + (1 - slope)


def flip_weight_1_to_0(ratio: float, occ: float, eps: float = 0.01) -> float:
    return flip_weight_0_to_1(1 - ratio, 1 - occ, eps)


def correct_half(bits: np.ndarray, occ: np.ndarray, target: int, stream: PCG64Stream) -> np.ndarray:
    """One half of ``_bipartite_bitstring_correcting`` (:231-265 / :268-304); ``bits`` is modified."""
    m = len(bits)
    ratio = target / m
    p = np.array([
        flip_weight_1_to_0(ratio, occ[j]) if bits[j] else flip_weight_0_to_1(ratio, occ[j])
        for j in range(m)
    ])
    p = np.minimum(1, np.maximum(0, p))
    if not np.any(p):
        return bits
    p = p / np.sum(p)
    n_diff = int(np.sum(bits)) - target
    if n_diff == 0:
        return bits
    cand = np.where(bits)[0] if n_diff > 0 else np.where(np.logical_not(bits))[0]
    pc = p[cand] / np.sum(p[cand])
    chosen = choice_without_replacement(stream, pc, abs(n_diff))
    bits[cand[chosen]] = np.logical_not(bits[cand[chosen]])
    return bits


def recover_configurations(bitstring_matrix, probabilities, avg_occupancies, num_elec_a, num_elec_b,
                           rand_seed=None):
    """Restatement of the reference function; advances ``rand_seed`` when it is a Generator."""
    rng = np.random.default_rng(rand_seed)
    if np.array(avg_occupancies).ndim == 1:  # deprecated 1-D form (:99-107)
        norb = bitstring_matrix.shape[1] // 2
        avg_occupancies = (np.flip(avg_occupancies[norb:]), np.flip(avg_occupancies[:norb]))
    if num_elec_a < 0 or num_elec_b < 0:
        raise ValueError("The numbers of electrons must be specified as non-negative integers.")
    stream = PCG64Stream.from_generator(rng)
    occs = np.flip(avg_occupancies).flatten()  # column order: [occ_b[N-1..0], occ_a[N-1..0]]
    out: dict[bytes, float] = {}
    bitstring_matrix = np.asarray(bitstring_matrix, dtype=bool)
    half = bitstring_matrix.shape[1] // 2 if bitstring_matrix.ndim == 2 else 0
    try:
        for row, freq in zip(bitstring_matrix, probabilities):
            row = row.copy()
            correct_half(row[:half], occs[:half], num_elec_b, stream)   # LEFT = beta first (:231)
            correct_half(row[half:], occs[half:], num_elec_a, stream)   # RIGHT = alpha (:268)
            key = row.tobytes()
            out[key] = out.get(key, 0.0) + freq
    finally:
        stream.to_generator(rng)
    mat = np.array([np.frombuffer(k, dtype=bool) for k in out])
    freqs = np.array(list(out.values()))
    freqs = np.abs(freqs) / np.sum(np.abs(freqs))
    return mat, freqs

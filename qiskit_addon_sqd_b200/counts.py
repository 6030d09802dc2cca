"""Input-format conversions used by the SQD loop (reference ``qiskit_addon_sqd/counts.py:45-61,
186-201``): sampled ``BitArray`` -> unique bitstring matrix + probabilities, bitstring matrix -> integers.

These are host-side format conversions on either side of the CUDA hot path; they are restated with
byte-level numpy operations (``packbits`` / ``unique`` on packed keys) instead of the reference's
per-bit Python loops, with identical outputs (row order, dtypes, probabilities).
"""

from __future__ import annotations

import numpy as np


def bitstring_matrix_to_integers(bitstring_matrix: np.ndarray) -> np.ndarray:
    """Big-endian integer of every row (column 0 is the most significant bit).

    ``int64`` for fewer than 64 bits, Python integers in an ``object`` array from 64 bits on
    (reference ``counts.py:186-201``)."""
    bits = np.asarray(bitstring_matrix)
    n_rows, n_bits = bits.shape
    pad = (-n_bits) % 8
    packed = np.packbits(np.pad(bits.astype(bool), ((0, 0), (pad, 0))), axis=1)  # left-padded bytes
    if n_bits < 64:
        wide = np.zeros((n_rows, 8), dtype=np.uint8)
        wide[:, 8 - packed.shape[1]:] = packed
        return wide.view(">u8").reshape(n_rows).astype(int)
    out = np.empty(n_rows, dtype=object)
    for i in range(n_rows):
        out[i] = int.from_bytes(packed[i].tobytes(), "big")
    return out


def bit_array_to_arrays(bit_array) -> tuple[np.ndarray, np.ndarray]:
    """Unique sampled bitstrings (rows in ascending lexicographic order) and their frequencies
    (reference ``counts.py:45-61``).  ``bit_array`` needs ``.array`` (``uint8 (shots, bytes)``, big-endian,
    left-padded), ``.num_bits`` and ``.num_shots``."""
    packed = np.ascontiguousarray(bit_array.array, dtype=np.uint8)
    num_bits = int(bit_array.num_bits)
    packed = packed.reshape(-1, packed.shape[-1])
    n_bytes = (num_bits + 7) // 8
    packed = np.ascontiguousarray(packed[:, packed.shape[1] - n_bytes:])
    spare = 8 * n_bytes - num_bits
    if spare:
        packed[:, 0] &= np.uint8(0xFF >> spare)  # bits beyond num_bits are not part of the sample
    # unique on the packed rows: byte-wise lexicographic order == row order of the bool matrix
    keys = packed.view(np.dtype((np.void, n_bytes))).reshape(-1)
    uniq, counts = np.unique(keys, return_counts=True)
    rows = np.frombuffer(uniq.tobytes(), dtype=np.uint8).reshape(len(uniq), n_bytes)
    bitstrings = np.unpackbits(rows, axis=1)[:, spare:].astype(bool)
    return bitstrings, counts / bit_array.num_shots

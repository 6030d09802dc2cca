"""CPU restatement of the reference's sample-format conversion ``bit_array_to_arrays``
(``qiskit_addon_sqd/counts.py:45-61``) -- TEST INFRASTRUCTURE, never imported by the product package.

The reference unpacks every shot to one bool per bit and calls ``np.unique(axis=0, return_counts=True)``;
byte-wise lexicographic order of the packed rows is the same order, so ``np.unique`` runs on the packed rows
here.  Pinned by the known answers of ``tests/test_sqd_loop.py`` and, through the SQD-loop goldens, by the
unmodified reference loop (``tests/golden/make_golden.py``).
"""

from __future__ import annotations

import numpy as np


def bit_array_to_arrays(bit_array) -> tuple[np.ndarray, np.ndarray]:
    packed = np.ascontiguousarray(bit_array.array, dtype=np.uint8)
    num_bits = int(bit_array.num_bits)
    packed = packed.reshape(-1, packed.shape[-1])
    n_bytes = (num_bits + 7) // 8
    packed = np.ascontiguousarray(packed[:, packed.shape[1] - n_bytes:])
    spare = 8 * n_bytes - num_bits
    if spare:
        packed[:, 0] &= np.uint8(0xFF >> spare)  # counts.py:57 keeps the last num_bits bits only
    keys = packed.view(np.dtype((np.void, n_bytes))).reshape(-1)
    uniq, counts = np.unique(keys, return_counts=True)
    rows = np.frombuffer(uniq.tobytes(), dtype=np.uint8).reshape(len(uniq), n_bytes)
    bitstrings = np.unpackbits(rows, axis=1)[:, spare:].astype(bool)
    return bitstrings, counts / bit_array.num_shots


def bit_array_to_arrays_literal(bit_array) -> tuple[np.ndarray, np.ndarray]:
    """The reference's own three numpy calls (counts.py:57-60), for cross-checking the packed version."""
    bool_array = np.unpackbits(bit_array.array, axis=-1)[..., -bit_array.num_bits:].astype(bool)
    bitstrings, counts = np.unique(bool_array, axis=0, return_counts=True)
    return bitstrings, counts / bit_array.num_shots

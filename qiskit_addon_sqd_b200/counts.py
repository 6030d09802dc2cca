"""Input-format conversions used by the SQD loop (reference ``qiskit_addon_sqd/counts.py:45-61,
186-201``): sampled ``BitArray`` -> unique bitstring matrix + probabilities, bitstring matrix -> integers.

``bit_array_to_arrays`` runs on the device (key packing, bitonic sort, run-length counts: ``csrc/sortuniq.cu``);
``bitstring_matrix_to_integers`` is a host byte-level conversion of arrays the caller holds on the host
(from 64 bits on it must return Python integers).  Outputs are identical to the reference's (row order,
dtypes, probabilities).
"""

from __future__ import annotations

import numpy as np


def bitstring_matrix_to_integers(bitstring_matrix: np.ndarray) -> np.ndarray:
    """Big-endian integer of every row (column 0 is the most significant bit).

    ``int64`` for fewer than 64 bits, Python integers in an ``object`` array from 64 bits on
    (reference ``counts.py:186-201``)."""
    bits = np.asarray(bitstring_matrix)
    n_rows, n_bits = bits.shape
    if 0 < n_bits < 64:
        # packbits fills the LAST byte with zero bits: the row is left-aligned in a big-endian 64-bit word,
        # one shift brings it down (no padded copy of the matrix: the loop calls this twice per batch)
        packed = np.packbits(bits.astype(bool, copy=False), axis=1)
        wide = np.zeros((n_rows, 8), dtype=np.uint8)
        wide[:, : packed.shape[1]] = packed
        return (wide.view(">u8").reshape(n_rows) >> np.uint64(64 - n_bits)).astype(int)
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
    left-padded), ``.num_bits`` and ``.num_shots``.

    Runs on the device (``csrc/sortuniq.cu``): the packed shots are uploaded as they are (``num_bits/8`` bytes
    per shot instead of the reference's one byte per BIT), turned into 128-bit keys, sorted and run-length
    counted; the distinct rows come back as the bool matrix.  Registers of more than 128 bits are beyond the
    key width and raise ``ValueError``."""
    from . import _lib

    packed = np.ascontiguousarray(bit_array.array, dtype=np.uint8)
    num_bits = int(bit_array.num_bits)
    packed = packed.reshape(-1, packed.shape[-1])
    n_shots, row_bytes = packed.shape
    if num_bits > 128:
        raise ValueError("qiskit_addon_sqd_b200 handles bit arrays of at most 128 bits (64 per spin).")
    if n_shots == 0 or num_bits == 0:
        empty = np.zeros((0 if n_shots == 0 else 1, num_bits), dtype=bool)
        return empty, np.full(empty.shape[0], n_shots, dtype=np.int64) / bit_array.num_shots
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device())
    st = _lib.stream_ptr(torch)
    raw = torch.from_numpy(packed).to(dev)
    wide = num_bits > 64
    lo = torch.empty(n_shots, dtype=torch.int64, device=dev)
    hi = torch.empty(n_shots, dtype=torch.int64, device=dev) if wide else None
    _lib.check(lib.sqd_bit_array_pack(_lib.ptr(raw), n_shots, row_bytes, num_bits, _lib.ptr(hi), _lib.ptr(lo), st),
               "sqd_bit_array_pack")
    uhi, ulo, cnt = _lib.sort_unique(torch, hi, lo, with_counts=True)
    nu = int(ulo.numel())
    bits = torch.empty((nu, num_bits), dtype=torch.uint8, device=dev)
    _lib.check(lib.sqd_keys_to_bits(_lib.ptr(uhi), _lib.ptr(ulo), nu, num_bits, _lib.ptr(bits), st),
               "sqd_keys_to_bits")
    bitstrings = _lib.download(torch, bits).view(bool)
    counts = _lib.download(torch, cnt).astype(np.int64)
    return bitstrings, counts / bit_array.num_shots

"""Configuration recovery on B200.

Host-side mirror of ``qiskit_addon_sqd/configuration_recovery.py:59-128`` (same signature, return
layout, deprecation warning and error messages).  The per-bitstring correction
(``_bipartite_bitstring_correcting`` :181-306, ``_p_flip_*`` :131-178 and the four
``numpy.random.Generator.choice`` call sites) runs in the ``sqd_recover`` CUDA kernels.

Randomness: with the default ``rng_mode="exact"`` the kernels consume the caller's PCG64 stream
exactly as the reference would, so seeded results (rows, probabilities AND the generator's state after
the call -- the SQD loop shares one generator between recovery and subsampling, ``fermion.py:374,
508, 517``) are bit-identical to the reference.  ``rng_mode="parallel"`` draws one substream per row
instead: same distribution, every row independent.
"""

from __future__ import annotations

import ctypes as C
import warnings
from collections.abc import Sequence

import numpy as np

from . import _lib


def _pcg64_words(rng: np.random.Generator) -> np.ndarray | None:
    st = rng.bit_generator.state
    if st.get("bit_generator") != "PCG64":
        return None
    s, inc = int(st["state"]["state"]), int(st["state"]["inc"])
    m = (1 << 64) - 1
    return np.array([s >> 64, s & m, inc >> 64, inc & m], dtype=np.uint64)


def _set_pcg64_words(rng: np.random.Generator, words: np.ndarray) -> None:
    st = rng.bit_generator.state
    st["state"]["state"] = (int(words[0]) << 64) | int(words[1])
    rng.bit_generator.state = st


def _unpack_halves(left: np.ndarray, right: np.ndarray, norb: int) -> np.ndarray:
    """Packed halves -> bool rows ``[left_{N-1}..left_0, right_{N-1}..right_0]`` (big-endian bytes + unpackbits)."""
    def half(words):
        b = np.ascontiguousarray(words, dtype=np.uint64).astype(">u8").view(np.uint8).reshape(-1, 8)
        return np.unpackbits(b, axis=1)[:, 64 - norb:]

    return np.concatenate([half(left), half(right)], axis=1).astype(bool)


def recover_configurations(
    bitstring_matrix: np.ndarray,
    probabilities: Sequence[float] | np.ndarray,
    avg_occupancies: tuple[np.ndarray, np.ndarray],
    num_elec_a: int,
    num_elec_b: int,
    rand_seed: np.random.Generator | int | None = None,
    *,
    rng_mode: str = "exact",
) -> tuple[np.ndarray, np.ndarray]:
    """Refine bitstrings based on average orbital occupancy and a target hamming weight.

    Same contract as the reference (``configuration_recovery.py:59-128``): bit ``i`` of a row is the
    spin-down orbital matching the spin-up orbital in bit ``i + N``; returns the refined bitstring
    matrix (duplicates merged, first-seen order) and the renormalised probabilities.
    """
    rng = np.random.default_rng(rand_seed)

    occ_dims = len(np.array(avg_occupancies).shape)
    bitstring_matrix = np.asarray(bitstring_matrix)
    if occ_dims == 1:
        warnings.warn(
            "Passing avg_occupancies as a 1D array is deprecated. Pass a length-2 tuple containing the spin-up and spin-down occupancies respectively.",
            DeprecationWarning,
            stacklevel=2,
        )
        norb = bitstring_matrix.shape[1] // 2
        avg_occupancies = (np.flip(avg_occupancies[norb:]), np.flip(avg_occupancies[:norb]))

    if num_elec_a < 0 or num_elec_b < 0:
        raise ValueError("The numbers of electrons must be specified as non-negative integers.")
    if rng_mode not in ("exact", "parallel"):
        raise ValueError("rng_mode must be 'exact' or 'parallel'")

    n = bitstring_matrix.shape[0]
    if n == 0:
        empty = np.array([])
        return np.array([]), np.abs(empty) / np.sum(np.abs(empty)) if empty.size else empty

    torch = _lib.require_cuda()
    lib = _lib.load()
    nbits = bitstring_matrix.shape[1]
    norb = nbits // 2
    if norb > 64 or nbits % 2:
        raise ValueError("qiskit_addon_sqd_b200 needs an even number of bits with at most 64 per spin.")
    occs = np.ascontiguousarray(np.flip(avg_occupancies).flatten(), dtype=np.float64)
    dev = torch.device("cuda", torch.cuda.current_device())
    st = _lib.stream_ptr(torch)

    bits = torch.from_numpy(np.ascontiguousarray(bitstring_matrix, dtype=np.uint8)).to(dev)
    left = torch.empty(n, dtype=torch.int64, device=dev)
    right = torch.empty(n, dtype=torch.int64, device=dev)
    _lib.check(lib.sqd_pack_bitstrings(_lib.ptr(bits), n, nbits, _lib.ptr(left), _lib.ptr(right), st),
               "sqd_pack_bitstrings")
    occ_dev = torch.from_numpy(occs).to(dev)
    left_out = torch.empty_like(left)
    right_out = torch.empty_like(right)
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    ws_bytes = lib.sqd_recover_workspace_bytes(n, norb)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)

    words = _pcg64_words(rng) if rng_mode == "exact" else None
    if words is not None:
        state_dev = torch.from_numpy(words.view(np.int64)).to(dev)
        mode, seed = 0, 0
    else:
        # non-PCG64 generators cannot be continued on the device: seed per-row substreams from it
        state_dev = None
        mode, seed = 1, int(rng.integers(0, 2**63 - 1))
    _lib.check(
        lib.sqd_recover(_lib.ptr(left), _lib.ptr(right), n, norb, _lib.ptr(occ_dev),
                        occ_dev.data_ptr() + 8 * norb, int(num_elec_b), int(num_elec_a), mode,
                        _lib.ptr(state_dev), seed, _lib.ptr(left_out), _lib.ptr(right_out),
                        _lib.ptr(status), _lib.ptr(ws), ws_bytes, st),
        "sqd_recover")
    status_h = status.cpu().numpy()
    if words is not None:
        # also when numpy would have raised: the generator has consumed the draws of the rows before the
        # failing ``choice`` call (configuration_recovery.py:247-249 raises before drawing for that call)
        _set_pcg64_words(rng, state_dev.cpu().numpy().view(np.uint64))
    if status_h[0] != 0:
        raise ValueError("Fewer non-zero entries in p than size")
    # merge duplicates on the device: distinct rows in first-seen order, probabilities of equal rows added in
    # input order (:112-126); only the distinct rows and their sums come back
    prob_dev = torch.from_numpy(np.ascontiguousarray(probabilities, dtype=np.float64).reshape(-1)).to(dev)
    if prob_dev.numel() != n:
        raise ValueError("probabilities must hold one entry per bitstring")
    mws_bytes = lib.sqd_merge_rows_workspace_bytes(n)
    mws = torch.empty(mws_bytes, dtype=torch.uint8, device=dev)
    uniq_l = torch.empty(n, dtype=torch.int64, device=dev)
    uniq_r = torch.empty(n, dtype=torch.int64, device=dev)
    sums_dev = torch.empty(n, dtype=torch.float64, device=dev)
    n_unique = C.c_int32(0)
    _lib.check(lib.sqd_merge_rows(_lib.ptr(left_out), _lib.ptr(right_out), n, _lib.ptr(prob_dev), _lib.ptr(uniq_l),
                                  _lib.ptr(uniq_r), _lib.ptr(sums_dev), C.byref(n_unique), _lib.ptr(mws), mws_bytes,
                                  st), "sqd_merge_rows")
    nu = int(n_unique.value)
    lo = uniq_l[:nu].cpu().numpy().view(np.uint64)
    ro = uniq_r[:nu].cpu().numpy().view(np.uint64)
    sums = sums_dev[:nu].cpu().numpy()
    bs_mat_out = _unpack_halves(lo, ro, norb)
    freqs_out = np.abs(sums) / np.sum(np.abs(sums))
    return bs_mat_out, freqs_out

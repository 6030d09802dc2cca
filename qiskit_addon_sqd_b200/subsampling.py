"""Post-selection and batch subsampling of sampled bitstrings (reference
``qiskit_addon_sqd/subsampling.py:96-211``): the step of the SQD loop between configuration recovery and
the subspace solver.

Both functions stay on the host.  ``subsample`` has to consume the caller's ``numpy.random.Generator``
exactly as the reference does (the loop shares one generator between recovery and subsampling,
``fermion.py:374, 508, 517``), and numpy's weighted ``choice`` without replacement is a sequential
``cumsum`` -- there is nothing for the GPU to win at the 2-3 ms this step costs.
"""

from __future__ import annotations

import numpy as np


def postselect_by_hamming_right_and_left(
    bitstring_matrix: np.ndarray,
    probabilities: np.ndarray,
    *,
    hamming_right: int,
    hamming_left: int,
) -> tuple[np.ndarray, np.ndarray]:
    """Keep the rows whose right (alpha) and left (beta) halves have the requested Hamming weights and
    renormalise their probabilities (reference ``subsampling.py:96-144``)."""
    if min(hamming_left, hamming_right) < 0:
        raise ValueError("Hamming weight must be specified with a non-negative integer.")
    n_rows, width = bitstring_matrix.shape
    if width % 2:
        raise ValueError(f"The length of the bitstrings must be even. Instead, got {width}.")
    if len(probabilities) != n_rows:
        raise ValueError(
            "The number of elements in the probabilities array must match the number of rows in the bitstring matrix."
        )
    half = width // 2
    weights = bitstring_matrix.reshape(n_rows, 2, half).sum(axis=2)  # [:, 0] left, [:, 1] right
    keep = (weights[:, 0] == hamming_left) & (weights[:, 1] == hamming_right)
    kept_probs = probabilities[keep]
    return bitstring_matrix[keep], kept_probs / np.sum(kept_probs)


def subsample(
    bitstring_matrix: np.ndarray,
    probabilities: np.ndarray,
    samples_per_batch: int,
    num_batches: int,
    rand_seed: np.random.Generator | int | None = None,
) -> list[np.ndarray]:
    """``num_batches`` batches of ``samples_per_batch`` rows, each drawn without replacement with the
    given weights (reference ``subsampling.py:147-211``); the generator stream is the reference's."""
    n_rows = bitstring_matrix.shape[0]
    if n_rows < 1:
        return [np.array([])] * num_batches
    if len(probabilities) != n_rows:
        raise ValueError(
            "The number of elements in the probabilities array must match the number of rows in the bitstring matrix."
        )
    if samples_per_batch < 1:
        raise ValueError("Samples per batch must be specified with a positive integer.")
    if num_batches < 1:
        raise ValueError("The number of batches must be specified with a positive integer.")
    rng = np.random.default_rng(rand_seed)
    everything = np.arange(n_rows).astype("int")
    if samples_per_batch >= n_rows:  # nothing to draw: every batch is the whole input, no random numbers used
        return [bitstring_matrix[everything] for _ in range(num_batches)]
    return [
        bitstring_matrix[rng.choice(everything, samples_per_batch, replace=False, p=probabilities)]
        for _ in range(num_batches)
    ]

"""The SQD loop (`fermion.diagonalize_fermionic_hamiltonian`), `counts` and `subsampling` mirrors against
goldens produced by the reference's own, unmodified loop (`tests/golden/make_golden.py::loop_goldens`).

CPU tests drive OUR loop logic with oracle stand-ins for the two GPU steps (dense solver, recovery
restatement) and demand the reference's strings in every iteration; the GPU test runs the product path.
"""

import functools
import os
import sys

import numpy as np
import pytest

from oracle import counts_oracle as co
from oracle import fermion_oracle as fo
from oracle import recovery_oracle as ro
from qiskit_addon_sqd_b200._synthetic import random_integrals

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
from make_golden_cases import LOOP_CASES  # noqa: E402


class PackedBits:
    """Duck-typed stand-in for qiskit's BitArray (`.array`, `.num_bits`, `.num_shots`)."""

    def __init__(self, packed: np.ndarray, num_bits: int):
        self.array, self.num_bits, self.num_shots = packed, num_bits, packed.shape[0]


def _load_case(ci):
    z = np.load(os.path.join(GOLD, "sqd_loop_golden.npz"))
    norb, nea, neb, n_iter, n_batches = (int(v) for v in z[f"c{ci}_meta"])
    h, g = random_integrals(norb, 50 + ci)
    kw = dict(LOOP_CASES[ci]["kw"])
    if ci == 1:
        kw["initial_occupancies"] = (np.linspace(0.9, 0.1, norb), np.linspace(0.8, 0.05, norb))
    return z, norb, (nea, neb), n_iter, n_batches, h, g, kw, PackedBits(z[f"c{ci}_packed"], 2 * norb)


def _check_history(z, ci, history, n_iter, n_batches, etol, otol):
    assert len(history) == n_iter
    for it, results in enumerate(history):
        assert len(results) == n_batches
        for k, r in enumerate(results):
            # string lists: bit-exact, every iteration (recovery + subsampling + ordering + carry-over)
            assert np.array_equal(np.asarray(r.sci_state.ci_strs_a), z[f"c{ci}_i{it}_b{k}_a"]), (it, k)
            assert np.array_equal(np.asarray(r.sci_state.ci_strs_b), z[f"c{ci}_i{it}_b{k}_b"]), (it, k)
            assert abs(r.energy - float(z[f"c{ci}_i{it}_b{k}_e"])) < etol, (it, k)
            assert np.abs(np.array(r.orbital_occupancies) - z[f"c{ci}_i{it}_b{k}_occ"]).max() < otol


def _oracle_solver(ci_strings, h, g, norb, nelec, *, spin_sq=None):
    from qiskit_addon_sqd_b200.fermion import SCIResult, SCIState

    out = []
    for sa, sb in ci_strings:
        if spin_sq is None:
            e, c, occ, _s2, _ = fo.solve_dense(sa, sb, h, g, norb)
        else:
            e, c, occ, _s2, _ = fo.solve_dense(sa, sb, h, g, norb, spin_sq=spin_sq, shift=0.2)
        out.append(SCIResult(e, SCIState(c, np.asarray(sa), np.asarray(sb), norb, tuple(nelec)),
                             orbital_occupancies=occ))
    return out


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_loop_logic_matches_reference_loop_on_cpu(ci, monkeypatch):
    from qiskit_addon_sqd_b200 import configuration_recovery, counts, fermion

    z, norb, nelec, n_iter, n_batches, h, g, kw, bits = _load_case(ci)
    monkeypatch.setattr(configuration_recovery, "recover_configurations", ro.recover_configurations)
    monkeypatch.setattr(counts, "bit_array_to_arrays", co.bit_array_to_arrays)  # device code in the product
    history = []
    solver = functools.partial(_oracle_solver, spin_sq=LOOP_CASES[ci]["spin_sq"])
    best = fermion.diagonalize_fermionic_hamiltonian(h, g, bits, norb=norb, nelec=nelec, sci_solver=solver,
                                                     callback=history.append, **kw)
    _check_history(z, ci, history, n_iter, n_batches, 1e-10, 1e-9)
    assert abs(best.energy - float(z[f"c{ci}_best_energy"])) < 1e-10


def test_loop_argument_errors_are_the_reference_messages(monkeypatch):
    from qiskit_addon_sqd_b200 import counts, fermion

    monkeypatch.setattr(counts, "bit_array_to_arrays", co.bit_array_to_arrays)  # device code in the product
    h, g = random_integrals(4, 1)
    bits = PackedBits(np.zeros((3, 1), dtype=np.uint8), 8)
    with pytest.raises(ValueError, match="Maximum number of iterations must be at least 1."):
        fermion.diagonalize_fermionic_hamiltonian(h, g, bits, 2, 4, (2, 2), max_iterations=0)
    with pytest.raises(ValueError, match="Spin symmetrization is only possible"):
        fermion.diagonalize_fermionic_hamiltonian(h, g, bits, 2, 4, (2, 1), symmetrize_spin=True)
    with pytest.raises(ValueError, match="the maximum dimension must be"):
        fermion.diagonalize_fermionic_hamiltonian(h, g, bits, 2, 4, (2, 2), symmetrize_spin=True, max_dim=(3, 4))
    with pytest.raises(ValueError, match="did not contain any valid bitstrings"):
        fermion.diagonalize_fermionic_hamiltonian(h, g, bits, 2, 4, (2, 2), sci_solver=_oracle_solver)


def test_counts_and_subsampling_known_answers():
    """Known answers of the reference's own tests (test_subsampling.py:53-75, test_counts.py) and docs
    (select_open_closed_shell.ipynb: counts -> integers)."""
    from qiskit_addon_sqd_b200 import counts, subsampling

    mat = np.array([[1, 0, 1, 0], [0, 1, 1, 0], [1, 1, 0, 0], [1, 0, 0, 1]], dtype=bool)
    p = np.array([0.1, 0.2, 0.3, 0.4])
    rows, probs = subsampling.postselect_by_hamming_right_and_left(mat, p, hamming_right=1, hamming_left=1)
    assert np.array_equal(rows, mat[[0, 1, 3]]) and np.allclose(probs, np.array([0.1, 0.2, 0.4]) / 0.7)
    assert p[0] == 0.1  # input untouched
    with pytest.raises(ValueError, match="non-negative integer"):
        subsampling.postselect_by_hamming_right_and_left(mat, p, hamming_right=-1, hamming_left=1)
    with pytest.raises(ValueError, match="must be even"):
        subsampling.postselect_by_hamming_right_and_left(mat[:, :3], p, hamming_right=1, hamming_left=1)
    assert all(b.size == 0 for b in subsampling.subsample(np.zeros((0, 4), dtype=bool), np.zeros(0), 2, 3))
    with pytest.raises(ValueError, match="Samples per batch"):
        subsampling.subsample(mat, p, 0, 1)
    with pytest.raises(ValueError, match="number of batches"):
        subsampling.subsample(mat, p, 1, 0)
    # fewer rows than requested: whole input, generator untouched
    gen = np.random.default_rng(3)
    full = subsampling.subsample(mat, p, 9, 2, gen)
    assert len(full) == 2 and np.array_equal(full[0], mat) and gen.random() == np.random.default_rng(3).random()
    # the generator stream is numpy's weighted choice
    g1, g2 = np.random.default_rng(11), np.random.default_rng(11)
    got = subsampling.subsample(mat, p, 2, 3, g1)
    want = [mat[g2.choice(np.arange(4), 2, replace=False, p=p)] for _ in range(3)]
    assert all(np.array_equal(a, b) for a, b in zip(got, want))
    # integers: big-endian, int64 below 64 bits, Python ints from 64 bits on (test_fermion.py:344-360)
    assert counts.bitstring_matrix_to_integers(np.array([[0, 0, 0, 1, 0, 0, 1, 0]], dtype=bool)).tolist() == [18]
    wide = np.zeros((1, 64), dtype=bool)
    wide[0, 0] = wide[0, -1] = True
    out = counts.bitstring_matrix_to_integers(wide)
    assert out.dtype == object and out[0] == (1 << 63) + 1
    arr = PackedBits(np.array([[0b00010010], [0b01001000], [0b00010010], [0b00010001]], dtype=np.uint8), 8)
    for fn in (co.bit_array_to_arrays, co.bit_array_to_arrays_literal):  # the oracle; the product runs on the GPU
        rows, probs = fn(arr)
        assert counts.bitstring_matrix_to_integers(rows).tolist() == [17, 18, 72]
        assert probs.tolist() == [0.25, 0.5, 0.25]
    rng = np.random.default_rng(0)
    odd = PackedBits(rng.integers(0, 256, (500, 3), dtype=np.uint8) & np.array([0x07, 0xFF, 0x81], dtype=np.uint8), 19)
    a, b = co.bit_array_to_arrays(odd), co.bit_array_to_arrays_literal(odd)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.gpu
@pytest.mark.parametrize("num_bits,row_bytes,shots", [(8, 1, 4), (19, 3, 500), (60, 8, 20_000), (64, 8, 5_000),
                                                      (65, 9, 3_000), (100, 13, 70_000), (128, 16, 2_500)])
def test_bit_array_to_arrays_on_gpu_matches_numpy(num_bits, row_bytes, shots):
    """Product path of ``counts.bit_array_to_arrays`` (device key packing + bitonic sort + run lengths) against the
    reference's own numpy calls (``counts.py:57-60``): rows, row order, dtypes and probabilities bit for bit."""
    from qiskit_addon_sqd_b200 import counts

    rng = np.random.default_rng(num_bits)
    pool = rng.integers(0, 256, (max(3, shots // 7), row_bytes), dtype=np.uint8)   # many repeated shots
    pool[0] = 0xFF                                                                   # all ones: equals the sort pad
    pool[1] = 0
    packed = pool[rng.integers(0, len(pool), shots)]
    packed[:, -1] ^= (rng.random(shots) < 0.2).astype(np.uint8)                     # and many singletons
    arr = PackedBits(packed, num_bits)
    rows, probs = counts.bit_array_to_arrays(arr)
    want_rows, want_probs = co.bit_array_to_arrays_literal(arr)
    assert rows.dtype == np.bool_ and rows.shape == want_rows.shape and np.array_equal(rows, want_rows)
    assert probs.dtype == want_probs.dtype and np.array_equal(probs, want_probs)
    assert probs.sum() == pytest.approx(1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("ci", [0, 1, 2])
def test_loop_on_gpu_matches_reference_loop(ci):
    """Product path: GPU configuration recovery + GPU subspace solves inside our loop reproduce the
    reference loop's string lists exactly and its energies to 1e-8 Ha in every iteration."""
    from qiskit_addon_sqd_b200 import fermion

    z, norb, nelec, n_iter, n_batches, h, g, kw, bits = _load_case(ci)
    history = []
    solver = functools.partial(fermion.solve_sci_batch, spin_sq=LOOP_CASES[ci]["spin_sq"])
    best = fermion.diagonalize_fermionic_hamiltonian(h, g, bits, norb=norb, nelec=nelec, sci_solver=solver,
                                                     callback=history.append, **kw)
    _check_history(z, ci, history, n_iter, n_batches, 1e-8, 1e-6)
    assert abs(best.energy - float(z[f"c{ci}_best_energy"])) < 1e-8
    assert np.abs(np.array(best.orbital_occupancies) - z[f"c{ci}_best_occ"]).max() < 1e-6


def test_subsample_edge_cases_of_the_reference_suite():
    """reference test/test_subsampling.py: 1-D empty input, mismatching probabilities, batch shapes."""
    from qiskit_addon_sqd_b200 import subsampling

    out = subsampling.subsample(np.array([]), np.array([]), 1, 1)
    assert len(out) == 1 and out[0].shape[0] == 0
    mat = np.array([[0, 1, 0, 1], [1, 0, 1, 0], [1, 1, 0, 0], [0, 0, 1, 1], [0, 1, 1, 0]], dtype=bool)
    uniform = np.full(5, 0.2)
    with pytest.raises(ValueError, match="number of elements in the probabilities array must match"):
        subsampling.subsample(mat, np.array([]), 1, 1)
    batches = subsampling.subsample(mat, uniform, 2, 10, rand_seed=4)
    assert len(batches) == 10 and all(b.shape == (2, 4) for b in batches)
    assert all(len({tuple(r) for r in b}) == 2 for b in batches)          # without replacement
    whole = subsampling.subsample(mat, uniform, 20, 1)
    assert len(whole) == 1 and np.array_equal(whole[0], mat)
    # an int seed and a Generator built from it give the same batches (np.random.default_rng semantics)
    a = subsampling.subsample(mat, uniform, 3, 4, rand_seed=9)
    b = subsampling.subsample(mat, uniform, 3, 4, rand_seed=np.random.default_rng(9))
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_digest_convergence_and_carryover_rules():
    """Second half of an iteration (reference fermion.py:563-640) on hand-made results."""
    from qiskit_addon_sqd_b200.fermion import SCIResult, SCIState, _SQDRun

    def result(e, amps, sa, sb, occ):
        return SCIResult(e, SCIState(np.asarray(amps, float), np.asarray(sa), np.asarray(sb), 3, (1, 1)),
                         orbital_occupancies=occ)

    run = _SQDRun(None, None, 3, (1, 1), 2, 1, False, (np.array([], int), np.array([], int)), (None, None),
                  1e-6, 1e-4, 0.3, np.random.default_rng(0), None)
    occ = (np.array([1.0, 0.0, 0.0]), np.array([0.0, 1.0, 0.0]))
    amps = [[0.9, 0.05], [0.31, -0.29]]
    first = result(-1.0, amps, [1, 4], [2, 4], occ)
    assert run.digest([first, result(-0.5, amps, [1, 2], [1, 2], occ)]) is False     # nothing to compare with yet
    assert run.best is first and run.reference is first
    # |amplitude| > 0.3: (0,0) and (1,0) -> alpha strings {1, 4}, beta string {2}; alpha ranked by row weight
    assert run.carry_a.tolist() == [1, 4] and run.carry_b.tolist() == [2]
    # same energy and occupancies within tolerance -> converged, best result kept
    again = result(-1.0 + 5e-7, amps, [1, 4], [2, 4], (occ[0] + 5e-5, occ[1]))
    assert run.digest([again]) is True and run.best is first
    # a lower energy in a later iteration replaces the best result even if not converged
    lower = result(-1.2, amps, [1, 4], [2, 4], (occ[0] * 0.5, occ[1]))
    assert run.digest([lower]) is False and run.best is lower


@pytest.mark.gpu
@pytest.mark.parametrize("na,nb,thr", [(5, 7, 0.5), (40, 101, 1.5), (316, 316, 2.5), (130, 1000, 3.0),
                                        (1100, 9, 2.0), (300, 300, 10.0)])
def test_carryover_on_gpu_matches_numpy(na, nb, thr):
    """``sqd_carryover`` against the reference's numpy expressions (``fermion.py:607-622``): selected rows / columns
    and marginal weights BIT-equal (numpy's pairwise summation order), so the ranking that follows is the same."""
    import torch

    from qiskit_addon_sqd_b200 import fermion

    rng = np.random.default_rng(na * 1000 + nb)
    amps = rng.standard_normal((na, nb))
    amps[rng.random((na, nb)) < 0.3] *= 1e-3
    ldc = (nb + 1) // 2 * 2
    x = torch.zeros((na, ldc), dtype=torch.float64, device="cuda")
    x[:, :nb] = torch.from_numpy(amps).cuda()
    if ldc > nb:
        x[:, nb:] = 99.0      # pad columns must not be looked at
    rows, cols, wa, wb = fermion._carryover_on_device(x, nb, torch.cuda.current_device(), thr)
    flat = np.abs(amps.reshape(-1))
    order = np.argsort(flat)
    big = order[np.searchsorted(flat, thr, sorter=order):]
    r, c = np.divmod(big, nb)
    r, c = np.unique(r), np.unique(c)
    assert np.array_equal(rows, r) and np.array_equal(cols, c)
    assert np.array_equal(wa, np.sum(np.abs(amps[r]) ** 2, axis=1))
    assert np.array_equal(wb, np.sum(np.abs(amps[:, c]) ** 2, axis=0))

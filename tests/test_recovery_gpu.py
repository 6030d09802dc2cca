"""Parity of the CUDA configuration recovery against the reference's goldens and known answers."""

import os

import numpy as np
import pytest

from oracle import recovery_oracle as ro

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "recovery_golden.npz")


def test_reference_known_answers(cuda_lib):
    """Ports of test/test_configuration_recovery.py:57-135 (deprecated 1-D occupancies included)."""
    from qiskit_addon_sqd_b200.configuration_recovery import recover_configurations

    with pytest.warns(DeprecationWarning):
        m, p = recover_configurations(np.empty((0, 6)), np.empty((0,)), [False] * 6, num_elec_a=0,
                                      num_elec_b=1)
    assert m.size == 0 and p.size == 0
    with pytest.warns(DeprecationWarning):
        m, p = recover_configurations(np.array([[False] * 4]), np.array([1.0]), [1.0] * 4, 2, 2,
                                      rand_seed=4224)
    assert (np.array([[True] * 4]) == m).all() and (np.array([1.0]) == p).all()
    with pytest.warns(DeprecationWarning):
        m, p = recover_configurations(np.array([[True] * 4]), np.array([1.0]), [0.0] * 4, 0, 0,
                                      rand_seed=4224)
    assert (np.array([[False] * 4]) == m).all() and (np.array([1.0]) == p).all()
    with pytest.warns(DeprecationWarning):
        m, p = recover_configurations(np.array([[True] * 4]), np.array([1.0]), [0.0, 1.0, 0.0, 0.0], 0,
                                      1, rand_seed=4224)
    assert (np.array([[False, True, False, False]]) == m).all()
    bs = np.random.default_rng(554).integers(2, size=(1, 74), dtype=bool)
    with pytest.warns(DeprecationWarning):
        m, p = recover_configurations(bs, np.array([1.0]), np.zeros(74), 0, 0, rand_seed=4224)
    assert (np.zeros((1, 74), dtype=bool) == m).all() and (np.array([1.0]) == p).all()
    with pytest.raises(ValueError) as e_info:
        with pytest.warns(DeprecationWarning):
            recover_configurations(np.array([[True] * 4]), np.array([1.0]), [0.0] * 4, 0, -1,
                                   rand_seed=4224)
    assert e_info.value.args[0] == "The numbers of electrons must be specified as non-negative integers."


@pytest.mark.parametrize("ci", range(6))
def test_exact_stream_matches_reference_golden(cuda_lib, ci):
    """Seeded run: rows, probabilities and the generator's state are bit-identical to the reference."""
    from qiskit_addon_sqd_b200.configuration_recovery import recover_configurations

    g = np.load(GOLD)
    norb, na, nb, n, seed = (int(v) for v in g[f"c{ci}_meta"])
    gen = np.random.default_rng(seed)
    bs = g[f"c{ci}_bs"]
    bs_copy = bs.copy()
    mat, freqs = recover_configurations(bs, g[f"c{ci}_probs"], (g[f"c{ci}_occ_a"], g[f"c{ci}_occ_b"]),
                                        na, nb, gen)
    assert np.array_equal(bs, bs_copy)  # inputs are never mutated
    assert mat.dtype == bool and np.array_equal(mat, g[f"c{ci}_mat"])
    assert np.array_equal(freqs, g[f"c{ci}_freqs"])
    assert np.array_equal(gen.random(4), g[f"c{ci}_next"])


@pytest.mark.parametrize("norb,nelec,n", [(12, (4, 6), 5000), (30, (15, 15), 4000), (40, (12, 12), 3000)])
def test_exact_stream_matches_oracle_at_larger_sizes(cuda_lib, norb, nelec, n):
    from qiskit_addon_sqd_b200.configuration_recovery import recover_configurations

    rng = np.random.default_rng(norb)
    bs = rng.integers(2, size=(n, 2 * norb), dtype=np.int64).astype(bool)
    probs = rng.random(n)
    occ = (rng.random(norb), rng.random(norb))
    g1, g2 = np.random.default_rng(99), np.random.default_rng(99)
    mat, freqs = recover_configurations(bs, probs, occ, nelec[0], nelec[1], g1)
    mat_ref, freqs_ref = ro.recover_configurations(bs, probs, occ, nelec[0], nelec[1], g2)
    assert np.array_equal(mat, mat_ref) and np.array_equal(freqs, freqs_ref)
    assert g1.bit_generator.state == g2.bit_generator.state
    assert (mat[:, :norb].sum(1) == nelec[1]).all() and (mat[:, norb:].sum(1) == nelec[0]).all()
    assert abs(freqs.sum() - 1) < 1e-12


def test_parallel_mode_properties(cuda_lib):
    """Per-row substreams: not stream-identical, but every invariant of the algorithm holds."""
    from qiskit_addon_sqd_b200.configuration_recovery import recover_configurations

    norb, na, nb, n = 20, 7, 9, 20000
    rng = np.random.default_rng(1)
    bs = rng.integers(2, size=(n, 2 * norb), dtype=np.int64).astype(bool)
    probs = np.full(n, 1.0 / n)
    occ = (rng.random(norb), rng.random(norb))
    mat, freqs = recover_configurations(bs, probs, occ, na, nb, 7, rng_mode="parallel")
    assert (mat[:, :norb].sum(1) == nb).all() and (mat[:, norb:].sum(1) == na).all()
    assert abs(freqs.sum() - 1) < 1e-12 and len(np.unique(mat, axis=0)) == len(mat)
    mat2, freqs2 = recover_configurations(bs, probs, occ, na, nb, 7, rng_mode="parallel")
    assert np.array_equal(mat, mat2) and np.array_equal(freqs, freqs2)  # same seed -> same result
    # rows that already have the right weights are untouched (idempotence)
    mat3, _ = recover_configurations(mat, freqs, occ, na, nb, 8, rng_mode="parallel")
    assert np.array_equal(mat3, mat)
    # flip statistics: an occupied orbital with occupancy ~0 is far more likely to be emptied than one ~1
    occ2 = (np.r_[np.zeros(norb // 2), np.ones(norb - norb // 2)], occ[1])
    ones = np.ones((4000, 2 * norb), dtype=bool)
    m4, f4 = recover_configurations(ones, np.full(4000, 1 / 4000), occ2, norb - norb // 2, norb, 3,
                                    rng_mode="parallel")
    # alpha column j <-> orbital norb-1-j; orbitals with occ 0 must be the ones emptied
    kept = (m4[:, norb:] * f4[:, None]).sum(0)[::-1]
    assert kept[: norb // 2].sum() < 0.05 * kept[norb // 2:].sum()


def test_too_few_candidates_raises_like_numpy(cuda_lib):
    from qiskit_addon_sqd_b200.configuration_recovery import recover_configurations

    # left half: 3 electrons too many but only one occupied bit has non-zero flip weight
    bs = np.array([[True, True, True, True, False, False, False, False]])
    occ_b = np.array([1.0, 1.0, 1.0, 0.0])   # beta orbitals 0..3 ; column j <-> orbital 3-j
    occ_a = np.zeros(4)
    with pytest.raises(ValueError, match="Fewer non-zero entries in p than size"):
        ro.recover_configurations(bs, np.array([1.0]), (occ_a, occ_b), 0, 1, 1)
    with pytest.raises(ValueError, match="Fewer non-zero entries in p than size"):
        recover_configurations(bs, np.array([1.0]), (occ_a, occ_b), 0, 1, 1)


@pytest.mark.parametrize("chain", ["0", "1"])
def test_error_in_the_middle_leaves_the_generator_where_numpy_does(cuda_lib, chain, monkeypatch):
    """Thousands of rows consume draws, then a row makes numpy raise: same exception, and the caller's
    generator has advanced exactly as far as numpy's (speculative walk and the one-warp chain)."""
    import subprocess
    import sys

    code = r'''
import numpy as np
from oracle import recovery_oracle as ro
from qiskit_addon_sqd_b200.configuration_recovery import recover_configurations
norb, n = 4, 6000
rng = np.random.default_rng(3)
bs = rng.integers(2, size=(n, 2 * norb), dtype=np.int64).astype(bool)
occ_b = np.array([1.0, 1.0, 1.0, 0.0]); occ_a = np.array([0.3, 0.6, 0.2, 0.9])
bad = 4321
bs[bad] = [True, True, True, True, False, True, False, False]   # beta half: 3 too many, one candidate
probs = np.full(n, 1.0 / n)
out = []
for fn in (ro.recover_configurations, recover_configurations):
    g = np.random.default_rng(11)
    try:
        fn(bs, probs, (occ_a, occ_b), 1, 1, g)
        out.append(("no error", None))
    except ValueError as e:
        out.append((str(e), g.bit_generator.state["state"]["state"]))
assert out[0][0] == out[1][0] == "Fewer non-zero entries in p than size", out
assert out[0][1] == out[1][1], out
print("ok")
'''
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SQD_RECOVER_CHAIN=chain, PYTHONPATH=root)
    p = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "ok" in p.stdout, p.stdout + p.stderr


def test_exact_stream_speculative_walk_equals_chain(cuda_lib):
    """1e5 half-rows with heavy collisions (few candidates, many draws): the windowed speculative walk, the
    round-1 one-warp chain and the oracle on a prefix agree bit for bit, generator state included."""
    import os
    import subprocess
    import sys

    code = r'''
import numpy as np, sys
from qiskit_addon_sqd_b200.configuration_recovery import recover_configurations
norb, n = 10, 50000
rng = np.random.default_rng(5)
bs = rng.integers(2, size=(n, 2 * norb), dtype=np.int64).astype(bool)
occ = (rng.random(norb) ** 3, rng.random(norb) ** 3)      # skewed weights: many collisions
g = np.random.default_rng(2024)
mat, freqs = recover_configurations(bs, np.full(n, 1.0 / n), occ, 3, 7, g)
np.save(sys.argv[1], mat); np.save(sys.argv[2], freqs)
print(g.bit_generator.state["state"]["state"])
'''
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    import tempfile

    with tempfile.TemporaryDirectory() as td:
        for chain in ("0", "1"):
            env = dict(os.environ, SQD_RECOVER_CHAIN=chain, PYTHONPATH=root)
            a, b = os.path.join(td, f"m{chain}.npy"), os.path.join(td, f"f{chain}.npy")
            p = subprocess.run([sys.executable, "-c", code, a, b], cwd=root, env=env, capture_output=True,
                               text=True, timeout=600)
            assert p.returncode == 0, p.stderr
            outs.append((np.load(a), np.load(b), p.stdout.strip()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    assert outs[0][2] == outs[1][2]

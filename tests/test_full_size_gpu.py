"""BASELINE.json's configurations at FULL size, checked through size-independent properties
(the dense oracle cannot go there) and, where it finishes in seconds, against the C oracle."""

import time

import numpy as np
import pytest

import bench
from qiskit_addon_sqd_b200._synthetic import PauliSum, random_pauli_operator

pytestmark = pytest.mark.gpu


def _properties(sub, ham, x, e, occ, nelec, tol=1e-6):
    """Eigen-residual, Rayleigh quotient, Hermiticity of the operator, particle number."""
    rng = np.random.default_rng(0)
    hx = sub.apply(ham, x)
    theta = sub.dot(x, hx) / sub.dot(x, x)
    assert abs(theta - e) < 1e-9
    r = hx - theta * x
    assert float(sub.dot(r, r)) ** 0.5 < tol * 10
    y = sub.upload_amplitudes(rng.standard_normal((sub.na, sub.nb)))
    z = sub.upload_amplitudes(rng.standard_normal((sub.na, sub.nb)))
    hy, hz = sub.apply(ham, y), sub.apply(ham, z)
    a, b = sub.dot(z, hy), sub.dot(hz, y)
    assert abs(a - b) < 1e-9 * max(1.0, abs(a))                      # <z|Hy> == <Hz|y>
    lin = sub.apply(ham, 2.0 * y - 3.0 * z)
    assert float((lin - (2.0 * hy - 3.0 * hz)).abs().max()) < 1e-9   # linearity
    assert abs(occ[0].sum() - nelec[0]) < 1e-9 and abs(occ[1].sum() - nelec[1]) < 1e-9
    assert np.all(occ[0] > -1e-12) and np.all(occ[0] < 1 + 1e-12)


@pytest.mark.parametrize("wl,check_oracle", [("c1", True), ("c2", True), ("t", True), ("c4", True), ("c5", True)])
def test_fermion_configs_full_size(cuda_lib, wl, check_oracle):
    import torch

    from qiskit_addon_sqd_b200 import fermion

    norb, nelec, h, g, batches = bench.make_batches(wl, 0, 1)
    sa, sb = batches[0]
    t0 = time.perf_counter()
    res = fermion.solve_sci_batch([(sa, sb)], h, g, norb, nelec)[0]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = fermion.last_solve_stats()[0]
    print(f"\n[{wl}] na x nb = {len(sa)} x {len(sb)}, {st.cycles} cycles, {dt * 1e3:.1f} ms end to end, "
          f"E = {res.energy:.10f}, residual {st.residual:.1e}")
    assert st.converged == 1
    sub = fermion._Subspace(sa, sb, norb, h, g)
    ham = sub.hamiltonian()
    x = sub.upload_amplitudes(res.sci_state.amplitudes)
    _properties(sub, ham, x, res.energy, res.orbital_occupancies, nelec)
    # same call twice -> bit-identical
    res2 = fermion.solve_sci_batch([(sa, sb)], h, g, norb, nelec)[0]
    assert res2.energy == res.energy and np.array_equal(res2.sci_state.amplitudes, res.sci_state.amplitudes)
    if check_oracle:
        from oracle import sci_cpu

        e_ref, amps_ref, occ_ref, info = sci_cpu.solve(sa, sb, h, g, algo=0, tol=1e-12, max_cycle=200)
        assert abs(res.energy - e_ref) < 1e-8                         # north_star tolerance (Ha)
        assert abs(abs(np.vdot(amps_ref, res.sci_state.amplitudes)) - 1) < 1e-6
        assert np.allclose(res.orbital_occupancies[0], occ_ref[0], atol=1e-6)


def test_qubit_config_full_size(cuda_lib):
    """C3: 40 qubits, 1e4 Pauli terms (2500 X masks x 4 Z masks), 1e5 sampled bitstrings."""
    from qiskit_addon_sqd_b200 import qubit

    nq, d0 = 40, 100_000
    rng = np.random.default_rng(103)
    base = rng.integers(0, 2, nq).astype(bool)
    rows = np.tile(base, (d0, 1))
    k = np.minimum(rng.geometric(0.18, d0) + 1, 14)
    for r in range(d0):
        rows[r, rng.choice(nq, k[r], replace=False)] ^= True
    x, z, c = random_pauli_operator(nq, 2500, 4, 4, 103)
    op = PauliSum(x, z, c)
    t0 = time.perf_counter()
    srt = qubit.sort_and_remove_duplicates(rows)
    t1 = time.perf_counter()
    A = qubit.project_operator_to_subspace(srt, op)
    t2 = time.perf_counter()
    e, v = qubit.solve_qubit(rows, op, k=1, which="SA")
    t3 = time.perf_counter()
    d = srt.shape[0]
    print(f"\n[c3] d = {d}, nnz = {A.nnz}, sort {1e3 * (t1 - t0):.0f} ms, project {1e3 * (t2 - t1):.0f} ms, "
          f"solve_qubit (project + Davidson) {1e3 * (t3 - t2):.0f} ms, E0 = {e[0]:.8f}")
    keys = (srt.astype(np.int64) * (1 << np.arange(nq - 1, -1, -1, dtype=np.int64))[None, :]).sum(1)
    assert np.all(np.diff(keys) > 0)                                  # sorted, unique
    assert A.shape == (d, d) and A.has_canonical_format
    assert abs(A - A.getH()).max() < 1e-12                            # real-weighted Paulis: Hermitian
    # brute-force rows: every term applied to a few source states with python integers
    xm = [sum(1 << q for q in range(nq) if xx[q]) for xx in x]
    zm = [sum(1 << q for q in range(nq) if zz[q]) for zz in z]
    ny = [int(np.count_nonzero(xx & zz)) for xx, zz in zip(x, z)]
    index = {int(kk): i for i, kk in enumerate(keys)}
    for i in rng.choice(d, 25, replace=False):
        expect = {}
        for t in range(len(xm)):
            j = index.get(int(keys[i]) ^ xm[t])
            if j is not None:
                amp = (-1) ** bin(int(keys[i]) & zm[t]).count("1") * (1j) ** ny[t]
                expect[j] = expect.get(j, 0) + c[t] * amp
        expect = {j: val for j, val in expect.items() if val != 0}
        row = A.getrow(int(i))
        assert sorted(expect) == row.indices.tolist()
        assert np.allclose([expect[j] for j in row.indices], row.data, atol=1e-12)
    r = A @ v[:, 0] - e[0] * v[:, 0]
    assert np.linalg.norm(r) < 1e-5 and abs(np.linalg.norm(v[:, 0]) - 1) < 1e-9


def test_recovery_full_size(cuda_lib):
    """1e5 sampled bitstrings, (30e,30o): exact-stream mode and substream mode."""
    from oracle import recovery_oracle as ro
    from qiskit_addon_sqd_b200.configuration_recovery import recover_configurations

    norb, na, nb, n = 30, 15, 15, 100_000
    rng = np.random.default_rng(7)
    bs = rng.integers(2, size=(n, 2 * norb), dtype=np.int64).astype(bool)
    probs = rng.random(n)
    occ = (rng.random(norb), rng.random(norb))
    for mode in ("exact", "parallel"):
        gen = np.random.default_rng(11)
        t0 = time.perf_counter()
        mat, freqs = recover_configurations(bs, probs, occ, na, nb, gen, rng_mode=mode)
        dt = time.perf_counter() - t0
        print(f"\n[recovery {mode}] {n} rows -> {len(mat)} unique in {dt * 1e3:.0f} ms "
              f"({dt / n * 1e9:.0f} ns per input bitstring)")
        assert (mat[:, :norb].sum(1) == nb).all() and (mat[:, norb:].sum(1) == na).all()
        assert abs(freqs.sum() - 1) < 1e-12 and len(np.unique(mat, axis=0)) == len(mat)
    # the exact stream is a prefix property: the first rows of a long run equal a short run
    g1, g2 = np.random.default_rng(5), np.random.default_rng(5)
    m1, f1 = recover_configurations(bs[:3000], probs[:3000], occ, na, nb, g1)
    m2, f2 = ro.recover_configurations(bs[:3000], probs[:3000], occ, na, nb, g2)
    assert np.array_equal(m1, m2) and np.array_equal(f1, f2)
    assert g1.bit_generator.state == g2.bit_generator.state


def test_sqd_loop_config2_full_size(cuda_lib):
    """BASELINE.json configs[1]: (10e,16o), 5 batches of ~1e4 determinants, 3 configuration-recovery
    iterations on one GPU -- the whole loop through the drop-in entry point.  Synthetic integrals and a
    synthetic noisy measurement record (no pyscf / N2 integrals offline)."""
    import functools

    from qiskit_addon_sqd_b200 import fermion
    from qiskit_addon_sqd_b200._synthetic import noisy_samples, random_integrals

    norb, nelec = 16, (5, 5)
    h, g = random_integrals(norb, 102)
    record = noisy_samples(norb, nelec, shots=10_000, n_strings=400, noise=0.04, seed=202)
    solver = functools.partial(fermion.solve_sci_batch, spin_sq=0.0)

    def run():
        history = []
        t0 = time.perf_counter()
        best = fermion.diagonalize_fermionic_hamiltonian(
            h, g, record, samples_per_batch=300, norb=norb, nelec=nelec, num_batches=5, max_iterations=3,
            max_dim=100, sci_solver=solver, symmetrize_spin=True, callback=history.append, seed=5)
        return best, history, time.perf_counter() - t0

    run()  # warm-up (allocator pools, module loads)
    best, history, dt = run()
    best2, history2, _ = run()
    assert len(history) == 3 and all(len(r) == 5 for r in history)
    dims = [r.sci_state.amplitudes.shape for it in history for r in it]
    assert all(d[0] <= 100 and d[1] <= 100 for d in dims) and max(d[0] * d[1] for d in dims) == 10_000
    for it in history:
        for r in it:
            sa = np.asarray(r.sci_state.ci_strs_a)
            assert np.all(np.diff(sa) > 0) and np.all(np.bitwise_count(sa.astype(np.uint64)) == 5)
            assert abs(r.orbital_occupancies[0].sum() - 5) < 1e-9 and abs(r.orbital_occupancies[1].sum() - 5) < 1e-9
            assert abs(np.linalg.norm(r.sci_state.amplitudes) - 1.0) < 1e-9
            e_rdm = np.einsum("pr,pr->", r.rdm1, h) + 0.5 * np.einsum("prqs,prqs->", r.rdm2, g)
            assert abs(e_rdm - r.energy) < 1e-8      # the reference's own energy formula (fermion.py:730-732)
    lowest = [min(r.energy for r in it) for it in history]
    assert best.energy == min(lowest)
    assert lowest[-1] <= lowest[0] + 1e-9            # carry-over + recovery do not lose the best state
    # same seed -> same strings, energies, amplitudes (reference test_fermion.py:285-342)
    assert best2.energy == best.energy and np.array_equal(best2.sci_state.amplitudes, best.sci_state.amplitudes)
    assert all(np.array_equal(a.sci_state.ci_strs_a, b.sci_state.ci_strs_a)
               for x, y in zip(history, history2) for a, b in zip(x, y))
    print(f"\nconfig 2 loop: 3 iterations x 5 subspaces of <= 1e4 dets in {dt*1e3:.1f} ms, "
          f"best energy per iteration {np.round(lowest, 8).tolist()}")


@pytest.mark.parametrize("spin_sq", [None, 0.0])
def test_solve_beyond_the_row_staging_limits(cuda_lib, spin_sq):
    """More beta strings than the staged sigma kernels take (v1: 5760, v2: 8192): the reference has no such
    limit (pyscf ``contract_2e`` behind ``fermion.py:721-723``), the solve must run -- on the wide kernel -- and
    agree with the C oracle (VERDICT r1 missing item 8)."""
    from oracle import sci_cpu
    from qiskit_addon_sqd_b200 import fermion
    from qiskit_addon_sqd_b200._synthetic import hf_centred_strings, random_integrals

    norb, nelec = 20, (5, 5)
    h, g = random_integrals(norb, 77)
    sa = hf_centred_strings(norb, 5, 40, 3)
    sb = hf_centred_strings(norb, 5, 9001, 4)
    res = fermion.solve_sci((sa.astype(np.int64), sb.astype(np.int64)), h, g, norb, nelec, spin_sq=spin_sq)
    st = fermion.last_solve_stats()[-1]
    assert st.sigma_path == 3 and st.converged == 1
    e_ref, amps_ref, occ_ref, info = sci_cpu.solve(sa, sb, h, g, algo=0, tol=1e-12, max_cycle=200,
                                                   spin_sq=spin_sq, shift=0.2)  # solve_sci: pyscf's default shift
    # Without the penalty the energy is an eigenvalue (second order in the eigenvector error): north_star
    # tolerance.  With it the REPORTED energy is <x|H|x> in an eigenvector of H + shift (S^2 - ss), first order
    # in the eigenvector error: the oracle's own answer moves by 3.5e-8 Ha between tol = 1e-12 and 1e-15 here,
    # so the two solvers are compared at 2e-7 and the energy is pinned through the oracle's sigma instead.
    assert abs(res.energy - e_ref) < (1e-8 if spin_sq is None else 2e-7)
    assert abs(abs(np.vdot(res.sci_state.amplitudes, amps_ref)) - 1.0) < 1e-7
    x = res.sci_state.amplitudes
    assert abs(np.vdot(x, sci_cpu.sigma(sa, sb, h, g, x)) / np.vdot(x, x) - res.energy) < 1e-10
    e_rdm = np.einsum("pr,pr->", res.rdm1, h) + 0.5 * np.einsum("prqs,prqs->", res.rdm2, g)
    assert abs(e_rdm - res.energy) < 1e-8

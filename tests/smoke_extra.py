"""Second half of ``__graft_entry__.smoke()``: qubit projection and configuration recovery on cuda:0,
each checked against its CPU oracle."""

import numpy as np


def smoke_extra() -> None:
    from oracle import qubit_oracle as qo
    from oracle import recovery_oracle as ro
    from qiskit_addon_sqd_b200 import configuration_recovery, qubit
    from qiskit_addon_sqd_b200._synthetic import PauliSum, random_pauli_operator

    nq, d0 = 20, 400
    rng = np.random.default_rng(5)
    base = rng.integers(0, 2, nq).astype(bool)
    rows = np.tile(base, (d0, 1))
    for r in range(d0):
        k = rng.integers(0, 4)
        if k:
            rows[r, rng.choice(nq, k, replace=False)] ^= True
    x, z, c = random_pauli_operator(nq, 30, 3, 3, 1)
    op = PauliSum(x, z, c)
    srt = qubit.sort_and_remove_duplicates(rows)
    proj = qubit.project_operator_to_subspace(srt, op)
    ref = qo.project_operator_to_subspace(qo.sort_and_remove_duplicates(rows), op)
    ref.sort_indices()
    assert np.array_equal(proj.indices, ref.indices) and np.array_equal(proj.data, ref.data)
    e, _ = qubit.solve_qubit(rows, op, k=1, which="SA")
    e_ref, _ = qo.solve_qubit(rows, op, k=1, which="SA")
    assert abs(e[0] - e_ref[0]) < 1e-8, (e, e_ref)
    print(f"[smoke] project_operator_to_subspace / solve_qubit ok: nnz={proj.nnz}, E0={e[0]:.10f}")

    norb, na, nb, n = 10, 4, 5, 500
    bs = rng.integers(2, size=(n, 2 * norb), dtype=np.int64).astype(bool)
    probs = rng.random(n)
    occ = (rng.random(norb), rng.random(norb))
    g1, g2 = np.random.default_rng(3), np.random.default_rng(3)
    mat, freqs = configuration_recovery.recover_configurations(bs, probs, occ, na, nb, g1)
    mat_ref, freqs_ref = ro.recover_configurations(bs, probs, occ, na, nb, g2)
    assert np.array_equal(mat, mat_ref) and np.array_equal(freqs, freqs_ref)
    assert g1.bit_generator.state == g2.bit_generator.state
    print(f"[smoke] recover_configurations ok: {n} rows -> {len(mat)} unique, stream-exact")

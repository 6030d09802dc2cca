"""Generate golden vectors by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

qiskit and jax are not installed here; ``oracle/shims`` provides the few names the reference imports
(``qiskit.quantum_info.{Pauli,SparsePauliOp}``, ``qiskit.primitives.BitArray``,
``qiskit.utils.deprecation.deprecate_func`` and a numpy-backed ``jax``).  pyscf is absent too: ``oracle/shims/pyscf``
stands in for ``pyscf.fci.selected_ci`` with the dense CPU oracle, which lets the reference's own
``fermion.py`` run -- its SQD loop, string handling and result conventions are the reference's code, only
the eigen-solver arithmetic is the oracle's (``sqd_loop_golden.npz``).
The outputs are committed as small ``.npz`` fixtures next to this script.
"""

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "shims"))
sys.path.insert(0, "/root/reference")

# import the submodules directly: the package __init__ pulls in fermion.py -> pyscf
import importlib.util


def _load(name):
    spec = importlib.util.spec_from_file_location(
        f"refsqd_{name}", f"/root/reference/qiskit_addon_sqd/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ref_qubit = _load("qubit")
ref_rec = _load("configuration_recovery")
from qiskit.quantum_info import Pauli, SparsePauliOp  # noqa: E402  (shim)


def qubit_goldens():
    out = {}
    rng = np.random.default_rng(2024)
    cases = [(6, 40, 12), (10, 150, 20), (17, 200, 25), (33, 120, 16), (63, 80, 10)]
    for ci, (nq, n_rows, n_terms) in enumerate(cases):
        # rows clustered so that X flips land inside the set
        base = rng.integers(0, 2, size=nq, dtype=np.int64).astype(bool)
        rows = np.tile(base, (n_rows, 1))
        for r in range(n_rows):
            k = rng.integers(0, 4)
            if k:
                rows[r, rng.choice(nq, k, replace=False)] ^= True
        rows = np.concatenate([rows, rows[: n_rows // 5]])  # duplicates for sort_and_remove_duplicates
        rng.shuffle(rows)
        srt = ref_qubit.sort_and_remove_duplicates(rows)
        labels = []
        for _ in range(n_terms):
            w = rng.integers(0, 4)
            lab = ["I"] * nq
            for q in (rng.choice(nq, w, replace=False) if w else []):
                lab[q] = "XY"[rng.integers(0, 2)]
            for q in range(nq):
                if lab[q] == "I" and rng.random() < 0.3:
                    lab[q] = "Z"
            labels.append("".join(lab))
        labels += labels[:3]  # repeated strings: exercises accumulation order / cancellation
        coeffs = rng.standard_normal(len(labels)) + 1j * rng.standard_normal(len(labels))
        coeffs[-1] = -coeffs[2]  # exact cancellation -> scipy drops the zero
        op = SparsePauliOp(labels, coeffs)
        proj = ref_qubit.project_operator_to_subspace(srt, op)
        proj.sort_indices()
        out[f"c{ci}_rows_in"] = rows
        out[f"c{ci}_rows_sorted"] = srt
        out[f"c{ci}_labels"] = np.array(labels)
        out[f"c{ci}_coeffs"] = coeffs
        out[f"c{ci}_data"] = np.asarray(proj.data)
        out[f"c{ci}_indices"] = np.asarray(proj.indices)
        out[f"c{ci}_indptr"] = np.asarray(proj.indptr)
        for ti in range(4):
            amp, r, c = ref_qubit.matrix_elements_from_pauli(srt, Pauli(labels[ti]))
            out[f"c{ci}_t{ti}_amp"] = np.asarray(amp)
            out[f"c{ci}_t{ti}_row"] = np.asarray(r)
            out[f"c{ci}_t{ti}_col"] = np.asarray(c)
        # Hermitian operator for the eigenvalue golden: real coefficients, each string once
        hop = SparsePauliOp(labels[:n_terms], np.real(coeffs[:n_terms]))
        e, v = ref_qubit.solve_qubit(rows, hop, k=1, which="SA")
        out[f"c{ci}_e0"] = e
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "qubit_golden.npz"), **out)
    print("qubit goldens:", len(out), "arrays")


def recovery_goldens():
    out = {}
    cases = [
        # norb, (nelec_a, nelec_b), rows, seed
        (4, (2, 2), 60, 11),
        (8, (3, 5), 200, 12),
        (16, (5, 5), 300, 13),
        (30, (15, 15), 200, 14),
        (37, (9, 12), 80, 15),
        (64, (20, 30), 40, 16),
    ]
    for ci, (norb, (na, nb), n, seed) in enumerate(cases):
        rng = np.random.default_rng(seed)
        bs = rng.integers(2, size=(n, 2 * norb), dtype=np.int64).astype(bool)
        bs[: n // 4] = bs[n // 4: 2 * (n // 4)]  # repeated rows -> merging of duplicates
        probs = rng.random(n)
        probs /= probs.sum()
        occ = (np.clip(rng.random(norb), 0.0, 1.0), np.clip(rng.random(norb), 0.0, 1.0))
        occ[0][rng.random(norb) < 0.1] = 0.0
        occ[1][rng.random(norb) < 0.1] = 1.0
        gen = np.random.default_rng(1000 + seed)
        mat, freqs = ref_rec.recover_configurations(bs, probs, occ, na, nb, rand_seed=gen)
        st = gen.bit_generator.state["state"]
        m = (1 << 64) - 1
        out[f"c{ci}_meta"] = np.array([norb, na, nb, n, 1000 + seed])
        out[f"c{ci}_bs"] = bs
        out[f"c{ci}_probs"] = probs
        out[f"c{ci}_occ_a"] = occ[0]
        out[f"c{ci}_occ_b"] = occ[1]
        out[f"c{ci}_mat"] = mat
        out[f"c{ci}_freqs"] = freqs
        out[f"c{ci}_state"] = np.array(
            [st["state"] >> 64, st["state"] & m, st["inc"] >> 64, st["inc"] & m], dtype=np.uint64)
        # the caller's stream continues: next doubles after the call
        out[f"c{ci}_next"] = gen.random(4)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "recovery_golden.npz"), **out)
    print("recovery goldens:", len(out), "arrays")


def loop_goldens():
    """The reference's ``diagonalize_fermionic_hamiltonian`` (unmodified source) on small noisy samples."""
    import functools
    import math

    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import qiskit_addon_sqd.fermion as ref_fermion  # whole package: fermion.py imports through the pyscf shim
    from qiskit.primitives import BitArray  # shim
    from qiskit_addon_sqd_b200._synthetic import hf_centred_strings, random_integrals

    from make_golden_cases import LOOP_CASES

    shapes = [dict(norb=6, nelec=(3, 3), shots=1500, noise=0.08),
              dict(norb=7, nelec=(3, 2), shots=1200, noise=0.05),
              dict(norb=5, nelec=(2, 2), shots=800, noise=0.10)]
    cases = [dict(shape, **extra) for shape, extra in zip(shapes, LOOP_CASES)]
    out = {"n_cases": np.array(len(cases))}
    for ci, case in enumerate(cases):
        norb, nelec = case["norb"], case["nelec"]
        h, g = random_integrals(norb, 50 + ci)
        rng = np.random.default_rng(900 + ci)
        pool_a = hf_centred_strings(norb, nelec[0], min(14, math.comb(norb, nelec[0])), 10 + ci)
        pool_b = hf_centred_strings(norb, nelec[1], min(14, math.comb(norb, nelec[1])), 20 + ci)
        wa = np.exp(-0.5 * np.arange(len(pool_a)))
        wb = np.exp(-0.5 * np.arange(len(pool_b)))
        ia = rng.choice(len(pool_a), case["shots"], p=wa / wa.sum())
        ib = rng.choice(len(pool_b), case["shots"], p=wb / wb.sum())
        bits = np.zeros((case["shots"], 2 * norb), dtype=bool)
        for k in range(norb):   # columns: [b_{N-1} .. b_0, a_{N-1} .. a_0]
            bits[:, norb - 1 - k] = (pool_b[ib].astype(np.uint64) >> np.uint64(k)) & np.uint64(1)
            bits[:, 2 * norb - 1 - k] = (pool_a[ia].astype(np.uint64) >> np.uint64(k)) & np.uint64(1)
        bits ^= rng.random(bits.shape) < case["noise"]
        bit_array = BitArray.from_bool_array(bits)
        kw = dict(case["kw"])
        if ci == 1:
            kw["initial_occupancies"] = (np.linspace(0.9, 0.1, norb), np.linspace(0.8, 0.05, norb))
        history = []
        solver = functools.partial(ref_fermion.solve_sci_batch, spin_sq=case["spin_sq"])
        best = ref_fermion.diagonalize_fermionic_hamiltonian(
            h, g, bit_array, norb=norb, nelec=nelec, sci_solver=solver, callback=history.append, **kw)
        out[f"c{ci}_meta"] = np.array([norb, nelec[0], nelec[1], len(history), kw["num_batches"]])
        out[f"c{ci}_packed"] = bit_array.array.reshape(case["shots"], -1)
        out[f"c{ci}_best_energy"] = np.array(best.energy)
        out[f"c{ci}_best_occ"] = np.array(best.orbital_occupancies)
        for it, results in enumerate(history):
            for k, r in enumerate(results):
                out[f"c{ci}_i{it}_b{k}_a"] = np.asarray(r.sci_state.ci_strs_a, dtype=np.int64)
                out[f"c{ci}_i{it}_b{k}_b"] = np.asarray(r.sci_state.ci_strs_b, dtype=np.int64)
                out[f"c{ci}_i{it}_b{k}_e"] = np.array(r.energy)
                out[f"c{ci}_i{it}_b{k}_occ"] = np.array(r.orbital_occupancies)
        print(f"loop case {ci}: {len(history)} iterations, best energy {best.energy:.10f}, dims",
              [tuple(r.sci_state.amplitudes.shape) for r in history[-1]])
    np.savez_compressed(os.path.join(HERE, "sqd_loop_golden.npz"), **out)


if __name__ == "__main__":
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        qubit_goldens()
        recovery_goldens()
        loop_goldens()

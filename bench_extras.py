"""Secondary measurements printed under ``extra`` in bench.py's JSON line.

Everything here goes through the product path (the C-ABI library behind ``qiskit_addon_sqd_b200``); the
oracle is used only for the bounded ``cpu_baseline`` samples, as in bench.py itself.

N = 1 (rank 0):
  fermion      sigma-build time, lone-solve latency and compulsory-byte GB/s at c4, t (north_star target
               shape) and c5
  qubit_c3     BASELINE.json configs[2]: 40 qubits, 1e4 Pauli terms, 1e5 configurations
  pauli_z40    the reference's published projection benchmark: one Z^(x)40 term, d ~ 5e7
               (docs/guides/benchmark_pauli_projection.ipynb, 4.174 s on the authors' CPU)
  recovery     recover_configurations on 1e5 sampled rows, exact-stream and parallel modes
N > 1 (all ranks take part, rank 0 reports):
  c5_sharded   configs[4]: ONE 1e6-determinant diagonalisation with the sigma build sharded over the ranks
  c4_strong    the literal configs[3]: 8 subspaces of 1e5 determinants in total, spread over the ranks
"""

from __future__ import annotations

import os
import time

import numpy as np


def _event_ms(torch, fn, reps: int, warm: int = 1):
    """Median device time of fn() over ``reps`` runs (CUDA events on the current stream)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts, out = [], None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), out


def _wall_ms(torch, fn, reps: int, warm: int = 1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts, out = [], None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts)), out


# ------------------------------------------------------------------------------------------------
# fermion: sigma build and lone solve per workload
# ------------------------------------------------------------------------------------------------
def fermion_extras(bench, torch, fermion, dev, peak_gbs: float, workloads=("c4", "t", "c5")) -> dict:
    out = {}
    for wl in workloads:
        norb, nelec, h, g, batches = bench.make_batches(wl, 0, 1)
        sa, sb = batches[0]
        ints = fermion._DeviceIntegrals(torch, h, g, dev)
        opts = fermion._solver_options({})
        sub = fermion._Subspace(sa, sb, norb, h, g)
        ham = sub.hamiltonian()
        x = sub.upload_amplitudes(np.random.default_rng(0).standard_normal((sub.na, sub.nb)))
        y = sub.new_vector()
        reps = 200 if wl != "c5" else 40

        def builds():
            for _ in range(reps):
                sub.apply(ham, x, y)

        ms, _ = _event_ms(torch, builds, 3)
        sigma_us = 1e3 * ms / reps
        del sub, ham, x, y

        def lone():
            return fermion._solve_on_device(sa, sb, norb, ints, None, 0.2, opts, want_spin=False,
                                            want_rdm=False, download=False, profile=False)

        solve_ms, r = _event_ms(torch, lone, 5, warm=2)
        st = r["stats"]
        b_sigma = bench.sigma_algorithmic_bytes(st)
        out[wl] = {
            "n_det": int(st.n_det), "norb": norb, "nelec": list(nelec),
            "sigma_us": sigma_us, "sigma_path": "v2" if getattr(st, "sigma_path", 1) == 2 else "v1",
            "sigma_bytes": b_sigma, "sigma_gbs": b_sigma / (sigma_us * 1e-6) / 1e9,
            "sigma_frac_of_hbm_peak": b_sigma / (sigma_us * 1e-6) / 1e9 / peak_gbs,
            "lone_solve_ms": solve_ms, "cycles": int(st.cycles), "energy": float(r["energy"]),
            "lone_mdet_per_s": st.n_det / (solve_ms * 1e-3) / 1e6,
        }
    out["note"] = ("sigma_us: CUDA-event time of back-to-back sigma builds on one stream (warm L2); "
                   "lone_solve_ms: one solve_sci-equivalent on an idle GPU (tables + Davidson to 1e-12 + "
                   "occupancies), inputs resident, median of 5")
    return out


# ------------------------------------------------------------------------------------------------
# qubit: configs[2] and the published Z^(x)40 projection benchmark
# ------------------------------------------------------------------------------------------------
def _c3_inputs(nq: int = 40, d: int = 100_000, seed: int = 103):
    from qiskit_addon_sqd_b200._synthetic import PauliSum, random_pauli_operator

    rng = np.random.default_rng(seed)
    base = rng.integers(0, 2, nq).astype(bool)
    n0 = 3 * d
    rows = np.tile(base, (n0, 1))
    k = rng.integers(0, 9, n0)
    cols = rng.integers(0, nq, (n0, 8))
    idx = np.arange(n0)
    for j in range(8):
        sel = k > j
        rows[idx[sel], cols[sel, j]] ^= True
    # first d distinct rows, in sampling order (the public entry points sort them themselves)
    w = (np.uint64(1) << np.arange(nq - 1, -1, -1, dtype=np.uint64))
    keys = (rows.astype(np.uint64) * w[None, :]).sum(axis=1, dtype=np.uint64)
    _, first = np.unique(keys, return_index=True)
    rows = rows[np.sort(first)[:d]]
    x, z, c = random_pauli_operator(nq, 2500, 4, 3, 7)
    return rows, PauliSum(x, z, c)


def qubit_extras(torch, dev, peak_gbs: float, cpu: bool = True) -> dict:
    from qiskit_addon_sqd_b200 import _lib, qubit

    lib = _lib.load()
    rows, op = _c3_inputs()
    d_in = rows.shape[0]
    keys = torch.unique(qubit._keys_device(torch, lib, rows))
    d = int(keys.numel())
    proj_ms, csr = _event_ms(torch, lambda: qubit._project_device(torch, lib, keys, op), 3)
    kt = {}
    qubit._project_device(torch, lib, keys, op, timings=kt)
    kern_ms = kt["table_and_count_ms"] + kt["fill_and_sort_ms"]
    nnz, T = csr.nnz, op.size
    # bytes model: keys read once, the term table once, CSR written once (int32 col + complex128 val), row_ptr
    proj_bytes = 8.0 * d + 36.0 * T + 20.0 * nnz + 4.0 * (d + 1)
    xv = torch.randn(2 * d, dtype=torch.float64, device=dev)
    yv = torch.empty_like(xv)
    st = _lib.stream_ptr(torch)

    def matvecs():
        for _ in range(50):
            lib.sqd_csr_matvec_c128(d, _lib.ptr(csr.row_ptr), _lib.ptr(csr.col), _lib.ptr(csr.val),
                                    _lib.ptr(xv), _lib.ptr(yv), st)

    mv_ms, _ = _event_ms(torch, matvecs, 3)
    mv_us = 1e3 * mv_ms / 50
    mv_bytes = 20.0 * nnz + 4.0 * (d + 1) + 32.0 * d
    e2e_proj_ms, _ = _wall_ms(torch, lambda: qubit.project_operator_to_subspace(
        qubit.sort_and_remove_duplicates(rows), op), 2)
    e2e_ms, (ev, vec) = _wall_ms(torch, lambda: qubit.solve_qubit(rows, op, k=1, which="SA"), 2)
    out = {
        "workload": f"configs[2]: {rows.shape[1]} qubits, {T} Pauli terms (2500 X masks x 4 Z masks, weight <= 3), "
                    f"{d_in} sampled configurations ({d} unique)",
        "d": d, "terms": T, "nnz": int(nnz),
        "project_ms_device": proj_ms, "project_kernels_ms": kern_ms, "project_kernels_split_ms": kt,
        "project_term_rows_per_s": T * d / (kern_ms * 1e-3),
        "project_bytes": proj_bytes, "project_gbs": proj_bytes / (kern_ms * 1e-3) / 1e9,
        "project_frac_of_hbm_peak": proj_bytes / (kern_ms * 1e-3) / 1e9 / peak_gbs,
        "project_note": "project_ms_device spans the whole call with the keys resident (host-side grouping of the "
                        "terms by X mask and the uploads of the term table included); project_kernels_ms is the "
                        "CUDA-event time of key table + count pass + fill pass + row sort; the bytes model counts "
                        "compulsory traffic only, the work is 2500 x 1e5 table probes per pass (L2-resident)",
        "matvec_us": mv_us, "matvec_gbs": mv_bytes / (mv_us * 1e-6) / 1e9,
        "matvec_frac_of_hbm_peak": mv_bytes / (mv_us * 1e-6) / 1e9 / peak_gbs,
        "e2e_sort_and_project_ms": e2e_proj_ms,
        "e2e_solve_qubit_ms": e2e_ms, "energy": float(ev[0]),
    }
    if cpu:
        from oracle import qubit_oracle as qo

        n_terms = 100
        from qiskit_addon_sqd_b200._synthetic import PauliSum

        sub_op = PauliSum(np.array([p.x for p in op.paulis[:n_terms]]),
                          np.array([p.z for p in op.paulis[:n_terms]]), op.coeffs[:n_terms])
        srt = qo.sort_and_remove_duplicates(rows)
        t0 = time.perf_counter()
        qo.project_operator_to_subspace(srt, sub_op)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {
            "kind": "port", "cores": 1, "sample": f"first {n_terms} of {T} terms on all {d} rows, "
            "oracle/qubit_oracle.py (vectorised numpy restatement of qubit.py:78-144)",
            "seconds": dt, "project_ms_extrapolated": 1e3 * dt * T / n_terms,
        }
    return out


def pauli_z40_extras(torch, dev, d_target: int = 50_000_000) -> dict:
    from qiskit_addon_sqd_b200 import qubit
    from qiskit_addon_sqd_b200._synthetic import PauliTerm

    nq = 40
    gen = torch.Generator(device=dev)
    gen.manual_seed(22)
    keys = torch.unique(torch.randint(0, 1 << nq, (d_target,), dtype=torch.int64, device=dev, generator=gen))
    d = int(keys.numel())
    # unpack on the device in slabs, into one pinned host matrix (the bench input lives on the host)
    host = torch.empty((d, nq), dtype=torch.bool, pin_memory=True)
    shifts = torch.arange(nq - 1, -1, -1, dtype=torch.int64, device=dev)
    slab = 1 << 22
    for lo in range(0, d, slab):
        hi = min(d, lo + slab)
        host[lo:hi].copy_(((keys[lo:hi, None] >> shifts[None, :]) & 1).to(torch.bool))
    del keys
    torch.cuda.synchronize()
    rows = host.numpy()
    pauli = PauliTerm(np.zeros(nq, dtype=bool), np.ones(nq, dtype=bool))
    ms, (amp, r, c) = _wall_ms(torch, lambda: qubit.matrix_elements_from_pauli(rows, pauli), 2)
    ok = bool(len(amp) == d and np.array_equal(r[:1000], c[:1000]) and
              np.array_equal(amp[:4096].real, 1.0 - 2.0 * (rows[:4096].sum(axis=1) % 2)))
    return {
        "workload": f"matrix_elements_from_pauli(Z^(x)40) on {d} sorted unique 40-bit configurations, host bool "
                    "matrix in, (amplitudes, rows, cols) host arrays out",
        "d": d, "e2e_ms": ms, "rows_per_s": d / (ms * 1e-3),
        "h2d_bytes": int(d * nq), "d2h_bytes": int(d * 32),
        "published_reference_s": 4.174, "published_source": "docs/guides/benchmark_pauli_projection.ipynb "
        "(authors' CPU, d = 49 998 839)", "speedup_vs_published": 4.174 / (ms * 1e-3),
        "spot_check_ok": ok,
    }


# ------------------------------------------------------------------------------------------------
# configuration recovery
# ------------------------------------------------------------------------------------------------
def recovery_extras(torch, dev, cpu: bool = True) -> dict:
    from qiskit_addon_sqd_b200 import configuration_recovery as cr
    from qiskit_addon_sqd_b200._synthetic import noisy_samples

    norb, nelec, n = 30, (15, 15), 100_000
    ba = noisy_samples(norb, nelec, n, 316, 0.03, 106)
    bits = np.unpackbits(ba.array, axis=1)[:, -2 * norb:].astype(bool)
    probs = np.full(n, 1.0 / n)
    occ_b = bits[:, :norb][:, ::-1].mean(axis=0)
    occ_a = bits[:, norb:][:, ::-1].mean(axis=0)
    occ = (occ_a, occ_b)
    wrong = int(np.count_nonzero((bits[:, :norb].sum(axis=1) != nelec[1]) |
                                 (bits[:, norb:].sum(axis=1) != nelec[0])))
    out = {"workload": f"recover_configurations: {n} sampled rows of 2 x {norb} bits, target ({nelec[0]},{nelec[1]}) "
                       f"electrons, {wrong} rows need repair (3 % bit-flip noise)", "rows": n,
           "rows_needing_repair": wrong}
    for mode in ("exact", "parallel"):
        ms, (mat, p) = _wall_ms(torch, lambda: cr.recover_configurations(
            bits, probs, occ, nelec[0], nelec[1], rand_seed=np.random.default_rng(7), rng_mode=mode), 2)
        good = bool(np.all(mat[:, :norb].sum(axis=1) == nelec[1]) and np.all(mat[:, norb:].sum(axis=1) == nelec[0])
                    and abs(p.sum() - 1.0) < 1e-12)
        out[mode] = {"e2e_ms": ms, "us_per_row": 1e3 * ms / n, "unique_out": int(mat.shape[0]),
                     "hamming_ok": good}
    if cpu:
        from oracle import recovery_oracle as ro

        ns = 2000
        t0 = time.perf_counter()
        ro.recover_configurations(bits[:ns], probs[:ns] * (n / ns), occ, nelec[0], nelec[1],
                                  np.random.default_rng(7))
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"kind": "port", "cores": 1, "sample": f"first {ns} rows, oracle/recovery_oracle.py "
                               "(restatement of configuration_recovery.py:59-304; the unmodified reference "
                               "measured 265-403 us/row, SURVEY 8d)", "seconds": dt, "us_per_row": 1e6 * dt / ns}
    return out


# ------------------------------------------------------------------------------------------------
# sample formats and loop glue: bit_array_to_arrays, carry-over selection, the whole SQD loop
# ------------------------------------------------------------------------------------------------
def formats_extras(torch, fermion, dev, cpu: bool = True) -> dict:
    from qiskit_addon_sqd_b200 import counts
    from qiskit_addon_sqd_b200._synthetic import noisy_samples

    out = {}
    norb, nelec, shots = 30, (15, 15), 1_000_000
    ba = noisy_samples(norb, nelec, shots, 3162, 0.03, 110)
    ms, (rows, probs) = _wall_ms(torch, lambda: counts.bit_array_to_arrays(ba), 3)
    rec = {"workload": f"bit_array_to_arrays: {shots} shots of {2 * norb} bits (packed bytes in, bool matrix of the "
                       "distinct rows + probabilities out, host arrays)", "e2e_ms": ms, "unique_rows": int(rows.shape[0]),
           "shots_per_s": shots / (ms * 1e-3)}
    if cpu:
        # the reference's own three numpy calls (counts.py:57-60) on a bounded sample of the same shots
        ns = 100_000
        sub = ba.array[:ns]
        t0 = time.perf_counter()
        bool_array = np.unpackbits(sub, axis=-1)[..., -2 * norb:].astype(bool)
        np.unique(bool_array, axis=0, return_counts=True)
        dt = time.perf_counter() - t0
        rec["cpu_baseline"] = {"kind": "reference", "cores": 1, "sample": f"first {ns} shots, the numpy calls of "
                               "counts.py:57-60 verbatim", "seconds": dt, "shots_per_s": ns / dt}
    out["bit_array_to_arrays"] = rec

    # carry-over selection at 1e5 and 1e6 amplitudes: device (resident amplitudes) vs the reference's numpy lines
    for na in (316, 1000):
        rng = np.random.default_rng(na)
        amps = rng.standard_normal((na, na)) * np.exp(-6.0 * rng.random((na, na)))
        amps /= np.linalg.norm(amps)
        x = torch.from_numpy(amps).to(dev)
        thr = 1e-4
        ms, sel = _wall_ms(torch, lambda: fermion._carryover_on_device(x, na, dev.index or 0, thr), 5)
        rec = {"n_det": na * na, "e2e_ms": ms, "rows_selected": int(len(sel[0])), "cols_selected": int(len(sel[1]))}
        if cpu:
            t0 = time.perf_counter()
            flat = np.abs(amps.reshape(-1))
            order = np.argsort(flat)
            big = order[np.searchsorted(flat, thr, sorter=order):]
            r, c = np.divmod(big, na)
            r, c = np.unique(r), np.unique(c)
            wa = np.sum(np.abs(amps[r]) ** 2, axis=1)
            wb = np.sum(np.abs(amps[:, c]) ** 2, axis=0)
            rec["cpu_reference_ms"] = 1e3 * (time.perf_counter() - t0)
            rec["bit_equal"] = bool(np.array_equal(sel[0], r) and np.array_equal(sel[1], c)
                                    and np.array_equal(sel[2], wa) and np.array_equal(sel[3], wb))
        out[f"carryover_{na * na}"] = rec
    out["carryover_note"] = ("sqd_carryover on amplitudes resident on the device + read-back of na+nb flags/weights, "
                             "against the reference's lines fermion.py:607-622 (argsort over all amplitudes) on the host")
    return out


def sqd_loop_extras(torch, fermion) -> dict:
    """BASELINE configs[1]: the whole SQD loop, (10e,16o), 5 batches x <= 1e4 determinants, 3 recovery iterations."""
    import functools

    from qiskit_addon_sqd_b200._synthetic import noisy_samples, random_integrals

    norb, nelec = 16, (5, 5)
    h, g = random_integrals(norb, 102)
    record = noisy_samples(norb, nelec, shots=10_000, n_strings=400, noise=0.04, seed=202)
    solver = functools.partial(fermion.solve_sci_batch, spin_sq=0.0, compute_rdms=False)

    def run():
        return fermion.diagonalize_fermionic_hamiltonian(
            h, g, record, samples_per_batch=300, norb=norb, nelec=nelec, num_batches=5, max_iterations=3,
            max_dim=100, sci_solver=solver, symmetrize_spin=True, seed=5)

    ms, best = _wall_ms(torch, run, 3)
    return {"workload": "configs[1]: diagonalize_fermionic_hamiltonian, (10e,16o), 10 000 shots, 3 iterations x 5 "
                        "subspaces of <= 100 x 100 strings, spin_sq = 0, host arrays in and out",
            "loop_ms": ms, "ms_per_iteration": ms / 3, "best_energy": float(best.energy),
            "n_det_best": int(best.sci_state.amplitudes.size)}


# ------------------------------------------------------------------------------------------------
# scale-up point of SURVEY 8(d): one subspace of 1e8 determinants (vectors >> L2, rows beyond the sigma staging)
# ------------------------------------------------------------------------------------------------
def s8_extras(torch, fermion, dev, peak_gbs: float, n: int = 10_000, cycles: int = 4) -> dict:
    from qiskit_addon_sqd_b200._synthetic import hf_centred_strings, random_integrals

    free, _total = torch.cuda.mem_get_info()
    need = 22.0 * 8.0 * n * n   # vectors of the Davidson workspace + diagonal + sigma + tables, with headroom
    if free < need:
        return {"skipped": f"needs {need / 1e9:.0f} GB of free device memory, {free / 1e9:.0f} GB available"}
    norb, ne = 30, 15
    h, g = random_integrals(norb, 108)
    sa = hf_centred_strings(norb, ne, n, 21)
    sb = hf_centred_strings(norb, ne, n, 22)
    ints = fermion._DeviceIntegrals(torch, h, g, dev)
    opts = fermion._solver_options({"max_cycle": cycles})
    t0 = time.perf_counter()
    r = fermion._solve_on_device(sa, sb, norb, ints, None, 0.2, opts, want_spin=False, want_rdm=False,
                                 download=False, profile=True)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    st = r["stats"]
    fma = (st.singles_a / st.na) * (st.singles_b / st.nb) * st.n_det + st.nnz_a * st.nb + st.nnz_b * st.na
    sigma_ms = st.sigma_ms / max(st.sigma_builds, 1)
    rest_ms = (st.davidson_ms - st.sigma_ms) / max(st.cycles, 1)
    # basis sizes m = 1 .. cycles (no restart within `cycles` <= max_space): gram m+1, residual 2m+3, ortho1/2 m+2 each
    ms_ = range(1, st.cycles + 1)
    vec_bytes = 8.0 * st.n_det * sum((m + 1) + (2 * m + 3) + 2 * (m + 2) for m in ms_) / max(st.cycles, 1)
    del r
    return {
        "workload": f"s8: ONE (30e,30o) subspace of {n} x {n} = {st.n_det} determinants, {st.cycles} Davidson cycles "
                    "(not converged: a timing of the kernels at the scale where vectors no longer fit in L2)",
        "sigma_path": {1: "v1", 2: "v2", 3: "wide"}.get(st.sigma_path, "?"),
        "nnz_a": st.nnz_a, "nnz_b": st.nnz_b, "singles_a": st.singles_a, "singles_b": st.singles_b,
        "sigma_ms_per_build": sigma_ms, "gathered_fma_per_build": fma,
        "sigma_gfma_per_s": fma / (sigma_ms * 1e-3) / 1e9,
        "vector_kernels_ms_per_cycle": rest_ms, "vector_bytes_per_cycle_model": vec_bytes,
        "vector_kernels_gbs": vec_bytes / (rest_ms * 1e-3) / 1e9,
        "vector_kernels_frac_of_hbm_peak": vec_bytes / (rest_ms * 1e-3) / 1e9 / peak_gbs,
        "wall_s_incl_tables": wall,
    }


# ------------------------------------------------------------------------------------------------
# N > 1: sharded single solve and strong scaling of the literal configs[3]
# ------------------------------------------------------------------------------------------------
def sharded_extras(bench, torch, dist, fermion, rank: int, world: int, dev, workload: str = "c5") -> dict:
    from qiskit_addon_sqd_b200._dispatch import max_over_ranks

    norb, nelec, h, g, batches = bench.make_batches(workload, 0, 1)
    sa, sb = batches[0]
    group = fermion.ShardGroup()

    def single():
        return fermion.solve_sci((sa, sb), h, g, norb, nelec)

    def sharded():
        return fermion.solve_sci_sharded((sa, sb), h, g, norb, nelec, group=group)

    ms1, ref = (None, None)
    if rank == 0:
        ms1, ref = _wall_ms(torch, single, 3)
    dist.barrier()
    sharded()
    ts = []
    for _ in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        res = sharded()
        torch.cuda.synchronize()
        ts.append(max_over_ranks(1e3 * (time.perf_counter() - t0), dev))
    ms_n = float(np.median(ts))
    e_all = [None] * world
    dist.all_gather_object(e_all, float(res.energy))
    group.close()
    if rank != 0:
        return {}
    n_det = len(sa) * len(sb)
    ldc = (len(sb) + 1) // 2 * 2
    gather = os.environ.get("SQD_SHARD_GATHER", "1") != "0"
    per_build = 8.0 * len(sa) * ldc * (world - 1) / world * (1 if gather else 2)
    dE = abs(float(res.energy) - float(ref.energy))
    assert len(set(e_all)) == 1, f"sharded energies differ between ranks: {e_all}"
    assert dE < 1e-9, f"sharded energy differs from the single-GPU solve by {dE}"
    return {
        "workload": f"{workload}: ONE ({sum(nelec)}e,{norb}o) diagonalisation of {n_det} determinants, host arrays in "
                    "and out (solve_sci vs solve_sci_sharded)",
        "ms_1gpu": ms1, "ms_sharded": ms_n, "speedup": ms1 / ms_n, "ranks": world,
        "exchange": "grouped ncclBroadcast of the disjoint row blocks" if gather else "ncclAllReduce of padded vectors",
        "exchange_bytes_per_rank_per_build": per_build, "allreduce_bytes": per_build,
        "dE": dE, "energies_equal_on_all_ranks": True, "energy": float(res.energy),
        "mdet_per_s_sharded": n_det / (ms_n * 1e-3) / 1e6,
    }


def strong_extras(bench, torch, dist, fermion, rank: int, world: int, dev, total: int = 8) -> dict:
    """The literal configs[3]: ``total`` subspaces of 1e5 determinants, subspace k on rank k mod N."""
    from qiskit_addon_sqd_b200._dispatch import max_over_ranks

    norb, nelec, h, g, batches = bench.make_batches("c4", 0, total)
    mine = [b for k, b in enumerate(batches) if k % world == rank]

    def step():
        return fermion.solve_sci_batch(mine, h, g, norb, nelec, compute_rdms=False) if mine else []

    step()
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        res = step()
        torch.cuda.synchronize()
        ts.append(max_over_ranks(1e3 * (time.perf_counter() - t0), dev))
    ms = float(np.median(ts))
    e_mine = {k: float(r.energy) for k, r in zip([k for k in range(total) if k % world == rank], res)}
    e_all = [None] * world
    dist.all_gather_object(e_all, e_mine)
    if rank != 0:
        return {}
    n_det = sum(len(a) * len(b) for a, b in batches)
    energies = {}
    for part in e_all:
        energies.update(part)
    return {"workload": f"configs[3] as written: {total} subspaces x ~1e5 determinants in total over {world} GPUs "
                        "(solve_sci_batch, host arrays, compute_rdms=False)",
            "ms": ms, "mdet_per_s": n_det / (ms * 1e-3) / 1e6, "subspaces_per_rank": (total + world - 1) // world,
            "energies": [energies[k] for k in range(total)]}

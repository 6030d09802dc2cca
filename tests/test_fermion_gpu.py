"""Parity of the CUDA fermion path (through the C-ABI) against the CPU oracle."""

import itertools

import numpy as np
import pytest

from oracle import fermion_oracle as fo
from qiskit_addon_sqd_b200._synthetic import (
    hf_centred_strings,
    random_integrals,
    strings_to_bitstring_matrix,
    uniform_strings,
)

pytestmark = pytest.mark.gpu

ETOL = 1e-8  # Ha, north_star tolerance for ground-state energies


def _all_strings(norb, nel):
    return np.array(sorted(sum(1 << i for i in c) for c in itertools.combinations(range(norb), nel)),
                    dtype=np.int64)


def _table_to_dense(tab):
    n = tab.n
    ptr = tab.row_ptr.cpu().numpy()
    col = tab.col.cpu().numpy()
    val = tab.val.cpu().numpy()
    H = np.diag(tab.diag.cpu().numpy())
    for i in range(n):
        for e in range(ptr[i], ptr[i + 1]):
            H[i, col[e]] = val[e]
    return H


CASES = [
    # norb, (n_alpha, n_beta), (na, nb), generator
    (4, (2, 2), (6, 6), "full"),
    (5, (2, 3), (7, 5), "hf"),
    (6, (3, 3), (20, 20), "full"),
    (8, (4, 3), (30, 21), "hf"),
    (12, (6, 6), (40, 33), "uniform"),
    (7, (1, 4), (7, 11), "hf"),
]


def _strings(norb, nel, n, kind, seed):
    if kind == "full":
        return _all_strings(norb, nel)
    if kind == "hf":
        return hf_centred_strings(norb, nel, n, seed)
    return uniform_strings(norb, nel, n, seed)


@pytest.mark.parametrize("norb,nelec,dims,kind", CASES)
def test_excitation_tables_match_slater_condon(cuda_lib, norb, nelec, dims, kind):
    from qiskit_addon_sqd_b200.fermion import _Subspace

    h, g = random_integrals(norb, 11 + norb)
    sa = _strings(norb, nelec[0], dims[0], kind, 1)
    sb = _strings(norb, nelec[1], dims[1], kind, 2)
    sub = _Subspace(sa, sb, norb, h, g)
    for tab, strs in ((sub.ta, sa), (sub.tb, sb)):
        ref = fo.same_spin_matrix(strs, h, g, norb)
        got = _table_to_dense(tab)
        assert np.abs(got - ref).max() < 1e-12
        # bit-exact structure: singles first (ascending), then doubles (ascending), signs and pq
        ptr = tab.row_ptr.cpu().numpy()
        ns = tab.n_single.cpu().numpy()
        col = tab.col.cpu().numpy()
        meta = tab.meta.cpu().numpy().view(np.uint32)
        links = {(t, s): (p, q, sg) for (t, s, p, q, sg) in
                 fo.single_excitation_links(strs, norb, include_diagonal=False)}
        assert int(ns.sum()) == len(links)
        for i in range(tab.n):
            singles = col[ptr[i]:ptr[i] + ns[i]]
            doubles = col[ptr[i] + ns[i]:ptr[i + 1]]
            assert list(singles) == sorted(singles) and list(doubles) == sorted(doubles)
            for e in range(ptr[i], ptr[i] + ns[i]):
                p, q, sg = links[(i, int(col[e]))]
                assert int(meta[e] & 0x7FFFFFFF) == p * norb + q
                assert (-1 if meta[e] >> 31 else 1) == sg
            for j in doubles:
                assert bin(int(strs[i]) ^ int(strs[j])).count("1") == 4


@pytest.mark.parametrize("norb,nelec,dims,kind", CASES)
def test_sigma_matches_dense_operator(cuda_lib, norb, nelec, dims, kind):
    import torch

    from qiskit_addon_sqd_b200.fermion import _Subspace

    h, g = random_integrals(norb, 21 + norb)
    sa = _strings(norb, nelec[0], dims[0], kind, 3)
    sb = _strings(norb, nelec[1], dims[1], kind, 4)
    na, nb = len(sa), len(sb)
    sub = _Subspace(sa, sb, norb, h, g)
    H = fo.projected_hamiltonian(sa, sb, h, g, norb)
    S2 = fo.spin_square_matrix(sa, sb, norb)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((na, nb))
    ham = sub.hamiltonian()
    hd = ham.diag.reshape(na, sub.ldc)[:, :nb].cpu().numpy()
    assert np.abs(hd.reshape(-1) - np.diag(H)).max() < 1e-12
    c = sub.upload_amplitudes(x)
    y = sub.download_amplitudes(sub.apply(ham, c))
    assert np.abs(y.reshape(-1) - H @ x.reshape(-1)).max() < 1e-11
    y2 = sub.download_amplitudes(sub.apply(sub.spin_operator(), c))
    assert np.abs(y2.reshape(-1) - S2 @ x.reshape(-1)).max() < 1e-12
    # linear spin penalty folded into the opposite-spin integrals
    pen = sub.hamiltonian(penalty_shift=0.37, penalty_ss=0.75)
    y3 = sub.download_amplitudes(sub.apply(pen, c))
    ref3 = (H + 0.37 * (S2 - 0.75 * np.eye(na * nb))) @ x.reshape(-1)
    assert np.abs(y3.reshape(-1) - ref3).max() < 1e-11
    # reproducibility: bit-identical on a second launch
    y_again = sub.download_amplitudes(sub.apply(ham, c))
    assert np.array_equal(y, y_again)
    torch.cuda.synchronize()


@pytest.mark.parametrize("norb,nelec,dims,kind", CASES)
@pytest.mark.parametrize("spin_sq", [None, 0.0])
def test_solve_fermion_matches_dense_eigh(cuda_lib, norb, nelec, dims, kind, spin_sq):
    from qiskit_addon_sqd_b200 import fermion

    h, g = random_integrals(norb, 31 + norb)
    sa = _strings(norb, nelec[0], dims[0], kind, 6)
    sb = _strings(norb, nelec[1], dims[1], kind, 7)
    if spin_sq is not None and nelec[0] != nelec[1]:
        spin_sq = 0.25 * (nelec[0] - nelec[1]) ** 2 + 0.5 * abs(nelec[0] - nelec[1])  # sz(sz+1)
    e, state, occ, s2 = fermion.solve_fermion((sa, sb), h, g, spin_sq=spin_sq, shift=0.1)
    e_ref, c_ref, occ_ref, s2_ref, _ = fo.solve_dense(sa, sb, h, g, norb, spin_sq=spin_sq, shift=0.1)
    assert abs(e - e_ref) < ETOL
    assert state.amplitudes.shape == (len(sa), len(sb))
    assert abs(np.linalg.norm(state.amplitudes) - 1.0) < 1e-10
    # eigenvector up to sign (non-degenerate synthetic spectra)
    ov = abs(np.vdot(state.amplitudes, c_ref))
    assert ov > 1 - 1e-8
    assert np.allclose(occ[0], occ_ref[0], atol=1e-6) and np.allclose(occ[1], occ_ref[1], atol=1e-6)
    assert abs(sum(occ[0]) - nelec[0]) < 1e-9 and abs(sum(occ[1]) - nelec[1]) < 1e-9
    assert abs(s2 - s2_ref) < 1e-6
    assert state.nelec == nelec and state.norb == norb


def test_quadratic_spin_penalty(cuda_lib):
    from qiskit_addon_sqd_b200 import fermion

    norb, nel = 6, 3
    h, g = random_integrals(norb, 77)
    sa = _all_strings(norb, nel)
    # target the triplet (ss = 2) in the Sz = 0 sector -> pyscf's quadratic branch.  The default start
    # vector (lowest closed-shell determinant) is a pure singlet and S^2 commutes with H, so -- as
    # pyscf's own comment says -- this branch "relies on the quality of the initial guess": pass ci0.
    ci0 = np.random.default_rng(3).standard_normal((len(sa), len(sa)))
    e, state, occ, s2 = fermion.solve_fermion((sa, sa), h, g, spin_sq=2.0, shift=0.5, ci0=ci0,
                                               max_cycle=1000)
    e_ref, c_ref, _, s2_ref, _ = fo.solve_dense(sa, sa, h, g, norb, spin_sq=2.0, shift=0.5)
    assert abs(e - e_ref) < ETOL
    assert abs(s2 - s2_ref) < 1e-6


def test_bitstring_matrix_entry_and_closed_shell_union(cuda_lib):
    from qiskit_addon_sqd_b200 import fermion

    norb, nel = 6, 3
    h, g = random_integrals(norb, 5)
    sa = hf_centred_strings(norb, nel, 9, 1)
    sb = hf_centred_strings(norb, nel, 9, 2)
    bs = strings_to_bitstring_matrix(sa, sb, norb)
    a, b = fermion.bitstring_matrix_to_ci_strs(bs, open_shell=True)
    assert np.array_equal(a, np.unique(sa)) and np.array_equal(b, np.unique(sb))
    a2, b2 = fermion.bitstring_matrix_to_ci_strs(bs, open_shell=False)
    assert np.array_equal(a2, np.union1d(sa, sb)) and a2 is b2
    e, state, occ, s2 = fermion.solve_fermion(bs, h, g, open_shell=False)
    u = np.union1d(sa, sb)
    e_ref = fo.solve_dense(u, u, h, g, norb)[0]
    assert abs(e - e_ref) < ETOL
    # notebook golden (docs/guides/select_open_closed_shell.ipynb:180,438)
    rows = ["00010010", "01001000", "00010001"]
    m = np.array([[c == "1" for c in r] for r in rows])
    ca, cb = fermion.bitstring_matrix_to_ci_strs(m, open_shell=False)
    assert list(ca) == [1, 2, 4, 8] and list(cb) == [1, 2, 4, 8]
    oa, ob = fermion.bitstring_matrix_to_ci_strs(m, open_shell=True)
    assert list(oa) == [1, 2, 8] or (list(oa), list(ob)) == ([2, 8], [1, 4])


def test_check_ci_strs_error_messages(cuda_lib):
    from qiskit_addon_sqd_b200 import fermion

    h, g = random_integrals(4, 1)
    with pytest.raises(ValueError, match=r"Spin-up CI string in index 0 has hamming weight 2, but CI "
                                         r"string in index 1 has hamming weight 1\."):
        fermion.solve_fermion((np.array([3, 4]), np.array([3, 5])), h, g)
    with pytest.raises(ValueError, match=r"Spin-down CI string in index 0 has hamming weight 2, but CI "
                                         r"string in index 2 has hamming weight 3\."):
        fermion.solve_fermion((np.array([3, 5]), np.array([3, 5, 7])), h, g)


def test_solve_sci_batch_plugin_contract(cuda_lib):
    from qiskit_addon_sqd_b200 import fermion

    norb, nel = 8, 4
    h, g = random_integrals(norb, 9)
    batches = []
    for k in range(5):
        s = hf_centred_strings(norb, nel, 20 + k, 40 + k)
        batches.append((s, s))  # symmetrize_spin: same array object for both spins (fermion.py:544)
    res = fermion.solve_sci_batch(batches, h, g, norb, (nel, nel), spin_sq=0.0)
    assert len(res) == 5
    for (s, _), r in zip(batches, res):
        e_ref, c_ref, occ_ref, _, _ = fo.solve_dense(s, s, h, g, norb, spin_sq=0.0, shift=0.2)
        assert abs(r.energy - e_ref) < ETOL
        assert r.sci_state.amplitudes.shape == (len(s), len(s))
        assert r.sci_state.amplitudes.flags["C_CONTIGUOUS"]
        assert np.allclose(r.orbital_occupancies[0], occ_ref[0], atol=1e-6)
    # same inputs twice -> identical results (reference test_fermion.py:285-342)
    res2 = fermion.solve_sci_batch(batches, h, g, norb, (nel, nel), spin_sq=0.0)
    for r, r2 in zip(res, res2):
        assert r.energy == r2.energy
        assert np.array_equal(r.sci_state.amplitudes, r2.sci_state.amplitudes)


def test_medium_subspace_against_sparse_oracle(cuda_lib):
    """(16e,30o)-like connectivity at a size the scipy oracle finishes in seconds."""
    from qiskit_addon_sqd_b200 import fermion

    norb, nel = 14, 5
    h, g = random_integrals(norb, 100)
    sa = hf_centred_strings(norb, nel, 120, 100)
    sb = hf_centred_strings(norb, nel, 97, 101)
    e, state, occ, s2 = fermion.solve_fermion((sa, sb), h, g)
    op = fo.SparseProjectedHamiltonian(sa, sb, h, g, norb)
    e_ref, c_ref = op.ground_state()
    assert abs(e - e_ref) < ETOL
    assert abs(abs(np.vdot(state.amplitudes, c_ref)) - 1) < 1e-7


def test_state_methods_and_npz_roundtrip(cuda_lib, tmp_path):
    from qiskit_addon_sqd_b200 import fermion

    norb, nel = 6, 3
    h, g = random_integrals(norb, 3)
    sa = hf_centred_strings(norb, nel, 10, 1)
    sb = hf_centred_strings(norb, 2, 8, 2)
    e, state, occ, s2 = fermion.solve_fermion((sa, sb), h, g)
    assert abs(state.spin_square() - s2) < 1e-10
    oa, ob = state.orbital_occupancies()
    assert np.allclose(oa, occ[0], atol=1e-12) and np.allclose(ob, occ[1], atol=1e-12)
    f = tmp_path / "state.npz"
    state.save(f)
    st2 = fermion.SCIState.load(f)
    assert np.array_equal(st2.amplitudes, state.amplitudes) and st2.nelec == state.nelec
    with pytest.raises(ValueError, match=r"'amplitudes' shape must be \(2, 2\) but got \(3, 2\)"):
        fermion.SCIState(np.zeros((3, 2)), np.array([1, 2]), np.array([1, 2]), 2, (1, 1))


def test_rdm1_matches_oracle(cuda_lib):
    from qiskit_addon_sqd_b200 import fermion

    norb = 8
    h, g = random_integrals(norb, 12)
    sa = hf_centred_strings(norb, 4, 30, 1)
    sb = hf_centred_strings(norb, 3, 21, 2)
    e, state, occ, s2 = fermion.solve_fermion((sa, sb), h, g)
    dma, dmb = state.rdm(rank=1, spin_summed=False)
    ra, rb = fo.rdm1s(state.amplitudes, sa, sb, norb)
    assert np.abs(dma - ra).max() < 1e-12 and np.abs(dmb - rb).max() < 1e-12
    assert np.allclose(np.diagonal(dma), occ[0], atol=1e-12) and np.allclose(np.diagonal(dmb), occ[1], atol=1e-12)
    assert abs(np.trace(dma) - 4) < 1e-10 and abs(np.trace(dmb) - 3) < 1e-10
    assert np.abs(state.rdm(rank=1, spin_summed=True) - (ra + rb)).max() < 1e-12
    with pytest.raises(NotImplementedError):
        state.rdm(rank=3)
    res = fermion.solve_sci((sa, sb), h, g, norb, (4, 3))
    assert np.abs(res.rdm1 - (ra + rb)).max() < 1e-6


@pytest.mark.parametrize("norb,nea,neb,na,nb", [(5, 2, 3, 7, 6), (6, 3, 3, 14, 14), (6, 1, 4, 5, 9), (7, 4, 2, 20, 12)])
def test_rdm2_matches_oracle(cuda_lib, norb, nea, neb, na, nb):
    """dm2aa / dm2ab / dm2bb (pyscf convention <p+ r+ s q>) against the operator-by-operator brute force,
    plus the reference's own energy formula (fermion.py:730-732) and the spin-summed form."""
    from qiskit_addon_sqd_b200 import fermion

    h, g = random_integrals(norb, 31 + norb)
    sa = hf_centred_strings(norb, nea, na, 1)
    sb = hf_centred_strings(norb, neb, nb, 2)
    e, state, occ, s2 = fermion.solve_fermion((sa, sb), h, g, open_shell=True)
    aa, ab, bb = state.rdm(rank=2, spin_summed=False)
    raa, rab, rbb = fo.rdm2s(state.amplitudes, sa, sb, norb)
    assert np.abs(aa - raa).max() < 1e-12
    assert np.abs(bb - rbb).max() < 1e-12
    assert np.abs(ab - rab).max() < 1e-12
    dm2 = state.rdm(rank=2, spin_summed=True)
    assert np.abs(dm2 - (raa + rbb + rab + rab.transpose(2, 3, 0, 1))).max() < 1e-12
    dm1 = state.rdm(rank=1, spin_summed=True)
    e_rdm = np.einsum("pr,pr->", dm1, h) + 0.5 * np.einsum("prqs,prqs->", dm2, g)
    assert abs(e_rdm - e) < 1e-10
    # traces: N(N-1) pairs
    n = nea + neb
    assert abs(np.einsum("pprr->", dm2) - n * (n - 1)) < 1e-10


def test_solve_sci_returns_rdms_like_the_reference(cuda_lib):
    """fermion.py:728-740: SCIResult carries the spin-summed rdm1 and rdm2 and the energy equals their
    contraction with the integrals."""
    from qiskit_addon_sqd_b200 import fermion

    norb = 10
    h, g = random_integrals(norb, 5)
    sa = hf_centred_strings(norb, 5, 60, 1)
    sb = hf_centred_strings(norb, 4, 45, 2)
    res = fermion.solve_sci((sa, sb), h, g, norb, (5, 4))
    assert res.rdm1.shape == (norb,) * 2 and res.rdm2.shape == (norb,) * 4
    e_rdm = np.einsum("pr,pr->", res.rdm1, h) + 0.5 * np.einsum("prqs,prqs->", res.rdm2, g)
    assert abs(e_rdm - res.energy) < 1e-9
    # symmetries of a real state: dm2[p,q,r,s] = dm2[q,p,s,r] = dm2[r,s,p,q]
    assert np.abs(res.rdm2 - res.rdm2.transpose(1, 0, 3, 2)).max() < 1e-12
    assert np.abs(res.rdm2 - res.rdm2.transpose(2, 3, 0, 1)).max() < 1e-12
    # bit-reproducible (no atomics)
    res2 = fermion.solve_sci((sa, sb), h, g, norb, (5, 4))
    assert np.array_equal(res.rdm2, res2.rdm2) and np.array_equal(res.rdm1, res2.rdm1)
    # opt-out used by throughput-oriented callers
    res3 = fermion.solve_sci((sa, sb), h, g, norb, (5, 4), compute_rdms=False)
    assert res3.rdm1 is None and res3.rdm2 is None and res3.energy == res.energy


def test_sigma_row_blocks_sum_to_full_sigma(cuda_lib):
    """The sharded build (one block of rows per rank, combined by an all-reduce) is exact: blocks written
    into zeroed buffers add up, bit for bit, to the unsharded sigma."""
    import ctypes as C

    import torch

    from qiskit_addon_sqd_b200 import _lib
    from qiskit_addon_sqd_b200.fermion import _Subspace

    norb = 10
    h, g = random_integrals(norb, 4)
    sa = hf_centred_strings(norb, 5, 70, 1)
    sb = hf_centred_strings(norb, 4, 55, 2)
    sub = _Subspace(sa, sb, norb, h, g)
    ham = sub.hamiltonian()
    x = sub.upload_amplitudes(np.random.default_rng(1).standard_normal((sub.na, sub.nb)))
    full = sub.apply(ham, x)
    total = torch.zeros_like(full)
    for lo, hi in ((0, 9), (9, 10), (10, 41), (41, 70)):
        part = torch.zeros_like(full)
        _lib.check(cuda_lib.sqd_sigma_rows(C.byref(ham.struct), _lib.ptr(x), _lib.ptr(part), lo, hi,
                                           _lib.stream_ptr(torch)))
        blk = part.reshape(sub.na, sub.ldc)
        assert float(blk[:lo].abs().max() if lo else 0.0) == 0.0 and float(blk[hi:].abs().max() if hi < sub.na else 0.0) == 0.0
        total += part
    assert torch.equal(total, full)


def test_solve_subspace_c_abi_direct(cuda_lib):
    """`sqd_solve_subspace` called the way a non-Python host would: raw device pointers in, caller-owned
    buffers out.  Checked against the dense oracle incl. the linear and quadratic spin penalties."""
    import ctypes as C

    import torch

    from qiskit_addon_sqd_b200 import _lib

    norb, nea, neb = 7, 3, 2
    h, g = random_integrals(norb, 77)
    sa = hf_centred_strings(norb, nea, 18, 1)
    sb = hf_centred_strings(norb, neb, 15, 2)
    dev = torch.device("cuda")
    d_sa = torch.from_numpy(sa.astype(np.uint64).view(np.int64)).to(dev)
    d_sb = torch.from_numpy(sb.astype(np.uint64).view(np.int64)).to(dev)
    d_h, d_g = torch.from_numpy(h).to(dev), torch.from_numpy(g).to(dev)
    na, nb = len(sa), len(sb)
    ldc = (nb + 1) // 2 * 2
    for spin_sq in (None, 0.75, 2.0):   # none / linear branch (sz(sz+1)+0.1 = 0.85) / quadratic branch
        x = torch.full((na * ldc,), float("nan"), dtype=torch.float64, device=dev)
        dm1 = torch.empty(norb**2, dtype=torch.float64, device=dev)
        dm2 = torch.empty(norb**4, dtype=torch.float64, device=dev)
        prm = _lib.SolveParams()
        prm.norb, prm.na, prm.nb, prm.n_alpha, prm.n_beta = norb, na, nb, nea, neb
        prm.d_strs_a, prm.d_strs_b = d_sa.data_ptr(), d_sb.data_ptr()
        prm.d_h, prm.d_g = d_h.data_ptr(), d_g.data_ptr()
        prm.penalty = 0 if spin_sq is None else 1
        prm.spin_sq, prm.shift = (0.0 if spin_sq is None else spin_sq), 0.1
        prm.want_spin = 1
        prm.max_space, prm.max_cycle = 12, 100
        prm.tol, prm.tol_residual, prm.lindep, prm.level_shift = 1e-12, 1e-6, 1e-14, 1e-4
        res = _lib.SolveResult()
        _lib.check(cuda_lib.sqd_solve_subspace(C.byref(prm), x.data_ptr(), dm1.data_ptr(), dm2.data_ptr(),
                                               C.byref(res), torch.cuda.current_stream().cuda_stream))
        if spin_sq is None:
            e_ref, c_ref, occ_ref, s2_ref, _ = fo.solve_dense(sa, sb, h, g, norb)
        else:
            e_ref, c_ref, occ_ref, s2_ref, _ = fo.solve_dense(sa, sb, h, g, norb, spin_sq=spin_sq, shift=0.1)
        assert res.ldc == ldc and res.info.converged in (1, 2), (spin_sq, res.info.converged, res.info.cycles,
                                                                 res.info.residual)
        assert abs(res.energy - e_ref) < ETOL
        assert res.have_spin_square == 1 and abs(res.spin_square - s2_ref) < 1e-6
        amps = x.cpu().numpy().reshape(na, ldc)
        assert np.all(amps[:, nb:] == 0.0)                       # pad columns stay zero
        assert np.abs(amps[:, :nb] - c_ref).max() < 1e-6         # same sign convention as the oracle
        assert np.allclose(np.array(res.occ_a[:norb]), occ_ref[0], atol=1e-7)
        assert np.allclose(np.array(res.occ_b[:norb]), occ_ref[1], atol=1e-7)
        r1 = dm1.cpu().numpy().reshape(norb, norb)
        r2 = dm2.cpu().numpy().reshape((norb,) * 4)
        e_rdm = np.einsum("pr,pr->", r1, h) + 0.5 * np.einsum("prqs,prqs->", r2, g)
        assert abs(e_rdm - res.energy) < 1e-9                    # fermion.py:730-732
    # error convention: <0 and a message, no exception across the C boundary
    prm.norb = 65
    assert cuda_lib.sqd_solve_subspace(C.byref(prm), x.data_ptr(), None, None, C.byref(res), None) < 0
    assert b"norb" in cuda_lib.sqd_last_error()


def test_fix_sign_and_read_back(cuda_lib):
    import ctypes as C

    import torch

    from qiskit_addon_sqd_b200 import _lib

    x = torch.tensor([0.1, -0.7, 0.7, 0.2], dtype=torch.float64, device="cuda")
    scratch = torch.empty(4096, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(cuda_lib.sqd_fix_sign(x.data_ptr(), 4, scratch.data_ptr(), st))
    assert x.cpu().tolist() == [-0.1, 0.7, -0.7, -0.2]            # first largest |x| becomes positive
    _lib.check(cuda_lib.sqd_fix_sign(x.data_ptr(), 4, scratch.data_ptr(), st))
    assert x.cpu().tolist() == [-0.1, 0.7, -0.7, -0.2]            # idempotent
    out = (C.c_double * 4)()
    _lib.check(cuda_lib.sqd_read_back(out, x.data_ptr(), 32, st))
    assert list(out) == [-0.1, 0.7, -0.7, -0.2]
    assert cuda_lib.sqd_read_back(out, x.data_ptr(), 1 << 20, st) < 0   # larger than the staging buffer


def test_solve_sci_batch_over_several_devices(cuda_lib):
    """One process driving every visible GPU (`devices="all"`): subspace k goes to device k mod G, no
    collective; results are identical to the single-device run."""
    import torch

    from qiskit_addon_sqd_b200 import fermion

    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least two GPUs")
    norb = 10
    h, g = random_integrals(norb, 9)
    batches = [(hf_centred_strings(norb, 5, 40 + 3 * k, 10 + k), hf_centred_strings(norb, 4, 35 + 2 * k, 20 + k))
               for k in range(5)]
    one = fermion.solve_sci_batch(batches, h, g, norb, (5, 4), spin_sq=0.75)
    many = fermion.solve_sci_batch(batches, h, g, norb, (5, 4), spin_sq=0.75, devices="all")
    for a, b in zip(one, many):
        assert a.energy == b.energy
        assert np.array_equal(a.sci_state.amplitudes, b.sci_state.amplitudes)
        assert np.array_equal(a.rdm2, b.rdm2)


# ---- external pin: hydrogen chains with textbook STO-3G integrals (oracle/sto3g.py) -------------------
def test_h2_sto3g_matches_szabo_ostlund(cuda_lib):
    """``solve_fermion`` on the full (1 alpha, 1 beta) space of minimal-basis H2 at R = 1.4 a0 reproduces the
    full-CI energy printed in Szabo & Ostlund (-1.1373 Ha with 1/R), and on the Hartree-Fock determinant
    alone the RHF energy (-1.1167 Ha).  Replaces the reference's pyscf pin (``test/test_fermion.py:54-125``)."""
    from oracle import sto3g
    from qiskit_addon_sqd_b200 import fermion

    h, g, en = sto3g.hydrogen_chain(2, sto3g.H2_R)
    bits = np.array([[0, 1, 0, 1], [1, 0, 1, 0], [0, 1, 1, 0], [1, 0, 0, 1]], dtype=bool)
    for spin_sq in (None, 0.0):
        e, state, occ, s2 = fermion.solve_fermion(bits, h, g, spin_sq=spin_sq)
        assert abs(e + en - sto3g.H2_E_FCI) < 5e-5
        assert state.amplitudes.shape == (2, 2) and abs(s2) < 1e-8
        assert abs(occ[0].sum() - 1.0) < 1e-9 and abs(occ[1].sum() - 1.0) < 1e-9
    S, hao, gao, _ = sto3g.hydrogen_chain_ao(2, sto3g.H2_R)
    C = np.stack([np.array([1.0, 1.0]) / np.sqrt(2 + 2 * S[0, 1]), np.array([1.0, -1.0]) / np.sqrt(2 - 2 * S[0, 1])],
                 axis=1)
    h_mo = C.T @ hao @ C
    g_mo = np.einsum("ap,bq,cr,ds,abcd->pqrs", C, C, C, C, gao)
    e_hf, *_ = fermion.solve_fermion((np.array([1]), np.array([1])), h_mo, g_mo)
    assert abs(e_hf + en - sto3g.H2_E_HF) < 5e-5


@pytest.mark.parametrize("n_atoms", [4, 6])
def test_hydrogen_chain_full_and_sampled_subspaces(cuda_lib, n_atoms):
    """H4 / H6 chains: full space == the recorded full-CI energy (cross-checked against Jordan-Wigner in
    tests/test_oracle_cpu.py); sampled subspaces == dense oracle within 1e-8 Ha, with and without the spin
    penalty; the variational bound E_sub >= E_FCI holds."""
    from oracle import sto3g
    from qiskit_addon_sqd_b200 import fermion

    h, g, en = sto3g.hydrogen_chain(n_atoms, 1.4)
    ne = n_atoms // 2
    full = _all_strings(n_atoms, ne)
    e_fci, state, occ, s2 = fermion.solve_fermion((full, full), h, g)
    expected = {4: -2.139442706994545, 6: -3.1435083836572515}[n_atoms]
    assert abs(e_fci + en - expected) < ETOL and abs(s2) < 1e-7
    rng = np.random.default_rng(n_atoms)
    for trial in range(3):
        sa = np.sort(rng.choice(full, size=max(2, len(full) * 2 // 3), replace=False))
        sb = np.sort(rng.choice(full, size=max(2, len(full) // 2), replace=False))
        for spin_sq in (None, 0.0):
            e, st, oc, ss = fermion.solve_fermion((sa, sb), h, g, spin_sq=spin_sq, shift=0.1)
            e_ref, *_ = fo.solve_dense(sa, sb, h, g, n_atoms, spin_sq=spin_sq, shift=0.1)
            assert abs(e - e_ref) < ETOL
            assert e >= e_fci - 1e-9

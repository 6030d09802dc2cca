"""Elementwise parity of ``sqd_sigma`` / ``sqd_sigma_rows`` with the C oracle at shapes that force EVERY
kernel instance the dispatcher can select.

v1 path (``fermion_sigma.cu``): ``CPT in {1, 2, 4, 8, 12}`` x ``STAGE_PACK in {T, F}``, long columns, split rows.
wide path (``sigma_wide_kernel``, no shared-memory staging: the only path beyond nb = 8192) at every shape.
v2 path (``fermion_sigma2.cu``): source-row grouped opposite-spin kernel (register-resident link lists,
``LMAX in {8, 16}``, one or several column groups) + dense same-spin tile kernel (with and without split-K).
The oracle is ``oracle/sci_cpu.c`` (one sigma build, direct excitation-table algorithm; validated against the
dense Slater-Condon oracle and the Jordan-Wigner construction in ``tests/test_oracle_cpu.py``).
"""

import ctypes as C

import numpy as np
import pytest

from oracle import sci_cpu
from qiskit_addon_sqd_b200._synthetic import hf_centred_strings, random_integrals, uniform_strings

pytestmark = pytest.mark.gpu

# (norb, n_alpha, n_beta, na, nb, kind, what it forces on the v1 path)
SHAPES = [
    (30, 15, 15, 316, 316, "hf", "CPT=1 staged table, long column, split rows (bench shape c4)"),
    (30, 8, 8, 316, 316, "hf", "CPT=1 staged table (north_star target shape t)"),
    (30, 15, 15, 48, 900, "hf", "CPT=1, SELL table > 24 KB -> STAGE_PACK=false with non-empty links"),
    (30, 15, 15, 40, 1010, "hf", "CPT=2 (993..1024 columns)"),
    (30, 15, 15, 36, 2000, "hf", "CPT=4"),
    (30, 15, 15, 24, 4000, "hf", "CPT=8"),
    (30, 15, 15, 16, 5700, "hf", "CPT=12"),
    (40, 12, 12, 64, 1000, "hf", "config-5 columns, 40 orbitals"),
    (24, 6, 6, 300, 280, "uniform", "sparse set: nearly empty link lists"),
    (12, 5, 7, 397, 211, "hf", "open shell, na != nb, odd nb (pad column)"),
]

TOL = 1e-11  # relative to max|sigma|


def _strings(norb, nel, n, kind, seed):
    return (hf_centred_strings if kind == "hf" else uniform_strings)(norb, nel, n, seed)


def _build(norb, nea, neb, na, nb, kind, sigma_path):
    from qiskit_addon_sqd_b200.fermion import _Subspace

    h, g = random_integrals(norb, 900 + norb)
    sa = _strings(norb, nea, na, kind, 11)
    sb = _strings(norb, neb, nb, kind, 12)
    sub = _Subspace(sa, sb, norb, h, g, sigma_path=sigma_path)
    return sub, sa, sb, h, g


# beyond the staged kernels (v1: nb <= 5760, v2: nb <= 8192) only the staging-free wide kernel applies: every
# path request must end up there
WIDE_ONLY = (20, 5, 5, 12, 9001, "hf", "nb > 8192: wide kernel, odd nb (pad column)")


@pytest.mark.parametrize("sigma_path", ["v1", "v2", "wide"])
@pytest.mark.parametrize("norb,nea,neb,na,nb,kind,why", SHAPES + [WIDE_ONLY])
def test_sigma_elementwise_vs_c_oracle(cuda_lib, norb, nea, neb, na, nb, kind, why, sigma_path):
    import torch

    from qiskit_addon_sqd_b200 import _lib

    sub, sa, sb, h, g = _build(norb, nea, neb, na, nb, kind, sigma_path)
    ham = sub.hamiltonian()
    if sigma_path == "v2" and not ham.uses_v2 and nb <= 8192:
        pytest.skip("v2 path not selected for this shape (sparse set: v1 is the product path)")
    assert ham.uses_wide == (sigma_path == "wide" or nb > 8192)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((sub.na, sub.nb))
    c = sub.upload_amplitudes(x)
    y = sub.download_amplitudes(sub.apply(ham, c))
    ref = sci_cpu.sigma(sa, sb, h, g, x)
    scale = np.abs(ref).max()
    assert np.abs(y - ref).max() <= TOL * scale, why
    # pads of sigma are written as zero
    full = sub.apply(ham, c).reshape(sub.na, sub.ldc)
    if sub.ldc > sub.nb:
        assert float(full[:, sub.nb:].abs().max()) == 0.0
    # bit-reproducible
    assert np.array_equal(y, sub.download_amplitudes(sub.apply(ham, c)))
    # row blocks (sharded builds) add up to the full vector bit for bit, rows outside stay untouched
    total = torch.zeros(sub.na * sub.ldc, dtype=torch.float64, device=c.device)
    cuts = sorted({0, 1, sub.na // 3, sub.na // 2 + 1, sub.na})
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        part = torch.zeros_like(total)
        _lib.check(cuda_lib.sqd_sigma_rows(C.byref(ham.struct), _lib.ptr(c), _lib.ptr(part), lo, hi,
                                           _lib.stream_ptr(torch)))
        blk = part.reshape(sub.na, sub.ldc)
        assert lo == 0 or float(blk[:lo].abs().max()) == 0.0
        assert hi == sub.na or float(blk[hi:].abs().max()) == 0.0
        total += part
    assert torch.equal(total.reshape(sub.na, sub.ldc), full)


@pytest.mark.parametrize("sigma_path", ["v1", "v2", "wide"])
def test_spin_operator_and_penalty_at_bench_shape(cuda_lib, sigma_path):
    """S^2 (opposite-spin tensor only, no same-spin part) and the linear spin penalty at 316 x 316."""
    from oracle import fermion_oracle as fo

    norb, ne, n = 14, 5, 150
    sub, sa, sb, h, g = _build(norb, ne, ne, n, n - 7, "hf", sigma_path)
    rng = np.random.default_rng(8)
    x = rng.standard_normal((sub.na, sub.nb))
    c = sub.upload_amplitudes(x)
    S2 = fo.spin_square_matrix(sa, sb, norb)
    y2 = sub.download_amplitudes(sub.apply(sub.spin_operator(), c))
    ref2 = (S2 @ x.reshape(-1)).reshape(x.shape)
    assert np.abs(y2 - ref2).max() < 1e-11 * max(1.0, np.abs(ref2).max())
    pen = sub.hamiltonian(penalty_shift=0.37, penalty_ss=0.75)
    y3 = sub.download_amplitudes(sub.apply(pen, c))
    ref3 = sci_cpu.sigma(sa, sb, h, g, x) + 0.37 * (ref2 - 0.75 * x)
    assert np.abs(y3 - ref3).max() < 1e-11 * np.abs(ref3).max()

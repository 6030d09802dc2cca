"""ctypes binding of ``libsqd_b200.so`` (C-ABI declared in ``include/sqd_b200.h``).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``build.sh``.  There is NO CPU
fallback: if the library is missing, or no CUDA device is visible, every product entry point raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsqd_b200.so")

MAX_SPACE = 32
MAX_LONG_COLUMNS = 64


class SpinTable(C.Structure):
    _fields_ = [
        ("n", C.c_int),
        ("strs", C.c_void_p),
        ("row_ptr", C.c_void_p),
        ("n_single", C.c_void_p),
        ("col", C.c_void_p),
        ("val", C.c_void_p),
        ("meta", C.c_void_p),
        ("pack", C.c_void_p),
    ]


class SigmaPlan(C.Structure):
    _fields_ = [
        ("n_chunks", C.c_int),
        ("n_slots", C.c_int),
        ("n_split", C.c_int),
        ("n_long", C.c_int),
        ("chunk_row", C.c_void_p),
        ("chunk_beg", C.c_void_p),
        ("chunk_end", C.c_void_p),
        ("chunk_slot", C.c_void_p),
        ("split_row", C.c_void_p),
        ("split_slot_beg", C.c_void_p),
        ("split_n", C.c_void_p),
        ("long_idx", C.c_void_p),
        ("long_cols", C.c_void_p),
        ("part", C.c_void_p),
    ]


class Sell(C.Structure):
    _fields_ = [
        ("n_slices", C.c_int),
        ("n_entries", C.c_int),
        ("perm", C.c_void_p),
        ("len", C.c_void_p),
        ("slice_ptr", C.c_void_p),
        ("pack", C.c_void_p),
        ("val", C.c_void_p),
    ]


class SigmaV2(C.Structure):
    """``sqd_sigma_v2`` of include/sqd_b200.h."""

    _fields_ = [
        ("enabled", C.c_int), ("lmax", C.c_int), ("n_groups", C.c_int), ("vc_pad", C.c_int),
        ("n_items", C.c_int), ("n_chunks", C.c_int), ("max_split", C.c_int),
        ("lda", C.c_int), ("ldb", C.c_int), ("ldp", C.c_int), ("ldq", C.c_int),
        ("vc_src", C.c_void_p), ("vc_off", C.c_void_p), ("vc_len", C.c_void_p), ("vc_q", C.c_void_p),
        ("col_seg", C.c_void_p), ("single_ptr", C.c_void_p), ("item_ptr", C.c_void_p),
        ("chunk_rec", C.c_void_p), ("item_tgt", C.c_void_p), ("item_gsel", C.c_void_p),
        ("item_pslot", C.c_void_p), ("heavy_rows", C.c_void_p), ("n_heavy", C.c_int),
        ("HaDT", C.c_void_p), ("HbDT", C.c_void_p), ("P", C.c_void_p), ("part", C.c_void_p),
    ]


V2_COUNTS = 16


class Operator(C.Structure):
    _fields_ = [
        ("a", SpinTable),
        ("b", SpinTable),
        ("norb", C.c_int),
        ("ldc", C.c_int),
        ("ldg", C.c_int),
        ("diag", C.c_void_p),
        ("gab", C.c_void_p),
        ("Wa", C.c_void_p),
        ("Wb", C.c_void_p),
        ("use_same_spin", C.c_int),
        ("plan", SigmaPlan),
        ("bd", Sell),
        ("bb", Sell),
        ("throughput_mode", C.c_int),
        ("v2", SigmaV2),
        ("wide", C.c_int),
    ]


class DavidsonParams(C.Structure):
    _fields_ = [
        ("max_space", C.c_int),
        ("max_cycle", C.c_int),
        ("tol", C.c_double),
        ("tol_residual", C.c_double),
        ("lindep", C.c_double),
        ("level_shift", C.c_double),
        ("check_every", C.c_int),
        ("ss_op", C.POINTER(Operator)),
        ("ss_shift", C.c_double),
        ("ss_value", C.c_double),
        ("profile", C.c_int),
        ("nccl_comm", C.c_void_p),
        ("row_begin", C.c_int),
        ("row_end", C.c_int),
        ("shard_bounds", C.c_void_p),
        ("shard_world", C.c_int),
        ("single_stream_ritz", C.c_int),
    ]


class DavidsonInfo(C.Structure):
    _fields_ = [
        ("converged", C.c_int),
        ("cycles", C.c_int),
        ("sigma_builds", C.c_int),
        ("theta", C.c_double),
        ("residual", C.c_double),
        ("sigma_ms", C.c_double),
        ("total_ms", C.c_double),
    ]


class SolveParams(C.Structure):
    """``sqd_solve_params`` of include/sqd_b200.h."""

    _fields_ = [
        ("norb", C.c_int), ("na", C.c_int), ("nb", C.c_int),
        ("n_alpha", C.c_int), ("n_beta", C.c_int),
        ("d_strs_a", C.c_void_p), ("d_strs_b", C.c_void_p),
        ("d_h", C.c_void_p), ("d_g", C.c_void_p),
        ("penalty", C.c_int), ("spin_sq", C.c_double), ("shift", C.c_double),
        ("want_spin", C.c_int),
        ("max_space", C.c_int), ("max_cycle", C.c_int),
        ("tol", C.c_double), ("tol_residual", C.c_double), ("lindep", C.c_double),
        ("level_shift", C.c_double),
        ("check_every", C.c_int),
        ("d_ci0", C.c_void_p),
        ("cost_per_chunk", C.c_int), ("long_threshold", C.c_int),
        ("profile", C.c_int),
        ("nccl_comm", C.c_void_p),
        ("row_begin", C.c_int), ("row_end", C.c_int),
        ("shard_rank", C.c_int), ("shard_world", C.c_int),
        ("throughput_mode", C.c_int),
        ("sigma_path", C.c_int), ("v2_lmax", C.c_int), ("v2_items_per_chunk", C.c_int),
    ]


class SolveResult(C.Structure):
    """``sqd_solve_result`` of include/sqd_b200.h."""

    _fields_ = [
        ("energy", C.c_double), ("spin_square", C.c_double), ("have_spin_square", C.c_int),
        ("occ_a", C.c_double * 64), ("occ_b", C.c_double * 64),
        ("info", DavidsonInfo),
        ("nnz_a", C.c_int64), ("nnz_b", C.c_int64),
        ("singles_a", C.c_int64), ("singles_b", C.c_int64),
        ("ldc", C.c_int),
        ("sigma_path", C.c_int),
    ]


_vp, _i, _i64, _u64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
_pi = C.POINTER(C.c_int)

# name -> (restype, argtypes).  Must list every function declared in include/sqd_b200.h.
SIGNATURES: dict[str, tuple] = {
    "sqd_version": (_i, []),
    "sqd_last_error": (C.c_char_p, []),
    "sqd_launch_count": (C.c_longlong, [_i]),
    "sqd_stream_wait": (_i, [_vp]),
    "sqd_download": (_i, [_vp, _vp, C.c_longlong, _vp]),
    "sqd_pack_bitstrings": (_i, [_vp, _i64, _i, _vp, _vp, _vp]),
    "sqd_excitation_count": (_i, [_vp, _i, _vp, _vp, _vp]),
    "sqd_exclusive_scan": (_i, [_vp, _vp, _i, _pi, _vp]),
    "sqd_excitation_fill": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sqd_make_gab": (_i, [_vp, _i, _d, _i, _vp, _i, _vp]),
    "sqd_opposite_spin_tables": (
        _i,
        [_vp, _i, _vp, _i, _i, _vp, _i, _vp, _vp, _d, _d, _vp, _vp, _vp, _i, _vp],
    ),
    "sqd_sigma_v1_supported": (_i, [_i, _i]),
    "sqd_sigma_smem_bytes": (_i64, [C.POINTER(Operator)]),
    "sqd_sigma": (_i, [C.POINTER(Operator), _vp, _vp, _vp]),
    "sqd_sigma_rows": (_i, [C.POINTER(Operator), _vp, _vp, _i, _i, _vp]),
    "sqd_nccl_unique_id": (_i, [C.c_char_p]),
    "sqd_nccl_init": (_i, [C.c_char_p, _i, _i, C.POINTER(C.c_void_p)]),
    "sqd_nccl_destroy": (_i, [_vp]),
    "sqd_allreduce_sum_f64": (_i, [_vp, _vp, _i64, _vp]),
    "sqd_sigma_profile": (_i, [C.POINTER(Operator), _vp, _vp, _vp, _vp]),
    "sqd_sell_build": (_i, [C.POINTER(SpinTable), _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sqd_sigma_plan_build": (
        _i,
        [C.POINTER(SpinTable), C.POINTER(SpinTable), _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
         _vp, _vp, _vp, _pi, _vp],
    ),
    "sqd_sigma_v2_recommended": (_i, [_i, _i, _i64, _i64]),
    "sqd_sigma_v2_plan_bytes": (_i64, [_i, _i, _i64, _i64, _i, _i]),
    "sqd_sigma_v2_plan": (
        _i, [C.POINTER(SpinTable), C.POINTER(SpinTable), _i, _i64, _i64, _i, _i, _vp, _i64, _pi, _vp]
    ),
    "sqd_sigma_v2_counts_ptr": (_vp, [_vp, _i, _i, _i64, _i64, _i, _i]),
    "sqd_sigma_v2_scratch_bytes": (_i64, [_pi, _i, _i, _i, _i, _i]),
    "sqd_sigma_v2_finish": (
        _i,
        [C.POINTER(SpinTable), C.POINTER(SpinTable), _i, _i64, _i64, _i, _i, _pi, _vp, _vp, _i64, _i,
         C.POINTER(SigmaV2), _vp],
    ),
    "sqd_davidson_workspace_bytes": (_i64, [_i, _i, _i]),
    "sqd_davidson": (
        _i,
        [C.POINTER(Operator), _vp, _vp, _vp, _vp, _i64, C.POINTER(DavidsonParams),
         C.POINTER(DavidsonInfo), _vp],
    ),
    "sqd_solve_subspace": (
        _i, [C.POINTER(SolveParams), _vp, _vp, _vp, C.POINTER(SolveResult), _vp]
    ),
    "sqd_init_guess": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "sqd_dot": (_i, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "sqd_occupancies": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "sqd_fix_sign": (_i, [_vp, _i64, _vp, _vp]),
    "sqd_read_back": (_i, [_vp, _vp, _i64, _vp]),
    "sqd_rdm1s_workspace_bytes": (_i64, [C.POINTER(Operator)]),
    "sqd_rdm1s": (_i, [C.POINTER(Operator), _vp, _i64, _i64, _vp, _vp, _vp, _vp]),
    "sqd_rdm2s_workspace_bytes": (_i64, [C.POINTER(Operator), _i64, _i64]),
    "sqd_rdm2s": (_i, [C.POINTER(Operator), _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "sqd_bits_to_keys": (_i, [_vp, _i64, _i, _vp, _vp]),
    "sqd_key_table_capacity": (_i64, [_i64]),
    "sqd_key_table_bytes": (_i64, [_i64]),
    "sqd_key_table_build": (_i, [_vp, _i64, _vp, _i64, _vp]),
    "sqd_pauli_connect": (_i, [_vp, _i64, _vp, _u64, _u64, _vp, _vp, _vp]),
    "sqd_pauli_elements": (_i, [_vp, _i64, _vp, _u64, _u64, _i, _vp, _vp, _vp, _vp]),
    "sqd_pauli_diag_group": (_i, [_vp, _i64, _vp, _vp, _vp, C.c_int32, C.c_int32, _vp, _vp]),
    "sqd_pauli_project_count": (
        _i, [_vp, _i64, _vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, C.c_int32, _vp, _vp, _vp]
    ),
    "sqd_pauli_project_fill": (
        _i, [_vp, _i64, _vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    ),
    "sqd_csr_matvec_c128": (_i, [_i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sqd_csr_davidson_workspace_bytes": (_i64, [_i64, _i, _i]),
    "sqd_csr_gershgorin": (_i, [_i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sqd_csr_components_workspace_bytes": (_i64, [_i64]),
    "sqd_csr_components": (_i, [_i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp]),
    "sqd_csr_component_start": (_i, [_i64, _vp, _i, _i, _d, _vp, _vp]),
    "sqd_csr_davidson": (
        _i, [_i64, _vp, _vp, _vp, _i, _i, _i, _d, _vp, _vp, _vp, _pi, _vp, _vp, _i64, _vp]
    ),
    "sqd_merge_rows_workspace_bytes": (_i64, [_i64]),
    "sqd_merge_rows": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "sqd_sort_unique_workspace_bytes": (_i64, [_i64]),
    "sqd_sort_unique": (_i, [_vp, _vp, _i64, _vp, _vp, _vp, C.POINTER(C.c_int64), _vp, _i64, _vp]),
    "sqd_bit_array_pack": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp]),
    "sqd_keys_to_bits": (_i, [_vp, _vp, _i64, _i, _vp, _vp]),
    "sqd_carryover": (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp]),
    "sqd_recover_workspace_bytes": (_i64, [_i64, _i]),
    "sqd_recover": (
        _i,
        [_vp, _vp, _i64, _i, _vp, _vp, _i, _i, _i, _vp, _u64, _vp, _vp, _vp, _vp, _i64, _vp],
    ),
}

_lib = None
_lock = threading.Lock()


class SqdCudaError(RuntimeError):
    """An ``sqd_*`` C-ABI call returned a non-zero status."""


def load():
    """Load the CUDA library (no compute call is made).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a).  qiskit_addon_sqd_b200 has no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def require_cuda():
    """Import torch and insist on a CUDA device: the product path never runs on the CPU."""
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError(
            "qiskit_addon_sqd_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists."
        )
    return torch


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().sqd_last_error()
        raise SqdCudaError(f"{what or 'sqd call'} failed ({status}): {msg.decode() if msg else ''}")


def ptr(t) -> int:
    """Device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr(torch) -> int:
    return torch.cuda.current_stream().cuda_stream


def read_back(torch, t):
    """Small device tensor -> numpy array through the library's per-thread pinned staging buffer
    (``sqd_read_back``); synchronises the current stream.  Unlike ``tensor.cpu()`` / ``.item()`` it never
    copies into pageable memory, which would stall the launches of the other solver threads."""
    import numpy as np

    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    buf = (C.c_char * max(nbytes, 1))()
    check(load().sqd_read_back(C.addressof(buf), ptr(t), nbytes, stream_ptr(torch)), "sqd_read_back")
    dt = {torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64}[t.dtype]
    return np.frombuffer(buf, dtype=dt, count=t.numel()).copy()




def download(torch, t):
    """Large device tensor -> fresh numpy array.  The copy is made by the library (``sqd_download``: one
    stream-ordered D2H into the array's own memory), i.e. outside the interpreter lock: the K solver threads of a
    batch fetch their results concurrently.  (Round 1 staged through a pinned buffer and copied with numpy, which
    serialised 58 MB per bench step on the lock; pinning the result arrays themselves was measured slower.)"""
    import numpy as np

    t = t.contiguous()
    out = np.empty(tuple(t.shape), dtype=_NP_DTYPES[str(t.dtype)])
    check(load().sqd_download(out.ctypes.data, t.data_ptr(), out.nbytes, stream_ptr(torch)), "sqd_download")
    return out


_NP_DTYPES = {"torch.float64": "float64", "torch.int64": "int64", "torch.int32": "int32", "torch.uint8": "uint8",
              "torch.float32": "float32", "torch.bool": "bool", "torch.int16": "int16", "torch.int8": "int8"}


def sort_unique(torch, hi, lo, with_counts: bool = False):
    """Distinct keys of ``(hi, lo)`` (int64 device tensors holding uint64 words; ``hi=None``: 64-bit keys) in
    ascending unsigned order, optionally with multiplicities (``sqd_sort_unique``: bitonic sort + run heads).
    Returns ``(hi_unique | None, lo_unique, counts | None)`` as device tensors of the distinct length."""
    lib = load()
    n = int(lo.numel())
    dev = lo.device
    out_lo = torch.empty(n, dtype=torch.int64, device=dev)
    out_hi = torch.empty(n, dtype=torch.int64, device=dev) if hi is not None else None
    cnt = torch.empty(n, dtype=torch.int32, device=dev) if with_counts else None
    ws_bytes = int(lib.sqd_sort_unique_workspace_bytes(n))
    if ws_bytes < 0:
        raise ValueError(f"sort_unique: {n} keys are beyond the 2^30 this library sorts")
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    n_unique = C.c_int64(0)
    check(lib.sqd_sort_unique(ptr(hi), ptr(lo), n, ptr(out_hi), ptr(out_lo), ptr(cnt), C.byref(n_unique),
                              ptr(ws), ws_bytes, stream_ptr(torch)), "sqd_sort_unique")
    nu = int(n_unique.value)
    return (out_hi[:nu] if out_hi is not None else None, out_lo[:nu], cnt[:nu] if cnt is not None else None)

#!/usr/bin/env python
"""Benchmark of the SQD subspace-diagonalisation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|t|c2|c5]

Metric (BASELINE.json): subspace-diagonalisation throughput -- determinants of subspace dimension
solved to convergence per second (wall time of one diagonalisation = ms_per_step / batches) -- with the
CI sigma-build's achieved HBM GB/s reported under "roofline".

A "step" is one ``solve_sci_batch`` over ``--batches`` independent subspaces per GPU (one SQD iteration's
worth of diagonalisations; BASELINE.json config "synthetic (30e,30o) random FP64 eri, batches x 1e5
dets, no NCCL").  Weak scaling: every rank solves its own ``--batches`` subspaces, no collective on the
data path; the only collectives are the barrier and the max-reduction of the timing.

  value : inputs (integrals, determinant strings) resident in HBM before the timed region; results
          stay on the device except the per-subspace energy scalar.
  e2e   : the same steps through the public API ``fermion.solve_sci_batch`` with HOST numpy inputs and
          outputs (H2D of hcore/eri/strings and D2H of amplitudes/occupancies inside the timed region).
  --impl reference : the CPU oracle port (oracle/sci_cpu.c, direct excitation-table algorithm, all
          host threads) on a bounded sample of the same workload; rank 0 only.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from qiskit_addon_sqd_b200._synthetic import hf_centred_strings, random_integrals  # noqa: E402

WORKLOADS = {
    # name: (norb, n_alpha, n_beta, na, nb, integral seed)
    "c4": (30, 15, 15, 316, 316, 104),   # BASELINE.json configs[3]: (30e,30o), 1e5 dets per batch
    "t": (30, 8, 8, 316, 316, 100),      # north_star target: (16e,30o), 1e5 dets
    "c2": (16, 5, 5, 100, 100, 102),     # configs[1]: (10e,16o), 1e4 dets per batch
    "c5": (40, 12, 12, 1000, 1000, 105), # configs[4]: (24e,40o), 1e6 dets
    "c1": (6, 3, 3, 20, 20, 101),        # configs[0]: (6e,6o) full space
    "s7": (30, 15, 15, 3162, 3162, 107), # scale-up point of SURVEY 8(d): 1e7 determinants, vectors >> L2
}
METRIC = "subspace_diag_throughput"
UNIT = "Mdet/s"


def make_batches(workload: str, rank: int, k_batches: int):
    norb, nea, neb, na, nb, seed = WORKLOADS[workload]
    h, g = random_integrals(norb, seed)
    import math

    na = min(na, math.comb(norb, nea))
    nb = min(nb, math.comb(norb, neb))
    batches = []
    for k in range(k_batches):
        s0 = 1000 * seed + 2 * (rank * k_batches + k)
        batches.append((hf_centred_strings(norb, nea, na, s0), hf_centred_strings(norb, neb, nb, s0 + 1)))
    return norb, (nea, neb), h, g, batches


# ------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# reference arm: CPU oracle port
# ------------------------------------------------------------------------------------------------
def run_reference(args, rank: int, world: int):
    if rank != 0:
        return
    from oracle import sci_cpu

    cores = os.cpu_count() or 1
    norb, nelec, h, g, batches = make_batches(args.workload, 0, max(1, args.ref_sample))
    n_det = sum(len(a) * len(b) for a, b in batches)

    def step():
        out = []
        for sa, sb in batches:
            e, amps, occ, info = sci_cpu.solve(sa, sb, h, g, algo=0, tol=1e-12, max_cycle=100,
                                               max_space=12, nthreads=cores)
            out.append((e, info["cycles"]))
        return out

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = step()
    dt = time.perf_counter() - t0
    value = n_det * args.steps / dt / 1e6
    sample = (f"{len(batches)} of {args.batches} subspaces per step, oracle/sci_cpu.c direct "
              f"excitation-table algorithm, Davidson tol 1e-12, {res[0][1]} cycles")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world: int) -> dict:
    norb, nea, neb, na, nb, seed = WORKLOADS[args.workload]
    return {
        "workload": f"{args.workload}: ({nea + neb}e,{norb}o) synthetic FP64 hcore/eri, "
                    f"{args.batches} subspaces x {na}x{nb}={na * nb} dets per GPU per step, HF-centred strings"
                    + (" [BASELINE.json configs[3], the per-GPU share of the 1->8 GPU scaling config; chosen "
                       "over configs[1] (5 x 1e4 dets) because its 1e5-determinant subspaces are the size the "
                       "north_star target is stated on and it is the configuration the multi-GPU numbers are "
                       "quoted on]" if args.workload == "c4" else ""),
        "batches_per_gpu": args.batches, "na": na, "nb": nb, "norb": norb, "nelec": [nea, neb],
        "davidson": {"tol": 1e-12, "max_cycle": 100,
                     "ours": "max_space 6, a restart keeps the Ritz vector and the previous one (locally optimal)",
                     "reference_arm": "pyscf's defaults: max_space 12, collapse onto the Ritz vector",
                     "note": "same convergence test on both arms; iteration counts agree within a few per cent"},
        "l2": "no explicit flush: the %d concurrent subspaces of a step hold ~%d MB of Davidson vectors, "
              "integrals and tables (> 126 MB L2), and every step rebuilds its tables and vectors from the "
              "resident inputs" % (args.batches, 35 * args.batches),
        "parallelism": f"{world} x independent ranks, no data-path collective",
    }


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def sigma_algorithmic_bytes(st) -> float:
    """B_sigma = 16 n_det + 8 norb^4 + 12 (nnz_a + nnz_b) + 8 (singles_a + singles_b)."""
    return 16.0 * st.n_det + 8.0 * st.norb**4 + 12.0 * (st.nnz_a + st.nnz_b) + 8.0 * (st.singles_a + st.singles_b)


def run_ours(args, rank: int, world: int, local_rank: int):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    from concurrent.futures import ThreadPoolExecutor

    from qiskit_addon_sqd_b200 import _lib, fermion
    from qiskit_addon_sqd_b200._dispatch import max_over_ranks

    lib = _lib.load()
    dev = torch.device("cuda", local_rank)
    norb, nelec, h, g, batches = make_batches(args.workload, rank, args.batches)
    K = len(batches)
    n_det_step = sum(len(a) * len(b) for a, b in batches)
    opts = fermion._solver_options({})

    # ---- resident inputs for the device-timed arm ----
    ints = fermion._DeviceIntegrals(torch, h, g, dev)
    strs_dev = [(torch.from_numpy(a.astype(np.uint64).view(np.int64)).to(dev),
                 torch.from_numpy(b.astype(np.uint64).view(np.int64)).to(dev)) for a, b in batches]
    # the same streams and host threads as the library's own batch driver: the stream-ordered memory pool
    # and the per-thread pinned staging buffers are then shared by the two arms
    streams = [fermion._stream_for(torch, local_rank, k) for k in range(K)]
    pool = fermion._worker_pool()

    def device_step(profile=False):
        main = torch.cuda.current_stream()

        def work(k):
            with torch.cuda.device(dev), torch.cuda.stream(streams[k]):
                streams[k].wait_stream(main)
                r = fermion._solve_on_device(batches[k][0], batches[k][1], norb, ints, None, 0.2, opts,
                                             want_spin=False, want_rdm=False, strs_dev=strs_dev[k],
                                             download=False, profile=profile,
                                             throughput=fermion._throughput_mode(K, n_det_step))
                return r

        res = list(pool.map(work, range(K)))
        for s in streams:
            main.wait_stream(s)
        return res

    # host inputs of the end-to-end arm live in pinned memory (numpy views of pinned tensors)
    h_pin = torch.from_numpy(h).pin_memory().numpy()
    g_pin = torch.from_numpy(g).pin_memory().numpy()
    batches_pin = [(torch.from_numpy(a.copy()).pin_memory().numpy(), torch.from_numpy(b.copy()).pin_memory().numpy())
                   for a, b in batches]

    def e2e_step():
        # the plugin call exactly as the SQD loop makes it: defaults, i.e. SCIResult.rdm1/rdm2 included
        return fermion.solve_sci_batch(batches_pin, h_pin, g_pin, norb, nelec)

    def e2e_loop_step():
        # what the loop consumes (energy, occupancies, amplitudes): RDMs skipped
        return fermion.solve_sci_batch(batches_pin, h_pin, g_pin, norb, nelec, compute_rdms=False)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        # the host drives K streams from K threads; the device-event span and the wall clock agree to
        # within launch latency -- report the larger so nothing is hidden
        ms = max(ms, wall * 1e3)
        return max_over_ranks(ms, dev), out

    # ---- warm-up (both arms) ----
    for _ in range(args.warmup):
        device_step()
    for _ in range(args.warmup):
        e2e_step()
        e2e_loop_step()

    sampler = ClockSampler(local_rank)
    if rank == 0 and not args.no_clock_sampler:
        sampler.start()
    lib.sqd_launch_count(1)
    ms_dev, res_dev = timed(device_step, args.steps)
    launches = int(lib.sqd_launch_count(1))
    ms_e2e, res_e2e = timed(e2e_step, args.steps)
    ms_e2e_loop, _ = timed(e2e_loop_step, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel: sigma build timed inside a real Davidson loop ----
    stats_prof = []
    if rank == 0:
        for k in range(min(K, 2)):
            best = None
            for _ in range(3):   # (the first profiled solve after the timed region sometimes runs into host noise)
                with torch.cuda.stream(streams[0]):
                    r = fermion._solve_on_device(batches[k][0], batches[k][1], norb, ints, None, 0.2, opts,
                                                 want_spin=False, want_rdm=False, strs_dev=strs_dev[k],
                                                 download=False, profile=True)
                    streams[0].synchronize()
                if best is None or r["stats"].davidson_ms < best.davidson_ms:
                    best = r["stats"]
            stats_prof.append(best)
    if dist is not None:
        dist.barrier()

    # ---- secondary measurements (bench_extras.py) ----
    extra = {}
    if args.extras != "none":
        import bench_extras as bx

        me = sys.modules[__name__]
        if world > 1:
            # all ranks take part: configs[4] sharded over the ranks, and the literal configs[3]
            extra["c5_sharded"] = bx.sharded_extras(me, torch, dist, fermion, rank, world, dev)
            extra["c4_strong"] = bx.strong_extras(me, torch, dist, fermion, rank, world, dev)
            dist.barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peaks = {}
    pk_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_file):
        peaks = json.load(open(pk_file))
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
    if args.extras != "none" and world == 1:
        with torch.cuda.stream(streams[0]):
            wls = ("c4", "t", "c5") if args.extras == "all" else ("c4", "t")
            extra["fermion"] = bx.fermion_extras(me, torch, fermion, dev, peak_gbs, wls)
            extra["t"] = extra["fermion"]["t"]
            if args.extras == "all":
                cpu = not args.no_cpu_baseline
                extra["qubit_c3"] = bx.qubit_extras(torch, dev, peak_gbs, cpu)
                extra["pauli_z40"] = bx.pauli_z40_extras(torch, dev, args.z40_rows)
                extra["recovery"] = bx.recovery_extras(torch, dev, cpu)
                extra["formats"] = bx.formats_extras(torch, fermion, dev, cpu)
                extra["sqd_loop_c2"] = bx.sqd_loop_extras(torch, fermion)
                extra["s8"] = bx.s8_extras(torch, fermion, dev, peak_gbs)
    sig_ms = np.mean([s.sigma_ms / max(s.cycles, 1) for s in stats_prof])
    sig_bytes = np.mean([sigma_algorithmic_bytes(s) for s in stats_prof])
    v2 = all(s.sigma_path == 2 for s in stats_prof)
    achieved = sig_bytes / (sig_ms * 1e-3) / 1e9
    # applied matrix elements per sigma build: opposite-spin (links+diagonal of both spins), same-spin doubles
    # and singles of each spin against every string of the other, operator diagonal
    applied = np.mean([(s.singles_a + s.na) * (s.singles_b + s.nb) + (s.nnz_a - s.singles_a) * s.nb
                       + (s.nnz_b - s.singles_b) * s.na + s.singles_a * s.nb + s.singles_b * s.na + s.n_det
                       for s in stats_prof])
    dav_ms = np.mean([s.davidson_ms for s in stats_prof])
    sig_share = np.mean([s.sigma_ms / max(s.davidson_ms, 1e-9) for s in stats_prof])
    cycles = [s.cycles for s in [r["stats"] for r in res_dev]]

    total_dets = n_det_step * args.steps * world
    value = total_dets / (ms_dev * 1e-3) / 1e6
    e2e_value = total_dets / (ms_e2e * 1e-3) / 1e6
    h2d = ints.h2d_bytes + sum(a.nbytes + b.nbytes for a, b in batches)
    d2h_loop = sum(len(a) * len(b) * 8 + 2 * norb * 8 + 64 for a, b in batches)
    d2h = d2h_loop + len(batches) * 8 * (norb**2 + norb**4)  # + spin-summed rdm1 and rdm2 per subspace

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "call": "fermion.solve_sci_batch(host arrays) with defaults: energy, amplitudes, occupancies, "
                        "rdm1 and rdm2 per subspace, as the reference's solve_sci returns them"},
        "e2e_loop_only": {"value": total_dets / (ms_e2e_loop * 1e-3) / 1e6, "unit": UNIT,
                          "ms_per_step": ms_e2e_loop / args.steps, "h2d_bytes_per_step": int(h2d),
                          "d2h_bytes_per_step": int(d2h_loop),
                          "call": "same with compute_rdms=False (the SQD loop never reads the RDMs)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {
            "kernel": ("sigma2_ab_kernel + sigma2_tile_kernel + sigma2_epilogue_kernel (one CI sigma-vector "
                       "build = these three launches, timed together inside the Davidson loop)"
                       if v2 else "sigma_b_kernel + sigma_a_kernel + sigma_combine_kernel (one sigma build)"),
            "bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
            "frac": achieved / peak_gbs, "peak_source": peak_src, "traffic": ncu_traffic(v2),
            "bytes_per_launch": sig_bytes, "ms_per_launch": sig_ms,
            "gflops": 2.0 * applied / (sig_ms * 1e-3) / 1e9, "matrix_elements_per_launch": applied,
            "dram_gbs_from_ncu_traffic": (ncu_traffic(v2) or 0.0) / (sig_ms * 1e-3) / 1e9,
            "share_of_davidson_loop": sig_share, "davidson_loop_ms": dav_ms,
            "note": "algorithmic bytes = 16 n_det + 8 norb^4 + 12 nnz + 8 links (SURVEY 8d); the working "
                    "set is L2-resident, the kernels are bound by shared-memory gathers, FP64 FMA issue and "
                    "latency, not by HBM (DESIGN.md 4); `traffic` is the COLD-cache DRAM traffic of the three "
                    "kernels (ncu flushes L2 before every kernel): 26 MB of it is the epilogue re-reading the "
                    "partial-product array the first kernel wrote, which stays in L2 in operation",
        },
        "davidson_cycles": {"min": int(min(cycles)), "max": int(max(cycles)),
                            "mean": float(np.mean(cycles))},
        "energies": [float(r["energy"]) for r in res_dev][:4],
    }
    if extra:
        line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, batches, h, g, res_dev)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def ncu_traffic(v2: bool = True):
    """DRAM bytes per sigma build (all kernels of one build) from the committed `ncu --set full` captures
    (profiles/r2_sigma2_{ab,tile,epilogue}_ncu_raw_c4.txt, or round 1's sigma_a/sigma_b for the v1 path);
    null when the summaries are not there."""
    names = [f"r2_sigma2_{k}_ncu_raw_c4.txt" for k in ("ab", "tile", "epilogue")] if v2 else \
            [f"r1_sigma_{k}_ncu_raw_c4.txt" for k in ("a", "b")]
    tot, found = 0.0, False
    for name in names:
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        found = True
        for ln in open(path):
            if ln.startswith(("dram__bytes_read.sum", "dram__bytes_write.sum")):
                val, unit = ln.split("=")[1].split()[:2]
                tot += float(val) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    return tot if found else None


def cpu_baseline(args, batches, h, g, res_dev) -> dict:
    """Oracle port timed on the host cores on a bounded sample (rank 0, N=1 only)."""
    from oracle import sci_cpu

    cores = os.cpu_count() or 1
    n_sample = min(len(batches), 4)
    t0 = time.perf_counter()
    dets, de_max, cyc = 0, 0.0, []
    for k in range(n_sample):
        sa, sb = batches[k]
        e, amps, occ, info = sci_cpu.solve(sa, sb, h, g, algo=0, tol=1e-12, max_cycle=100, max_space=12,
                                           nthreads=cores)
        dets += len(sa) * len(sb)
        de_max = max(de_max, abs(e - float(res_dev[k]["energy"])))
        cyc.append(info["cycles"])
    dt = time.perf_counter() - t0
    out = {"value": dets / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{n_sample} of {len(batches)} subspaces of the step, oracle/sci_cpu.c direct "
                     f"excitation-table algorithm (OpenMP), Davidson tol 1e-12, cycles {cyc}",
           "seconds": dt, "max_abs_dE_vs_gpu_Ha": de_max}
    # context: cost of ONE sigma build with pyscf's published algorithm (gather -> dgemm -> scatter
    # through N-2 intermediates), which is what the reference's CPU path actually runs
    if not args.skip_pyscf_style:
        sa, sb = batches[0]
        t0 = time.perf_counter()
        sci_cpu.solve(sa, sb, h, g, algo=1, max_cycle=-1, nthreads=cores)
        out["pyscf_style_sigma_build_s"] = time.perf_counter() - t0
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--batches", type=int, default=8, help="subspaces per GPU per step")
    ap.add_argument("--ref-sample", type=int, default=2, help="subspaces per step on the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--skip-pyscf-style", action="store_true")
    ap.add_argument("--extras", default="all", choices=["all", "fermion", "none"],
                    help="secondary measurements under the 'extra' key (bench_extras.py)")
    ap.add_argument("--z40-rows", type=int, default=50_000_000, help="rows of the Z^(x)40 projection extra")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostics: do not poll nvidia-smi")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

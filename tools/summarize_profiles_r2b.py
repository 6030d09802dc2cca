"""Summaries of tools/gpu_profile_r2b.sh's `ncu --set full` captures (gpurun_out/r2prof_{qubit,recover,wide}.ncu-rep)
-> profiles/r2_other_kernels_ncu.md (one row per captured launch) and profiles/r2_<name>_ncu_raw.csv.

    python tools/summarize_profiles_r2b.py
"""
import csv
import io
import os
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go = os.path.join(root, "gpurun_out")
out = os.path.join(root, "profiles")

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/inst"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
]

with open(os.path.join(out, "r2_other_kernels_ncu.md"), "w") as md:
    md.write("# ncu `--set full` summaries: kernels outside the fermion sigma build (round 2)\n\n"
             "Captured by `tools/gpu_profile_r2b.sh` (`--clock-control none`, cold cache per launch, every launch "
             "replayed ~40 times: durations here are NOT bench values).  Workloads: `qubit` = BASELINE configs[2] "
             "(40 qubits, 1e4 Pauli terms, 3e5 sampled rows, 74 637 unique: the first 14 launches are the key sort of "
             "`sqd_sort_unique`; `pauli2` = the projection kernels of the same workload, count and fill pass), `recover` = 1e5 "
             "bitstrings of (30e,30o), exact numpy stream, `wide` = the staging-free sigma kernel at one "
             "9000 x 9000 = 8.1e7-determinant subspace.  Full metric tables: `profiles/r2_<name>_ncu_raw.csv`.\n")
    for name in ("qubit", "pauli2", "recover", "wide"):
        rep = os.path.join(go, f"r2prof_{name}.ncu-rep")
        if not os.path.exists(rep):
            continue
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        with open(os.path.join(out, f"r2_{name}_ncu_raw.csv"), "w") as f:
            f.write(raw)
        rows = list(csv.reader(io.StringIO(raw)))
        head, units, vals = rows[0], rows[1], rows[2:]
        ki = head.index("Kernel Name")
        md.write(f"\n## {name}\n\n| kernel | " + " | ".join(c[1] for c in COLS) + " |\n|---|" + "---:|" * len(COLS) + "\n")
        for v in vals:
            cells = []
            for metric, _ in COLS:
                if metric in head:
                    i = head.index(metric)
                    x = v[i]
                    try:
                        x = f"{float(x.replace(',', '')):.4g}"
                    except ValueError:
                        pass
                    u = units[i]
                    cells.append(f"{x} {u}".strip() if u not in ("", "%", "inst/cycle", "register/thread") else x)
                else:
                    cells.append("-")
            kn = v[ki].split("(")[0].replace("void ", "")
            md.write(f"| `{kn}` | " + " | ".join(cells) + " |\n")
print(open(os.path.join(out, "r2_other_kernels_ncu.md")).read())

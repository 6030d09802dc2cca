"""Turn round 2's raw ncu outputs under gpurun_out/ (tools/gpu_profile_r2.sh) into the committed summaries
under profiles/.

    python tools/summarize_profiles_r2.py
"""
import collections
import csv
import os
import shutil
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go = os.path.join(root, "gpurun_out")
out = os.path.join(root, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list of the bench command ----
src = os.path.join(go, "r2_launches_c4.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(out, "r2_launches_c4.csv"))
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
        name = row["Kernel Name"].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    sig = sum(v[1] for k, v in agg.items() if k.startswith(("sigma2_", "void sigma2_", "sqd::sigma2_", "void sqd::sigma2_")))
    with open(os.path.join(out, "r2_launch_summary_c4.md"), "w") as f:
        f.write("# ncu launch list summary (round 2, workload c4)\n\n"
                "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2100 -c 1000 "
                "python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extras none` (the bench command itself; "
                "ncu serialises the 8 solver threads; the window is 1000 launches of the second warm-up step, "
                "i.e. about half of one step of 8 solves; every launch runs cold-cache and alone: compare shares, "
                "not absolute times).\n\n"
                "| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {100 * v[1] / tot:.1f}% |\n")
        f.write(f"\nTotal device time in the window: {tot:.1f} us; the three sigma kernels together: "
                f"{sig:.1f} us = {100 * sig / max(tot, 1e-9):.1f}% (bench.py reports `share_of_davidson_loop` from "
                "CUDA events inside a real solve for comparison).\n")

WANT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors.sum",
        "sm__cycles_active.avg", "sm__cycles_active.max", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fp64.sum",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct")
SHORT = {"sigma2_ab_kernel": "ab", "sigma2_tile_kernel": "tile", "sigma2_epilogue_kernel": "epilogue",
         "sigma_a_kernel": "a", "sigma_b_kernel": "b", "sigma_combine_kernel": "combine"}

index = ["# Round 2 `ncu --set full` captures of the sigma kernels\n",
         "Command per workload: `ncu --set full --clock-control none --import-source on -k regex:sigma2_ -s 9 -c 3 "
         "python tests/gpu_sigma_bench.py <wl> 10 v2` (v1 at s7: `-k regex:sigma_(a|b|combine)_kernel`).  "
         "ncu's default cache control flushes L2 before every replay pass, so durations here are COLD-cache; the "
         "warm per-kernel split (`--cache-control none`) and the CUDA-event times are in DESIGN.md.\n",
         "| workload | kernel (instance) | grid x block | regs | dyn smem | time us (cold) | DRAM rd+wr MB | warps active % | "
         "smem wavefronts | file |", "|---|---|---|---:|---:|---:|---:|---:|---:|---|"]
for wl, gen in (("c4", "v2"), ("t", "v2"), ("c5", "v2"), ("c2", "v2"), ("s7", "v1")):
    rep = os.path.join(go, f"r2prof_{gen}_{wl}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        kname_full = r[col["Kernel Name"]]
        base = kname_full.split("<")[0].split("(")[0].replace("void ", "").replace("sqd::", "")
        short = SHORT.get(base, base)
        prefix = "sigma2" if gen == "v2" else "sigma"
        fname = f"r2_{prefix}_{short}_ncu_raw_{wl}.txt"
        with open(os.path.join(out, fname), "w") as f:
            f.write(f"# {kname_full}  grid {r[col['Grid Size']]} block {r[col['Block Size']]}  (workload {wl})\n")
            for w in WANT:
                if w in col:
                    f.write(f"{w} = {r[col[w]]} {units[col[w]]}\n")

        def val(name, scale=1.0):
            try:
                return float(r[col[name]].replace(",", "")) * scale
            except (KeyError, ValueError):
                return float("nan")

        def to_bytes(name):
            u = units[col[name]]
            return val(name) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)

        tu = units[col["gpu__time_duration.sum"]]
        t_us = val("gpu__time_duration.sum") * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(tu, 1.0)
        inst = kname_full.split("(")[0].replace("void ", "")
        index.append(f"| {wl} | `{inst}` | {r[col['Grid Size']]} x {r[col['Block Size']]} | "
                     f"{r[col['launch__registers_per_thread']]} | {r[col['launch__shared_mem_per_block_dynamic']]} "
                     f"{units[col['launch__shared_mem_per_block_dynamic']]} | {t_us:.1f} | "
                     f"{(to_bytes('dram__bytes_read.sum') + to_bytes('dram__bytes_write.sum')) / 1e6:.2f} | "
                     f"{val('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
                     f"{r[col['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']]} | `{fname}` |")
    # the details page of the main kernel (first sigma2_ab / sigma_a row)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    keep = [l for l in txt.splitlines() if not l.strip().startswith(("OPT", "INF")) or "Est." in l]
    open(os.path.join(out, f"r2_sigma_{gen}_ncu_details_{wl}.txt"), "w").write("\n".join(keep) + "\n")
open(os.path.join(out, "r2_sigma_instances.md"), "w").write("\n".join(index) + "\n")
print("\n".join(index))

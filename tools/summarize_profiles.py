"""Turn the raw ncu outputs under gpurun_out/ into the committed summaries under profiles/.

    python tools/summarize_profiles.py r1 c4

Inputs (written by tools/gpu_profile.sh on the GPU box):
  gpurun_out/launches_<wl>.csv        ncu --metrics gpu__time_duration.sum --clock-control none
  gpurun_out/prof_sigma_<wl>.ncu-rep  ncu --set full --clock-control none --import-source on (sigma_a)
"""
import collections
import csv
import os
import shutil
import subprocess
import sys

tag, wl = sys.argv[1], sys.argv[2]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
os.makedirs(out, exist_ok=True)
src = os.path.join(root, "gpurun_out", f"launches_{wl}.csv")
shutil.copy(src, os.path.join(out, f"{tag}_launches_{wl}.csv"))
lines = [l for l in open(src) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
    name = row["Kernel Name"].split("(")[0]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out, f"{tag}_launch_summary_{wl}.md"), "w") as f:
    f.write(f"# ncu launch list summary ({tag}, workload {wl})\n\n"
            "Command: `ncu --metrics gpu__time_duration.sum --clock-control none python bench.py "
            f"--workload {wl} --steps 1 --warmup 3 --no-cpu-baseline` (the bench command itself; ncu serialises "
            "the 8 solver threads; `--launch-skip 2000 -c 2100` records one whole step of 8 solves (about 2000 launches), "
            "set-up kernels included, after the first warm-up step; every launch runs cold-cache: "
            "compare shares, not absolute times).\n\n"
            "| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {100 * v[1] / tot:.1f}% |\n")
    f.write(f"\nTotal device time: {tot:.1f} us\n")
for kname, rep in (("sigma_a", os.path.join(root, "gpurun_out", f"prof_sigma_{wl}.ncu-rep")),
                   ("sigma_b", os.path.join(root, "gpurun_out", f"prof_sigmab_{wl}.ncu-rep"))):
    if not os.path.exists(rep):
        continue
    txt = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
    keep = [l for l in txt.splitlines() if not l.strip().startswith(("OPT", "INF")) or "Est." in l]
    open(os.path.join(out, f"{tag}_{kname}_ncu_details_{wl}.txt"), "w").write("\n".join(keep) + "\n")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    want = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__cycles_active.avg", "sm__cycles_active.max", "smsp__inst_executed.sum",
            "sm__inst_executed.avg.per_cycle_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sectors.sum")
    with open(os.path.join(out, f"{tag}_{kname}_ncu_raw_{wl}.txt"), "w") as f:
        for w in want:
            for i, h in enumerate(rows[0]):
                if h == w:
                    f.write(f"{w} = {rows[2][i]} {rows[1][i]}\n")
print("profiles written to", out)

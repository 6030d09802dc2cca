#!/bin/bash
# ncu evidence (run under gpurun, ONE GPU).  Numbers printed by these runs are never bench values.
#   1. launch list of the bench command itself; ncu serialises the 8 solver threads (~0.3 s per launch),
#      so the capture is bounded: skip the first warm-up step, record the next $NL launches
#   2. full captures of the two sigma kernels inside a real single-stream solve
mkdir -p gpurun_out
WL=${1:-c4}
NL=${2:-700}
if [ "$NL" -gt 0 ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2800 -c $NL --csv \
    --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${WL}.out 2>&1
fi
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sigma_a_kernel -s 20 -c 1 -f \
    -o gpurun_out/prof_sigma_${WL} python tests/gpu_profile_driver.py $WL 2 > gpurun_out/prof_sigma_${WL}.out 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sigma_b_kernel -s 20 -c 1 -f \
    -o gpurun_out/prof_sigmab_${WL} python tests/gpu_profile_driver.py $WL 2 > gpurun_out/prof_sigmab_${WL}.out 2>&1
ls -la gpurun_out | tail -8

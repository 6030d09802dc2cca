#!/bin/bash
mkdir -p gpurun_out
WL=${1:-c4}
# launch list: every kernel of one warm solve with its device time
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${WL}.csv python tests/gpu_profile_driver.py $WL 2 > gpurun_out/launches_${WL}.out 2>&1
# full capture of the sigma kernel (skip the first solve's launches)
ncu --set full --clock-control none --import-source on -k regex:sigma_a_kernel -s 20 -c 1 -f -o gpurun_out/prof_sigma_${WL} python tests/gpu_profile_driver.py $WL 2 > gpurun_out/prof_sigma_${WL}.out 2>&1
ls -la gpurun_out | tail -8
tail -3 gpurun_out/launches_${WL}.out

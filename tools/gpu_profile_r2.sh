#!/bin/bash
# Round-2 ncu evidence (run under gpurun, ONE GPU).  Numbers printed by these runs are never bench values.
#   1. launch list of the bench command itself (ncu serialises the 8 solver threads): skip the first warm-up
#      step (~2000 launches of 8 solves), record the next $NL launches of the second one
#   2. `--set full` captures of every sigma kernel the planner selects: the three v2 kernels at c4, t, c5, c2
#      and the v1 kernels at the streaming scale-up point s7 (1e7 determinants, where the planner keeps v1)
mkdir -p gpurun_out
NL=${1:-1000}
if [ "$NL" -gt 0 ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2100 -c $NL --csv \
    --log-file gpurun_out/r2_launches_c4.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2_launches_c4.out 2>&1
fi
for WL in c4 t c5 c2; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sigma2_" -s 9 -c 3 -f \
        -o gpurun_out/r2prof_v2_${WL} python tests/gpu_sigma_bench.py $WL 10 v2 > gpurun_out/r2prof_v2_${WL}.out 2>&1
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"sigma_(a|b|combine)_kernel" -s 6 -c 3 -f \
    -o gpurun_out/r2prof_v1_s7 python tests/gpu_sigma_bench.py s7 4 v1 > gpurun_out/r2prof_v1_s7.out 2>&1
ls -la gpurun_out | grep r2prof
tail -2 gpurun_out/r2prof_v1_s7.out

#!/bin/bash
# warm-cache launch list of ONE lone solve (single-pass ncu, caches not flushed between launches)
WL=${1:-c4}
timeout 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum --csv \
    --log-file gpurun_out/r2_warm_${WL}.csv python tests/gpu_profile_driver.py $WL 3 > gpurun_out/r2_warm_${WL}.out 2>&1
tail -1 gpurun_out/r2_warm_${WL}.out

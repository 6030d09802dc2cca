# warm-cache per-kernel split of one sigma build (single-pass ncu, caches not flushed between launches)
WL=${1:-c4}
timeout 300 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum -k regex:"sigma" -s 30 -c 12 python tests/gpu_sigma_bench.py $WL 30 ${2:-v2} 2>&1 | grep -E "^  [a-z].*\(|gpu__time" | paste - - | awk '{print $1, $2, $NF}' 

mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sigma_instances_gpu.py -x -q -m gpu 2>&1 | tail -5
python tests/gpu_sigma_bench.py c4 200
python tests/gpu_sigma_bench.py t 200
python tests/gpu_sigma_bench.py c5 50
timeout 300 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size --clock-control none -k regex:sigma2 -s 15 -c 3 python tests/gpu_sigma_bench.py c4 20 v2 2>&1 | grep -E "sigma2|gpu__time|sm__cycles|inst_executed|wavefronts|warps_active|grid_size|block_size"

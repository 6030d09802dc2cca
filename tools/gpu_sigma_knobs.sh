mkdir -p gpurun_out
python -m pytest "tests/test_sigma_instances_gpu.py" -x -q -m gpu -k "v2 and c4" 2>&1 | grep -E "assert|Error|error|^E " | head -20
for st in 2 4 8; do for ct in 2 3 4; do
echo "stages $st ctas/SM $ct"; SQD_V2_STAGES=$st SQD_V2_CTAS_PER_SM=$ct python tests/gpu_sigma_bench.py c4 100 v2 2>&1 | tail -1
done; done
SQD_V2_STAGES=8 SQD_V2_CTAS_PER_SM=3 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sigma2 -s 15 -c 3 python tests/gpu_sigma_bench.py c4 20 v2 2>&1 | grep -E "sigma2|gpu__time"

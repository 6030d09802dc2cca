#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/sweep.txt
python -m pytest tests/test_fermion_gpu.py -q -m gpu -x 2>&1 | tail -2 >> gpurun_out/sweep.txt
for cfg in "0 24576 128" "0 24576 256" "0 24576 512" "0 24576 1024"; do
  set -- $cfg
  SQD_SIGMA_STAGES=$1 SQD_SIGMA_PACK_BYTES=$2 SQD_SIGMA_CHUNK_COST=$3 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sw.json 2>gpurun_out/sw.err
  python - "$cfg" <<'PY' >> gpurun_out/sweep.txt
import json,sys
l=[x for x in open('gpurun_out/sw.json') if x.startswith('{')]
if not l: print(sys.argv[1], "FAILED"); raise SystemExit
d=json.loads(l[-1]); r=d['roofline']
print(sys.argv[1], "value %.2f e2e %.2f sigma_us %.1f loop_ms %.2f"%(d['value'], d['e2e']['value'], r['ms_per_launch']*1e3, r['davidson_loop_ms']))
PY
done
cat gpurun_out/sweep.txt

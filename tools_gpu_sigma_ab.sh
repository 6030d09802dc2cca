#!/bin/bash
# parity first, then sigma timing under two register caps
python -m pytest tests/test_fermion_gpu.py -q -m gpu -x 2>&1 | tail -3
for m in 4 3; do
  echo "SQD_SIGMA_MINB=$m"
  SQD_SIGMA_MINB=$m python tests/gpu_sigma_concurrency.py c4 8 2>&1 | grep "K="
  SQD_SIGMA_MINB=$m python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step',d['ms_per_step'],'value', d['value'],'e2e', d['e2e']['value'],'sigma_us',d['roofline']['ms_per_launch']*1e3)"
done

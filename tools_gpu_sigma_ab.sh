#!/bin/bash
python -m pytest tests/test_fermion_gpu.py -q -m gpu -x 2>&1 | tail -2
for m in 3 1; do
  echo "SQD_SIGMA_MINB=$m"
  for rep in 1 2; do
  SQD_SIGMA_MINB=$m python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step',d['ms_per_step'],'value', d['value'],'e2e', d['e2e']['value'],'sigma_us',d['roofline']['ms_per_launch']*1e3)"
  done
done

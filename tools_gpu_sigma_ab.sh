#!/bin/bash
for m in 1 4 3; do
  echo "SQD_SIGMA_MINB=$m"
  SQD_SIGMA_MINB=$m python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step',d['ms_per_step'],'value', d['value'],'e2e', d['e2e']['value'], d['e2e_loop_only']['value'],'sigma_us',d['roofline']['ms_per_launch']*1e3)"
done

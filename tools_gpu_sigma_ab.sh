#!/bin/bash
for c in 0 1 0 1; do
  echo "SQD_RITZ_SIDE=$c"
  SQD_RITZ_SIDE=$c python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms_per_step',d['ms_per_step'],'value', d['value'],'e2e', d['e2e']['value'], d['e2e_loop_only']['value'],'loop_ms',d['roofline']['davidson_loop_ms'])"
done

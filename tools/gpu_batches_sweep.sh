#!/bin/bash
# how does a step scale with the number of concurrent subspaces?  (host-launch-bound vs device-bound)
mkdir -p gpurun_out; : > gpurun_out/batches.txt
for b in 1 2 4 8 16; do
  python bench.py --steps 3 --warmup 3 --batches $b --no-cpu-baseline > gpurun_out/sw.json 2>gpurun_out/sw.err
  python - "$b" <<'PY' >> gpurun_out/batches.txt
import json,sys
l=[x for x in open('gpurun_out/sw.json') if x.startswith('{')]
if not l: print(sys.argv[1], "FAILED"); raise SystemExit
d=json.loads(l[-1]); r=d['roofline']
print("batches", sys.argv[1], "ms_per_step %.2f value %.2f e2e %.2f sigma_us %.1f loop_ms %.2f launches %d"%(d['ms_per_step'], d['value'], d['e2e']['value'], r['ms_per_launch']*1e3, r['davidson_loop_ms'], d['gpu_launches']))
PY
done
cat gpurun_out/batches.txt

# step time of the bench workload under different knob settings (each line: env -> value, ms/step, sigma ms)
mkdir -p gpurun_out
run() { env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extras none 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*', '| Mdet/s %.1f step %.2f ms e2e %.1f sigma %.1f us lone-loop %.2f ms launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], 1e3*d['roofline']['ms_per_launch'], d['roofline']['davidson_loop_ms'], d['gpu_launches']))"; }
for cfg in "$@"; do run $cfg; done

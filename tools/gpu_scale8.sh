# weak-scaling check on N GPUs of one box: bench.py under torchrun, one JSON summary line per setting
N=${1:-8}
shift
for cfg in "$@"; do
env $cfg python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    ln=ln.strip()
    if ln.startswith('{'):
        d=json.loads(ln); print('$cfg N=$N', '| Mdet/s %.1f step %.2f ms e2e %.1f (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
done

set -x
N=$1
for ch in default 4 8 16; do
if [ "$ch" != "default" ]; then export NCCL_MAX_CTAS=$ch NCCL_MAX_NCHANNELS=$ch; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --dtype fp32 --steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-latency 2>/dev/null | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nccl channels $ch: N=$N', round(d['value'],1), 'seq/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1))
except Exception as e: print('ERR', e)"
done

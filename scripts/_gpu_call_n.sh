N=$1
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --dtype fp32 --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu_fp32.json 2> gpurun_out/r02_bench_${N}gpu_fp32.err
python - <<PY
import json
f="gpurun_out/r02_bench_${N}gpu_fp32.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"],1), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), d["launches_per_step"], d.get("clocks"))
except Exception as e:
    print(f, "ERR", e)
PY

set -x
N=$1
for dt in fp32 bf16; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --dtype $dt --steps 20 --warmup 3 > gpurun_out/r02_bench_${N}gpu_$dt.json 2> gpurun_out/r02_bench_${N}gpu_$dt.err
python - <<PY
import json
f="gpurun_out/r02_bench_${N}gpu_$dt.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"], (d.get("e2e_fp32_frames") or {}).get("value"), d["launches_per_step"], d.get("clocks"))
except Exception as e:
    print(f, "ERR", e); import subprocess; print(subprocess.run("grep -a 'rank0\|Error\|error' "+f.replace('.json','.err')+" | tail -12", shell=True, capture_output=True, text=True).stdout[-2500:])
PY
done

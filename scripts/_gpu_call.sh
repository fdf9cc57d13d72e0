set -x
python -m pytest tests/test_gpu_trajectory.py tests/test_gpu_module.py -m gpu -q --tb=short 2>&1 | tail -5 | cut -c1-300
for dt in fp32 bf16; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --dtype $dt --steps 10 --warmup 3 > gpurun_out/r02g_bench_2gpu_$dt.json 2> gpurun_out/r02g_bench_2gpu_$dt.err
python - <<PY
import json
f="gpurun_out/r02g_bench_2gpu_$dt.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"], d["e2e_fp32_frames"]["value"], d["launches_per_step"])
except Exception as e:
    print(f, "ERR", e); import subprocess; print(subprocess.run("grep -a 'rank0' "+f.replace('.json','.err')+" | tail -12", shell=True, capture_output=True, text=True).stdout[-2500:])
PY
done

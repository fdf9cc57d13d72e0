set -x
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_conv_bf16.py tests/test_gpu_bf16_step.py tests/test_gpu_parity.py -m gpu -q --tb=short -x 2>&1 | tail -4 | cut -c1-300
for dt in bf16 fp32; do
timeout 600 python bench.py --dtype $dt --steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02l_bench_$dt.json 2> gpurun_out/r02l_bench_$dt.err
python - <<PY
import json
f="gpurun_out/r02l_bench_$dt.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["launches_per_step"], d["loss"]); print({k: v["ms"] for k, v in d["roofline"]["kernels"].items()}); print({k: d["roofline"][k] for k in ("kernel","bound","achieved","peak","frac","traffic","share_of_step")})
except Exception as e:
    print(f, "ERR", e); import subprocess; print(subprocess.run("tail -20 "+f.replace('.json','.err'), shell=True, capture_output=True, text=True).stdout[-2500:])
PY
done

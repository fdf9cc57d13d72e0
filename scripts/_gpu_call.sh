set -x
python -m pytest tests/test_gpu_conv_bf16.py -m gpu -q --tb=short 2>&1 | tail -60 | cut -c1-400
HULC_B200_BF16_CONV=0 python -m pytest tests/test_gpu_bf16_step.py -m gpu -q --tb=short 2>&1 | tail -30 | cut -c1-600
python -m pytest tests/test_gpu_bf16_step.py -m gpu -q --tb=short -s 2>&1 | tail -40 | cut -c1-1200
python bench.py --dtype bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02d_bench_bf16.json 2> gpurun_out/r02d_bench_bf16.err; python - <<'PY'
import json
for f in ("gpurun_out/r02d_bench_bf16.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_uint8_frames"]["value"], d["launches_per_step"]); print({k:(v["ms"],v["tflops"],v["gbs"]) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e: print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY

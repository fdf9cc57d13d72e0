set -x
python -m pytest tests/test_gpu_bf16_step.py tests/test_gpu_bf16.py tests/test_param_store.py -m gpu -q 2>&1 | tail -30
python bench.py --dtype bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02c_bench_bf16.json 2> gpurun_out/r02c_bench_bf16.err; python - <<'PY'
import json
for f in ("gpurun_out/r02c_bench_bf16.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_uint8_frames"]["value"], d["launches_per_step"]); print({k:(v["ms"],v["tflops"]) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e: print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02c_launches_bf16.csv python scripts/profile_step.py --steps 2 --precision bf16 > gpurun_out/r02c_prof.log 2>&1; tail -2 gpurun_out/r02c_prof.log
HULC_SANITIZE_TIMEOUT=240 scripts/sanitize.sh memcheck racecheck

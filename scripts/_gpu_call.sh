set -x
python -m pytest tests/test_gpu_rnn_tc.py tests/test_gpu_trajectory.py -m gpu -q --tb=short -s 2>&1 | tail -12 | cut -c1-400
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02f_bench_fp32.json 2> gpurun_out/r02f_bench_fp32.err
python bench.py --dtype bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02f_bench_bf16.json 2> gpurun_out/r02f_bench_bf16.err
python - <<'PY'
import json
for f in ("gpurun_out/r02f_bench_fp32.json","gpurun_out/r02f_bench_bf16.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_fp32_frames"]["value"], d["e2e_lightning_contract"]["value"], d["launches_per_step"]); print({k:(v["ms"],v["tflops"],v["gbs"]) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e: print(f, "ERR", e); print(open(f.replace(".json",".err")).read()[-2500:])
PY

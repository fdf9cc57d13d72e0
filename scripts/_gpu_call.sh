set -x
timeout 120 python scripts/time_rnn.py 2>&1 | tail -2
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -8 | cut -c1-300 ) 2>&1
for dt in bf16 fp32; do
timeout 600 python bench.py --dtype $dt --steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02i_bench_$dt.json 2> gpurun_out/r02i_bench_$dt.err
python - <<PY
import json
f="gpurun_out/r02i_bench_$dt.json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["launches_per_step"], d["loss"])
except Exception as e:
    print(f, "ERR", e); import subprocess; print(subprocess.run("tail -20 "+f.replace('.json','.err'), shell=True, capture_output=True, text=True).stdout[-2500:])
PY
done

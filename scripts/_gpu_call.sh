set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 | cut -c1-300 ) 2>&1 | grep -v "^$"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_fp32.json 2> gpurun_out/r02_bench_fp32.err; tail -c 300 gpurun_out/r02_bench_fp32.err
timeout 600 python bench.py --dtype fp32 --steps 200 --warmup 5 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02_bench_long_200steps_fp32.json 2>/dev/null
for cfg in mcil gcbc64 gcbc64-gru; do
timeout 600 python bench.py --config $cfg --dtype fp32 --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02_bench_${cfg}_fp32.json 2> gpurun_out/r02_bench_${cfg}_fp32.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*fp32.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d.get("launches_per_step"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_tf32.csv python scripts/profile_step.py --steps 2 --precision tf32 > /dev/null 2>&1
ncu --set full --clock-control none -k regex:'gemm_bf16_kernel' -c 100 -o gpurun_out/r02_full_gemm_tf32 python scripts/profile_step.py --steps 1 > /dev/null 2>&1
ncu -i gpurun_out/r02_full_gemm_tf32.ncu-rep --page raw --csv > gpurun_out/r02_full_gemm_tf32_raw.csv 2>/dev/null; rm -f gpurun_out/r02_full_gemm_tf32.ncu-rep
du -sh gpurun_out

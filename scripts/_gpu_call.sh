set -x
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 | cut -c1-300 ) 2>&1 | grep -v "^$"
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_fp32.json 2> gpurun_out/r02_bench_fp32.err; tail -c 600 gpurun_out/r02_bench_fp32.err
timeout 900 python bench.py --dtype bf16 --steps 20 --warmup 3 > gpurun_out/r02_bench_bf16.json 2> gpurun_out/r02_bench_bf16.err; tail -c 600 gpurun_out/r02_bench_bf16.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null
for cfg in mcil gcbc64 gcbc64-gru; do for dt in fp32 bf16; do
timeout 600 python bench.py --config $cfg --dtype $dt --steps 10 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02_bench_${cfg}_$dt.json 2> gpurun_out/r02_bench_${cfg}_$dt.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d.get("launches_per_step"), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "ERR", e)
PY
for prec in bf16 tf32; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r02_launches_$prec.csv python scripts/profile_step.py --steps 2 --precision $prec > /dev/null 2>&1
done
ncu --set full --clock-control none -k regex:'conv1_band_wgrad|spatial_softmax|conv_band_kernel<128' -c 16 -o gpurun_out/r02b_full_bf16 python scripts/profile_step.py --steps 1 --precision bf16 > /dev/null 2>&1
ncu -i gpurun_out/r02b_full_bf16.ncu-rep --page raw --csv > gpurun_out/r02b_full_bf16_raw.csv 2>/dev/null; rm -f gpurun_out/r02b_full_bf16.ncu-rep
ncu --set full --clock-control none -k regex:'conv1_band_wgrad|spatial_softmax|conv_band_kernel<128' -c 16 -o gpurun_out/r02b_full_tf32 python scripts/profile_step.py --steps 1 > /dev/null 2>&1
ncu -i gpurun_out/r02b_full_tf32.ncu-rep --page raw --csv > gpurun_out/r02b_full_tf32_raw.csv 2>/dev/null; rm -f gpurun_out/r02b_full_tf32.ncu-rep
HULC_SANITIZE_TIMEOUT=420 bash scripts/sanitize.sh memcheck racecheck 2>&1 | tail -3
du -sh gpurun_out

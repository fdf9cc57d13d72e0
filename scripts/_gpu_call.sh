set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 | cut -c1-300 ) 2>&1 | grep -v "^$"
for dt in fp32 bf16; do
timeout 600 python bench.py --dtype $dt --steps 200 --warmup 5 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02_bench_long_200steps_$dt.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r02_bench_long_200steps_$dt.json').read().strip().splitlines()[-1]); print('200 steps $dt', round(d['value'],1), round(d['ms_per_step'],3), d['clocks'], 'e2e', round(d['e2e']['value'],1), d['roofline'].get('note'))"
done

set -x
timeout 600 python bench.py --dtype fp32 --steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-latency 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fp32', d['value'], d['ms_per_step'], d['loss'], d['launches_per_step'])"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trajectory.py tests/test_validation.py tests/test_inference.py tests/test_gpu_module.py tests/test_module_surface.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -6 | cut -c1-300

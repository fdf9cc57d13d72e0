set -x
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gemm.py -m gpu -q --tb=short -x 2>&1 | grep -v "^$" | tail -12 | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_trajectory.py tests/test_validation.py tests/test_inference.py tests/test_gpu_module.py -m gpu -q --tb=short 2>&1 | grep -v "^$" | tail -6 | cut -c1-300
for v in 3 1; do
HULC_B200_GEMM_TF32_TMA=$v timeout 600 python bench.py --dtype fp32 --steps 20 --warmup 3 --no-cpu-baseline --no-eager --no-latency 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fp32 tma=$v', d['value'], d['ms_per_step'], d['loss'], {k: v['ms'] for k, v in d['roofline']['kernels'].items() if 'dense' in k})"
done

timeout 900 python -m pytest tests/test_gpu_module.py -m gpu -q --tb=short -x -k "built_on_cpu" 2>&1 | grep -v "^$" | tail -25 | cut -c1-400
HULC_B200_GEMM_TF32_TMA=0 timeout 900 python -m pytest tests/test_gpu_module.py -m gpu -q --tb=line -x -k "built_on_cpu" 2>&1 | tail -3 | cut -c1-300

set -x
timeout 900 python -m pytest tests/test_gpu_bf16_step.py -m gpu -q --tb=long -x -k graph_replay 2>&1 | grep -v "^$" | head -120 | cut -c1-300
HULC_B200_OVERLAP_WGRAD=0 timeout 900 python -m pytest tests/test_gpu_bf16_step.py -m gpu -q --tb=short -x -k graph_replay 2>&1 | tail -5 | cut -c1-300

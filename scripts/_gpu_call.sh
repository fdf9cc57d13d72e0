timeout 60 python -m pytest tests/test_kernels.py -m gpu -x -q -k "cosine or bce" 2>&1 | tail -3 | cut -c1-200

import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from hulc_b200 import ops
def timeit(fn, n=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
x = torch.zeros(1 << 20, device="cuda")
print("empty-ish kernel (scale_ 1M)", f"{timeit(lambda: ops.scale_(x, 1.0)):.1f} us")
for (M, N, K) in [(64, 2048, 256), (64, 2048, 512), (64, 2048, 1024), (64, 2048, 2048), (64, 2048, 4096), (64, 16384, 256), (64, 16384, 2048), (128, 16384, 128), (64, 1024, 2048), (64, 128, 128)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    print((M, N, K), " ".join(f"tc{tc}: {timeit(lambda: ops.gemm(A, B, C, transB=True, tc=tc), 30):6.1f} us" for tc in (1, 3)))

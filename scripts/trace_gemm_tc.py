"""Timeline of one tf32 product of hulc_gemm_tc (CTA 0) from the globaltimer stamps compiled in with -DHULC_TC_TRACE
(scripts/build_trace.sh).  Slots (tc_pipeline.cuh): 0 kernel entry | 1 prologue done | 2 first stage landed | 3 last stage landed |
4 accumulator ready | 5 epilogue role loop done | 6 cluster partials in place | 7 cluster reduction stored | 8 end."""
import ctypes
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from hulc_b200 import _lib, ops  # noqa: E402

real = _lib.Library(ROOT / "hulc_b200" / "lib" / "libhulc_trace.so", allow_missing=True)
_lib.lib = lambda: real
lib = real.cdll
lib.hulc_tc_trace_read.argtypes = [ctypes.c_void_p]
for (M, N, K, tc, tA, tB) in ((64, 128, 128, 3, False, True), (64, 2048, 2048, 3, False, True), (2048, 2048, 128, 1, False, True), (2048, 2048, 2048, 1, True, False),
                              (2048, 2048, 2048, 3, False, True)):
    A = torch.randn((K, M) if tA else (M, K), device="cuda")
    B = torch.randn((N, K) if tB else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    for _ in range(4):
        ops.gemm(A, B, C, transA=tA, transB=tB, tc=tc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(A, B, C, transA=tA, transB=tB, tc=tc)
    e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 64)()
    assert lib.hulc_tc_trace_read(ctypes.addressof(buf)) == 0
    t = [buf[i] for i in range(9)]
    print(f"{M}x{N}x{K} tc={tc} transA={int(tA)} transB={int(tB)}: {e0.elapsed_time(e1) * 100:.1f} us per launch | " +
          "  ".join(f"[{i}] +{(t[i] - t[0]) / 1e3:5.2f}" for i in range(9) if t[i] >= t[0]))

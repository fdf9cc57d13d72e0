import sys, ctypes
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from hulc_b200 import _lib, ops
_lib._LIB = _lib.Library(_lib.PKG / "lib" / "libhulc_trace.so")
rd = _lib._LIB.cdll.hulc_tc_trace_read
names = ["entry", "setup done", "first stage", "last stage", "acc ready", "role loop done", "cluster sync1", "reduce done", "exit"]
for (M, N, K) in [(64, 128, 128), (128, 16384, 128)]:
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda"); C = torch.empty(M, N, device="cuda")
    for tc, act in ((1, 0), (1, 256), (1, 512)):
        for _ in range(3):
            ops.gemm(A, B, C, transB=True, tc=tc, act=act)
        buf = (ctypes.c_ulonglong * 64)()
        rd(buf)
        t0 = buf[0]
        print((M, N, K), "tc", tc, "act", act, " ".join(f"{names[i]}={(buf[i]-t0)/1e3:.1f}us" for i in range(1, 9) if buf[i] >= t0))

"""Per-step timeline of the bf16 persistent recurrence kernel (CTA 0), from clock64 stamps compiled in with -DHULC_RNN_TRACE
(scripts/build_trace.sh builds hulc_b200/lib/libhulc_trace.so).  Slots: 0 producer warp 0 starts polling | 1 its pieces have arrived |
2 issuer 0 sees k-block 0 staged | 3 issuer 0 has committed | 4 epilogue sees the accumulator | 5 partials pushed | 6 the four partials
of the own rows have landed | 7 new state stored."""
import ctypes
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
lib = ctypes.CDLL(str(ROOT / "hulc_b200" / "lib" / "libhulc_trace.so"))
H, B, S = 2048, 64, 32
transW = int(sys.argv[1]) if len(sys.argv) > 1 else 0
tf32 = len(sys.argv) > 2 and sys.argv[2] == "tf32"  # the fp32-state variant behind hulc_rnn_tc_seq (needs HULC_B200_RNN_GEN=2)
g = torch.Generator().manual_seed(0)
W16 = ((torch.rand(H, H, generator=g) * 2 - 1) / H ** 0.5).cuda().to(torch.bfloat16)
pre = (torch.randn(S, B, H, generator=g) * 0.5).cuda()
hbuf = torch.zeros(S + 2, B, H, device="cuda")
x16 = torch.empty((S + 1) * B * H, dtype=torch.bfloat16, device="cuda")
vp, ll, ci = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
fn = lib.hulc_rnn_seq_bf16
fn.argtypes = [vp, ci, ci, vp, ci, vp, vp, ll, ci, vp, ll, ci, vp, ll, ci, ci, ci, ci, ci, vp]
fn.restype = ci
W32 = W16.float()
ws = torch.zeros(1 << 20, device="cuda")
fn32 = lib.hulc_rnn_tc_seq
fn32.argtypes = [vp, ci, ci, vp, ll, ci, vp, ll, ci, vp, ll, ci, vp, ll, ci, ci, ci, ci, ci, vp, ctypes.c_size_t, vp]
fn32.restype = ci
for it in range(3):
    if tf32:
        hbuf.zero_()
        rc = fn32(W32.data_ptr(), H, transW, hbuf[0].data_ptr(), hbuf.stride(0), H, hbuf[1].data_ptr(), hbuf.stride(0), H, pre[0].data_ptr(), pre.stride(0), H,
                  None, 0, 0, 1, B, H, S, ws.data_ptr(), ws.numel() * 4, None)
        assert rc == 0, rc
        continue
    rc = fn(W16.data_ptr(), H, transW, hbuf[0].data_ptr(), H, x16.data_ptr(), hbuf[1].data_ptr(), hbuf.stride(0), H, pre[0].data_ptr(), pre.stride(0), H,
            None, 0, 0, 1, B, H, S, None)
    assert rc == 0, rc
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (64 * 16))()
lib.hulc_rnn_push_trace_read.argtypes = [vp]
assert lib.hulc_rnn_push_trace_read(ctypes.addressof(buf)) == 0
t = np.array(buf, dtype=np.int64).reshape(64, 16)[:S, :8].astype(np.float64)
clk = 1.965e3  # cycles per us
step = np.diff(t[:, 7])
print("us per step (stores done -> stores done): median %.2f  min %.2f  max %.2f" % (np.median(step) / clk, step.min() / clk, step.max() / clk))
names = ["poll start", "pieces arrived", "issuer: kb0 staged", "issuer: committed", "epi: acc ready", "epi: pushed", "epi: partials landed", "epi: stored"]
rel = (t[4:] - t[3:-1, 7:8]) / clk  # relative to the previous step's "stored"
for i, n in enumerate(names):
    print(f"  {n:22s} +{np.median(rel[:, i]):6.2f} us after the previous step's store (min {rel[:, i].min():6.2f}, max {rel[:, i].max():6.2f})")

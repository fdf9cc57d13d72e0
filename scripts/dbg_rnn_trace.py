"""Per-step timeline of the persistent recurrence kernel (CTA 0), from a -DHULC_RNN_TRACE build:
    nvcc ... -DHULC_RNN_TRACE (scripts/build_trace.sh) -> hulc_b200/lib/libhulc_trace.so;  python scripts/dbg_rnn_trace.py"""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from hulc_b200 import _lib, ops

_lib._LIB = _lib.Library(_lib.PKG / "lib" / "libhulc_trace.so")
rd = _lib._LIB.cdll.hulc_rnn_trace_read
H, B, S = 2048, 64, 32
names = ["prod: step start", "prod: flags seen", "prod: loads landed", "prod: last stage published", "mma: first stage", "mma: last commit issued",
         "epi: tmem_full", "epi: parked", "epi: cluster barrier done", "epi: stores issued", "epi: fence+bar done", "epi: flag added",
         "prod: kb5 start", "prod: kb10 start", "prod(w12): loads landed", "prod(w12): last published"]
for transW, dbg in ((False, 0), (False, 64)):
    W = (torch.rand(H, H, device="cuda") * 2 - 1) / H ** 0.5
    pre = torch.randn(S, B, H, device="cuda") * 0.5
    hb = torch.zeros(S + 2, B, H, device="cuda")
    for _ in range(3):
        ops.rnn_tc_seq(W, hb[0], hb[1], pre[0], S, prev_step=hb.stride(0), out_step=hb.stride(0), add_step=pre.stride(0), act=1 | dbg,
                       gate0=hb[1] if transW else None, gate_step=hb.stride(0) if transW else 0, transW=transW)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.rnn_tc_seq(W, hb[0], hb[1], pre[0], S, prev_step=hb.stride(0), out_step=hb.stride(0), add_step=pre.stride(0), act=1 | dbg,
                   gate0=hb[1] if transW else None, gate_step=hb.stride(0) if transW else 0, transW=transW)
    e1.record()
    torch.cuda.synchronize()
    print(f"transW={transW} dbg={dbg}: {e0.elapsed_time(e1) * 1e3:.1f} us for {S} steps")
    buf = (ctypes.c_longlong * (64 * 64))()
    rd(buf)
    for s in (0, 1, 10, 11, 31):
        t = [buf[s * 64 + i] for i in range(16)]
        base = t[0]
        print(f"  step {s:2d}: " + " | ".join(f"{names[i].split(':')[1].strip()[:14]}={(t[i] - base) / 1.9e3:6.2f}us" for i in range(16)))
    b10 = buf[10 * 64]
    print("  step 10, per k-block: empty seen by producer | published by producer | full seen by MMA  (us after step start)")
    for kb in range(16):
        print(f"    kb {kb:2d}: {(buf[10 * 64 + 48 + kb] - b10) / 1.9e3:6.2f} | {(buf[10 * 64 + 32 + kb] - b10) / 1.9e3:6.2f} | {(buf[10 * 64 + 16 + kb] - b10) / 1.9e3:6.2f}")
    print(f"  step period (start 10 -> start 11): {(buf[11 * 64] - buf[10 * 64]) / 1.9e3:.2f} us")

"""Timing of the CUDA-core (NCHW, exact fp32) vs tensor-core (NHWC, tf32) convolution kernels at the BASELINE config-2 shapes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from hulc_b200 import ops

def t(fn, n=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for (cin, cout, ks, st, hw, n) in [(3, 32, 8, 4, 200, N // 2), (32, 64, 4, 2, 49, N), (64, 64, 3, 1, 23, N), (3, 32, 8, 4, 84, N // 2), (32, 64, 4, 2, 20, N), (64, 64, 3, 1, 9, N)]:
    ho = (hw - ks) // st + 1
    x = torch.randn(n, cin, hw, hw, device="cuda"); w = torch.randn(cout, cin, ks, ks, device="cuda") * 0.05; b = torch.zeros(cout, device="cuda")
    xh = x if cin == 3 else x.permute(0, 2, 3, 1).contiguous()
    y = torch.empty(n, cout, ho, ho, device="cuda"); yh = torch.empty(n, ho, ho, cout, device="cuda")
    dy = torch.randn_like(y); dyh = torch.randn_like(yh); dw = torch.empty_like(w)
    dx = torch.empty_like(x); dxh = torch.empty_like(xh) if cin != 3 else None
    fl = 2.0 * n * ho * ho * cout * cin * ks * ks
    r = {}
    ms = t(lambda: ops.conv2d_fwd(x, w, b, st, y)); r["fwd simt"] = f"{ms:6.3f}ms {fl/ms/1e9:6.1f}TF"
    ms = t(lambda: ops.conv2d_tc_fwd(xh, w, b, st, yh)); r["fwd tc"] = f"{ms:6.3f}ms {fl/ms/1e9:6.1f}TF"
    ms = t(lambda: ops.conv2d_wgrad(x, dy, dw, st)); r["wgrad simt"] = f"{ms:6.3f}ms {fl/ms/1e9:6.1f}TF"
    ms = t(lambda: ops.conv2d_tc_wgrad(xh, dyh, dw, st)); r["wgrad tc"] = f"{ms:6.3f}ms {fl/ms/1e9:6.1f}TF"
    if cin != 3:
        ms = t(lambda: ops.conv2d_dgrad(dy, w, x.shape, st, gate=x, dx=dx)); r["dgrad simt"] = f"{ms:6.3f}ms {fl/ms/1e9:6.1f}TF"
        ms = t(lambda: ops.conv2d_tc_dgrad(dyh, w, dxh, st, gate=xh)); r["dgrad tc"] = f"{ms:6.3f}ms {fl/ms/1e9:6.1f}TF"
    torch.backends.cudnn.allow_tf32 = True
    xc = x.contiguous(memory_format=torch.channels_last)
    ms = t(lambda: torch.nn.functional.conv2d(xc, w, b, stride=st)); r["fwd cudnn"] = f"{ms:6.3f}ms {fl/ms/1e9:6.1f}TF"
    print((cin, cout, ks, st, hw, n), r, flush=True)

import sys, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
import torch, torch.nn.functional as F
from hulc_b200 import ops
torch.manual_seed(0)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm())
for (cin, cout, ks, st, hw, n) in [(3, 32, 8, 4, 200, 16), (3, 32, 8, 4, 200, 5), (32, 64, 4, 2, 49, 16), (64, 64, 3, 1, 23, 16)]:
    x = (torch.rand(n, cin, hw, hw, device="cuda") * 2 - 1)
    if cin != 3: x = x.relu()
    ho = (hw - ks) // st + 1
    dy = torch.randn(n, cout, ho, ho, device="cuda") * (torch.rand(n, cout, ho, ho, device="cuda") > 0.5)
    xh = x if cin == 3 else x.permute(0, 2, 3, 1).contiguous()
    dyh = dy.permute(0, 2, 3, 1).contiguous()
    w = torch.randn(cout, cin, ks, ks, device="cuda", dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, stride=st).backward(dy.double())
    dw = torch.zeros(cout, cin, ks, ks, device="cuda"); ops.conv2d_tc_wgrad(xh, dyh, dw, st)
    dw0 = torch.zeros(cout, cin, ks, ks, device="cuda"); ops.conv2d_wgrad(x, dy, dw0, st)
    # truncation-only reference
    tr = lambda t: (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
    w2 = torch.randn(cout, cin, ks, ks, device="cuda", dtype=torch.float64, requires_grad=True)
    F.conv2d(tr(x).double(), w2, stride=st).backward(tr(dy).double())
    print((cin, cout, ks, st, hw, n), "tc rel", rel(dw, w.grad), "simt rel", rel(dw0, w.grad), "trunc-model rel", rel(w2.grad, w.grad), "tc vs trunc-model", rel(dw, w2.grad))
    e = (dw.double() - w.grad).abs()
    print("   per-ci err norms", [float(e[:, c].norm() / w.grad[:, c].norm()) for c in range(min(cin, 3))], "per-ky", [round(float(e[:, :, k].norm() / w.grad[:, :, k].norm()), 4) for k in range(ks)])

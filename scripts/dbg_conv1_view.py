"""conv1 forward (3->32, 8x8, stride 4) on the tensor cores against torch fp64, with timing.
    python scripts/dbg_conv1_view.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from hulc_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(0)
for N, H in ((3, 200), (5, 84), (1024, 200), (1024, 84)):
    x = (torch.rand(N, 3, H, H, generator=g) * 2 - 1).cuda()
    w = (torch.randn(32, 3, 8, 8, generator=g) * 0.07).cuda()
    b = torch.randn(32, generator=g).cuda()
    HO = (H - 8) // 4 + 1
    y = torch.full((N, HO, HO, 32), float("nan"), device="cuda")
    bits = torch.zeros(N, HO, HO, 1, dtype=torch.int32, device="cuda")
    ops.conv2d_tc_fwd(x, w, b, 4, y, relu=True, relu_bits=bits)
    torch.cuda.synchronize()
    n_ref = min(N, 8)
    ref = torch.relu(torch.nn.functional.conv2d(x[:n_ref].double(), w.double(), b.double(), stride=4)).permute(0, 2, 3, 1)
    err = float((y[:n_ref].double() - ref).abs().max())
    nan = int(torch.isnan(y).sum())
    sign = ((y > 0).to(torch.int64) << torch.arange(32, device="cuda")).sum(-1)
    bits_ok = bool(torch.equal(sign & 0xFFFFFFFF, bits[..., 0].to(torch.int64) & 0xFFFFFFFF))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.conv2d_tc_fwd(x, w, b, 4, y, relu=True, relu_bits=bits)
    e0.record()
    for _ in range(10):
        ops.conv2d_tc_fwd(x, w, b, 4, y, relu=True, relu_bits=bits)
    e1.record()
    torch.cuda.synchronize()
    print(f"N={N} H={H}: max err {err:.3e} (ref max {float(ref.abs().max()):.2f}) nan {nan} bits_ok {bits_ok}  {e0.elapsed_time(e1) / 10 * 1e3:.1f} us")

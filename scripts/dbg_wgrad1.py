"""conv1 weight + bias gradient on the tensor cores against torch fp64, with timing.
    python scripts/dbg_wgrad1.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from hulc_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(1)
for n, hw in ((5, 200), (7, 84), (1024, 200), (1024, 84)):
    x = torch.randn(n, 3, hw, hw, generator=g).cuda()
    ho = (hw - 8) // 4 + 1
    dy = torch.randn(n, ho, ho, 32, generator=g).cuda()
    dw = torch.zeros(32, 3, 8, 8, device="cuda")
    db = torch.zeros(32, device="cuda")
    ops.conv2d_tc_wgrad(x, dy, dw, 4, db=db)
    torch.cuda.synchronize()
    m = min(n, 16)
    if m < n:
        dw.zero_(); db.zero_()
        ops.conv2d_tc_wgrad(x[:m].contiguous(), dy[:m].contiguous(), dw, 4, db=db)
    wref = torch.nn.grad.conv2d_weight(x[:m].double(), (32, 3, 8, 8), dy[:m].double().permute(0, 3, 1, 2).contiguous(), stride=4)
    bref = dy[:m].double().sum((0, 1, 2))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.conv2d_tc_wgrad(x, dy, dw, 4, db=db)
    e0.record()
    for _ in range(10):
        ops.conv2d_tc_wgrad(x, dy, dw, 4, db=db)
    e1.record()
    torch.cuda.synchronize()
    dw.zero_(); db.zero_()
    ops.conv2d_tc_wgrad(x[:m].contiguous(), dy[:m].contiguous(), dw, 4, db=db)
    print(f"n={n} hw={hw}: dw err {float((dw.double() - wref).abs().max()):.3e} (scale {float(wref.abs().max()):.1f})  "
          f"db err {float((db.double() - bref).abs().max()):.3e} (scale {float(bref.abs().max()):.1f})  {e0.elapsed_time(e1) / 10 * 1e3:.1f} us (incl. reduce)")

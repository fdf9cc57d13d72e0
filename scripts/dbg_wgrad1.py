import sys, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.nn.functional as F
from hulc_b200 import ops
n, hw = 5, 200
g = torch.Generator().manual_seed(1)
x = torch.randn(n, 3, hw, hw, generator=g).cuda()
ho = (hw - 8) // 4 + 1
dy = torch.randn(n, ho, ho, 32, generator=g).cuda()
dw = torch.zeros(32, 3, 8, 8, device="cuda"); db = torch.zeros(32, device="cuda")
ops.conv2d_tc_wgrad(x, dy, dw, 4, db=db)
torch.cuda.synchronize()
print("db", db[:6].tolist())
print("ref", dy.sum((0, 1, 2))[:6].tolist())
ws = ops.workspace(x.device)
ctas = min(148, n * ho)
part = ws[1024 + ctas * 192 * 32 : 1024 + ctas * 192 * 32 + ctas * 32].view(ctas, 32)
print("bias partial rows 0..2", part[:3, :4].tolist(), "sum", part.sum(0)[:4].tolist())
wpart = ws[1024 : 1024 + ctas * 192 * 32].view(ctas, 192, 32)
print("dw partial row 191 cta0", wpart[0, 191, :4].tolist())
xr = x.double().requires_grad_(False)
wref = torch.nn.grad.conv2d_weight(x, (32, 3, 8, 8), dy.permute(0, 3, 1, 2).contiguous(), stride=4)
print("dw err", float((dw - wref).abs().max()), "scale", float(wref.abs().max()), "dw absmax", float(dw.abs().max()))
print("launch count", ops.launch_count())

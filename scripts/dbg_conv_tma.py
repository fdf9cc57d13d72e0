import sys, math
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch, torch.nn.functional as F
from hulc_b200 import ops
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
for (cin, cout, ks, st, hw, n) in [(32, 64, 4, 2, 49, 40), (32, 64, 4, 2, 49, 300), (64, 64, 3, 1, 23, 300)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, hw, hw, generator=g).cuda().relu()
    w = (torch.randn(cout, cin, ks, ks, generator=g) / math.sqrt(cin * ks * ks)).cuda()
    b = torch.randn(cout, generator=g).cuda()
    ref = F.relu(F.conv2d(x, w, b, stride=st))
    ho = ref.shape[-1]
    try:
        y = ops.conv2d_tc_fwd(nhwc(x), w, b, st, torch.empty(n, ho, ho, cout, device="cuda"))
        torch.cuda.synchronize()
        print((cin, ks, n), "fwd ok, err", float((y - nhwc(ref)).abs().max()))
    except Exception as e:
        print((cin, ks, n), "fwd FAILED", str(e)[:100]); break
    dy = torch.randn_like(ref)
    try:
        dx = ops.conv2d_tc_dgrad(nhwc(dy), w, torch.empty(n, hw, hw, cin, device="cuda"), st)
        torch.cuda.synchronize()
        refdx = torch.nn.grad.conv2d_input(x.shape, w, dy, stride=st)
        print((cin, ks, n), "dgrad ok, err", float((dx - nhwc(refdx)).abs().max()))
    except Exception as e:
        print((cin, ks, n), "dgrad FAILED", str(e)[:100]); break

import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import bench
from hulc_b200.engine import HulcEngine
from hulc_b200.utils import synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
eng = HulcEngine("hulc", "rnn_decoder", device="cuda", dropout_p=0.1)
eng.load_state_dict(synthetic.make_state_dict("hulc"))
batch = synthetic.make_batch(B, 32, seed=1, device="cuda")
for i in range(2):
    out = eng.step(batch, seed=i)
    eng.optimizer_step()
torch.cuda.synchronize()
print("steps ok", float(out["total_loss"]))
r = bench.dominant_kernel_roofline(torch, eng, batch, bench.measured_peaks(), 40.0)
print(r["kernel"], r["kernels"])

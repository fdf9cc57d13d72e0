"""Run a few full-size training steps (BASELINE config 2 shape) — the target command for ncu captures.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --steps 2
"""
import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from hulc_b200.engine import HulcEngine  # noqa: E402
from hulc_b200.utils import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--seq", type=int, default=32)
ap.add_argument("--model", default="hulc")
ap.add_argument("--rnn", default="rnn_decoder")
ap.add_argument("--precision", default="tf32")
args = ap.parse_args()
eng = HulcEngine(args.model, args.rnn, device="cuda", dropout_p=0.1, precision=args.precision)
eng.load_state_dict(synthetic.make_state_dict(args.model, args.rnn))
batch = synthetic.make_batch(args.batch, args.seq, seed=1, device="cuda")
for i in range(args.steps):
    torch.cuda.synchronize()
    out = eng.step(batch, seed=i)
    eng.optimizer_step()
torch.cuda.synchronize()
print("loss", float(out["total_loss"]))

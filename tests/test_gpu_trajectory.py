"""Loss-curve parity: K = 20 optimizer steps of the engine in its tensor-core modes against the oracle driven by
torch.optim.Adam on the same batches, weights and sampled plan classes.  Backs the claim that the tf32 / bf16 gradient error is
unbiased noise: a biased gradient would bend the trajectory away from the reference's within a few steps of Adam (which normalises every
coordinate to steps of ~lr regardless of the gradient's size)."""
import numpy as np
import pytest
import torch

from hulc_b200.utils import synthetic
from oracle import hulc_oracle as O

pytestmark = pytest.mark.gpu
K, B, S = 20, 2, 8


@pytest.fixture(scope="module")
def reference_run():
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked tests need a CUDA device; hulc_b200 has no CPU fallback")
    sd = synthetic.make_state_dict("hulc")
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=2e-4)
    batches = [synthetic.make_batch(B, S, seed=1 + i) for i in range(4)]
    noise = [{m: synthetic.plan_noise(B, S, m, seed=1 + i) for m in batches[0]} for i in range(K)]
    losses, idx = [], []
    for i in range(K):
        opt.zero_grad(set_to_none=True)
        out = O.training_step(params, batches[i % 4], plan_u={m: noise[i][m]["u"] for m in batches[0]})
        out["total_loss"].backward()
        opt.step()
        losses.append(float(out["total_loss"].detach()))
        idx.append({m: out[f"plan_idx_{m}"].clone() for m in batches[0]})
    return sd, batches, idx, losses


# (median, max) relative deviation of the loss allowed over the K steps.  Adam's first steps move EVERY coordinate by ~lr whatever the size of
# its gradient, so coordinates whose gradient is rounding noise take sign-random steps in any arithmetic (fp32 included); with two sequences per
# modality the loss itself moves 2-8 % per step, and single steps of the two runs differ visibly.  What must hold: the deviations stay a
# small fraction of the loss's own movement and do not grow along the trajectory (a biased gradient would make them grow).
# Measured (B200): fp32 median 1.4e-4 / max 3.2e-4; tf32 1.3e-2 / 6.4e-2; bf16 9.5e-3 / 9.2e-2 — first-five vs last-five means 1.06e-2 vs 1.03e-2
# (tf32) and 6.3e-3 vs 7.8e-3 (bf16): the spikes sit at the same steps in both modes (the batch whose loss the reference itself moves by 8 %).
@pytest.mark.parametrize("precision,tol", [("fp32", (1e-3, 3e-3)), ("tf32", (3e-2, 1.5e-1)), ("bf16", (3e-2, 1.5e-1))])
def test_loss_trajectory_follows_the_reference(reference_run, precision, tol):
    from hulc_b200.engine import HulcEngine

    sd, batches, idx, ref = reference_run
    eng = HulcEngine("hulc", "rnn_decoder", device="cuda", dropout_p=0.0, precision=precision)
    eng.load_state_dict(sd)
    dev_batches = [synthetic._to(b, "cuda") for b in batches]
    losses = []
    for i in range(K):
        out = eng.step(dev_batches[i % 4], plan_idx={m: v.cuda() for m, v in idx[i].items()})
        eng.optimizer_step()
        losses.append(out["total_loss"].item())
    eng.check_nan_flag()
    rel = np.abs(np.array(losses) - np.array(ref)) / np.abs(np.array(ref))
    print("TRAJ", precision, "median / max relative deviation of the loss over", K, "steps:", float(np.median(rel)), rel.max(), "at step", int(rel.argmax()),
          "| first five / last five mean:", rel[:5].mean(), rel[-5:].mean(), "| reference first/last loss", ref[0], ref[-1], "| deviations", np.round(rel, 5).tolist())
    assert min(ref[-4:]) < min(ref[:4])  # the reference itself learns on these batches
    assert float(np.median(rel)) < tol[0] and rel.max() < tol[1], (float(np.median(rel)), rel.max())
    # no drift: the deviation over the last five steps is not larger than 3x that over the first five (+ the mode's resolution)
    assert rel[-5:].mean() <= 3 * rel[:5].mean() + tol[0]

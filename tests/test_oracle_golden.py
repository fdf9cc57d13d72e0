"""The oracle (oracle/hulc_oracle.py) against the fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py → tests/golden/*.npz).  CPU only."""
import numpy as np
import pytest
import torch

from hulc_b200.utils import synthetic
from oracle import hulc_oracle as O

CASES = {
    "hulc_b2s8": ("hulc", "rnn_decoder", 2, 8, 0.0),
    "hulc_b2s8_drop": ("hulc", "rnn_decoder", 2, 8, 0.1),
    "hulc_gru_b2s8": ("hulc", "gru_decoder", 2, 8, 0.0),
    "gcbc_b2s8": ("gcbc", "rnn_decoder", 2, 8, 0.0),
    "mcil_b2s8": ("mcil", "rnn_decoder", 2, 8, 0.0),
    "hulc_aux_b4s32": ("hulc", "rnn_decoder", 4, 32, 0.0),  # BC-Z + MIA auxiliary heads on (ablation configs, hulc.py:567-648)
}


def run_oracle(model, rnn_model, B, S, p, plan_idx=None, aux=False):
    from hulc_b200.spec import ModelDims

    sd = synthetic.make_state_dict(model, rnn_model, dims=ModelDims.shipped(model, rnn_model, bc_z=True, mia=True) if aux else None)
    for v in sd.values():
        v.requires_grad_(True)
    batch = synthetic.make_batch(B, S, seed=1)
    noise = {m: synthetic.plan_noise(B, S, m) for m in batch}
    masks = {m: synthetic.dropout_masks(B, S, m, p) for m in batch} if p > 0 else None
    out = O.training_step(
        sd, batch, model=model, rnn_model=rnn_model, dropout_p=p, plan_idx=plan_idx,
        plan_u={m: noise[m]["u"] for m in batch}, plan_eps={m: noise[m]["eps"] for m in batch}, dropout_masks=masks,
        bc_z_beta=1.0 if aux else None, mia_beta=1.0 if aux else None,
    )
    out["total_loss"].backward()
    return sd, out


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_fixture(name, golden_dir):
    model, rnn_model, B, S, p = CASES[name]
    fx = np.load(golden_dir / f"{name}.npz")
    sd, out = run_oracle(model, rnn_model, B, S, p, aux="_aux_" in name)
    np.testing.assert_allclose(out["total_loss"].item(), fx["total_loss"], rtol=1e-5, atol=1e-6)
    if "pred_lang" in fx.files:  # the reference logs beta * loss (beta = 1, conf/loss/default.yaml)
        np.testing.assert_allclose(out["lang_pred_loss"].item(), fx["pred_lang"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(out["lang_contrastive_loss"].item(), fx["lang_contrastive"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["action_loss"].item(), fx["action_loss"], rtol=1e-5, atol=1e-6)
    if "kl_loss" in fx:
        np.testing.assert_allclose(out["kl_loss"].item(), fx["kl_loss"], rtol=1e-5, atol=1e-7)
    if "lang_clip_loss" in fx:
        np.testing.assert_allclose(3.0 * out["lang_clip_loss"].item(), fx["lang_clip_loss"], rtol=1e-5, atol=1e-6)
    for m in ("vis", "lang"):
        for k in ("logit_probs", "log_scales", "means", "gripper_act", "actions_tcp"):
            key = f"{k}_{m}"
            if key in fx.files:
                ref = fx[key]
                np.testing.assert_allclose(out[key].detach().numpy()[: ref.shape[0]], ref, rtol=1e-4, atol=2e-5, err_msg=key)
        if f"plan_idx_{m}" in fx.files:
            assert np.array_equal(out[f"plan_idx_{m}"].numpy(), fx[f"plan_idx_{m}"])
    # every parameter gradient: norm + leading entries; parameters the reference leaves without a gradient stay so
    n_checked = 0
    for k, v in sd.items():
        gn = float(fx[f"gradnorm/{k}"])
        if gn < 0:
            assert v.grad is None or float(v.grad.abs().max()) == 0.0, k
            continue
        np.testing.assert_allclose(float(v.grad.norm()), gn, rtol=2e-4, atol=1e-8, err_msg=k)
        head = fx[f"gradhead/{k}"]
        np.testing.assert_allclose(v.grad.reshape(-1)[: head.size].numpy(), head, rtol=1e-3, atol=1e-6 + 1e-4 * gn, err_msg=k)
        n_checked += 1
    assert n_checked > 50


def test_param_spec_matches_reference_counts():
    """Parameter totals quoted from the reference in SURVEY.md §8c."""
    count = lambda **kw: sum(int(np.prod(s)) for s in synthetic.param_spec(**kw).values())
    assert count(model="hulc") == 47_053_559
    assert count(model="hulc", rnn_model="gru_decoder") == 76_823_287
    assert count(model="mcil") == 74_362_066
    assert count(model="gcbc") == 44_956_407


@pytest.mark.slow
def test_oracle_full_window_fixture(golden_dir):
    fx = np.load(golden_dir / "hulc_b4s32.npz")
    _, out = run_oracle("hulc", "rnn_decoder", 4, 32, 0.0)
    np.testing.assert_allclose(out["total_loss"].item(), fx["total_loss"], rtol=1e-5)


def test_tcp_frame_for_zero_base_rotation():
    a = torch.rand(2, 3, 7) * 2 - 1
    obs = torch.zeros(2, 3, 15)
    out = O.world_to_tcp_frame(a, obs)
    torch.testing.assert_close(out[..., :3], a[..., :3])
    # tcp_new_T_tcp_old = R(0.01 a)^-1: to first order the negated relative rotation
    torch.testing.assert_close(out[..., 3:6], -a[..., 3:6], rtol=0, atol=2e-2)
    torch.testing.assert_close(out[..., 6], a[..., 6])

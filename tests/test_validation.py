"""Validation path: Hulc.validation_step / lmp_val (hulc/models/hulc.py:301-388, 739-841).

* the oracle's restatement against the fixtures written from the UNMODIFIED reference (oracle/make_golden.py, VAL_CASES);
* HulcEngine.validation_step on the host SIMT emulator against the oracle (reduced frames);
* (gpu) HulcEngine.validation_step on the B200 against the same fixtures at the reference's frame sizes, both precision modes.
Randomness (latent-plan sample, Gumbel / inverse-CDF uniforms of LogisticDecoderRNN._sample) is injected from seeds."""
from pathlib import Path

import numpy as np
import pytest
import torch

from hulc_b200.utils import synthetic
from oracle import hulc_oracle as O
from oracle.make_golden import VAL_CASES

GOLDEN = Path(__file__).resolve().parent / "golden"
RTOL, ATOL = 1e-3, 1e-4


def _inputs(model, B, S, fx=None, hw=(200, 84)):
    batch = synthetic.make_batch(B, S, seed=1, static_hw=hw[0], gripper_hw=hw[1])
    n_dims = 6 if model != "mcil" else 7
    noise = {w: {m: synthetic.validation_noise(B, S, m, w, n_dims=n_dims) for m in batch} for w in ("pp", "pr")}
    kw = dict(sample_u={w: {m: (noise[w][m]["u_mix"], noise[w][m]["u_inv"]) for m in batch} for w in ("pp", "pr")})
    if model == "mcil":
        kw["plan_eps"] = {w: {m: noise[w][m]["eps"] for m in batch} for w in ("pp", "pr")}
    elif fx is not None and model != "gcbc":
        kw["plan_idx"] = {w: {m: torch.from_numpy(fx[f"plan_idx_{w}_{m}"]) for m in batch} for w in ("pp", "pr")}
    return batch, noise, kw


def _check(out, fx, mods, S, mae_atol=2e-3, sr_slack=1):
    cpu = lambda t: t.detach().float().cpu().numpy()
    for m in mods:
        for w in ("pp", "pr"):
            np.testing.assert_allclose(cpu(out[f"action_loss_{w}_{m}"]), fx[f"action_loss_{w}_{m}"], rtol=RTOL, atol=ATOL, err_msg=f"action_loss_{w}_{m}")
            # the sampled actions pass through exp(log_scale) * logit(u) and the frame change: absolute tolerance on the L1 error
            np.testing.assert_allclose(cpu(out[f"mae_{w}_{m}"]), fx[f"mae_{w}_{m}"], rtol=RTOL, atol=mae_atol, err_msg=f"mae_{w}_{m}")
            assert abs(float(out[f"gripper_sr_{w}_{m}"]) - float(fx[f"gripper_sr_{w}_{m}"])) <= sr_slack / S + 1e-6, f"gripper_sr_{w}_{m}"
        np.testing.assert_allclose(cpu(out[f"kl_loss_{m}"]), fx[f"kl_loss_{m}"], rtol=RTOL, atol=1e-6, err_msg=f"kl_loss_{m}")
    if "val_pred_clip_loss" in fx.files:
        np.testing.assert_allclose(cpu(out["val_pred_clip_loss"]), fx["val_pred_clip_loss"], rtol=RTOL, atol=ATOL)


def _check_gcbc(out, fx, mods, S, mae_atol=2e-3, sr_slack=1):
    cpu = lambda t: t.detach().float().cpu().numpy()
    for m in mods:
        np.testing.assert_allclose(cpu(out[f"action_loss_{m}"]), fx[f"action_loss_{m}"], rtol=RTOL, atol=ATOL, err_msg=f"action_loss_{m}")
        np.testing.assert_allclose(cpu(out[f"mae_{m}"]), fx[f"mae_{m}"], rtol=RTOL, atol=mae_atol, err_msg=f"mae_{m}")
        assert abs(float(out[f"gripper_sr_{m}"]) - float(fx[f"gripper_sr_{m}"])) <= sr_slack / S + 1e-6
    np.testing.assert_allclose(cpu(out["val_pred_clip_loss"]), fx["val_pred_clip_loss"], rtol=RTOL, atol=ATOL)


@pytest.mark.parametrize("name", list(VAL_CASES))
def test_oracle_validation_matches_reference_fixture(name):
    model, rnn_model, B, S = VAL_CASES[name]
    if B * S > 64:
        pytest.skip("the CPU suite keeps to the small case; the full window is checked on the GPU")
    fx = np.load(GOLDEN / f"{name}.npz")
    batch, noise, kw = _inputs(model, B, S, fx)
    sd = synthetic.make_state_dict(model, rnn_model)
    if "plan_idx" in kw:
        kw["plan_idx"] = {w: {m: v.long() for m, v in d.items()} for w, d in kw["plan_idx"].items()}
    out = O.validation_step(sd, batch, model=model, rnn_model=rnn_model, **kw)
    if model == "gcbc":
        _check_gcbc(out, fx, list(batch), S, mae_atol=1e-4)
        return
    _check(out, fx, list(batch), S, mae_atol=1e-4)
    # the logged aggregates of validation_step (hulc.py:806-829)
    for m in batch:
        np.testing.assert_allclose(float(out[f"mae_pp_{m}"][..., :3].mean()), fx[f"logged/val_pos_mae/{m}_pos_mae_pp"], rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(float(out[f"mae_pr_{m}"].mean()), fx[f"logged/val_total_mae/{m}_total_mae_pr"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("model", ["hulc", "mcil", "gcbc"])
def test_emu_validation_step_matches_oracle(emu, monkeypatch, model):
    from hulc_b200 import engine
    from hulc_b200.engine import HulcEngine, ParamStore

    monkeypatch.setattr(engine, "_POISON", True)
    B, S, hw = 2, 4, (64, 44)
    sd = synthetic.make_state_dict(model, "rnn_decoder")
    k = ((((hw[1] - 8) // 4 + 1) - 4) // 2 + 1) - 2
    key = "perceptual_encoder.rgb_gripper_encoder.conv_model.7.weight"
    sd[key] = sd[key][:, : 64 * k * k].contiguous()
    batch, noise, kw = _inputs(model, B, S, hw=hw)
    eng = HulcEngine(model, "rnn_decoder", device="cpu", dropout_p=0.1, precision="fp32")
    eng.spec[key] = tuple(sd[key].shape)
    eng.ps = ParamStore(eng.spec, "cpu")
    eng.load_state_dict(sd)
    if model == "gcbc":
        out = eng.validation_step(batch, sample_u={"pr": kw["sample_u"]["pr"]})
        ref = O.validation_step(sd, batch, model="gcbc", sample_u={"pr": kw["sample_u"]["pr"]})
        for m in batch:
            torch.testing.assert_close(out[f"action_loss_{m}"], ref[f"action_loss_{m}"], rtol=RTOL, atol=ATOL)
            torch.testing.assert_close(out[f"sample_act_{m}"], ref[f"sample_act_{m}"], rtol=1e-3, atol=2e-3)
            torch.testing.assert_close(out[f"mae_{m}"], ref[f"mae_{m}"], rtol=1e-3, atol=1e-3)
            assert abs(float(out[f"gripper_sr_{m}"]) - float(ref[f"gripper_sr_{m}"])) < 1e-6
        torch.testing.assert_close(out["val_pred_clip_loss"], ref["val_pred_clip_loss"], rtol=RTOL, atol=ATOL)
        return
    if model == "hulc":  # the engine samples by inverse CDF from injected uniforms; hand the same classes to the oracle
        o1 = eng.step(batch, backward=False, plan_u={m: noise["pr"][m]["u"] for m in batch})
        idx_pr = o1["plan_idx"].clone()
        o2 = eng.step(batch, backward=False, plan_u={m: noise["pp"][m]["u"] for m in batch}, plan_from="prior")
        idx_pp = o2["plan_idx"].clone()
        kw["plan_idx"] = {"pp": {m: idx_pp[i * B : (i + 1) * B] for i, m in enumerate(batch)}, "pr": {m: idx_pr[i * B : (i + 1) * B] for i, m in enumerate(batch)}}
    p_before = eng.dropout_p
    out = eng.validation_step(batch, **kw)
    assert eng.dropout_p == p_before
    okw = dict(kw)
    if model == "hulc":
        okw["plan_idx"] = {w: {m: v.long() for m, v in d.items()} for w, d in kw["plan_idx"].items()}
    ref = O.validation_step(sd, batch, model=model, rnn_model="rnn_decoder", **okw)
    for m in batch:
        for w in ("pp", "pr"):
            torch.testing.assert_close(out[f"action_loss_{w}_{m}"], ref[f"action_loss_{w}_{m}"], rtol=RTOL, atol=ATOL)
            torch.testing.assert_close(out[f"sample_act_{w}_{m}"], ref[f"sample_act_{w}_{m}"], rtol=1e-3, atol=2e-3)
            torch.testing.assert_close(out[f"mae_{w}_{m}"], ref[f"mae_{w}_{m}"], rtol=1e-3, atol=1e-3)
            assert abs(float(out[f"gripper_sr_{w}_{m}"]) - float(ref[f"gripper_sr_{w}_{m}"])) < 1e-6
            torch.testing.assert_close(out[f"sampled_plan_{w}_{m}"], ref[f"sampled_plan_{w}_{m}"], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(out[f"kl_loss_{m}"], ref[f"kl_loss_{m}"], rtol=RTOL, atol=1e-6)
    if model == "hulc":
        torch.testing.assert_close(out["val_pred_clip_loss"], ref["val_pred_clip_loss"], rtol=RTOL, atol=ATOL)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(VAL_CASES))
@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_gpu_validation_step_matches_reference_fixture(name, precision):
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    from hulc_b200.engine import HulcEngine

    model, rnn_model, B, S = VAL_CASES[name]
    fx = np.load(GOLDEN / f"{name}.npz")
    batch, noise, kw = _inputs(model, B, S, fx)
    eng = HulcEngine(model, rnn_model, device="cuda", dropout_p=0.1, precision=precision)
    eng.load_state_dict(synthetic.make_state_dict(model, rnn_model))
    to = lambda t: t.cuda()
    dkw = {"sample_u": {w: {m: (to(a), to(b)) for m, (a, b) in d.items()} for w, d in kw["sample_u"].items()}}
    if "plan_eps" in kw:
        dkw["plan_eps"] = {w: {m: to(v) for m, v in d.items()} for w, d in kw["plan_eps"].items()}
    if "plan_idx" in kw:
        dkw["plan_idx"] = {w: {m: to(v) for m, v in d.items()} for w, d in kw["plan_idx"].items()}
    if model == "gcbc":
        dkw = {"sample_u": {"pr": dkw["sample_u"]["pr"]}}
    out = eng.validation_step(synthetic._to(batch, "cuda"), **dkw)
    eng.check_nan_flag()
    if model == "gcbc":
        _check_gcbc(out, fx, list(batch), S, mae_atol=0.3 if precision == "tf32" else 2e-3, sr_slack=3 if precision == "tf32" else 1)
        return
    # tf32 products move the action logits by up to ~1e-4: enough to flip a Gumbel-max choice between two near-tied mixture components
    # for a handful of the B*S*6 draws, which moves that sample (not the distribution) by O(1) -> the L1 error is only held loosely there;
    # the exact-fp32 mode pins the whole chain (sampling kernel, frame change, reductions) tightly.
    _check(out, fx, list(batch), S, mae_atol=0.3 if precision == "tf32" else 2e-3, sr_slack=3 if precision == "tf32" else 1)


@pytest.mark.gpu
def test_gpu_module_validation_step_logs_reference_keys():
    """hulc_b200.models.hulc.Hulc.validation_step: every key the reference logs in validation_step (hulc.py:806-834; all but the
    dataset-bound clip_groundtruth metrics) with the reference's values, and the returned sampled plans / episode indices."""
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    from hulc_b200.models.hulc import Hulc

    name = "val_hulc_b2s8"
    model_name, rnn_model, B, S = VAL_CASES[name]
    fx = np.load(GOLDEN / f"{name}.npz")
    batch, noise, kw = _inputs(model_name, B, S, fx)
    cfg = synthetic.model_config("hulc", target_root="hulc_b200")
    cfg.pop("_target_"); cfg.pop("_recursive_")
    model = Hulc(**cfg, device=torch.device("cuda"), precision="fp32")  # exact products: no near-tie flips in the sampled actions (see above)
    model.load_state_dict(synthetic.make_state_dict("hulc"), strict=False)
    to = lambda t: t.cuda()
    out = model.validation_step(
        synthetic._to(batch, "cuda"), 0,
        sample_u={w: {m: (to(a), to(b)) for m, (a, b) in d.items()} for w, d in kw["sample_u"].items()},
        plan_idx={w: {m: to(v) for m, v in d.items()} for w, d in kw["plan_idx"].items()})
    ref_keys = [k[len("logged/"):] for k in fx.files if k.startswith("logged/")]
    assert ref_keys, "fixture holds the reference's logged values"
    for k in ref_keys:
        assert k in model.logged, f"reference logs {k}"
        tol = 5e-3 if "mae" in k else 1e-4
        np.testing.assert_allclose(float(model.logged[k]), float(fx["logged/" + k]), rtol=1e-3, atol=tol if "grip" not in k else 1.0 / S + 1e-6, err_msg=k)
    for m in batch:
        assert torch.equal(out[f"idx_{m}"].cpu(), batch[m]["idx"])
        for w in ("pp", "pr"):
            plan = out[f"sampled_plan_{w}_{m}"].view(B, 32, 32)
            assert torch.equal(plan.argmax(-1).cpu(), torch.from_numpy(fx[f"plan_idx_{w}_{m}"]).long())


def test_emu_module_validation_logs_reference_values(emu):
    """The same module-level check on the host emulator, at the reference's frame sizes (B=2, S=8): all 24 values the reference logs."""
    from hulc_b200.models.hulc import Hulc

    name = "val_hulc_b2s8"
    model_name, rnn_model, B, S = VAL_CASES[name]
    fx = np.load(GOLDEN / f"{name}.npz")
    batch, noise, kw = _inputs(model_name, B, S, fx)
    cfg = synthetic.model_config("hulc", target_root="hulc_b200")
    cfg.pop("_target_"); cfg.pop("_recursive_")
    model = Hulc(**cfg, device=torch.device("cpu"), precision="fp32")  # exact-fp32 kernels: the tcgen05 ones are not emulated
    model.load_state_dict(synthetic.make_state_dict("hulc"), strict=False)
    out = model.validation_step(batch, 0, sample_u=kw["sample_u"], plan_idx=kw["plan_idx"])
    ref_keys = [k[len("logged/"):] for k in fx.files if k.startswith("logged/")]
    assert len(ref_keys) == 24
    for k in ref_keys:
        np.testing.assert_allclose(float(model.logged[k]), float(fx["logged/" + k]), rtol=1e-3, atol=1e-4, err_msg=k)
    for m in batch:
        assert torch.equal(out[f"idx_{m}"], batch[m]["idx"])
        for w in ("pp", "pr"):
            assert torch.equal(out[f"sampled_plan_{w}_{m}"].view(B, 32, 32).argmax(-1), torch.from_numpy(fx[f"plan_idx_{w}_{m}"]).long())


def test_emu_gcbc_module_validation_logs_reference_values(emu):
    """GCBC.validation_step (gcbc.py:183-281) through hulc_b200.models.gcbc.GCBC on the emulator: the reference's logged keys and values."""
    from hulc_b200.models.gcbc import GCBC

    name = "val_gcbc_b2s8"
    _, rnn_model, B, S = VAL_CASES[name]
    fx = np.load(GOLDEN / f"{name}.npz")
    batch, noise, kw = _inputs("gcbc", B, S, fx)
    cfg = synthetic.model_config("gcbc", target_root="hulc_b200")
    cfg.pop("_target_"); cfg.pop("_recursive_")
    model = GCBC(**cfg, device=torch.device("cpu"), precision="fp32")
    model.load_state_dict(synthetic.make_state_dict("gcbc"), strict=False)
    out = model.validation_step(batch, 0, sample_u={"pr": kw["sample_u"]["pr"]})
    ref_keys = [k[len("logged/"):] for k in fx.files if k.startswith("logged/")]
    assert len(ref_keys) == 12
    for k in ref_keys:
        np.testing.assert_allclose(float(model.logged[k]), float(fx["logged/" + k]), rtol=1e-3, atol=1e-4, err_msg=k)
    assert all(torch.equal(out[f"idx_{m}"], batch[m]["idx"]) for m in batch)

"""Inference path: Hulc.step / get_pp_plan_{lang,vision} / predict_with_plan (hulc/models/hulc.py:843-957) and
LogisticDecoderRNN.act with the carried hidden state (logistic_decoder_rnn.py:104-119).

Fixtures (`tests/golden/infer_*.npz`) hold the actions of rollouts of the UNMODIFIED reference (oracle/make_golden.py, INFER_CASES: a language
goal re-planned every 3 steps over 7 steps, a goal image re-planned every 4 over 5) with the randomness injected from seeds.
* the oracle's restatement against the fixtures; * hulc_b200.models.hulc.Hulc.step on the host emulator; * (gpu) on the B200."""
from pathlib import Path

import numpy as np
import pytest
import torch

from hulc_b200.utils import synthetic
from oracle import hulc_oracle as O
from oracle.make_golden import INFER_CASES

GOLDEN = Path(__file__).resolve().parent / "golden"


def _obs(r, t, dev="cpu"):
    return {"rgb_obs": {"rgb_static": r["rgb_static"][t][None, None].to(dev), "rgb_gripper": r["rgb_gripper"][t][None, None].to(dev)}, "depth_obs": {},
            "robot_obs": r["robot_obs"][t][None, None].to(dev), "robot_obs_raw": r["robot_obs_raw"][t][None, None].to(dev)}


def _goal(r, kind, dev="cpu"):
    if kind == "lang":
        return "the task"
    return {"rgb_obs": {"rgb_static": r["goal_static"][None].to(dev), "rgb_gripper": r["goal_gripper"][None].to(dev)}, "depth_obs": {},
            "robot_obs": r["goal_robot_obs"][None].to(dev)}


def _rollout(model, name, dev):
    kind, T, replan = INFER_CASES[name]
    fx = np.load(GOLDEN / f"{name}.npz")
    r = synthetic.rollout_inputs(T, kind)
    model.replan_freq = replan
    model.lang_embeddings = {"the task": r["lang"].numpy()[None]}
    model.reset()
    acts = []
    for t in range(T):
        idx = torch.from_numpy(fx["plan_idx"][t // replan]).to(dev) if t % replan == 0 else None
        a = model.step(_obs(r, t, dev), _goal(r, kind, dev), plan_idx=idx, sample_u=(r["u_mix"][t].to(dev), r["u_inv"][t].to(dev)))
        assert tuple(a.shape) == (1, 1, 7)
        acts.append(a.reshape(7).float().cpu())
    assert model.rollout_step_counter == T
    return torch.stack(acts).numpy(), fx["actions"]


@pytest.mark.parametrize("name", list(INFER_CASES))
def test_oracle_rollout_matches_reference_fixture(name):
    kind, T, replan = INFER_CASES[name]
    fx = np.load(GOLDEN / f"{name}.npz")
    r = synthetic.rollout_inputs(T, kind)
    sd = synthetic.make_state_dict("hulc")
    hidden = plan = goal = None
    for t in range(T):
        if t % replan == 0:
            idx = torch.from_numpy(fx["plan_idx"][t // replan])[None]
            if kind == "lang":
                plan, goal, _ = O.inference_plan(sd, r["rgb_static"][t : t + 1], r["rgb_gripper"][t : t + 1], lang=r["lang"], plan_idx=idx)
            else:
                plan, goal, _ = O.inference_plan(sd, torch.cat([r["rgb_static"][t : t + 1], r["goal_static"]]), torch.cat([r["rgb_gripper"][t : t + 1], r["goal_gripper"]]),
                                                 plan_idx=idx)
            hidden = torch.zeros(2, 1, 2048)
        a, hidden = O.inference_act(sd, r["rgb_static"][t : t + 1], r["rgb_gripper"][t : t + 1], r["robot_obs_raw"][t : t + 1], plan, goal, hidden, r["u_mix"][t], r["u_inv"][t])
        np.testing.assert_allclose(a.reshape(7).numpy(), fx["actions"][t], rtol=1e-4, atol=1e-4, err_msg=f"step {t}")


@pytest.mark.parametrize("name", list(INFER_CASES))
def test_emu_module_rollout_matches_reference_fixture(emu, name):
    from hulc_b200.models.hulc import Hulc

    cfg = synthetic.model_config("hulc", target_root="hulc_b200")
    cfg.pop("_target_"); cfg.pop("_recursive_")
    model = Hulc(**cfg, device=torch.device("cpu"), precision="fp32")
    model.load_state_dict(synthetic.make_state_dict("hulc"), strict=False)
    got, want = _rollout(model, name, "cpu")
    np.testing.assert_allclose(got, want, rtol=1e-3, atol=2e-3)
    with pytest.raises(RuntimeError):
        model.reset()
        model.engine.infer_act(torch.zeros(1, 3, 200, 200), torch.zeros(1, 3, 84, 84), torch.zeros(1, 15))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(INFER_CASES))
@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_gpu_module_rollout_matches_reference_fixture(name, precision):
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    from hulc_b200.models.hulc import Hulc

    cfg = synthetic.model_config("hulc", target_root="hulc_b200")
    cfg.pop("_target_"); cfg.pop("_recursive_")
    model = Hulc(**cfg, device=torch.device("cuda"), precision=precision)
    model.load_state_dict(synthetic.make_state_dict("hulc"), strict=False)
    got, want = _rollout(model, name, "cuda")
    model.engine.check_nan_flag()
    if precision == "fp32":
        np.testing.assert_allclose(got, want, rtol=1e-3, atol=2e-3)
    else:
        # tf32 products can flip a Gumbel-max choice between near-tied mixture components (see tests/test_validation.py): most sampled
        # dimensions must agree tightly, none may be non-finite
        ok = np.isclose(got, want, rtol=2e-2, atol=2e-2)
        assert np.isfinite(got).all() and ok.mean() >= 0.9, f"{ok.mean():.2f} of the sampled action dims agree"


@pytest.mark.parametrize("model,rnn_model", [("hulc", "gru_decoder"), ("mcil", "rnn_decoder"), ("gcbc", "rnn_decoder")])
def test_emu_engine_rollout_variants_match_oracle(emu, monkeypatch, model, rnn_model):
    """GRU decoder, MCIL (continuous plan, 7 logistic dims, no frame change) and GCBC (no plan) through HulcEngine.infer_plan / infer_act on the
    emulator against the oracle's restatement (which the two reference rollouts above pin for the default model).  Reduced frames."""
    from hulc_b200 import engine
    from hulc_b200.engine import HulcEngine, ParamStore

    monkeypatch.setattr(engine, "_POISON", True)
    hw, T, replan = (64, 44), 4, 2
    sd = synthetic.make_state_dict(model, rnn_model)
    k = ((((hw[1] - 8) // 4 + 1) - 4) // 2 + 1) - 2
    key = "perceptual_encoder.rgb_gripper_encoder.conv_model.7.weight"
    sd[key] = sd[key][:, : 64 * k * k].contiguous()
    eng = HulcEngine(model, rnn_model, device="cpu", dropout_p=0.1, precision="fp32")
    eng.spec[key] = tuple(sd[key].shape)
    eng.ps = ParamStore(eng.spec, "cpu")
    eng.load_state_dict(sd)
    d = synthetic.make_modality("vis", 1, T, seed=9, static_hw=hw[0], gripper_hw=hw[1])
    lang = synthetic.make_modality("lang", 1, 1, seed=9, static_hw=hw[0], gripper_hw=hw[1])["lang"]
    g = torch.Generator().manual_seed(4)
    n_dims = 6 if model != "mcil" else 7
    hidden = plan = goal = None
    for t in range(T):
        st, gr = d["rgb_obs"]["rgb_static"][0, t : t + 1], d["rgb_obs"]["rgb_gripper"][0, t : t + 1]
        raw = d["state_info"]["robot_obs"][0, t : t + 1]
        if t % replan == 0:
            idx = torch.randint(0, 32, (1, 32), generator=g)
            eps = torch.randn(1, 256, generator=g)
            use_lang = t == 0
            frames = (st, gr) if use_lang else (torch.cat([st, d["rgb_obs"]["rgb_static"][0, -1:]]), torch.cat([gr, d["rgb_obs"]["rgb_gripper"][0, -1:]]))
            plan, goal, _ = O.inference_plan(sd, *frames, lang=lang if use_lang else None, plan_idx=idx, plan_eps=eps, model=model)
            hidden = torch.zeros(2, 1, 2048)
            p_e, g_e = eng.infer_plan(frames[0].contiguous(), frames[1].contiguous(), lang=lang if use_lang else None, plan_idx=idx if model == "hulc" else None,
                                      plan_eps=eps if model == "mcil" else None)
            torch.testing.assert_close(g_e, goal, rtol=1e-3, atol=1e-4)
            if model != "gcbc":
                torch.testing.assert_close(p_e, plan, rtol=1e-3, atol=1e-4)
        u_mix, u_inv = torch.rand(1, 1, n_dims, 10, generator=g), torch.rand(1, 1, n_dims, generator=g)
        a, hidden = O.inference_act(sd, st, gr, raw, plan, goal, hidden, u_mix, u_inv, model=model, rnn_model=rnn_model)
        a_e = eng.infer_act(st.contiguous(), gr.contiguous(), raw, sample_u=(u_mix, u_inv))
        torch.testing.assert_close(a_e, a, rtol=1e-3, atol=2e-3)
        torch.testing.assert_close(eng._infer_state["hidden"], hidden, rtol=1e-3, atol=1e-4)


def test_emu_gcbc_module_step_matches_oracle(emu):
    """GCBC.step (gcbc.py:287-317): the goal is encoded once per rollout, the decoder sees an empty plan."""
    from hulc_b200.models.gcbc import GCBC

    cfg = synthetic.model_config("gcbc", target_root="hulc_b200")
    cfg.pop("_target_"); cfg.pop("_recursive_")
    model = GCBC(**cfg, device=torch.device("cpu"), precision="fp32")
    sd = synthetic.make_state_dict("gcbc")
    model.load_state_dict(sd, strict=False)
    r = synthetic.rollout_inputs(3, "lang")
    model.lang_embeddings = {"the task": r["lang"].numpy()[None]}
    model.reset()
    hidden = torch.zeros(2, 1, 2048)
    plan, goal, _ = O.inference_plan(sd, r["rgb_static"][0:1], r["rgb_gripper"][0:1], lang=r["lang"], model="gcbc")
    for t in range(3):
        a = model.step(_obs(r, t), "the task", sample_u=(r["u_mix"][t], r["u_inv"][t]))
        ref, hidden = O.inference_act(sd, r["rgb_static"][t : t + 1], r["rgb_gripper"][t : t + 1], r["robot_obs_raw"][t : t + 1], plan, goal, hidden, r["u_mix"][t],
                                      r["u_inv"][t], model="gcbc")
        torch.testing.assert_close(a, ref, rtol=1e-3, atol=2e-3)

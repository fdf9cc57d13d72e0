"""The training-time image pipeline on the device: RandomShiftsAug + ScaleImageTensor + Normalize of
conf/datamodule/transforms/rand_shift.yaml:2-22 fused into one pass over the stored uint8 frames (hulc_frames_u8_shift_to_f32), against the
oracle's restatement and the fixture written from the unmodified reference class (oracle/make_golden.py::run_aug_case)."""
import numpy as np
import pytest
import torch

from oracle import hulc_oracle as O

CAMS = {"static": (200, 10, 3), "gripper": (84, 4, 4)}


def _frames(cam):
    h, pad, n = CAMS[cam]
    g = torch.Generator().manual_seed(17 + h)
    x = torch.randint(0, 256, (n, 3, h, h), generator=g, dtype=torch.uint8)
    sh = torch.randint(0, 2 * pad + 1, (n, 2), generator=g)
    return x, sh, pad


@pytest.mark.parametrize("cam", list(CAMS))
def test_oracle_matches_reference_fixture(cam, golden_dir):
    fx = np.load(golden_dir / "aug_random_shifts.npz")
    x, sh, pad = _frames(cam)
    assert np.array_equal(sh.numpy(), fx[f"{cam}_shifts"])
    out = O.random_shifts_aug(x, sh, pad)
    np.testing.assert_allclose(out[:, :, ::7, ::5].numpy(), fx[f"{cam}_out_sub"], rtol=0, atol=1e-4)  # bilinear sampling at pixel centres: 6e-5


@pytest.mark.parametrize("cam", list(CAMS))
def test_kernel_matches_oracle_and_fixture(K, cam, golden_dir):
    fx = np.load(golden_dir / "aug_random_shifts.npz")
    x, sh, pad = _frames(cam)
    out = K.frames_u8_shift_to_f32(x, torch.empty(x.shape), pad, shifts=sh.to(torch.int32))
    assert torch.equal(out, O.random_shifts_aug(x, sh, pad))  # the same IEEE operations in the same order
    np.testing.assert_allclose(out[:, :, ::7, ::5].numpy(), fx[f"{cam}_out_sub"], rtol=0, atol=1e-4)
    # pad 0 / zero shift at pad: the plain normalisation
    same = K.frames_u8_shift_to_f32(x, torch.empty(x.shape), pad, shifts=torch.full((x.shape[0], 2), pad, dtype=torch.int32))
    assert torch.equal(same, K.frames_u8_to_f32(x, torch.empty(x.shape)))


def test_philox_shifts_are_per_frame_uniform_integers(K):
    h, pad, n = 36, 4, 64
    ramp_y = (torch.arange(h).view(1, 1, h, 1) * 6).expand(n, 3, h, h).to(torch.uint8).contiguous()  # value = 6 y
    ramp_x = (torch.arange(h).view(1, 1, 1, h) * 6).expand(n, 3, h, h).to(torch.uint8).contiguous()  # value = 6 x
    px = lambda out: ((out * 0.5 + 0.5) * 255).round().long()
    oy = K.frames_u8_shift_to_f32(ramp_y, torch.empty(ramp_y.shape), pad, seed=5, site=9)
    ox = K.frames_u8_shift_to_f32(ramp_x, torch.empty(ramp_x.shape), pad, seed=5, site=9)  # same stream: the same shifts
    cy, cx = px(oy)[:, :, h // 2, h // 2], px(ox)[:, :, h // 2, h // 2]  # far from the borders: 6 (y + sy - pad), 6 (x + sx - pad)
    for c in (cy, cx):
        assert torch.equal(c[:, 0], c[:, 1]) and torch.equal(c[:, 0], c[:, 2])  # one shift per frame, all channels alike
    sy, sx = cy[:, 0] // 6 - h // 2 + pad, cx[:, 0] // 6 - h // 2 + pad
    assert int(sx.min()) >= 0 and int(sx.max()) <= 2 * pad and int(sy.min()) >= 0 and int(sy.max()) <= 2 * pad
    assert len(set(zip(sx.tolist(), sy.tolist()))) > 20  # 64 frames, 81 possible shifts
    assert len(set(sx.tolist())) == 2 * pad + 1 and len(set(sy.tolist())) == 2 * pad + 1  # every offset occurs
    # the borders replicate the edge pixel
    top = px(oy)[:, 0, 0, 0]
    assert torch.equal(top, 6 * (sy - pad).clamp(min=0))
    again = K.frames_u8_shift_to_f32(ramp_y, torch.empty(ramp_y.shape), pad, seed=5, site=9)
    assert torch.equal(oy, again)
    other = K.frames_u8_shift_to_f32(ramp_y, torch.empty(ramp_y.shape), pad, seed=6, site=9)
    assert not torch.equal(oy, other)


def test_engine_trains_on_shifted_frames(emu):
    """A training step on uint8 frames with device augmentation == the same step on fp32 frames augmented by the oracle; validation never shifts."""
    from engine_check import run_pair  # noqa: F401  (path set-up)
    from hulc_b200.engine import HulcEngine, ParamStore
    from hulc_b200.utils import synthetic

    hw = (64, 44)
    B, S = 2, 3
    sd = synthetic.make_state_dict("hulc")
    k = ((((hw[1] - 8) // 4 + 1) - 4) // 2 + 1) - 2
    key = "perceptual_encoder.rgb_gripper_encoder.conv_model.7.weight"
    sd[key] = sd[key][:, : 64 * k * k].contiguous()

    def engine():
        e = HulcEngine("hulc", device="cpu", dropout_p=0.0, precision="fp32")
        e.spec[key] = tuple(sd[key].shape)
        e.ps = ParamStore(e.spec, "cpu")
        e.load_state_dict(sd)
        return e

    batch = synthetic.make_batch(B, S, seed=1, static_hw=hw[0], gripper_hw=hw[1])
    u8 = {m: {kk: (dict(v) if isinstance(v, dict) else v) for kk, v in d.items()} for m, d in batch.items()}
    for m in u8:
        u8[m]["rgb_obs"] = {kk: ((v * 0.5 + 0.5) * 255).round().clamp(0, 255).to(torch.uint8) for kk, v in batch[m]["rgb_obs"].items()}
    g = torch.Generator().manual_seed(3)
    shifts = {"static": torch.randint(0, 21, (2 * B * S, 2), generator=g, dtype=torch.int32), "gripper": torch.randint(0, 9, (2 * B * S, 2), generator=g, dtype=torch.int32)}
    aug = {m: {kk: (dict(v) if isinstance(v, dict) else v) for kk, v in d.items()} for m, d in u8.items()}
    for i, m in enumerate(aug):
        sl = slice(i * B * S, (i + 1) * B * S)
        aug[m]["rgb_obs"] = {
            "rgb_static": O.random_shifts_aug(u8[m]["rgb_obs"]["rgb_static"].flatten(0, 1), shifts["static"][sl].long(), 10).view(B, S, 3, hw[0], hw[0]),
            "rgb_gripper": O.random_shifts_aug(u8[m]["rgb_obs"]["rgb_gripper"].flatten(0, 1), shifts["gripper"][sl].long(), 4).view(B, S, 3, hw[1], hw[1]),
        }
    noise = {m: synthetic.plan_noise(B, S, m)["u"] for m in batch}
    a, b = engine(), engine()
    a.set_augmentation(10, 4)
    oa = a.step(u8, plan_u=noise, aug_shifts=shifts)
    ob = b.step(aug, plan_u=noise)
    assert float(oa["total_loss"]) == float(ob["total_loss"])
    assert torch.equal(a.ps.grad, b.ps.grad)
    # Philox-drawn shifts: a different loss than the unshifted frames, reproducible for a seed
    l1 = float(a.step(u8, plan_u=noise, seed=7)["total_loss"])
    l2 = float(a.step(u8, plan_u=noise, seed=7)["total_loss"])
    l0 = float(b.step(u8, plan_u=noise, seed=7)["total_loss"])
    assert l1 == l2 and l1 != l0
    # forward-only (validation) passes do not augment
    assert float(a.step(u8, plan_u=noise, backward=False)["total_loss"]) == float(b.step(u8, plan_u=noise, backward=False)["total_loss"])

"""GPU-only behaviour of the module: moving a built module between devices (what Lightning does to every DDP rank), CUDA-graph
replay across batch shapes, KL-schedule changes under replay, the bounded graph cache."""
import numpy as np
import pytest
import torch

from hulc_b200.utils import synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _need_gpu():
    if not torch.cuda.is_available():
        pytest.fail("gpu-marked tests need a CUDA device; hulc_b200 has no CPU fallback")


def _model(device, model="hulc", dropout_p=0.0, precision="tf32"):
    from hulc_b200.models.hulc import Hulc

    cfg = synthetic.model_config(model, target_root="hulc_b200", dropout_p=dropout_p)
    cfg.pop("_target_"), cfg.pop("_recursive_")
    m = Hulc(**cfg, device=device, precision=precision)
    m.load_state_dict(synthetic.make_state_dict(model), strict=False)
    return m


def _same_training(losses, sd, ref_losses, ref_sd, n=2, lr=2e-4):
    """Two runs of the same steps: losses equal; parameters identical up to the order of the fp32 atomics in the bias / LayerNorm-gain gradient
    sums, which Adam's normalisation turns into at most a few times lr on a handful of near-zero-gradient elements."""
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-5)
    for k in ref_sd:
        diff = (sd[k].float() - ref_sd[k].float()).abs()
        # (a count, not only a fraction: one moved element of a 32-element bias is already 3 % of it)
        assert float(diff.max()) <= 3 * n * lr and int((diff > 1e-6).sum()) <= max(2, int(1e-2 * diff.numel())), k


def _steps(m, dev, n=2, B=2, S=8):
    opt = m.configure_optimizers()["optimizer"]
    losses = []
    for i in range(n):
        batch = synthetic._to(synthetic.make_batch(B, S, seed=1 + i), dev)
        loss = m.training_step(batch, i, seed=100 + i)
        loss.backward()
        opt.step()
        opt.zero_grad()
        losses.append(float(loss.detach()))
    return losses, {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}


def test_built_on_cpu_then_moved_to_gpu():
    """hulc/training.py builds the model before Lightning picks the device: construct on the CPU, `.to(cuda)`, train — identical to a
    module built on the device (the Adam step counter, RNG seed and NaN flag move with the parameters)."""
    ref_losses, ref_sd = _steps(_model("cuda"), "cuda")
    m = _model("cpu")
    with pytest.raises(Exception):  # no CPU path: the kernels refuse host tensors
        m.training_step(synthetic.make_batch(2, 8), 0)
    m = m.to("cuda")
    assert m.engine.ps.step_dev.is_cuda and m.engine.rng_dev.is_cuda and m.engine.nan_flag.is_cuda
    losses, sd = _steps(m, "cuda")
    _same_training(losses, sd, ref_losses, ref_sd)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_moved_between_gpus():
    """Every DDP rank > 0 builds on cuda:0 and is moved to cuda:<local_rank>: nothing may keep pointing at cuda:0."""
    ref_losses, ref_sd = _steps(_model("cuda:0"), "cuda:0")
    m = _model("cuda:0")
    _steps(m, "cuda:0", n=1)  # leaves activation buffers, a step count and RNG state on cuda:0
    m = _model("cuda:0").to("cuda:1")
    for t in (m.engine.ps.flat, m.engine.ps.grad, m.engine.ps.step_dev, m.engine.rng_dev, m.engine.nan_flag):
        assert t.device == torch.device("cuda:1")
    with torch.cuda.device(1):
        losses, sd = _steps(m, "cuda:1")
    _same_training(losses, sd, ref_losses, ref_sd)
    # a tensor on the wrong GPU is refused instead of dereferenced
    from hulc_b200 import _lib, ops

    with torch.cuda.device(1), pytest.raises(_lib.HulcError):
        ops.scale_(torch.ones(8, device="cuda:0"), 2.0)


def test_graph_replay_survives_other_batch_shapes_and_kl_schedule():
    """CUDA-graph replay of the training step: (1) a batch of another shape in between (the last, partial batch of an epoch) must not
    invalidate the buffers an earlier graph was captured on; (2) set_kl_beta (called once per epoch by the KL schedules) takes effect;
    (3) the cache of captured graphs is bounded."""
    dev = "cuda"
    m, e = _model(dev), _model(dev)  # graphs vs eager twin, same seeds
    m.enable_cuda_graphs()
    full = synthetic._to(synthetic.make_batch(2, 8, seed=1), dev)
    part = synthetic._to(synthetic.make_batch(1, 8, seed=2), dev)
    seq = [full, full, part, full, part, full]
    with torch.no_grad():
        for i, b in enumerate(seq):
            if i == 3:
                m.set_kl_beta(0.05), e.set_kl_beta(0.05)
            # eager twin: the same device-side RNG progression (one seed increment per step)
            lg = m.training_step(b, i)
            le = e.training_step(b, i)
            np.testing.assert_allclose(float(lg), float(le), rtol=1e-6, err_msg=f"step {i}")
            for k in ("plan_proposal.fc_state.0.weight", "perceptual_encoder.rgb_static_encoder.conv_model.0.weight"):
                torch.testing.assert_close(m.engine.ps.g[k], e.engine.ps.g[k], rtol=1e-5, atol=1e-7)
    assert len(m._graphs) == 2
    # bounded cache: fresh input tensors every step (a loader that does not reuse its staging buffers) must not grow it without bound
    m.max_graphs = 2
    with torch.no_grad():
        for i in range(4):
            m.training_step(synthetic._to(synthetic.make_batch(1, 8, seed=10 + i), dev), i)
    assert len(m._graphs) == 2

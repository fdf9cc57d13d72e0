"""The whole training step (hulc_b200.engine: forward, hand-written backward) executed on the host SIMT emulator build
of the kernel sources, against the oracle (CPU autograd).  Tiny shapes, reduced frame sizes; every buffer the engine
allocates is poisoned with NaN first so a read of a never-written element fails the test.  The GPU parity tests proper
are tests/test_gpu_*.py."""
import pytest

from engine_check import compare, run_pair

VARIANTS = [
    ("hulc", "rnn_decoder", 0.0),
    ("hulc", "rnn_decoder", 0.1),
    ("hulc", "gru_decoder", 0.0),
    ("gcbc", "rnn_decoder", 0.0),
    ("mcil", "rnn_decoder", 0.0),
]


@pytest.mark.parametrize("model,rnn_model,p", VARIANTS)
def test_emu_step_matches_oracle(emu, monkeypatch, model, rnn_model, p):
    from hulc_b200 import engine

    monkeypatch.setattr(engine, "_POISON", True)
    res = run_pair(model, rnn_model, B=2, S=4, p=p, device="cpu", hw=(64, 44))
    rep = compare(res)
    assert rep["worst_grad"][1] < 1e-3


@pytest.mark.parametrize("mask", ["all", "none"])
def test_emu_step_with_bc_z_and_mia_heads(emu, monkeypatch, mask):
    """The ablation configs' auxiliary heads (hulc.py:567-648): BC-Z language regression (cosine distance) and the MIA discriminator (BCE over
    matching / rolled pairs) next to the CLIP loss — losses and every gradient against the oracle, whose restatement is pinned to the
    unmodified reference by the fixture hulc_aux_b4s32.  An all-false use_for_aux_lang_loss switches both off (loss * 0 in the reference)."""
    import torch

    from hulc_b200 import engine

    monkeypatch.setattr(engine, "_POISON", True)
    B = 3
    res = run_pair("hulc", "rnn_decoder", B=B, S=4, p=0.0, device="cpu", hw=(64, 44), aux=True,
                   aux_mask=None if mask == "all" else torch.zeros(B, dtype=torch.bool))
    rep = compare(res)
    assert rep["worst_grad"][1] < 1e-3
    if mask == "none":
        assert rep["lang_pred_loss"][0] == 0.0 and rep["lang_contrastive_loss"][0] == 0.0


def test_emu_partial_aux_mask_is_refused(emu):
    """A mask that keeps part of the batch: the reference's BC-Z loss fails on it (mask applied twice, hulc.py:592-594); refused here too."""
    import torch

    with pytest.raises((NotImplementedError, IndexError)):
        run_pair("hulc", "rnn_decoder", B=3, S=4, p=0.0, device="cpu", hw=(64, 44), aux=True, aux_mask=torch.tensor([True, False, True]))

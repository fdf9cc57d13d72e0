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

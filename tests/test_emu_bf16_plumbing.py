"""The operand-twin plumbing of the bf16 path (HulcEngine precision="bf16") on the CPU: the SIMT kernels run on the host emulator, the
tcgen05 bf16 product (not emulated) is replaced by a torch stand-in with the same contract (bf16 operands, fp32 accumulation, the fused
epilogue) — test infrastructure only.  Checks that every product receives fresh bf16 twins, that bf16-only activations are never read
as fp32, and that losses / gradients stay within bf16 distance of the oracle."""
import pytest
import torch

from engine_check import compare, run_pair


def _fake_gemm_bf16(A, B, C=None, Cb=None, *, transA=False, transB=False, alpha=1.0, beta=0.0, bias=None, addend=None, add_mod=0, act=0, gate=None, drop=None):
    assert A.dtype == torch.bfloat16 and B.dtype == torch.bfloat16
    a = (A.t() if transA else A).float()
    b = (B.t() if transB else B).float()
    v = alpha * (a @ b)
    M, N = v.shape
    if bias is not None:
        v = v + bias
    if addend is not None:
        rows = torch.arange(M) % add_mod if add_mod else torch.arange(M)
        v = v + addend[rows]
    if beta != 0.0:
        v = v + beta * C
    if act & 3 == 1:
        v = v.relu()
    elif act & 3 == 2:
        v = v.tanh()
    if gate is not None:
        g = gate.float()
        v = v * (1 - g * g) if act & 4 else torch.where(g > 0, v, torch.zeros_like(v))
    if drop is not None and drop.p > 0:
        assert drop.keep is not None, "the CPU stand-in only takes injected masks"
        v = v * drop.keep.view(M, N).float() / (1 - drop.p)
    if C is not None:
        C.copy_(v)
    if Cb is not None:
        Cb.copy_(v.to(torch.bfloat16))
    return C if C is not None else Cb


@pytest.mark.parametrize("model,rnn_model,p", [("hulc", "rnn_decoder", 0.1), ("hulc", "gru_decoder", 0.0), ("gcbc", "rnn_decoder", 0.0), ("mcil", "rnn_decoder", 0.0)])
def test_bf16_plumbing_on_the_emulator(emu, monkeypatch, model, rnn_model, p):
    from hulc_b200 import engine, ops

    monkeypatch.setattr(engine, "_POISON", True)  # any read of a never-written fp32 buffer (a bf16-only activation) surfaces as NaN
    monkeypatch.setattr(ops, "gemm_bf16", _fake_gemm_bf16)
    monkeypatch.setattr(ops, "cast_bf16", lambda x, out=None: out.copy_(x.to(torch.bfloat16)) if out is not None else x.to(torch.bfloat16))
    monkeypatch.setattr(ops, "gemm_bf16_ok", lambda A, B: A.stride(-1) == 1 and B.stride(-1) == 1)
    orig_adam = ops.adam_step

    def adam(p_, g, m, v, **kw):
        pb = kw.pop("p_bf16", None)
        orig_adam(p_, g, m, v, **kw)
        if pb is not None:
            pb.copy_(p_.to(torch.bfloat16))

    monkeypatch.setattr(ops, "adam_step", adam)
    orig_init = engine.HulcEngine.__init__

    def init(self, *a, **kw):
        orig_init(self, *a, **kw)
        self.tc = self.bf16_conv = False  # the tensor-core convolutions / persistent recurrence do not exist on the emulator: SIMT convs, per-step products

    monkeypatch.setattr(engine.HulcEngine, "__init__", init)
    res = run_pair(model, rnn_model, B=2, S=4, p=p, device="cpu", hw=(64, 44), precision="bf16", use_idx=True)
    eng = res["eng"]
    assert eng.bf16 and eng._bf16_only, "hidden activations should have been produced as bf16 only"
    rep = compare(res, rtol=2e-2, atol=2e-2, grad_rtol=2e-1, inter_rtol=3e-2, inter_atol=3e-2,
                  skip_grads=("logit_scale",))  # B = 2: the scalar CLIP temperature gradient is a difference of near-equal terms
    # a second step and an optimizer step: twins are per step, the parameter copy follows Adam
    eng.optimizer_step()
    assert torch.equal(eng.ps.flat_bf16, eng.ps.flat.to(torch.bfloat16))
    print(model, rnn_model, rep["total_loss"], rep["worst_grad"])

"""ParamStore layout (hulc_b200/engine.py): the decoder-head group is padded to a multiple of four rows so that the fused head products
have 16-byte-aligned rows; the pad rows belong to no state_dict key, start at zero and stay zero under Adam."""
import torch

from hulc_b200.engine import ParamStore
from hulc_b200.utils.synthetic import param_spec


def test_head_group_is_padded_and_invisible():
    spec = param_spec("hulc", "rnn_decoder")
    ps = ParamStore(spec, "cpu")
    assert ps.n_heads == 182 and ps.n_heads_padded == 184
    assert tuple(ps.heads_w.shape) == (184, 2048) and tuple(ps.heads_b.shape) == (184,)
    # the four reference tensors are consecutive row blocks of the fused view, in the order the loss kernel expects
    row = 0
    for h in ("prob_fc", "mean_fc", "log_scale_fc", "gripper_fc"):
        w = ps.p[f"action_decoder.{h}.weight"]
        assert w.data_ptr() == ps.heads_w[row].data_ptr()
        assert ps.p[f"action_decoder.{h}.bias"].data_ptr() == ps.heads_b[row:].data_ptr()
        row += w.shape[0]
    assert row == 182
    # state_dict round trip ignores the padding; every offset is 16-byte aligned per group
    sd = {k: torch.randn(spec[k]) for k in spec}
    ps.load_state_dict(sd)
    out = ps.state_dict()
    assert set(out) == set(spec) and all(torch.equal(out[k], sd[k]) for k in spec)
    assert float(ps.heads_w[182:].abs().max()) == 0.0 and float(ps.heads_b[182:].abs().max()) == 0.0
    assert sum(int(torch.tensor(spec[k]).prod()) if len(spec[k]) else 1 for k in spec) <= ps.numel


def test_mcil_head_group():
    ps = ParamStore(param_spec("mcil", "rnn_decoder"), "cpu")
    assert ps.n_heads % 4 != 0 or ps.n_heads == ps.n_heads_padded
    assert ps.n_heads_padded % 4 == 0 and ps.heads_w.shape[0] == ps.n_heads_padded

"""Pin the oracle (oracle/slot_oracle.py) against outputs of the real reference.

The reference has no tests or golden vectors; tests/golden/*.npz are outputs of
the reference modules themselves, frozen by oracle/make_golden.py.  Where /root/reference is
present (build container) the oracle is additionally checked against the live reference.
"""
import pytest
import torch

from oracle import reference_bridge as rb
from oracle import slot_oracle as so
from tests.golden_io import load_case, load_json, rel_err

SA_CASES = ["sa_small_grad", "sa_slate_grad", "sa_k11_t5_ragged", "sa_k16_t7", "sa_k1_t1", "sa_sharp"]


@pytest.mark.parametrize("name", SA_CASES)
def test_slot_attention_forward_matches_reference(name):
    meta, g = load_case(name)
    slots, attn = so.slot_attention(g["in"]["inputs"], g["in"]["slots0"], g["p"], meta["T"], meta["eps"])
    assert slots.shape == g["out"]["slots"].shape and attn.shape == g["out"]["attn"].shape
    assert rel_err(slots, g["out"]["slots"]) < 2e-6
    assert rel_err(attn, g["out"]["attn"]) < 2e-6
    # slot-to-token assignment: identical wherever the fp64 top-2 margin is decidable in fp32
    p64 = so.to_dtype(g["p"], torch.float64)
    _, attn64 = so.slot_attention(g["in"]["inputs"].double(), g["in"]["slots0"].double(), p64, meta["T"], meta["eps"])
    ties = so.tie_mask(attn64, 1e-5)
    same = attn.argmax(-1) == g["out"]["attn"].argmax(-1)
    assert bool((same | ties).all())


@pytest.mark.parametrize("name", ["sa_small_grad", "sa_slate_grad", "sa_k11_t5_ragged", "sa_k1_t1"])
def test_slot_attention_gradients_match_reference(name):
    meta, g = load_case(name)
    x = g["in"]["inputs"].clone().requires_grad_(True)
    s0 = g["in"]["slots0"].clone().requires_grad_(True)
    p = {k: v.clone().requires_grad_(True) for k, v in g["p"].items()}
    slots, attn = so.slot_attention(x, s0, p, meta["T"], meta["eps"])
    loss = (slots * g["g.out"]["slots"]).sum() + (attn * g["g.out"]["attn"]).sum()
    loss.backward()
    assert rel_err(x.grad, g["g.in"]["inputs"]) < 2e-5
    assert rel_err(s0.grad, g["g.in"]["slots0"]) < 2e-5
    for k, gv in g["g.p"].items():
        # d/d(norm_slots.bias) is analytically zero (a shift common to all slots cancels in the softmax)
        assert rel_err(p[k].grad, gv) < 5e-5 or float((p[k].grad - gv).abs().max()) < 1e-5, k


def test_encoder_matches_reference():
    meta, g = load_case("encoder_slate")
    slots, attn = so.slot_attention_encoder(g["in"]["x"], g["in"]["noise"], g["p"], meta["T"])
    assert rel_err(slots, g["out"]["slots"]) < 2e-6
    assert rel_err(attn, g["out"]["attn"]) < 2e-6


@pytest.mark.parametrize("name", ["slate_encode_64", "bcdec_encode_32"])
def test_slate_encode_matches_reference(name):
    meta, g = load_case(name)
    obs = g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0
    slots, attn = so.slate_encode(obs, g["in"]["noise"], g["p"], meta["T"])
    assert rel_err(slots, g["out"]["slots"]) < 5e-6
    masks = so.masks_from_attn(attn, obs, with_attns=False)
    assert masks.shape == g["out"]["masks"].shape
    assert rel_err(masks, g["out"]["masks"]) < 5e-6
    attns = so.masks_from_attn(attn, obs, with_attns=True)
    assert rel_err(attns, g["out"]["attns"]) < 5e-6


def test_path_gradients_match_reference():
    """d(loss)/d(all hot-path parameters) through CNN + pos-emb + token MLP + slot attention."""
    meta, g = load_case("slate_path_grad_16")
    obs = g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0
    p = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "linear_position" not in k else v)
         for k, v in g["p"].items()}
    slots, attn = so.slate_encode(obs, g["in"]["noise"], p, meta["T"])
    assert rel_err(slots, g["out"]["slots"]) < 5e-6
    loss = (slots * g["g.out"]["slots"]).sum() + (attn * g["g.out"]["attn"]).sum()
    loss.backward()
    for k, gv in g["g.p"].items():
        assert rel_err(p[k].grad, gv) < 1e-4 or float((p[k].grad - gv).abs().max()) < 1e-5, k


def test_survey_known_answer():
    """SURVEY.md 8(c) KAT: seed-defined module and inputs; answers frozen in survey_kat.json."""
    kat = load_json("survey_kat.json")
    torch.manual_seed(0)
    # same draw order as SlotAttention.__init__ (slot_attn.py:30-45, networks.py:56-74)
    import torch.nn as nn

    def lin(i, o, bias=True, kaiming=False):
        m = nn.Linear(i, o, bias)
        (nn.init.kaiming_uniform_(m.weight, nonlinearity="relu") if kaiming else nn.init.xavier_uniform_(m.weight))
        if bias:
            nn.init.zeros_(m.bias)
        return m

    C, D, H = 64, 192, 192
    ln_i, ln_s, ln_m = nn.LayerNorm(C), nn.LayerNorm(D), nn.LayerNorm(D)
    pq, pk, pv = lin(D, D, False), lin(C, D, False), lin(C, D, False)
    gru = nn.GRUCell(D, D)
    nn.init.xavier_uniform_(gru.weight_ih)
    nn.init.orthogonal_(gru.weight_hh)
    nn.init.zeros_(gru.bias_ih)
    nn.init.zeros_(gru.bias_hh)
    m0, m2 = lin(D, H, kaiming=True), lin(H, D)
    p = {"norm_inputs.weight": ln_i.weight, "norm_inputs.bias": ln_i.bias, "norm_slots.weight": ln_s.weight,
         "norm_slots.bias": ln_s.bias, "norm_mlp.weight": ln_m.weight, "norm_mlp.bias": ln_m.bias,
         "project_q.weight": pq.weight, "project_k.weight": pk.weight, "project_v.weight": pv.weight,
         "gru.weight_ih": gru.weight_ih, "gru.weight_hh": gru.weight_hh, "gru.bias_ih": gru.bias_ih,
         "gru.bias_hh": gru.bias_hh, "mlp.0.weight": m0.weight, "mlp.0.bias": m0.bias,
         "mlp.2.weight": m2.weight, "mlp.2.bias": m2.bias}
    p = {k: v.detach() for k, v in p.items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4, 4096, 64, generator=g)
    s0 = torch.randn(4, 6, 192, generator=g)
    slots, attn = so.slot_attention(x, s0, p, 3)
    assert abs(float(slots.sum()) - kat["slots_sum"]) < 2e-3
    assert abs(float(slots.abs().mean()) - kat["slots_abs_mean"]) < 1e-5
    assert abs(float(attn.sum()) - kat["attn_sum"]) < 1e-1
    assert torch.allclose(attn[0, 0], torch.tensor(kat["attn00"]), atol=1e-6)
    hist = torch.bincount(attn.argmax(-1).flatten(), minlength=6).tolist()
    assert sum(abs(a - b) for a, b in zip(hist, kat["argmax_hist"])) <= 2


@pytest.mark.skipif(not rb.available(), reason="reference tree not present (GPU box)")
def test_oracle_matches_live_reference():
    ref = rb.load()
    torch.manual_seed(5)
    sa = ref.SlotAttention(3, 6, 64, 192, 192, 1)
    g = torch.Generator().manual_seed(6)
    x, s0 = torch.randn(2, 512, 64, generator=g), torch.randn(2, 6, 192, generator=g)
    with torch.no_grad():
        rs, ra = sa(x, s0)
    p = {k: v.detach() for k, v in sa.state_dict().items()}
    s, a = so.slot_attention(x, s0, p, 3)
    assert rel_err(s, rs) < 2e-6 and rel_err(a, ra) < 2e-6
    assert torch.equal(so.position_grid(64), ref.PositionalEmbedding(64, 64).linear_position_embedding)


def _pool_module(meta):
    """ocrl_b200's drop-in pooling transformer with the fixture's seeded parameters (same draws as the reference)."""
    from ocrl_b200.pooling import Transformer
    from oracle.make_golden import pool_params

    torch.manual_seed(meta["seed"])
    t = Transformer(meta["Din"], meta["d_model"], meta["nhead"], 1, None, False)
    t.eval()
    pool_params(t, meta["seed"], meta["d_model"])
    psum = float(sum(p.detach().double().sum() for p in t.parameters()))
    assert abs(psum - meta["param_sum"]) < 1e-6 * meta["param_abs_sum"]
    assert list(t.state_dict().keys()) == meta["state_keys"]  # reference checkpoints load strictly
    return t


def test_pooling_oracle_and_module_match_reference():
    """PPO consumer (poolings/common/transformer.py:9-33): the oracle restatement and the drop-in module's torch path
    against the output frozen from the real reference module."""
    from oracle import pool_oracle as po

    meta, g = load_case("pool_transformer")
    t = _pool_module(meta)
    p = {k: v.detach() for k, v in t.state_dict().items()}
    assert rel_err(po.transformer_pool(g["in"]["slots"], p, meta["nhead"]), g["out"]["pooled"]) < 2e-6
    with torch.no_grad():
        assert rel_err(t(g["in"]["slots"]), g["out"]["pooled"]) < 2e-6

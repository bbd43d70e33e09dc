"""GPU parity of the fused backward (through the C ABI / autograd Function) against the
reference's autograd gradients frozen in tests/golden and against the CPU oracle."""
import pytest
import torch

from oracle import slot_oracle as so
from tests.golden_io import load_case, rel_err

pytestmark = pytest.mark.gpu
GRAD_CASES = ["sa_small_grad", "sa_slate_grad", "sa_k11_t5_ragged", "sa_k1_t1"]
TOL = 2e-4  # gradients: 1e-4-level agreement in fp32 with a different summation order


def _close(a, b, tol=TOL):
    return rel_err(a, b) < tol or float((a - b).abs().max()) < 2e-5


@pytest.mark.parametrize("name", GRAD_CASES)
def test_slot_attention_gradients_match_reference(name):
    from ocrl_b200 import functional as F

    meta, g = load_case(name)
    x = g["in"]["inputs"].cuda().requires_grad_(True)
    s0 = g["in"]["slots0"].cuda().requires_grad_(True)
    p = {k: v.cuda().requires_grad_(True) for k, v in g["p"].items()}
    slots, attn = F.SlotAttentionFunction.apply(x, s0, meta["T"], meta["eps"], "fp32",
                                                *[p[n] for n in F.SA_PARAM_ORDER])
    assert rel_err(slots.detach().cpu(), g["out"]["slots"]) < 1e-4
    loss = (slots * g["g.out"]["slots"].cuda()).sum() + (attn * g["g.out"]["attn"].cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    bad = []
    if not _close(x.grad.cpu(), g["g.in"]["inputs"]):
        bad.append(("inputs", rel_err(x.grad.cpu(), g["g.in"]["inputs"])))
    if not _close(s0.grad.cpu(), g["g.in"]["slots0"]):
        bad.append(("slots0", rel_err(s0.grad.cpu(), g["g.in"]["slots0"])))
    for k, gv in g["g.p"].items():
        if not _close(p[k].grad.cpu(), gv):
            bad.append((k, rel_err(p[k].grad.cpu(), gv)))
    assert not bad, bad


def test_gradients_without_attention_gradient():
    """d_attn_vis = None path (loss on slots only), checked against oracle autograd."""
    from ocrl_b200 import functional as F

    meta, g = load_case("sa_small_grad")
    x = g["in"]["inputs"].clone().requires_grad_(True)
    s0 = g["in"]["slots0"].clone().requires_grad_(True)
    p = {k: v.clone().requires_grad_(True) for k, v in g["p"].items()}
    slots, _ = so.slot_attention(x, s0, p, meta["T"], meta["eps"])
    (slots * g["g.out"]["slots"]).sum().backward()
    xc = g["in"]["inputs"].cuda().requires_grad_(True)
    sc = g["in"]["slots0"].cuda().requires_grad_(True)
    pc = {k: v.cuda().requires_grad_(True) for k, v in g["p"].items()}
    slots_c, _ = F.SlotAttentionFunction.apply(xc, sc, meta["T"], meta["eps"], "fp32", *[pc[n] for n in F.SA_PARAM_ORDER])
    (slots_c * g["g.out"]["slots"].cuda()).sum().backward()
    assert _close(xc.grad.cpu(), x.grad) and _close(sc.grad.cpu(), s0.grad)
    for k in p:
        assert _close(pc[k].grad.cpu(), p[k].grad), k


def test_bf16_kv_gradients_within_tolerance():
    from ocrl_b200 import functional as F

    meta, g = load_case("sa_slate_grad")
    x = g["in"]["inputs"].cuda().requires_grad_(True)
    s0 = g["in"]["slots0"].cuda().requires_grad_(True)
    p = {k: v.cuda().requires_grad_(True) for k, v in g["p"].items()}
    slots, attn = F.SlotAttentionFunction.apply(x, s0, meta["T"], meta["eps"], "bf16", *[p[n] for n in F.SA_PARAM_ORDER])
    loss = (slots * g["g.out"]["slots"].cuda()).sum() + (attn * g["g.out"]["attn"].cuda()).sum()
    loss.backward()
    assert rel_err(x.grad.cpu(), g["g.in"]["inputs"]) < 5e-2
    assert rel_err(p["project_q.weight"].grad.cpu(), g["g.p"]["project_q.weight"]) < 5e-2
    assert rel_err(p["gru.weight_ih"].grad.cpu(), g["g.p"]["gru.weight_ih"]) < 5e-2


def test_module_training_path_matches_reference_gradients():
    """Whole hot path through SLATE_Module._get_slots with autograd (CNN + pos + token MLP via torch,
    slot attention via the fused kernels) against the reference's gradients."""
    import ocrl_b200
    from ocrl_b200.config import slate_config

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    meta, g = load_case("slate_path_grad_16")
    model = ocrl_b200.SLATE(*slate_config(num_slots=6, num_iterations=3, slot_size=64, mlp_hidden_size=128, obs_size=16))
    mod = model._module
    sd = mod.state_dict()
    sd.update(g["p"])
    mod.load_state_dict(sd)
    mod.cuda()
    mod.eval()
    enc = mod._slotattn
    noise = g["in"]["noise"].cuda()
    enc.init_slots = lambda batch, like: enc.slot_mu + torch.exp(enc.slot_log_sigma) * noise
    obs = (g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0).cuda()
    slots, attn = mod._get_slots(obs, with_attns=True)
    assert rel_err(slots.detach().cpu(), g["out"]["slots"]) < 1e-4
    loss = (slots * g["g.out"]["slots"].cuda()).sum() + (attn * g["g.out"]["attn"].cuda()).sum()
    loss.backward()
    named = dict(mod.named_parameters())
    bad = [(k, rel_err(named[k].grad.cpu(), gv)) for k, gv in g["g.p"].items()
           if not _close(named[k].grad.cpu(), gv, 5e-4)]
    assert not bad, bad

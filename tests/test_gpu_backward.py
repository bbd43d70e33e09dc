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


@pytest.mark.parametrize("kv", ["fp32", "bf16"])
def test_gradients_at_baseline_size(kv):
    """BASELINE.json size (B = 64, N = 4096, K = 6, T = 3, D = 192): the persistent-cluster walk (more images than
    clusters), the rank-T*K expansion and the per-cluster partial reduction of the fused backward against oracle
    autograd on the CPU -- input gradients of every image, d slots0 and all 17 parameter gradients."""
    from ocrl_b200 import functional as F

    B, N, K, T, D = 64, 4096, 6, 3, 192
    p = so.random_sa_params(K, 64, D, D, seed=5)
    gen = torch.Generator().manual_seed(77)
    x = torch.randn(B, N, 64, generator=gen)
    s0 = torch.randn(B, K, D, generator=gen)
    g_slots = torch.randn(B, K, D, generator=gen)
    g_attn = torch.randn(B, N, K, generator=gen) * 0.1
    # oracle autograd (fp32, CPU)
    xo, so0 = x.clone().requires_grad_(True), s0.clone().requires_grad_(True)
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    slots_o, attn_o = so.slot_attention(xo, so0, po, T, 1e-8)
    ((slots_o * g_slots).sum() + (attn_o * g_attn).sum()).backward()
    # CUDA path through the C ABI
    xc, sc = x.cuda().requires_grad_(True), s0.cuda().requires_grad_(True)
    pc = {k: v.cuda().requires_grad_(True) for k, v in p.items()}
    slots_c, attn_c = F.SlotAttentionFunction.apply(xc, sc, T, 1e-8, kv, *[pc[n] for n in F.SA_PARAM_ORDER])
    ((slots_c * g_slots.cuda()).sum() + (attn_c * g_attn.cuda()).sum()).backward()
    torch.cuda.synchronize()
    tol = TOL if kv == "fp32" else 5e-2
    assert rel_err(slots_c.detach().cpu(), slots_o.detach()) < (1e-4 if kv == "fp32" else 2e-2)
    errs = {"inputs": rel_err(xc.grad.cpu(), xo.grad), "slots0": rel_err(sc.grad.cpu(), so0.grad)}
    per_image = ((xc.grad.cpu() - xo.grad).flatten(1).norm(dim=1) / xo.grad.flatten(1).norm(dim=1)).tolist()
    for k in p:
        errs[k] = rel_err(pc[k].grad.cpu(), po[k].grad)
    print(f"{kv}: " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    print(f"{kv}: input gradient per image: worst {max(per_image):.1e}, median {sorted(per_image)[B // 2]:.1e}")
    # every image, first and last cluster rounds alike (bf16 k/v: single images scatter around the batch figure)
    assert max(per_image) < (tol if kv == "fp32" else 2 * tol), (kv, "worst image", max(per_image))
    assert len(errs) == 19
    # d/d(norm_slots.bias) is analytically ZERO: the bias adds the same vector W_q beta to every slot's query, i.e. the same
    # number to all K logits of a token, and the softmax over slots ignores it -- both sides hold rounding noise only,
    # so that entry is bounded against the scale of its sibling instead of compared
    zero = errs.pop("norm_slots.bias")
    scale = float(po["norm_slots.weight"].grad.norm())
    noise_o, noise_c = float(po["norm_slots.bias"].grad.norm()), float(pc["norm_slots.bias"].grad.norm())
    print(f"norm_slots.bias (analytically zero): oracle {noise_o:.1e}, cuda {noise_c:.1e}, sibling scale {scale:.1e}, rel {zero:.1e}")
    assert noise_o < 1e-3 * scale and noise_c < (1e-3 if kv == "fp32" else 2e-2) * scale
    # bf16 mode against the EXACT fp32 reference: k, v stored in bf16 and the forward's tensor-core operands (queries,
    # softmax weights, update weights) rounded to bf16; through T softmax iterations that is 3e-2 .. 6e-2 on the gradients
    # at this size (forward 2e-2, asserted above).  All 19 tensors are reported and bounded at twice the small-case 5e-2.
    bad = {k: v for k, v in errs.items() if not v < (tol if kv == "fp32" else 2 * tol)}
    assert not bad, (kv, bad)


@pytest.mark.parametrize("kv", ["fp32", "bf16"])
@pytest.mark.parametrize("B,N,K,T,D,H", [(3, 320, 16, 3, 192, 192), (20, 1024, 11, 5, 192, 192), (2, 100, 9, 2, 192, 192)])
def test_nine_to_sixteen_slots_train_at_slot_size_192(B, N, K, T, D, H, kv):
    """BASELINE.json config 5 (num_slots up to 16, D = 192) through forward AND backward: the slot-sized state of the
    backward cluster kernel used to exceed the shared memory from K = 9 on (OCRL_E_SHAPE).  Against oracle autograd."""
    from ocrl_b200 import functional as F

    if kv == "fp32" and K > 12:
        pytest.skip("the fp32 parity-mode FORWARD (FFMA cluster kernel) holds its state for K <= 12 at D = 192")
    p = so.random_sa_params(K, 64, D, H, seed=21)
    gen = torch.Generator().manual_seed(K * 100 + N)
    x = torch.randn(B, N, 64, generator=gen)
    s0 = torch.randn(B, K, D, generator=gen)
    g_slots = torch.randn(B, K, D, generator=gen)
    g_attn = torch.randn(B, N, K, generator=gen) * 0.1
    xo, so0 = x.clone().requires_grad_(True), s0.clone().requires_grad_(True)
    po = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    slots_o, attn_o = so.slot_attention(xo, so0, po, T, 1e-8)
    ((slots_o * g_slots).sum() + (attn_o * g_attn).sum()).backward()
    xc, sc = x.cuda().requires_grad_(True), s0.cuda().requires_grad_(True)
    pc = {k: v.cuda().requires_grad_(True) for k, v in p.items()}
    slots_c, attn_c = F.SlotAttentionFunction.apply(xc, sc, T, 1e-8, kv, *[pc[n] for n in F.SA_PARAM_ORDER])
    ((slots_c * g_slots.cuda()).sum() + (attn_c * g_attn.cuda()).sum()).backward()
    torch.cuda.synchronize()
    tol = TOL if kv == "fp32" else 1e-1
    assert rel_err(slots_c.detach().cpu(), slots_o.detach()) < (1e-4 if kv == "fp32" else 2e-2)
    errs = {"inputs": rel_err(xc.grad.cpu(), xo.grad), "slots0": rel_err(sc.grad.cpu(), so0.grad)}
    for k in p:
        if k != "norm_slots.bias":  # analytically zero (see test_gradients_at_baseline_size)
            errs[k] = rel_err(pc[k].grad.cpu(), po[k].grad)
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, (kv, bad)

"""GPU parity of the CUDA forward path (through the C ABI) against the frozen reference outputs
(tests/golden) and the CPU oracle.  fp32 mode: 1e-4 relative on slots and attention, argmax
identical wherever the fp64 top-2 margin is decidable; bf16 k/v storage: 2e-2 relative."""
import pytest
import torch

from oracle import slot_oracle as so
from tests.golden_io import load_case, rel_err

pytestmark = pytest.mark.gpu

SA_CASES = ["sa_small_grad", "sa_slate_grad", "sa_k11_t5_ragged", "sa_k16_t7", "sa_k1_t1", "sa_sharp"]
FP32_TOL = 1e-4
BF16_TOL = 2e-2


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("name", SA_CASES)
def test_kv_projection_matches_oracle(name):
    from ocrl_b200 import functional as F

    meta, g = load_case(name)
    k_ref, v_ref = so.kv_project(g["in"]["inputs"], g["p"])
    k, v, _ = F.kv_project(g["in"]["inputs"].cuda(), _cuda(g["p"]), kv="fp32")
    torch.cuda.synchronize()
    assert rel_err(k.cpu(), k_ref) < 1e-5 and rel_err(v.cpu(), v_ref) < 1e-5
    kb, vb, _ = F.kv_project(g["in"]["inputs"].cuda(), _cuda(g["p"]), kv="bf16")
    assert kb.dtype == torch.bfloat16
    assert rel_err(kb.float().cpu(), k_ref) < 8e-3 and rel_err(vb.float().cpu(), v_ref) < 8e-3


@pytest.mark.parametrize("name", SA_CASES)
def test_iteration_kernel_matches_oracle_on_same_kv(name):
    """The fused loop alone, fed with the oracle's k and v."""
    from ocrl_b200 import functional as F

    meta, g = load_case(name)
    k_ref, v_ref = so.kv_project(g["in"]["inputs"], g["p"])
    s_ref, a_ref = so.iterate(k_ref, v_ref, g["in"]["slots0"], g["p"], meta["T"], meta["eps"])
    s, a, _ = F.iterate(k_ref.cuda(), v_ref.cuda(), g["in"]["slots0"].cuda(), _cuda(g["p"]), meta["T"],
                        epsilon=meta["eps"])
    torch.cuda.synchronize()
    assert rel_err(s.cpu(), s_ref) < FP32_TOL, rel_err(s.cpu(), s_ref)
    assert rel_err(a.cpu(), a_ref) < FP32_TOL, rel_err(a.cpu(), a_ref)


@pytest.mark.parametrize("name", SA_CASES)
def test_slot_attention_fp32_matches_reference(name):
    from ocrl_b200 import functional as F

    meta, g = load_case(name)
    s, a = F.slot_attention(g["in"]["inputs"].cuda(), g["in"]["slots0"].cuda(), _cuda(g["p"]), meta["T"],
                            epsilon=meta["eps"], kv="fp32")
    torch.cuda.synchronize()
    s, a = s.cpu(), a.cpu()
    assert s.shape == g["out"]["slots"].shape and a.shape == g["out"]["attn"].shape
    assert rel_err(s, g["out"]["slots"]) < FP32_TOL, rel_err(s, g["out"]["slots"])
    assert rel_err(a, g["out"]["attn"]) < FP32_TOL, rel_err(a, g["out"]["attn"])
    assert torch.allclose(a.sum(-1), torch.ones_like(a.sum(-1)), atol=1e-5)
    # slot-to-token assignment: bit-exact wherever fp32 can decide it
    p64 = so.to_dtype(g["p"], torch.float64)
    _, a64 = so.slot_attention(g["in"]["inputs"].double(), g["in"]["slots0"].double(), p64, meta["T"], meta["eps"])
    ties = so.tie_mask(a64, 1e-5)
    same = a.argmax(-1) == g["out"]["attn"].argmax(-1)
    assert bool((same | ties).all()), int((~(same | ties)).sum())
    # how much the tie rule excuses: tokens whose fp64 top-2 margin is below 1e-5, and how many of those really differ
    n_tok, n_ties, n_excused = same.numel(), int(ties.sum()), int((~same & ties).sum())
    print(f"{name}: {n_tok} tokens, {n_ties} with fp64 top-2 margin < 1e-5, {n_excused} argmax differences excused")
    assert n_ties <= max(2, n_tok // 2000) or meta["K"] == 1, (n_ties, n_tok)  # a handful, never a loophole
    assert n_excused <= max(1, n_tok // 5000), (n_excused, n_tok)


@pytest.mark.parametrize("name", SA_CASES)
def test_slot_attention_bf16_kv_within_tolerance(name):
    from ocrl_b200 import functional as F

    meta, g = load_case(name)
    s, a = F.slot_attention(g["in"]["inputs"].cuda(), g["in"]["slots0"].cuda(), _cuda(g["p"]), meta["T"],
                            epsilon=meta["eps"], kv="bf16")
    torch.cuda.synchronize()
    assert rel_err(s.cpu(), g["out"]["slots"]) < BF16_TOL, rel_err(s.cpu(), g["out"]["slots"])
    assert rel_err(a.cpu(), g["out"]["attn"]) < BF16_TOL, rel_err(a.cpu(), g["out"]["attn"])
    if meta["K"] > 1:
        flips = (a.cpu().argmax(-1) != g["out"]["attn"].argmax(-1)).float().mean().item()
        assert flips < 0.03  # reported, not promised to be zero


def test_encoder_token_mlp_fused():
    """SlotAttentionEncoder: token LN+MLP fused in front of the projection."""
    from ocrl_b200 import functional as F

    meta, g = load_case("encoder_slate")
    p = g["p"]
    enc = {k: p[k] for k in ("layer_norm.weight", "layer_norm.bias", "mlp.0.weight", "mlp.0.bias", "mlp.2.weight",
                             "mlp.2.bias")}
    sa = {k[len("slot_attention."):]: v for k, v in p.items() if k.startswith("slot_attention.")}
    slots0 = so.init_slots(g["in"]["noise"], p)
    k, v, y = F.kv_project(g["in"]["x"].cuda(), _cuda(sa), kv="fp32", enc=_cuda(enc), want_y=True)
    assert rel_err(y.cpu(), so.token_mlp(g["in"]["x"], p)) < 1e-5
    s, a, _ = F.iterate(k, v, slots0.cuda(), _cuda(sa), meta["T"])
    assert rel_err(s.cpu(), g["out"]["slots"]) < FP32_TOL
    assert rel_err(a.cpu(), g["out"]["attn"]) < FP32_TOL


@pytest.mark.parametrize("name", ["slate_encode_64", "bcdec_encode_32"])
def test_feature_map_ingest_with_position_table(name):
    """NCHW feature map + position table -> tokens -> whole path (CNN taken from the oracle)."""
    from ocrl_b200 import functional as F

    meta, g = load_case(name)
    p = g["p"]
    obs = g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0
    fmap = so.cnn_encoder(obs, p)
    pos = so.position_table(p)
    pre = "_slotattn."
    enc = {k: p[pre + k] for k in ("layer_norm.weight", "layer_norm.bias", "mlp.0.weight", "mlp.0.bias",
                                   "mlp.2.weight", "mlp.2.bias")}
    sa = {k[len(pre + "slot_attention."):]: v for k, v in p.items() if k.startswith(pre + "slot_attention.")}
    slots0 = so.init_slots(g["in"]["noise"], p, pre)
    k, v, _ = F.kv_project(fmap.cuda(), _cuda(sa), kv="fp32", enc=_cuda(enc), pos_table=pos.flatten(1).cuda())
    s, a, _ = F.iterate(k, v, slots0.cuda(), _cuda(sa), meta["T"])
    assert rel_err(s.cpu(), g["out"]["slots"]) < FP32_TOL
    masks = so.masks_from_attn(a.cpu(), obs, with_attns=False)
    assert rel_err(masks, g["out"]["masks"]) < FP32_TOL


def test_full_size_properties():
    """BASELINE size (B=64, N=4096, K=6, T=3, D=192): size-independent properties.
    attention rows sum to 1; images are independent (a permutation of the batch permutes the
    outputs bit-exactly); results do not depend on the batch they were computed in."""
    from ocrl_b200 import functional as F

    torch.manual_seed(0)
    p = _cuda(so.random_sa_params(6, 64, 192, 192, seed=3))
    x = torch.randn(64, 4096, 64, device="cuda")
    s0 = torch.randn(64, 6, 192, device="cuda")
    s, a = F.slot_attention(x, s0, p, 3)
    assert torch.isfinite(s).all() and torch.isfinite(a).all()
    assert torch.allclose(a.sum(-1), torch.ones(64, 4096, device="cuda"), atol=1e-5)
    perm = torch.randperm(64, device="cuda")
    s2, a2 = F.slot_attention(x[perm].contiguous(), s0[perm].contiguous(), p, 3)
    assert torch.equal(s2, s[perm]) and torch.equal(a2, a[perm])
    s3, a3 = F.slot_attention(x[:3].contiguous(), s0[:3].contiguous(), p, 3)
    assert torch.equal(s3, s[:3]) and torch.equal(a3, a[:3])
    # spot-check two images against the oracle
    pc = {k: v.cpu() for k, v in p.items()}
    sr, ar = so.slot_attention(x[:2].cpu(), s0[:2].cpu(), pc, 3)
    assert rel_err(s[:2].cpu(), sr) < FP32_TOL and rel_err(a[:2].cpu(), ar) < FP32_TOL


def test_error_conventions():
    from ocrl_b200 import functional as F

    p = _cuda(so.random_sa_params(6, 64, 192, 192, seed=3))
    with pytest.raises(RuntimeError):
        F.slot_attention(torch.randn(1, 16, 64), torch.randn(1, 6, 192), p, 3)  # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        F.iterate(torch.randn(1, 64, 192, device="cuda"), torch.randn(1, 64, 192, device="cuda"),
                  torch.randn(1, 17, 192, device="cuda"), p, 3)  # K > 16

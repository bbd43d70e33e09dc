"""GPU parity of the bf16 / tensor-core path: tcgen05+TMA token stage (kv_proj_tc.cu) and the
mma.sync iteration kernel (sa_iter_fwd_tc.cu) against the CPU oracle, 2e-2 relative (north star bf16 mode)."""
import pytest
import torch

from oracle import slot_oracle as so
from tests.golden_io import load_case, rel_err

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _enc(seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    return {"layer_norm.weight": 1 + 0.1 * r(64), "layer_norm.bias": 0.1 * r(64), "mlp.0.weight": 0.2 * r(64, 64),
            "mlp.0.bias": 0.1 * r(64), "mlp.2.weight": 0.2 * r(64, 64), "mlp.2.bias": 0.1 * r(64)}


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


@pytest.mark.parametrize("D", [64, 128, 192])
@pytest.mark.parametrize("B,N", [(2, 256), (3, 100), (1, 1), (5, 1024)])
def test_tcgen05_token_stage(D, B, N):
    """LayerNorm + [token MLP] + k/v projection on tcgen05 with TMA-fed tiles; ragged and partial tiles."""
    from ocrl_b200 import functional as F

    p, enc = so.random_sa_params(6, 64, D, D, seed=5), _enc(7)
    g = torch.Generator().manual_seed(B * 1000 + N)
    x = torch.randn(B, N, 64, generator=g) + 0.3
    k_ref, v_ref = so.kv_project(x, p)
    k, v, _ = F.kv_project(x.cuda(), _cuda(p), kv="bf16")
    assert k.dtype == torch.bfloat16 and k.shape == (B, N, D)
    assert rel_err(k.float().cpu(), k_ref) < 8e-3 and rel_err(v.float().cpu(), v_ref) < 8e-3
    y_ref = so.token_mlp(x, enc)
    k_ref, v_ref = so.kv_project(y_ref, p)
    k, v, y = F.kv_project(x.cuda(), _cuda(p), kv="bf16", enc=_cuda(enc), want_y=True)
    assert rel_err(y.cpu(), y_ref) < 8e-3
    assert rel_err(k.float().cpu(), k_ref) < 1e-2 and rel_err(v.float().cpu(), v_ref) < 1e-2


@pytest.mark.parametrize("S", [16, 32])
def test_tcgen05_feature_map_ingest(S):
    """NCHW feature map + position table through the TMA transposing load."""
    from ocrl_b200 import functional as F

    p, enc = so.random_sa_params(6, 64, 192, 192, seed=5), _enc(7)
    g = torch.Generator().manual_seed(S)
    fmap, pos = torch.randn(3, 64, S, S, generator=g), torch.randn(64, S * S, generator=g)
    tok = (fmap.flatten(2) + pos.unsqueeze(0)).permute(0, 2, 1).contiguous()
    k_ref, v_ref = so.kv_project(so.token_mlp(tok, enc), p)
    k, v, _ = F.kv_project(fmap.cuda(), _cuda(p), kv="bf16", enc=_cuda(enc), pos_table=pos.cuda())
    assert rel_err(k.float().cpu(), k_ref) < 1e-2 and rel_err(v.float().cpu(), v_ref) < 1e-2


@pytest.mark.parametrize("B,H,W", [(3, 16, 16), (2, 32, 64), (5, 8, 16), (3, 12, 12), (2, 10, 16)])
def test_tcgen05_bf16_channels_last_ingest(B, H, W):
    """bf16 channels-last feature map (the conv stack's output) + position table.  When H*W is a multiple
    of the 128-token tile the table rides through TMA as bf16 tiles, otherwise it is read per row."""
    from ocrl_b200 import functional as F

    p, enc = so.random_sa_params(6, 64, 192, 192, seed=5), _enc(7)
    g = torch.Generator().manual_seed(H * 100 + W)
    fmap = torch.randn(B, 64, H, W, generator=g).bfloat16()
    pos = torch.randn(64, H * W, generator=g)
    tok = (fmap.float().flatten(2) + pos.unsqueeze(0)).permute(0, 2, 1).contiguous()
    k_ref, v_ref = so.kv_project(so.token_mlp(tok, enc), p)
    fm = fmap.cuda().contiguous(memory_format=torch.channels_last)
    k, v, _ = F.kv_project(fm, _cuda(p), kv="bf16", enc=_cuda(enc), pos_table=pos.cuda())
    assert rel_err(k.float().cpu(), k_ref) < 1e-2 and rel_err(v.float().cpu(), v_ref) < 1e-2
    k2, v2, _ = F.kv_project(fm, _cuda(p), kv="bf16", enc=_cuda(enc))
    k_ref, v_ref = so.kv_project(so.token_mlp(tok - pos.t().unsqueeze(0), enc), p)
    assert rel_err(k2.float().cpu(), k_ref) < 1e-2 and rel_err(v2.float().cpu(), v_ref) < 1e-2


def test_tensor_core_iteration_kernel_on_exact_bf16_inputs():
    """The mma.sync loop fed with bf16-rounded k, v: compared with the oracle run on the same rounded
    values, so only the kernel's own arithmetic (bf16 q / weights, fp32 accumulate) is measured."""
    from ocrl_b200 import functional as F

    for name in ("sa_slate_grad", "sa_k16_t7", "sa_k11_t5_ragged", "sa_sharp"):
        meta, g = load_case(name)
        k_ref, v_ref = so.kv_project(g["in"]["inputs"], g["p"])
        kb, vb = k_ref.bfloat16(), v_ref.bfloat16()
        s_ref, a_ref = so.iterate(kb.float(), vb.float(), g["in"]["slots0"], g["p"], meta["T"], meta["eps"])
        s, a, _ = F.iterate(kb.cuda(), vb.cuda(), g["in"]["slots0"].cuda(), _cuda(g["p"]), meta["T"], epsilon=meta["eps"])
        assert rel_err(s.cpu(), s_ref) < BF16_TOL, (name, rel_err(s.cpu(), s_ref))
        assert rel_err(a.cpu(), a_ref) < BF16_TOL, (name, rel_err(a.cpu(), a_ref))
        assert torch.allclose(a.sum(-1), torch.ones_like(a.sum(-1)), atol=1e-4)


def test_bf16_full_size_properties():
    """BASELINE size in bf16 mode, k/v form (the factored form has its own copy of this test): finite, rows sum to one,
    batch-permutation equivariant bit-exactly."""
    from ocrl_b200 import functional as F

    torch.manual_seed(0)
    p = _cuda(so.random_sa_params(6, 64, 192, 192, seed=3))
    x = torch.randn(64, 4096, 64, device="cuda")
    s0 = torch.randn(64, 6, 192, device="cuda")
    s, a = F.slot_attention(x, s0, p, 3, kv="bf16", factored=False)
    assert torch.isfinite(s).all() and torch.isfinite(a).all()
    assert torch.allclose(a.sum(-1), torch.ones(64, 4096, device="cuda"), atol=1e-4)
    perm = torch.randperm(64, device="cuda")
    s2, a2 = F.slot_attention(x[perm].contiguous(), s0[perm].contiguous(), p, 3, kv="bf16", factored=False)
    assert torch.equal(s2, s[perm]) and torch.equal(a2, a[perm])
    pc = {k: v.cpu() for k, v in p.items()}
    sr, ar = so.slot_attention(x[:2].cpu(), s0[:2].cpu(), pc, 3)
    assert rel_err(s[:2].cpu(), sr) < BF16_TOL and rel_err(a[:2].cpu(), ar) < BF16_TOL

"""GPU parity of the FACTORED inference path (ocrl_xhat_fwd + ocrl_sa_iter_fwd_xhat, include/ocrl_sa.h): the loop streams
x^ = norm_inputs(x) instead of k, v and applies project_k / project_v through folded update weights
(k . q = x^ . (s W_k^T q),  sum_n w_n v_n = W_v sum_n w_n x^_n; reference slot_attn.py:54-102).
Compared with the CPU oracle on the frozen reference goldens, 2e-2 relative (north-star bf16 mode), and with the
k/v form of the same library."""
import pytest
import torch

from oracle import slot_oracle as so
from tests.golden_io import load_case, rel_err

pytestmark = pytest.mark.gpu
BF16_TOL = 2e-2


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def _enc(seed):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    return {"layer_norm.weight": 1 + 0.1 * r(64), "layer_norm.bias": 0.1 * r(64), "mlp.0.weight": 0.2 * r(64, 64),
            "mlp.0.bias": 0.1 * r(64), "mlp.2.weight": 0.2 * r(64, 64), "mlp.2.bias": 0.1 * r(64)}


@pytest.mark.parametrize("B,N", [(2, 256), (3, 100), (1, 1), (5, 1024)])
@pytest.mark.parametrize("with_mlp", [False, True])
def test_xhat_token_stage(B, N, with_mlp):
    """ocrl_xhat_fwd: [token LN + MLP] + norm_inputs, bf16 output; ragged and partial tiles."""
    from ocrl_b200 import functional as F

    p, enc = so.random_sa_params(6, 64, 192, 192, seed=5), _enc(7)
    g = torch.Generator().manual_seed(B * 1000 + N)
    x = torch.randn(B, N, 64, generator=g) + 0.3
    y_ref = so.token_mlp(x, enc) if with_mlp else x
    xh_ref = so.layer_norm(y_ref, p["norm_inputs.weight"], p["norm_inputs.bias"])
    xh, v, y = F.kv_project(x.cuda(), _cuda(p), kv="bf16", enc=_cuda(enc) if with_mlp else None, want_y=with_mlp,
                            xhat_only=True)
    assert v is None and xh.dtype == torch.bfloat16 and xh.shape == (B, N, 64)
    assert rel_err(xh.float().cpu(), xh_ref) < (1e-2 if with_mlp else 4e-3)
    if with_mlp:
        assert rel_err(y.cpu(), y_ref) < 8e-3


def test_xhat_token_stage_from_feature_maps():
    """The same ingest formats as the k/v token stage: NCHW fp32 map and channels-last bf16 map, with the position table."""
    from ocrl_b200 import functional as F

    p, enc = so.random_sa_params(6, 64, 192, 192, seed=5), _enc(7)
    g = torch.Generator().manual_seed(3)
    fmap, pos = torch.randn(3, 64, 16, 16, generator=g), torch.randn(64, 256, generator=g)
    for bf in (False, True):
        fm = fmap.bfloat16().float() if bf else fmap
        tok = (fm.flatten(2) + pos.unsqueeze(0)).permute(0, 2, 1).contiguous()
        xh_ref = so.layer_norm(so.token_mlp(tok, enc), p["norm_inputs.weight"], p["norm_inputs.bias"])
        src = fmap.cuda().bfloat16().contiguous(memory_format=torch.channels_last) if bf else fmap.cuda()
        xh, _, _ = F.kv_project(src, _cuda(p), kv="bf16", enc=_cuda(enc), pos_table=pos.cuda(), xhat_only=True)
        assert rel_err(xh.float().cpu(), xh_ref) < 1e-2


@pytest.mark.parametrize("name", ["sa_slate_grad", "sa_sharp", "sa_k16_t7"])
def test_factored_loop_matches_reference_goldens(name):
    """SlotAttention.forward through the factored kernels against the frozen reference outputs."""
    from ocrl_b200 import functional as F

    meta, g = load_case(name)
    assert meta["D"] == 192
    s, a = F.slot_attention(g["in"]["inputs"].cuda(), g["in"]["slots0"].cuda(), _cuda(g["p"]), meta["T"],
                            epsilon=meta["eps"], kv="bf16", factored=True)
    assert F.last_kernel() == "tcgen05_xhat"
    assert rel_err(s.cpu(), g["out"]["slots"]) < BF16_TOL, rel_err(s.cpu(), g["out"]["slots"])
    assert rel_err(a.cpu(), g["out"]["attn"]) < BF16_TOL, rel_err(a.cpu(), g["out"]["attn"])
    assert torch.allclose(a.sum(-1), torch.ones_like(a.sum(-1)), atol=1e-4)
    # ... and against the k/v form: two bf16 evaluations of the same function
    s2, a2 = F.slot_attention(g["in"]["inputs"].cuda(), g["in"]["slots0"].cuda(), _cuda(g["p"]), meta["T"],
                              epsilon=meta["eps"], kv="bf16", factored=False)
    assert F.last_kernel() == "tcgen05"
    assert rel_err(s, s2) < BF16_TOL and rel_err(a, a2) < BF16_TOL
    # arg-max masks: report the flip rate against the fp64 oracle (bf16 mode is not promised to be 0)
    flips = (a.cpu().argmax(-1) != g["out"]["attn"].argmax(-1)).float().mean().item()
    print(f"{name}: factored argmax flip rate {flips:.2e}")
    assert flips < 2e-2


def test_factored_loop_on_exact_bf16_inputs():
    """Fed with the oracle's own x^ rounded to bf16: only the kernel's arithmetic differs from the oracle."""
    from ocrl_b200 import functional as F

    for name in ("sa_slate_grad", "sa_k16_t7"):
        meta, g = load_case(name)
        p = g["p"]
        xh = so.layer_norm(g["in"]["inputs"], p["norm_inputs.weight"], p["norm_inputs.bias"]).bfloat16()
        D = p["project_k.weight"].shape[0]
        k, v = (xh.float() @ p["project_k.weight"].t()) * D ** -0.5, xh.float() @ p["project_v.weight"].t()
        s_ref, a_ref = so.iterate(k, v, g["in"]["slots0"], p, meta["T"], meta["eps"])
        s, a = F.iterate_xhat(xh.cuda(), g["in"]["slots0"].cuda(), _cuda(p), meta["T"], epsilon=meta["eps"])
        assert rel_err(s.cpu(), s_ref) < BF16_TOL and rel_err(a.cpu(), a_ref) < BF16_TOL


@pytest.mark.parametrize("K,T,B,N", [(1, 1, 1, 64), (3, 2, 7, 320), (6, 3, 40, 256), (8, 5, 2, 1000), (9, 3, 3, 512),
                                     (16, 2, 2, 300), (6, 7, 80, 128),
                                     # from 2048 tokens up the pass runs 256-token steps (ragged last tile, 8 and 16 slot columns)
                                     (6, 3, 3, 2048), (8, 2, 2, 2500), (11, 5, 2, 4096), (16, 3, 3, 2304)])
def test_factored_loop_shapes(K, T, B, N):
    """Slot counts 1..16, ragged token counts, more images than lanes x clusters, T up to 7."""
    from ocrl_b200 import functional as F

    p = so.random_sa_params(K, 64, 192, 192, seed=11 + K)
    g = torch.Generator().manual_seed(K * 100 + T)
    x = torch.randn(B, N, 64, generator=g)
    s0 = torch.randn(B, K, 192, generator=g)
    s, a = F.slot_attention(x.cuda(), s0.cuda(), _cuda(p), T, kv="bf16", factored=True)
    assert F.last_kernel() == "tcgen05_xhat"
    idx = sorted({0, B // 2, B - 1})
    s_ref, a_ref = so.slot_attention(x[idx], s0[idx], p, T)
    assert rel_err(s[idx].cpu(), s_ref) < BF16_TOL, rel_err(s[idx].cpu(), s_ref)
    assert rel_err(a[idx].cpu(), a_ref) < BF16_TOL
    assert torch.isfinite(s).all() and torch.allclose(a.sum(-1), torch.ones_like(a.sum(-1)), atol=1e-4)


def test_factored_loop_beyond_the_op_table():
    """More ops per cluster than the kernel's op table holds (2048): the per-op bookkeeping falls back to divisions."""
    from ocrl_b200 import abi, functional as F

    B, N, K, T = 1500, 64, 4, 3   # two clusters: 750 images x 3 iterations = 2250 ops each
    p = so.random_sa_params(K, 64, 192, 192, seed=21)
    g = torch.Generator().manual_seed(9)
    x, s0 = torch.randn(B, N, 64, generator=g), torch.randn(B, K, 192, generator=g)
    s, a = F.slot_attention(x.cuda(), s0.cuda(), _cuda(p), T, kv="bf16", factored=True, opts=abi.launch_opts(max_clusters=2))
    assert F.last_kernel() == "tcgen05_xhat"
    idx = [0, 1, 749, 750, 1498, 1499]
    s_ref, a_ref = so.slot_attention(x[idx], s0[idx], p, T)
    assert rel_err(s[idx].cpu(), s_ref) < BF16_TOL and rel_err(a[idx].cpu(), a_ref) < BF16_TOL
    s2, _ = F.slot_attention(x.cuda(), s0.cuda(), _cuda(p), T, kv="bf16", factored=True)  # table path, every cluster
    assert torch.equal(s, s2)


def test_factored_full_size_properties():
    """BASELINE size (B = 64, N = 4096, K = 6, T = 3): finite, rows sum to one, batch-permutation equivariant bit for
    bit, independent of the cluster cap, within tolerance of the oracle and of the k/v form."""
    from ocrl_b200 import abi, functional as F

    torch.manual_seed(0)
    p = _cuda(so.random_sa_params(6, 64, 192, 192, seed=3))
    x = torch.randn(64, 4096, 64, device="cuda")
    s0 = torch.randn(64, 6, 192, device="cuda")
    s, a = F.slot_attention(x, s0, p, 3, kv="bf16", factored=True)
    assert F.last_kernel() == "tcgen05_xhat"
    assert torch.isfinite(s).all() and torch.isfinite(a).all()
    assert torch.allclose(a.sum(-1), torch.ones(64, 4096, device="cuda"), atol=1e-4)
    perm = torch.randperm(64, device="cuda")
    s2, a2 = F.slot_attention(x[perm].contiguous(), s0[perm].contiguous(), p, 3, kv="bf16", factored=True)
    assert torch.equal(s2, s[perm]) and torch.equal(a2, a[perm])
    s3, a3 = F.slot_attention(x, s0, p, 3, kv="bf16", factored=True, opts=abi.launch_opts(max_clusters=5))
    assert torch.equal(s3, s) and torch.equal(a3, a)
    pc = {k: v.cpu() for k, v in p.items()}
    sr, ar = so.slot_attention(x[:2].cpu(), s0[:2].cpu(), pc, 3)
    assert rel_err(s[:2].cpu(), sr) < BF16_TOL and rel_err(a[:2].cpu(), ar) < BF16_TOL
    sk, ak = F.slot_attention(x, s0, p, 3, kv="bf16", factored=False)
    assert rel_err(s, sk) < BF16_TOL and rel_err(a, ak) < BF16_TOL


def test_factored_prepared_weights_follow_parameter_updates():
    """The folded weights (W_q'' = s W_k^T W_q, W_ih'' = W_ih W_v) are prepared once per parameter version: changing
    project_k in place must change the result of the next call."""
    from ocrl_b200 import functional as F

    p = _cuda(so.random_sa_params(6, 64, 192, 192, seed=4))
    g = torch.Generator().manual_seed(1)
    x, s0 = torch.randn(2, 256, 64, generator=g).cuda(), torch.randn(2, 6, 192, generator=g).cuda()
    prep = F.PreparedWeights()
    s1, _ = F.slot_attention(x, s0, p, 3, kv="bf16", prepared=prep)
    s1b, _ = F.slot_attention(x, s0, p, 3, kv="bf16", prepared=prep)  # second call: prepared = 1
    assert torch.equal(s1, s1b)
    p["project_k.weight"].mul_(1.5)
    s2, _ = F.slot_attention(x, s0, p, 3, kv="bf16", prepared=prep)
    pc = {k: v.cpu() for k, v in p.items()}
    s_ref, _ = so.slot_attention(x.cpu(), s0.cpu(), pc, 3)
    assert rel_err(s2.cpu(), s_ref) < BF16_TOL and not torch.equal(s1, s2)


def test_factored_rejects_what_it_does_not_cover():
    """fp32 parity mode and training never take the factored kernels; the C entry points say so instead of guessing."""
    import ctypes

    from ocrl_b200 import abi, functional as F

    assert not F.factored_covers(64, 64, 128, 6, "bf16") and not F.factored_covers(64, 192, 192, 6, "fp32")
    dims = abi.make_dims(1, 64, 64, 64, 128, 6, 3, kv_dtype=abi.DT_BF16, math_mode=abi.MATH_TENSOR)
    t = torch.zeros(1 << 20, device="cuda")
    w = abi.SaWeights(*[abi.ptr(t)] * 13)
    rc = abi.lib().ocrl_sa_iter_fwd_xhat(ctypes.byref(dims), abi.ptr(t), abi.ptr(t), abi.ptr(t), abi.ptr(t), ctypes.byref(w),
                                         abi.ptr(t), None, abi.ptr(t), None, abi.stream_ptr())
    assert rc == -1 and b"not instantiated" in abi.lib().ocrl_last_error()
    p = _cuda(so.random_sa_params(6, 64, 192, 192, seed=4))
    with pytest.raises(RuntimeError):
        F.kv_project(torch.randn(1, 64, 64, device="cuda"), p, kv="fp32", xhat_only=True)

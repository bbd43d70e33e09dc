"""GPU parity of the drop-in modules (SLATE / SLATE_Module / SlotAttentionEncoder / SlotAttention)
against the frozen reference outputs."""
import os

import pytest
import torch

import ocrl_b200
from ocrl_b200.config import slate_config, slot_attention_config
from tests.golden_io import load_case, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(autouse=True)
def _fp32_convs():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield


def _inject_noise(encoder, noise):
    def init_slots(batch, like):
        return encoder.slot_mu + torch.exp(encoder.slot_log_sigma) * noise.to(like.device)

    encoder.init_slots = init_slots


def _load_hot(module, params):
    sd = module.state_dict()
    for k, v in params.items():
        assert k in sd and sd[k].shape == v.shape, k
        sd[k] = v
    module.load_state_dict(sd)


@pytest.mark.parametrize("name,cfg", [("slate_encode_64", slate_config()),
                                      ("bcdec_encode_32", slot_attention_config(obs_size=32))])
def test_slate_call_matches_reference(name, cfg):
    meta, g = load_case(name)
    model = ocrl_b200.SLATE(*cfg)
    _load_hot(model._module, g["p"])
    model.to("cuda")
    model.eval()
    _inject_noise(model._module._slotattn, g["in"]["noise"])
    obs = (g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0).cuda()
    with torch.no_grad():
        slots = model(obs)
        slots_m, masks = model(obs, with_masks=True)
        _, attns = model(obs, with_attns=True)
    assert slots.shape == g["out"]["slots"].shape and masks.shape == g["out"]["masks"].shape
    assert rel_err(slots.cpu(), g["out"]["slots"]) < TOL
    assert torch.equal(slots, slots_m)
    assert rel_err(masks.cpu(), g["out"]["masks"]) < TOL
    assert rel_err(attns.cpu(), g["out"]["attns"]) < TOL
    # argmax segmentation (what calculate_ari consumes, utils/tools.py:309-320)
    mine, ref = masks.cpu().flatten(2).argmax(1), g["out"]["masks"].flatten(2).argmax(1)
    top2 = g["out"]["masks"].flatten(2).topk(2, dim=1).values
    decidable = (top2[:, 0] - top2[:, 1]) > 1e-4
    assert bool(((mine == ref) | ~decidable).all())
    # the bare module also works with a single argument (sb3s/ocr_extractor.py:45)
    with torch.no_grad():
        assert torch.equal(model._module(obs), slots)


def test_encoder_module_matches_reference():
    meta, g = load_case("encoder_slate")
    enc = ocrl_b200.SlotAttentionEncoder(meta["T"], meta["K"], meta["C"], meta["D"], meta["H"], 4, 1)
    enc.load_state_dict(g["p"])
    enc.cuda()
    _inject_noise(enc, g["in"]["noise"])
    with torch.no_grad():
        slots, attn = enc(g["in"]["x"].cuda())
    assert rel_err(slots.cpu(), g["out"]["slots"]) < TOL and rel_err(attn.cpu(), g["out"]["attn"]) < TOL


@pytest.mark.parametrize("name", ["sa_small_grad", "sa_k11_t5_ragged"])
def test_slot_attention_module_matches_reference(name):
    meta, g = load_case(name)
    sa = ocrl_b200.SlotAttention(meta["T"], meta["K"], meta["C"], meta["D"], meta["H"], 1, meta["eps"])
    sa.load_state_dict(g["p"])
    sa.cuda()
    x, s0 = g["in"]["inputs"].cuda(), g["in"]["slots0"].cuda()
    with torch.no_grad():
        slots, attn = sa(x, s0)
    assert rel_err(slots.cpu(), g["out"]["slots"]) < TOL and rel_err(attn.cpu(), g["out"]["attn"]) < TOL
    # inputs are not mutated, outputs are fresh contiguous fp32 tensors
    assert torch.equal(x.cpu(), g["in"]["inputs"]) and slots.is_contiguous() and slots.dtype == torch.float32


def test_bf16_mode_through_the_module(monkeypatch):
    meta, g = load_case("slate_encode_64")
    monkeypatch.setenv("OCRL_KV_DTYPE", "bf16")
    model = ocrl_b200.SLATE(*slate_config())
    assert model._module._slotattn.slot_attention.kv_dtype == "bf16"
    _load_hot(model._module, g["p"])
    model.to("cuda")
    _inject_noise(model._module._slotattn, g["in"]["noise"])
    obs = (g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0).cuda()
    with torch.no_grad():
        slots, masks = model(obs, with_masks=True)
    assert rel_err(slots.cpu(), g["out"]["slots"]) < 2e-2
    assert rel_err(masks.cpu(), g["out"]["masks"]) < 2e-2


@pytest.mark.parametrize("conv", ["tf32", "bf16", "bf16-unfused", "bf16-ocrlepilogue"])
def test_bf16_mode_with_tensor_core_convs(monkeypatch, conv):
    """bf16 mode end to end: channels-last cuDNN convs (TF32, bf16 with fused bias+ReLU and the last bias folded
    into the position table, or plain bf16 autocast) feeding the tcgen05 token stage (bf16 tokens when the convs
    are bf16) and the tensor-core iteration kernel; 2e-2 tolerance."""
    meta, g = load_case("slate_encode_64")
    monkeypatch.setenv("OCRL_KV_DTYPE", "bf16")
    monkeypatch.setenv("OCRL_CONV_FUSED", "0" if conv.endswith("unfused") else "1")
    monkeypatch.setenv("OCRL_CONV_EPILOGUE", "ocrl" if conv.endswith("ocrlepilogue") else "cudnn")
    conv = conv.split("-")[0]
    monkeypatch.setenv("OCRL_CONV_DTYPE", conv)
    model = ocrl_b200.SLATE(*slate_config())
    _load_hot(model._module, g["p"])
    model.to("cuda")
    model.eval()
    _inject_noise(model._module._slotattn, g["in"]["noise"])
    obs = (g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0).cuda()
    with torch.no_grad():
        slots, masks = model(obs, with_masks=True)
    es, em = rel_err(slots.cpu(), g["out"]["slots"]), rel_err(masks.cpu(), g["out"]["masks"])
    print(f"bf16 mode, {conv} convs: slots rel err {es:.2e}, masks rel err {em:.2e}")
    assert es < 2e-2 and em < 2e-2


@pytest.mark.parametrize("kv", ["fp32", "bf16"])
@pytest.mark.parametrize("iter_clusters", [None, 2, 8, "auto"])
def test_streamed_encoder_matches_direct_calls(iter_clusters, kv):
    """StreamedEncoder (double-buffered H2D / graph replay / D2H) returns, batch by batch, what the model returns for
    the same frames and the same slot-initialisation noise -- in the exact-parity mode and in the bf16 mode the
    bench runs (persistent tcgen05 iteration kernel, whose grid the cluster cap changes) -- and the golden output."""
    from ocrl_b200 import functional as F

    meta, g = load_case("slate_encode_64")
    model = ocrl_b200.SLATE(*slate_config(kv_dtype=kv))
    _load_hot(model._module, g["p"])
    model.to("cuda")
    model.eval()
    _inject_noise(model._module._slotattn, g["in"]["noise"].cuda())  # device tensor: the call is graph-captured
    obs = (g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0)
    batches = [obs, obs.flip(0), obs.roll(1, 0), obs]
    with torch.no_grad():
        want = [model(b.cuda()).cpu() for b in batches]
    env_before = dict(os.environ)
    enc = ocrl_b200.StreamedEncoder(model, obs.cuda(), iter_clusters=iter_clusters)  # cap changes the grid, not the result
    assert dict(os.environ) == env_before  # the cap travels in ocrl_sa_launch_opts, not in the environment
    pinned = [b.contiguous().pin_memory() for b in batches]
    outs = [torch.empty_like(want[0]).pin_memory() for _ in batches]
    for b, o in zip(pinned, outs):
        enc.submit(b, o)
    enc.synchronize()
    for w, o in zip(want, outs):
        assert torch.equal(w, o)
    tol = 2e-2 if kv == "bf16" else 1e-4
    assert rel_err(outs[0], g["out"]["slots"]) < tol and rel_err(outs[3], g["out"]["slots"]) < tol
    if kv == "bf16":  # the module's bf16 inference = the factored form of the tcgen05 kernel (tests/test_gpu_factored.py)
        assert F.last_kernel() == "tcgen05_xhat"


@pytest.mark.parametrize("name", ["loss_slate_16", "loss_bcdec_16"])
def test_get_loss_matches_reference_on_gpu(name):
    """``get_loss`` on the GPU (fused slot-attention kernels inside, dVAE / decoders through torch) against the frozen
    reference numbers: ``dvae_mse`` / ``cross_entropy`` / ``loss`` for SLATE, ``mse`` for the broadcast decoder
    (slate_module.py:198-233), gumbel and slot noise of the reference run, dropout off."""
    from oracle.make_golden import preset_exponential

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    meta, g = load_case(name)
    torch.manual_seed(meta["seed"])
    model = ocrl_b200.SLATE(*slate_config(num_slots=meta["K"], num_iterations=meta["T"], slot_size=meta["D"],
                                          mlp_hidden_size=meta["H"], obs_size=meta["S"], use_bcdec=meta["use_bcdec"]))
    psum = float(sum(p.detach().double().sum() for p in model._module.parameters()))
    assert abs(psum - meta["param_sum"]) < 1e-6 * meta["param_abs_sum"]
    model.to("cuda")
    model.eval()
    _inject_noise(model._module._slotattn, g["in"]["noise"].cuda())
    obs = (g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0).cuda()
    with preset_exponential([g["in"][f"exponential{i}"] for i in range(meta["n_draws"])]):
        m = model.get_loss(obs, None)
    for k, want in g["out"].items():
        got = m[k].detach().cpu().reshape(1)
        assert rel_err(got, want) < 1e-4, (k, float(got), float(want))
    m["loss"].backward()  # the loss carries the graph through the fused backward
    assert all(torch.isfinite(p.grad).all() for p in model._module._slotattn.parameters() if p.grad is not None)


@pytest.mark.parametrize("B,S", [(5, 6), (1, 1), (32, 16), (4, 7)])
def test_pooling_transformer_kernel_matches_reference(B, S):
    """The fused rollout forward of the PPO consumer's slot pooling (ocrl_pool_transformer_fwd) against the frozen
    reference output (B = 5, S = 6) and, on other batch / slot counts, against the oracle: 1e-4."""
    from oracle import pool_oracle as po
    from tests.test_oracle_golden import _pool_module

    meta, g = load_case("pool_transformer")
    t = _pool_module(meta).cuda()
    p = {k: v.detach().cpu() for k, v in t.state_dict().items()}
    if (B, S) == (meta["B"], meta["S"]):
        slots, want = g["in"]["slots"], g["out"]["pooled"]
    else:
        slots = torch.randn(B, S, meta["Din"], generator=torch.Generator().manual_seed(B * 100 + S))
        want = po.transformer_pool(slots, p, meta["nhead"])
    with torch.no_grad():
        got = t(slots.cuda())  # no autograd graph, eval mode -> the fused kernel
        torch_path = t._trans(torch.cat([t._cls_token().repeat(B, 1, 1), t._linear(slots.cuda())], dim=1).permute(1, 0, 2))[0]
    assert got.shape == (B, meta["d_model"])
    assert rel_err(got.cpu(), want) < 1e-4, rel_err(got.cpu(), want)
    assert rel_err(got.cpu(), torch_path.cpu()) < 1e-4
    # training pass (PPO minibatch update): gradients flow through the torch path of the same module
    t.train()
    out = t(slots.cuda().requires_grad_(True))
    assert out.requires_grad


def test_rollout_extractor_graph_follows_pooling_updates():
    """pooling(ocr(obs)) captured in one CUDA graph (RolloutExtractor + GraphedEncoder): equal to the eager calls, and
    captured again after an optimizer-style in-place update of the pooling's weights (PPO trains the pooling)."""
    from types import SimpleNamespace as NS

    meta, g = load_case("slate_encode_64")
    model = ocrl_b200.SLATE(*slate_config(kv_dtype="bf16"))
    _load_hot(model._module, g["p"])
    model.to("cuda")
    model.eval()
    _inject_noise(model._module._slotattn, g["in"]["noise"].cuda())
    pcfg = NS(d_model=128, nhead=8, num_layers=1, pos_emb="None", norm_first=False, use_mlp1=False, use_mlp2=False,
              cw_embedding=False, push_embedding=False)
    torch.manual_seed(3)
    pool = ocrl_b200.Transformer_Module(model.rep_dim, model.num_slots, pcfg).cuda().eval()
    obs = (g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0).contiguous().cuda()
    enc = ocrl_b200.GraphedEncoder(ocrl_b200.RolloutExtractor(model, pool), obs)
    with torch.no_grad():
        want = pool(model(obs))
        got = enc(obs).clone()
        assert got.shape == (obs.shape[0], 128) and torch.equal(got, want)
        pool._trans._linear.weight.mul_(1.5)  # an optimizer step on the pooling
        assert enc.stale()
        want2 = pool(model(obs))
        got2 = enc(obs).clone()
    assert not torch.equal(want2, want) and torch.equal(got2, want2)


def test_graphs_survive_calls_with_another_batch_size():
    """A captured graph keeps reading the prepared weight copies of ITS batch size while the same module serves another
    batch size (rollout batch next to a minibatch): the per-shape workspaces of functional.PreparedWeights must not be
    replaced under a live graph (regression: a single-entry cache freed the workspace the graph pointed at)."""
    meta, g = load_case("slate_encode_64")
    model = ocrl_b200.SLATE(*slate_config(kv_dtype="bf16"))
    _load_hot(model._module, g["p"])
    model.to("cuda")
    model.eval()
    obs = (g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0).contiguous().cuda()
    big = torch.cat([obs, obs.flip(0), obs], dim=0)  # batch 6
    noise_big, noise_small = torch.zeros(6, 6, 192, device="cuda"), torch.zeros(2, 6, 192, device="cuda")  # kept alive: the graphs read them
    _inject_noise(model._module._slotattn, noise_big)
    enc_big = ocrl_b200.GraphedEncoder(model, big)
    want_big = enc_big(big).clone()
    _inject_noise(model._module._slotattn, noise_small)
    with torch.no_grad():
        for _ in range(3):
            small = model(obs)  # another shape: its own prepared workspace
            junk = [torch.randn(1 << 20, device="cuda") for _ in range(4)]  # churn the allocator over freed blocks
            del junk
    enc_small = ocrl_b200.GraphedEncoder(model, obs)
    assert torch.equal(enc_small(obs), small)
    assert torch.equal(enc_big(big), want_big)

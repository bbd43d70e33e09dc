"""Host-logic tests on the CPU: the CUDA operators are replaced (in the test only) by the oracle,
so that the Python plumbing around them -- module wiring, shapes, get_loss / update / save /
load, the data-parallel gradient exchange -- is exercised without a GPU."""
import pytest
import torch

import ocrl_b200
from ocrl_b200 import functional as F
from ocrl_b200 import slot_attn
from ocrl_b200.config import slate_config, slot_attention_config
from oracle import slot_oracle as so
from tests.golden_io import load_case, rel_err


def _oracle_slot_attention(inputs, slots0, p, T, *, epsilon=1e-8, kv="fp32", enc=None, pos_table=None,
                           want_attn=True, opts=None, prepared=None):
    if pos_table is not None:
        B, C, H, W = inputs.shape
        inputs = (inputs + pos_table.view(1, C, H, W)).permute(0, 2, 3, 1).flatten(1, 2)
    if enc is not None:
        inputs = so.token_mlp(inputs, enc)
    return so.slot_attention(inputs, slots0, p, T, epsilon)


class _OracleFn:
    @staticmethod
    def apply(inputs, slots0, T, epsilon, kv, *params):
        return so.slot_attention(inputs, slots0, dict(zip(F.SA_PARAM_ORDER, params)), T, epsilon)


@pytest.fixture
def cpu_kernels(monkeypatch):
    monkeypatch.setattr(F, "slot_attention", _oracle_slot_attention)
    monkeypatch.setattr(F, "SlotAttentionFunction", _OracleFn)
    orig = slot_attn.SlotAttention._check

    def check(self, inputs, slots, fmap=False):
        class _Fake:  # pretend the tensors are on the GPU for the device check only
            is_cuda = True
            dim = inputs.dim
            shape = inputs.shape
        return orig(self, _Fake(), slots, fmap)

    monkeypatch.setattr(slot_attn.SlotAttention, "_check", check)


def _inject_noise(encoder, noise):
    encoder.init_slots = lambda batch, like: encoder.slot_mu + torch.exp(encoder.slot_log_sigma) * noise


def test_slate_call_plumbing(cpu_kernels):
    meta, g = load_case("slate_encode_64")
    model = ocrl_b200.SLATE(*slate_config())
    sd = model._module.state_dict()
    sd.update(g["p"])
    model._module.load_state_dict(sd)
    model.eval()
    _inject_noise(model._module._slotattn, g["in"]["noise"])
    obs = g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0
    with torch.no_grad():
        slots = model(obs)
        _, masks = model(obs, with_masks=True)
        _, attns = model(obs, with_attns=True)
    assert rel_err(slots, g["out"]["slots"]) < 1e-5
    assert rel_err(masks, g["out"]["masks"]) < 1e-5 and rel_err(attns, g["out"]["attns"]) < 1e-5


@pytest.mark.parametrize("cfg", [slate_config(obs_size=16), slot_attention_config(obs_size=16)])
def test_get_loss_update_and_checkpoint_round_trip(cpu_kernels, cfg):
    torch.manual_seed(0)
    model = ocrl_b200.SLATE(*cfg)
    model.train()
    obs = torch.rand(2, 3, 16, 16)
    metrics = model.get_loss(obs, None)
    assert metrics["loss"].requires_grad
    expected = {"loss", "mse", "ari"} if cfg[0].use_bcdec else {"loss", "dvae_mse", "cross_entropy", "tau"}
    assert expected | {"lr_dvae", "lr_enc", "lr_dec"} == set(metrics)
    before = {k: v.clone() for k, v in model._module.state_dict().items()}
    out = model.update(obs, None, step=0)
    assert "norm" in out and torch.isfinite(out["loss"])
    moved = [k for k, v in model._module.state_dict().items() if v.is_floating_point() and not torch.equal(v, before[k])]
    assert any(k.startswith("_slotattn.slot_attention.") for k in moved)
    assert any(k.startswith("_enc.") for k in moved)
    ckpt = model.save()
    clone = ocrl_b200.SLATE(*cfg)
    clone.load(ckpt)
    for (k, a), (_, b) in zip(model._module.state_dict().items(), clone._module.state_dict().items()):
        assert torch.equal(a, b), k
    with torch.no_grad():
        masks = torch.zeros(2, 4, 1, 16, 16)
        masks[:, -1] = 1.0
        m2 = model.get_loss(obs, masks)
    assert "loss" in m2


def test_get_samples_shapes(cpu_kernels):
    model = ocrl_b200.SLATE(*slot_attention_config(obs_size=16))
    model.eval()
    with torch.no_grad():
        out = model.get_samples(torch.rand(2, 3, 16, 16))
    assert out["samples"].dtype.name == "uint8" and out["samples"].shape == (2, 16, 16 * (2 + 6), 3)


@pytest.mark.parametrize("name", ["loss_slate_16", "loss_bcdec_16"])
def test_get_loss_matches_reference(cpu_kernels, name):
    """``get_loss`` (slate_module.py:198-233) with the gumbel and slot noise of the frozen reference run: dVAE,
    transformer decoder / broadcast decoder and the loss arithmetic of ocrl_b200.adjacent against the reference's numbers."""
    from oracle.make_golden import preset_exponential

    meta, g = load_case(name)
    torch.manual_seed(meta["seed"])
    model = ocrl_b200.SLATE(*slate_config(num_slots=meta["K"], num_iterations=meta["T"], slot_size=meta["D"],
                                          mlp_hidden_size=meta["H"], obs_size=meta["S"], use_bcdec=meta["use_bcdec"]))
    psum = float(sum(p.detach().double().sum() for p in model._module.parameters()))
    assert abs(psum - meta["param_sum"]) < 1e-6 * meta["param_abs_sum"]  # same seeded parameters as the reference run
    model.eval()
    _inject_noise(model._module._slotattn, g["in"]["noise"])
    obs = g["in"]["frames_u8"].permute(0, 3, 1, 2).float() / 255.0
    with preset_exponential([g["in"][f"exponential{i}"] for i in range(meta["n_draws"])]):
        m = model.get_loss(obs, None)
    for k, want in g["out"].items():
        assert rel_err(m[k].detach().reshape(1), want) < 1e-5, (k, float(m[k]), float(want))

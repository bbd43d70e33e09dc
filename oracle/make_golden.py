"""Freeze outputs of the REAL reference into tests/golden/ (run in the build container).

    python -m oracle.make_golden            # needs /root/reference (read-only)

TEST INFRASTRUCTURE.  The reference has no tests or golden vectors of its own, so these fixtures
are outputs of the reference modules themselves (imported through oracle/reference_bridge.py)
on seeded inputs; CPU fp32, torch as installed in the image.  Each .npz holds the parameters
(reference state_dict names), the inputs, the outputs and, where stated, autograd gradients.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import reference_bridge as rb  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _perturb(module: torch.nn.Module, seed: int) -> None:
    """Zero biases / unit LayerNorm affines hide terms; give them seeded non-trivial values."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, prm in module.named_parameters():
            if name.endswith("bias"):
                prm.add_(0.1 * torch.randn(prm.shape, generator=g))
            elif "norm" in name and name.endswith("weight"):
                prm.add_(0.1 * torch.randn(prm.shape, generator=g))


def _np(d):
    return {k: v.detach().cpu().numpy() for k, v in d.items()}


def make_sa_case(ref, name, *, T, K, C, D, H, B, N, seed, with_grads, input_scale=1.0):
    torch.manual_seed(seed)
    sa = ref.SlotAttention(T, K, C, D, H, 1).double().float()
    _perturb(sa, seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    x = (input_scale * torch.randn(B, N, C, generator=g) + 0.3).requires_grad_(with_grads)
    s0 = torch.randn(B, K, D, generator=g).requires_grad_(with_grads)
    slots, attn = sa(x, s0)
    out = {"in.inputs": x, "in.slots0": s0, "out.slots": slots, "out.attn": attn}
    out.update({"p." + k: v for k, v in sa.state_dict().items()})
    meta = dict(T=T, K=K, C=C, D=D, H=H, B=B, N=N, eps=1e-8)
    if with_grads:
        gs = torch.randn(slots.shape, generator=g)
        ga = torch.randn(attn.shape, generator=g)
        loss = (slots * gs).sum() + (attn * ga).sum()
        params = dict(sa.named_parameters())
        grads = torch.autograd.grad(loss, [x, s0] + list(params.values()))
        out["g.out.slots"], out["g.out.attn"] = gs, ga
        out["g.in.inputs"], out["g.in.slots0"] = grads[0], grads[1]
        for (k, _), gv in zip(params.items(), grads[2:]):
            out["g.p." + k] = gv
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), **_np(out))
    print(name, "slots.sum", float(slots.sum()), "attn.sum", float(attn.sum()))


def make_encoder_case(ref, name, *, T, K, C, D, H, B, N, seed):
    torch.manual_seed(seed)
    enc = ref.SlotAttentionEncoder(T, K, C, D, H, 4, 1)
    _perturb(enc, seed + 1)
    g = torch.Generator().manual_seed(seed + 2)
    x = torch.randn(B, N, C, generator=g)
    torch.manual_seed(seed + 3)
    noise = torch.empty(B, K, D).normal_()  # the draw slot_attn.py:155 will make
    torch.manual_seed(seed + 3)
    with torch.no_grad():
        slots, attn = enc(x)
    out = {"in.x": x, "in.noise": noise, "out.slots": slots, "out.attn": attn}
    out.update({"p." + k: v for k, v in enc.state_dict().items()})
    meta = dict(T=T, K=K, C=C, D=D, H=H, B=B, N=N, eps=1e-8)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), **_np(out))
    print(name, "slots.sum", float(slots.sum()))


def _frames(B, S, seed):
    """A few flat-colour rectangles on black; uint8 HWC like the HDF5 datasets (utils/datasets.py:17)."""
    rng = np.random.RandomState(seed)
    out = np.zeros((B, S, S, 3), np.uint8)
    cols = np.array([[0, 0, 255], [0, 128, 0], [255, 255, 0], [255, 0, 0]], np.uint8)
    for b in range(B):
        for _ in range(5):
            w, h = rng.randint(S // 8, S // 4, size=2)
            x0, y0 = rng.randint(0, S - w), rng.randint(0, S - h)
            out[b, y0 : y0 + h, x0 : x0 + w] = cols[rng.randint(4)]
    return out


def make_slate_case(ref, name, *, B, S, seed, K=6, T=3, D=192, H=192, use_bcdec=False):
    ocr, env = rb.slate_config(num_slots=K, num_iterations=T, slot_size=D, mlp_hidden_size=H,
                               use_bcdec=use_bcdec, obs_size=S)
    torch.manual_seed(seed)
    model = ref.SLATE(ocr, env)
    model.eval()
    _perturb(model._module._enc, seed + 1)
    _perturb(model._module._slotattn, seed + 2)
    frames = _frames(B, S, seed + 3)
    obs = torch.from_numpy(frames).permute(0, 3, 1, 2).float() / 255.0
    torch.manual_seed(seed + 4)
    noise = torch.empty(B, K, D).normal_()
    with torch.no_grad():
        torch.manual_seed(seed + 4)
        slots = model(obs)
        torch.manual_seed(seed + 4)
        slots_m, masks = model(obs, with_masks=True)
        torch.manual_seed(seed + 4)
        _, attns = model(obs, with_attns=True)
    assert torch.equal(slots, slots_m)
    sd = model._module.state_dict()
    hot = {k: v for k, v in sd.items() if k.startswith(("_enc.", "_enc_pos.", "_slotattn."))}
    out = {"in.frames_u8": torch.from_numpy(frames), "in.noise": noise, "out.slots": slots,
           "out.masks": masks, "out.attns": attns}
    out.update({"p." + k: v for k, v in hot.items()})
    meta = dict(T=T, K=K, C=64, D=D, H=H, B=B, N=S * S, S=S, eps=1e-8)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), **_np(out))
    keys = {k: list(v.shape) for k, v in sd.items()}
    tag = "bcdec" if use_bcdec else "slate"
    with open(os.path.join(GOLDEN, f"state_dict_{tag}.json"), "w") as f:
        json.dump(dict(keys=keys, rep_dim=model.rep_dim, num_slots=model.num_slots,
                       n_params=sum(p.numel() for p in model._module.parameters())), f, indent=0)
    print(name, "slots.sum", float(slots.sum()), "n_state", len(keys))
    return model


def make_survey_kat(ref):
    """SURVEY.md section 8(c) known-answer test, regenerated (seed-defined, no stored tensors)."""
    torch.manual_seed(0)
    sa = ref.SlotAttention(3, 6, 64, 192, 192, 1).eval()
    g = torch.Generator().manual_seed(1234)
    x = torch.randn(4, 4096, 64, generator=g)
    s0 = torch.randn(4, 6, 192, generator=g)
    with torch.no_grad():
        slots, attn = sa(x, s0)
    kat = dict(slots_sum=float(slots.sum()), slots_abs_mean=float(slots.abs().mean()),
               attn_sum=float(attn.sum()), attn00=[float(a) for a in attn[0, 0]],
               argmax_hist=torch.bincount(attn.argmax(-1).flatten(), minlength=6).tolist())
    with open(os.path.join(GOLDEN, "survey_kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print("survey_kat", kat)


def make_training_case(ref, name, *, B, S, seed, steps, use_bcdec):
    """A few reference SLATE.update() steps with dropout off and tau/gumbel noise frozen out of the
    compared quantities: records the hot-path gradient of step 0 for a fixed loss on slots."""
    K, T, D, H = (6, 3, 64, 128)
    ocr, env = rb.slate_config(num_slots=K, num_iterations=T, slot_size=D, mlp_hidden_size=H,
                               use_bcdec=use_bcdec, obs_size=S)
    torch.manual_seed(seed)
    model = ref.SLATE(ocr, env)
    mod = model._module
    mod.eval()
    _perturb(mod._enc, seed + 1)
    _perturb(mod._slotattn, seed + 2)
    frames = _frames(B, S, seed + 3)
    obs = torch.from_numpy(frames).permute(0, 3, 1, 2).float() / 255.0
    torch.manual_seed(seed + 4)
    noise = torch.empty(B, K, D).normal_()
    torch.manual_seed(seed + 4)
    slots, attns = mod._get_slots(obs, with_attns=True)
    g = torch.Generator().manual_seed(seed + 5)
    gs, ga = torch.randn(slots.shape, generator=g), torch.randn(attns.shape, generator=g)
    loss = (slots * gs).sum() + (attns * ga).sum()
    names = [n for n, _ in mod.named_parameters() if n.startswith(("_enc.", "_enc_pos.", "_slotattn."))]
    prms = [dict(mod.named_parameters())[n] for n in names]
    grads = torch.autograd.grad(loss, prms)
    out = {"in.frames_u8": torch.from_numpy(frames), "in.noise": noise, "out.slots": slots,
           "out.attn": attns, "g.out.slots": gs, "g.out.attn": ga}
    sd = mod.state_dict()
    out.update({"p." + k: v for k, v in sd.items() if k.startswith(("_enc.", "_enc_pos.", "_slotattn."))})
    out.update({"g.p." + n: gv for n, gv in zip(names, grads)})
    meta = dict(T=T, K=K, C=64, D=D, H=H, B=B, N=S * S, S=S, eps=1e-8)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), **_np(out))
    print(name, "loss", float(loss))


class preset_exponential:
    """Context manager: ``Tensor.exponential_`` returns the given tensors in order (the gumbel noise of
    ocrs/common/utils.py:75-85 is ``-log(exponential_())``) -- the one source of randomness of ``get_loss`` that cannot
    be frozen by a seed across CPU and CUDA generators.  With ``record=True`` it keeps the draws instead."""

    def __init__(self, draws=None, record=False):
        self.draws, self.record, self.i = ([] if draws is None else list(draws)), record, 0

    def __enter__(self):
        self._orig = torch.Tensor.exponential_
        outer = self

        def fake(t, *a, **k):
            if outer.record:
                outer._orig(t, *a, **k)
                outer.draws.append(t.detach().clone().cpu())
                return t
            d = outer.draws[outer.i]
            outer.i += 1
            assert tuple(d.shape) == tuple(t.shape), (d.shape, t.shape)
            return t.copy_(d.to(t.device))

        torch.Tensor.exponential_ = fake
        return self

    def __exit__(self, *exc):
        torch.Tensor.exponential_ = self._orig
        return False


def make_loss_case(ref, name, *, B, S, seed, use_bcdec):
    """``SLATE.get_loss`` of the reference (slate_module.py:198-233) with dropout off (eval mode) and the gumbel noise and
    the slot-initialisation noise frozen: pins ``dvae_mse`` / ``cross_entropy`` (SLATE) or ``mse`` (broadcast decoder).
    The 5.6 M parameters are not stored: both sides build the module from ``torch.manual_seed(seed)`` (the init
    families are equal draw for draw, tests/test_interface.py) and the fixture carries a checksum."""
    K, T, D, H = 6, 3, 64, 128
    ocr, env = rb.slate_config(num_slots=K, num_iterations=T, slot_size=D, mlp_hidden_size=H,
                               use_bcdec=use_bcdec, obs_size=S)
    torch.manual_seed(seed)
    model = ref.SLATE(ocr, env)
    model.eval()
    frames = _frames(B, S, seed + 3)
    obs = torch.from_numpy(frames).permute(0, 3, 1, 2).float() / 255.0
    torch.manual_seed(seed + 4)
    noise = torch.empty(B, K, D).normal_()
    with torch.no_grad(), preset_exponential(record=True) as rec:  # first run: the gumbel draws get_loss makes
        model.get_loss(obs, None)
    torch.manual_seed(seed + 4)  # second run: the draws replayed, so the generator is consumed by the slot noise only
    with preset_exponential(rec.draws):
        metrics = model.get_loss(obs, None)
    out = {"in.frames_u8": torch.from_numpy(frames), "in.noise": noise}
    for i, d in enumerate(rec.draws):
        out[f"in.exponential{i}"] = d
    for k in ("loss", "dvae_mse", "cross_entropy", "mse"):
        if k in metrics:
            out["out." + k] = metrics[k].detach().reshape(1)
    checksum = float(sum(p.detach().double().sum() for p in model._module.parameters()))
    abs_sum = float(sum(p.detach().double().abs().sum() for p in model._module.parameters()))
    meta = dict(T=T, K=K, C=64, D=D, H=H, B=B, N=S * S, S=S, eps=1e-8, seed=seed, use_bcdec=use_bcdec,
                n_draws=len(rec.draws), param_sum=checksum, param_abs_sum=abs_sum, tau=float(model._module._tau))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), **_np(out))
    print(name, {k: float(v) for k, v in metrics.items() if torch.is_tensor(v) and v.numel() == 1}, "draws", len(rec.draws))


def pool_params(module, seed, d_model):
    """The seeded, perturbed parameters of a pooling Transformer (reference's or the drop-in: same construction order,
    hence the same draws): biases / LayerNorm weights perturbed, CLS token non-zero."""
    _perturb(module, seed + 1)
    with torch.no_grad():
        module._cls_token._cls_token.add_(0.3 * torch.randn(d_model, generator=torch.Generator().manual_seed(seed + 2)))


def make_pool_case(name, *, B, S, Din, d_model, nhead, seed):
    """The PPO consumer's slot pooling (poolings/common/transformer.py:9-33, one post-norm layer as in
    configs/pooling/transformer.yaml) on seeded slots; eval mode (dropout inactive).  The 0.6 M parameters are not
    stored: both sides build the module from ``torch.manual_seed(seed)`` + ``pool_params`` and the fixture carries a
    checksum."""
    import importlib

    mod = importlib.import_module("poolings.common.transformer")
    torch.manual_seed(seed)
    t = mod.Transformer(Din, d_model, nhead, 1, None, False)
    t.eval()
    pool_params(t, seed, d_model)
    g = torch.Generator().manual_seed(seed + 3)
    slots = 1.5 * torch.randn(B, S, Din, generator=g)
    with torch.no_grad():
        out = t(slots)
    data = {"in.slots": slots, "out.pooled": out}
    meta = dict(B=B, S=S, Din=Din, d_model=d_model, nhead=nhead, eps=1e-5, seed=seed,
                param_sum=float(sum(p.detach().double().sum() for p in t.parameters())),
                param_abs_sum=float(sum(p.detach().double().abs().sum() for p in t.parameters())),
                state_keys=list(t.state_dict().keys()))
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), meta=json.dumps(meta), **_np(data))
    print(name, "out.sum", float(out.sum()), tuple(out.shape))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref = rb.load()
    torch.set_num_threads(8)
    # SlotAttention.forward(inputs, slots) with gradients
    make_sa_case(ref, "sa_small_grad", T=3, K=6, C=64, D=64, H=128, B=2, N=256, seed=11, with_grads=True)
    make_sa_case(ref, "sa_slate_grad", T=3, K=6, C=64, D=192, H=192, B=2, N=320, seed=12, with_grads=True)
    # ragged N, other K / T
    make_sa_case(ref, "sa_k11_t5_ragged", T=5, K=11, C=64, D=64, H=128, B=3, N=100, seed=13, with_grads=True)
    make_sa_case(ref, "sa_k16_t7", T=7, K=16, C=64, D=192, H=192, B=1, N=1024, seed=14, with_grads=False)
    make_sa_case(ref, "sa_k1_t1", T=1, K=1, C=64, D=64, H=128, B=2, N=64, seed=15, with_grads=True)
    make_sa_case(ref, "sa_sharp", T=3, K=6, C=64, D=192, H=192, B=1, N=1024, seed=16, with_grads=False,
                 input_scale=6.0)
    # SlotAttentionEncoder.forward(x) (token LN+MLP, Gaussian slot init with the draw frozen)
    make_encoder_case(ref, "encoder_slate", T=3, K=6, C=64, D=192, H=192, B=2, N=1024, seed=21)
    # whole hot path through SLATE.__call__ on 64x64 and 32x32 frames
    make_slate_case(ref, "slate_encode_64", B=2, S=64, seed=31)
    make_slate_case(ref, "bcdec_encode_32", B=3, S=32, seed=32, K=6, T=7, D=64, H=128, use_bcdec=True)
    make_training_case(ref, "slate_path_grad_16", B=2, S=16, seed=41, steps=1, use_bcdec=False)
    make_survey_kat(ref)
    # get_loss of the adjacent training modules (dVAE, transformer decoder, broadcast decoder), noise frozen
    make_loss_case(ref, "loss_slate_16", B=2, S=16, seed=51, use_bcdec=False)
    make_loss_case(ref, "loss_bcdec_16", B=2, S=16, seed=52, use_bcdec=True)
    # PPO consumer: slot pooling transformer (rollout path)
    make_pool_case("pool_transformer", B=5, S=6, Din=192, d_model=128, nhead=8, seed=61)


if __name__ == "__main__":
    main()

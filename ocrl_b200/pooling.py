"""Drop-in slot pooling of the PPO consumer (reference: poolings/common/transformer.py:9-44,
poolings/transformer/transformer_module.py:29-133, used at sb3s/ocr_extractor.py:33-45).

``Transformer_Module(ocr_rep_dim, ocr_num_slots, config)`` keeps the reference's constructor, sub-module names and
``state_dict`` (``_trans._linear``, ``_trans._cls_token._cls_token``, ``_trans._trans.layers.0.*``), so checkpoints of
the reference load strictly.  During rollouts / evaluation (no autograd graph, dropout inactive) the forward is ONE
hand-written kernel behind the C ABI (``ocrl_pool_transformer_fwd``: only the CLS row of the layer's output is formed);
when gradients are needed (PPO's minibatch pass trains the pooling) or for configurations the kernel does not cover the
module runs the same arithmetic through torch, like the reference.
"""
from __future__ import annotations

import torch
from torch import nn

from . import abi


class ClsToken(nn.Module):
    def __init__(self, emb_size):
        super().__init__()
        self._cls_token = nn.Parameter(torch.zeros(emb_size))

    def forward(self):
        return self._cls_token


def pool_transformer(slots: torch.Tensor, p: dict, nhead: int, ln_eps: float = 1e-5) -> torch.Tensor:
    """The fused inference forward.  slots [B,S,Din] fp32 CUDA; p: the parameters of ``Transformer`` by state_dict name
    (``_linear.weight`` ... ``_trans.layers.0.norm2.bias``).  Returns [B, d_model]."""
    if not slots.is_cuda:
        raise RuntimeError("pool_transformer: ocrl_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    import ctypes

    B, S, Din = slots.shape
    L = "_trans.layers.0."
    names = ["_linear.weight", "_linear.bias", "_cls_token._cls_token", L + "self_attn.in_proj_weight",
             L + "self_attn.in_proj_bias", L + "self_attn.out_proj.weight", L + "self_attn.out_proj.bias",
             L + "linear1.weight", L + "linear1.bias", L + "linear2.weight", L + "linear2.bias",
             L + "norm1.weight", L + "norm1.bias", L + "norm2.weight", L + "norm2.bias"]
    keep = [p[n].detach().float().contiguous() for n in names]
    dm, dff = keep[0].shape[0], keep[7].shape[0]
    w = abi.PoolWeights(*[abi.ptr(t) for t in keep])
    x = slots.detach().float().contiguous()
    out = torch.empty(B, dm, device=slots.device, dtype=torch.float32)
    abi.check(abi.lib().ocrl_pool_transformer_fwd(abi.ptr(x), ctypes.byref(w), abi.ptr(out), B, S, Din, dm, nhead, dff,
                                                  ln_eps, abi.stream_ptr()), "ocrl_pool_transformer_fwd")
    return out


class Transformer(nn.Module):
    """poolings/common/transformer.py:9-33: Linear -> CLS token -> TransformerEncoder -> CLS row."""

    def __init__(self, in_dim, d_model, nhead, num_layers, pos=None, norm_first=False):
        super().__init__()
        self._linear = nn.Linear(in_dim, d_model)
        self._cls_token = ClsToken(d_model)
        self._pos = pos
        self._nhead, self._num_layers = nhead, num_layers
        layer = nn.TransformerEncoderLayer(d_model, nhead)  # post-norm, ReLU, dim_feedforward 2048 (the reference's defaults)
        self._trans = nn.TransformerEncoder(layer, num_layers, enable_nested_tensor=False)

    def _fused_ok(self, state):
        lay = self._trans.layers[0]
        needs_grad = torch.is_grad_enabled() and (state.requires_grad or any(p.requires_grad for p in self.parameters()))
        return (state.is_cuda and not needs_grad and self._pos is None and self._num_layers == 1 and not lay.norm_first
                and (not self.training or lay.dropout.p == 0.0) and state.shape[1] <= 16 and state.shape[2] % 32 == 0
                and self._linear.out_features % 32 == 0 and self._linear.out_features <= 128
                and lay.linear1.out_features % 128 == 0 and lay.linear1.out_features <= 4096
                and lay.activation_relu_or_gelu == 1)

    def forward(self, state):
        if state.dim() == 3 and self._fused_ok(state):
            p = dict(self.named_parameters())
            return pool_transformer(state, p, self._nhead, self._trans.layers[0].norm1.eps)
        B, S, D = state.shape
        x = self._linear(state.reshape(-1, D)).reshape(B, S, -1)
        x = torch.cat([self._cls_token().repeat(B, 1, 1), x], dim=1).permute(1, 0, 2)
        x = x if self._pos is None else self._pos(x)
        return self._trans(x)[0]


class Transformer_Module(nn.Module):
    """poolings/transformer/transformer_module.py:29-133 for the shipped configuration (configs/pooling/transformer.yaml:
    pos_emb None, no extra MLP / embedding front ends); other configurations raise -- use the reference's module there."""

    def __init__(self, ocr_rep_dim: int, ocr_num_slots: int, config, num_stacked_obss: int = 1) -> None:
        super().__init__()
        self.rep_dim = d_model = config.d_model
        self.config = config
        for flag in ("use_mlp1", "use_mlp2", "cw_embedding", "push_embedding"):
            if getattr(config, flag, False):
                raise NotImplementedError(f"ocrl_b200.Transformer_Module: pooling.{flag} is not covered (reference module only)")
        if num_stacked_obss > 1 or str(getattr(config, "pos_emb", "None")) != "None":
            raise NotImplementedError("ocrl_b200.Transformer_Module: positional encodings are not covered (reference module only)")
        self._trans = Transformer(ocr_rep_dim, d_model, config.nhead, config.num_layers, None, config.norm_first)

    def forward(self, state):
        return self._trans(state)


class RolloutExtractor(nn.Module):
    """``pooling(ocr(obs))`` -- what sb3s/ocr_extractor.py:45 evaluates per environment step -- as one callable whose
    parameters are those of both parts, so that ``GraphedEncoder(RolloutExtractor(ocr, pooling), example_obs)`` captures
    the whole rollout feature path (convolutions, token stage, iteration kernel, pooling kernel) in ONE CUDA graph and
    re-captures it when the PPO update changes the pooling's weights."""

    def __init__(self, ocr, pooling):
        super().__init__()
        self._ocr = ocr  # the OCR wrapper is a plain object (frozen, ocr_extractor.py:29-31); its module is registered below
        self._ocr_module = getattr(ocr, "_module", ocr)
        self._pooling = pooling

    @property
    def _module(self):  # GraphedEncoder watches the parameter versions of this module
        return self

    def forward(self, obs):
        return self._pooling(self._ocr(obs))

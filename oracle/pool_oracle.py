"""CPU restatement of the PPO consumer's slot pooling -- TEST INFRASTRUCTURE (imported by tests/ only).

Follows poolings/common/transformer.py:20-33 (Linear, CLS token, nn.TransformerEncoder -> row 0) with the
nn.TransformerEncoderLayer the reference builds at :15-19 written out: post-norm, ReLU, one head-split scaled
dot-product attention, dropout inactive.  Pinned against the real reference module by tests/golden/pool_transformer.npz
(oracle/make_golden.py::make_pool_case)."""
import torch


def _ln(x, w, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def transformer_pool(slots, p, nhead, eps=1e-5):
    """slots [B,S,Din]; p: state_dict of poolings.common.transformer.Transformer (one layer).  Returns [B, d_model]."""
    L = "_trans.layers.0."
    B, S, _ = slots.shape
    x = slots @ p["_linear.weight"].T + p["_linear.bias"]                       # transformer.py:23-24
    x = torch.cat([p["_cls_token._cls_token"].expand(B, 1, -1), x], dim=1)        # :26-28  [B, S+1, dm]
    dm = x.shape[-1]
    hd = dm // nhead
    qkv = x @ p[L + "self_attn.in_proj_weight"].T + p[L + "self_attn.in_proj_bias"]
    q, k, v = qkv.split(dm, dim=-1)
    sh = lambda t: t.reshape(B, S + 1, nhead, hd).transpose(1, 2)               # [B, h, S+1, hd]
    att = torch.softmax(sh(q) @ sh(k).transpose(-1, -2) / hd ** 0.5, dim=-1) @ sh(v)
    att = att.transpose(1, 2).reshape(B, S + 1, dm)
    y = att @ p[L + "self_attn.out_proj.weight"].T + p[L + "self_attn.out_proj.bias"]
    x = _ln(x + y, p[L + "norm1.weight"], p[L + "norm1.bias"], eps)              # post-norm layer
    f = torch.relu(x @ p[L + "linear1.weight"].T + p[L + "linear1.bias"]) @ p[L + "linear2.weight"].T + p[L + "linear2.bias"]
    x = _ln(x + f, p[L + "norm2.weight"], p[L + "norm2.bias"], eps)
    return x[:, 0]                                                               # :33 (row 0 = CLS)

"""CPU oracle for the slot-attention hot path (TEST INFRASTRUCTURE, not product code).

A plain restatement, in explicit tensor arithmetic, of what the reference computes on the
path ``SLATE.__call__ -> SLATE_Module._get_slots -> SlotAttentionEncoder -> SlotAttention``.
Every function cites the reference lines it follows (paths relative to /root/reference).
It is dtype-generic (fp32 for parity, fp64 for tie analysis) and differentiable, so the same
code is the gradient oracle (torch autograd over these explicit formulas).

Parity pinning: checked against outputs of the imported reference (tests/golden/*.npz, made
by oracle/make_golden.py) and, where /root/reference exists, against the live reference.

Parameters are passed as a flat ``dict[str, Tensor]`` using the reference's state_dict names,
addressed with a prefix (e.g. ``"_slotattn.slot_attention."``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------
def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last axis: biased variance, eps inside the sqrt, affine.
    Used at ocrs/common/slot_attn.py:30-32,125 (torch default eps = 1e-5)."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + eps) * weight + bias


def gru_cell(x: Tensor, h: Tensor, w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor) -> Tensor:
    """nn.GRUCell as built by ocrs/common/networks.py:67-74; gate rows ordered [r; z; n].
    r = s(W_ir x + b_ir + W_hr h + b_hr); z likewise; n = tanh(W_in x + b_in + r*(W_hn h + b_hn));
    h' = (1 - z) * n + z * h."""
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    D = h.shape[-1]
    r = torch.sigmoid(gi[..., :D] + gh[..., :D])
    z = torch.sigmoid(gi[..., D : 2 * D] + gh[..., D : 2 * D])
    n = torch.tanh(gi[..., 2 * D :] + r * gh[..., 2 * D :])
    return (1.0 - z) * n + z * h


# --------------------------------------------------------------------------------------
# SlotAttention (ocrs/common/slot_attn.py:47-102), heads == 1
# --------------------------------------------------------------------------------------
def kv_project(inputs: Tensor, p: Params, pre: str = "") -> Tuple[Tensor, Tensor]:
    """slot_attn.py:54-61: x^ = LN_in(x); k = D^-1/2 * W_k x^; v = W_v x^  (heads = 1)."""
    xh = layer_norm(inputs, p[pre + "norm_inputs.weight"], p[pre + "norm_inputs.bias"])
    wk, wv = p[pre + "project_k.weight"], p[pre + "project_v.weight"]
    D = wk.shape[0]
    k = (xh @ wk.t()) * (D ** -0.5)
    v = xh @ wv.t()
    return k, v


def slot_update(updates: Tensor, slots_prev: Tensor, p: Params, pre: str = "") -> Tensor:
    """slot_attn.py:96-100: GRUCell(updates, slots_prev) then residual MLP over LN_m."""
    B, K, D = slots_prev.shape
    h = gru_cell(
        updates.reshape(-1, D),
        slots_prev.reshape(-1, D),
        p[pre + "gru.weight_ih"],
        p[pre + "gru.weight_hh"],
        p[pre + "gru.bias_ih"],
        p[pre + "gru.bias_hh"],
    ).reshape(B, K, D)
    m = layer_norm(h, p[pre + "norm_mlp.weight"], p[pre + "norm_mlp.bias"])
    m = torch.relu(m @ p[pre + "mlp.0.weight"].t() + p[pre + "mlp.0.bias"])
    m = m @ p[pre + "mlp.2.weight"].t() + p[pre + "mlp.2.bias"]
    return h + m


def iterate(
    k: Tensor,
    v: Tensor,
    slots: Tensor,
    p: Params,
    num_iterations: int,
    epsilon: float = 1e-8,
    pre: str = "",
    return_trace: bool = False,
):
    """The T-iteration loop, slot_attn.py:64-100, on already projected k (pre-scaled), v.

    k, v: [B, N, D]; slots: [B, K, D].  Returns (slots [B,K,D], attn_vis [B,N,K]).
    """
    trace = []
    attn_vis = None
    for _ in range(num_iterations):
        slots_prev = slots
        s = layer_norm(slots, p[pre + "norm_slots.weight"], p[pre + "norm_slots.bias"])  # :66
        q = s @ p[pre + "project_q.weight"].t()  # :69-71
        logits = k @ q.transpose(-1, -2)  # [B,N,K]   :72-74
        attn = torch.softmax(logits, dim=-1)  # over the K slots   :75-82
        attn_vis = attn  # heads == 1 -> sum over heads is identity   :83
        a = attn + epsilon  # :86
        a = a / a.sum(dim=-2, keepdim=True)  # normalise over the N tokens   :87
        updates = a.transpose(-1, -2) @ v  # [B,K,D]   :88-93
        slots = slot_update(updates, slots_prev, p, pre)  # :96-100
        if return_trace:
            trace.append(dict(q=q, logits=logits, updates=updates, slots=slots))
    if return_trace:
        return slots, attn_vis, trace
    return slots, attn_vis


def slot_attention(inputs: Tensor, slots: Tensor, p: Params, num_iterations: int,
                   epsilon: float = 1e-8, pre: str = "") -> Tuple[Tensor, Tensor]:
    """SlotAttention.forward(inputs, slots), slot_attn.py:47-102."""
    k, v = kv_project(inputs, p, pre)
    return iterate(k, v, slots, p, num_iterations, epsilon, pre)


# --------------------------------------------------------------------------------------
# SlotAttentionEncoder (ocrs/common/slot_attn.py:147-161)
# --------------------------------------------------------------------------------------
def token_mlp(x: Tensor, p: Params, pre: str = "") -> Tensor:
    """slot_attn.py:151: x = W2 relu(W1 LN(x) + b1) + b2 (C_in -> C_in -> C_in)."""
    y = layer_norm(x, p[pre + "layer_norm.weight"], p[pre + "layer_norm.bias"])
    y = torch.relu(y @ p[pre + "mlp.0.weight"].t() + p[pre + "mlp.0.bias"])
    return y @ p[pre + "mlp.2.weight"].t() + p[pre + "mlp.2.bias"]


def init_slots(noise: Tensor, p: Params, pre: str = "") -> Tensor:
    """slot_attn.py:155-156: slots = mu + exp(log_sigma) * N(0,1); noise is the N(0,1) draw."""
    return p[pre + "slot_mu"] + torch.exp(p[pre + "slot_log_sigma"]) * noise


def slot_attention_encoder(x: Tensor, noise: Tensor, p: Params, num_iterations: int,
                           pre: str = "") -> Tuple[Tensor, Tensor]:
    """SlotAttentionEncoder.forward with the Gaussian draw made explicit (``noise`` [B,K,D])."""
    y = token_mlp(x, p, pre)
    slots0 = init_slots(noise, p, pre)
    return slot_attention(y, slots0, p, num_iterations, 1e-8, pre + "slot_attention.")


# --------------------------------------------------------------------------------------
# feature stage (ocrs/common/models.py:96-107, ocrs/common/utils.py:10-33)
# --------------------------------------------------------------------------------------
def position_grid(size: int, dtype=torch.float32) -> Tensor:
    """utils.py:13-26: [1,4,S,S] ramps ordered (north, south, west, east);
    east[y,x] = x/(S-1), west = 1-east, south[y,x] = y/(S-1), north = 1-south."""
    ramp = torch.linspace(0, 1, size, dtype=dtype)
    east = ramp.view(1, size).expand(size, size)
    south = ramp.view(size, 1).expand(size, size)
    # torch.linspace(1, 0, S) is not bitwise 1 - linspace(0, 1, S); build it the way the reference does
    rev = torch.linspace(1, 0, size, dtype=dtype)
    west = rev.view(1, size).expand(size, size)
    north = rev.view(size, 1).expand(size, size)
    return torch.stack([north, south, west, east], dim=0).unsqueeze(0).contiguous()


def cnn_encoder(obs: Tensor, p: Params, pre: str = "_enc.") -> Tensor:
    """models.py:99-107: 4x conv5x5 stride 1 pad 2; ReLU after the first three."""
    x = obs
    for i in range(3):
        x = F.conv2d(x, p[f"{pre}_encoder.{i}.m.weight"], p[f"{pre}_encoder.{i}.m.bias"], padding=2)
        x = torch.relu(x)
    return F.conv2d(x, p[f"{pre}_encoder.3.weight"], p[f"{pre}_encoder.3.bias"], padding=2)


def position_table(p: Params, pre: str = "_enc_pos.") -> Tensor:
    """utils.py:28-33: the 1x1 conv over the constant ramp grid, i.e. a [C,S,S] table."""
    grid = p[pre + "linear_position_embedding"]  # [1,4,S,S]
    w = p[pre + "channels_map.weight"].flatten(1)  # [C,4]
    b = p[pre + "channels_map.bias"]
    return torch.einsum("cj,jhw->chw", w, grid[0]) + b.view(-1, 1, 1)


def feature_tokens(obs: Tensor, p: Params) -> Tensor:
    """slate_module.py:132-133: emb = enc_pos(enc(obs)); NCHW -> [B, H*W, C]."""
    f = cnn_encoder(obs, p) + position_table(p).unsqueeze(0)
    return f.permute(0, 2, 3, 1).flatten(1, 2)


def slate_encode(obs: Tensor, noise: Tensor, p: Params, num_iterations: int) -> Tuple[Tensor, Tensor]:
    """SLATE.__call__(obs) hot path (slate.py:36-37 -> slate_module.py:130-136, 181-196).
    Returns (slots [B,K,D], attn [B,N,K])."""
    emb = feature_tokens(obs, p)
    return slot_attention_encoder(emb, noise, p, num_iterations, "_slotattn.")


def masks_from_attn(attn: Tensor, obs: Tensor, with_attns: bool) -> Tensor:
    """slate_module.py:189-194: [B,N,K] -> [B,K,1,H,W]; with_attns blends with the frame."""
    B, _, H, W = obs.shape
    a = attn.transpose(-1, -2).reshape(B, attn.shape[-1], 1, H, W)
    if with_attns:
        a = obs.unsqueeze(1) * a + (1.0 - a)
    return a


# --------------------------------------------------------------------------------------
# helpers for tests
# --------------------------------------------------------------------------------------
def tie_mask(attn64: Tensor, margin: float) -> Tensor:
    """Tokens whose top-1/top-2 attention gap (computed in fp64) is below ``margin``: the
    argmax of such tokens is not decidable in fp32 and is excluded from bit-exact checks."""
    if attn64.shape[-1] < 2:
        return torch.zeros(attn64.shape[:-1], dtype=torch.bool, device=attn64.device)
    top2 = attn64.topk(2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]) < margin


def to_dtype(p: Params, dtype) -> Params:
    return {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in p.items()}


def random_sa_params(num_slots: int, input_size: int, slot_size: int, mlp_hidden: int,
                     seed: int, nontrivial: bool = True) -> Params:
    """Seeded parameters with the reference's init families (networks.py:56-74) but, when
    ``nontrivial``, non-zero biases and non-unit LayerNorm affines so that every term of the
    path is exercised by parity tests."""
    g = torch.Generator().manual_seed(seed)
    D, C, H = slot_size, input_size, mlp_hidden

    def xavier(o, i):
        a = math.sqrt(6.0 / (i + o))
        return (torch.rand(o, i, generator=g) * 2 - 1) * a

    def kaiming(o, i):
        a = math.sqrt(6.0 / i)
        return (torch.rand(o, i, generator=g) * 2 - 1) * a

    def small(n):
        return 0.1 * torch.randn(n, generator=g) if nontrivial else torch.zeros(n)

    def gamma(n):
        return 1.0 + (0.1 * torch.randn(n, generator=g) if nontrivial else torch.zeros(n))

    q, _ = torch.linalg.qr(torch.randn(3 * D, D, generator=g))
    p = {
        "norm_inputs.weight": gamma(C), "norm_inputs.bias": small(C),
        "norm_slots.weight": gamma(D), "norm_slots.bias": small(D),
        "norm_mlp.weight": gamma(D), "norm_mlp.bias": small(D),
        "project_q.weight": xavier(D, D),
        "project_k.weight": xavier(D, C),
        "project_v.weight": xavier(D, C),
        "gru.weight_ih": xavier(3 * D, D),
        "gru.weight_hh": q.contiguous(),
        "gru.bias_ih": small(3 * D), "gru.bias_hh": small(3 * D),
        "mlp.0.weight": kaiming(H, D), "mlp.0.bias": small(H),
        "mlp.2.weight": xavier(D, H), "mlp.2.bias": small(D),
    }
    return p

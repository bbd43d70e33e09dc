"""Drop-in ``SlotAttention`` / ``SlotAttentionEncoder`` (reference: ocrs/common/slot_attn.py:9-161).

Same constructor signatures, attribute names, parameter names/shapes (so reference checkpoints
load strictly) and return values; the arithmetic runs in the sm_100a kernels of libocrl_sa.so.
CPU tensors raise -- there is no fallback path.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import functional as F
from .networks import gru_cell, linear

_ENC_KEYS = ("layer_norm.weight", "layer_norm.bias", "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias")


def default_kv_dtype() -> str:
    """'fp32' (parity mode, default) or 'bf16' (k/v stored in bf16, fp32 accumulate)."""
    kv = os.environ.get("OCRL_KV_DTYPE", "fp32")
    if kv not in ("fp32", "bf16"):
        raise ValueError(f"OCRL_KV_DTYPE must be fp32 or bf16, got {kv}")
    return kv


class SlotAttention(nn.Module):
    def __init__(self, num_iterations, num_slots, input_size, slot_size, mlp_hidden_size, heads, epsilon=1e-8,
                 kv_dtype=None, launch_opts=None):
        """Reference signature (slot_attn.py:10-19) plus two keyword-only extras of this implementation:
        ``kv_dtype`` 'fp32' (exact-parity mode) | 'bf16' (bf16 k/v, tensor cores); None reads OCRL_KV_DTYPE once, here;
        ``launch_opts`` an ``abi.launch_opts(...)`` applied to every forward of this module (None: library default)."""
        super().__init__()
        self.num_iterations = num_iterations
        self.num_slots = num_slots
        self.input_size = input_size
        self.slot_size = slot_size
        self.mlp_hidden_size = mlp_hidden_size
        self.epsilon = epsilon
        self.num_heads = heads
        self.kv_dtype = default_kv_dtype() if kv_dtype is None else kv_dtype
        if self.kv_dtype not in ("fp32", "bf16"):
            raise ValueError(f"kv_dtype must be fp32 or bf16, got {self.kv_dtype}")
        self.launch_opts = launch_opts

        self.norm_inputs = nn.LayerNorm(input_size)
        self.norm_slots = nn.LayerNorm(slot_size)
        self.norm_mlp = nn.LayerNorm(slot_size)
        self.project_q = linear(slot_size, slot_size, bias=False)
        self.project_k = linear(input_size, slot_size, bias=False)
        self.project_v = linear(input_size, slot_size, bias=False)
        self.gru = gru_cell(slot_size, slot_size)
        self.mlp = nn.Sequential(
            linear(slot_size, mlp_hidden_size, weight_init="kaiming"),
            nn.ReLU(),
            linear(mlp_hidden_size, slot_size),
        )

    # -- parameter views in the order the C ABI wants them -----------------------------------------
    def _params(self):
        return {
            "norm_inputs.weight": self.norm_inputs.weight, "norm_inputs.bias": self.norm_inputs.bias,
            "norm_slots.weight": self.norm_slots.weight, "norm_slots.bias": self.norm_slots.bias,
            "norm_mlp.weight": self.norm_mlp.weight, "norm_mlp.bias": self.norm_mlp.bias,
            "project_q.weight": self.project_q.weight, "project_k.weight": self.project_k.weight,
            "project_v.weight": self.project_v.weight,
            "gru.weight_ih": self.gru.weight_ih, "gru.weight_hh": self.gru.weight_hh,
            "gru.bias_ih": self.gru.bias_ih, "gru.bias_hh": self.gru.bias_hh,
            "mlp.0.weight": self.mlp[0].weight, "mlp.0.bias": self.mlp[0].bias,
            "mlp.2.weight": self.mlp[2].weight, "mlp.2.bias": self.mlp[2].bias,
        }

    def _check(self, inputs, slots, fmap=False):
        if self.num_heads != 1:
            raise NotImplementedError("ocrl_b200.SlotAttention supports num_slot_heads=1 (all shipped configs)")
        if not inputs.is_cuda:
            raise RuntimeError("ocrl_b200.SlotAttention runs on CUDA (sm_100a) only; there is no CPU fallback")
        if inputs.dim() != (4 if fmap else 3) or slots.dim() != 3 or inputs.shape[0] != slots.shape[0]:
            raise ValueError(f"expected inputs [B,N,C] and slots [B,K,D], got {tuple(inputs.shape)}, {tuple(slots.shape)}")

    def forward(self, inputs, slots, *, _enc=None, _pos_table=None):
        """inputs [B,N,C_in], slots [B,K,D] -> (slots [B,K,D], attn_vis [B,N,K]).

        ``_enc`` / ``_pos_table`` are private hooks used by the encoder / SLATE module to fuse the
        token LayerNorm+MLP and the position-table add into the projection kernel (inference only).
        """
        self._check(inputs, slots, fmap=inputs.dim() == 4)
        p = self._params()
        needs_grad = torch.is_grad_enabled() and (
            inputs.requires_grad or slots.requires_grad or any(t.requires_grad for t in p.values()))
        if not needs_grad:
            with torch.no_grad():
                prep = self.__dict__.get("_prepared")
                if prep is None:
                    prep = self.__dict__["_prepared"] = F.PreparedWeights()
                return F.slot_attention(inputs, slots, p, self.num_iterations, epsilon=self.epsilon,
                                        kv=self.kv_dtype, enc=_enc, pos_table=_pos_table, opts=self.launch_opts,
                                        prepared=prep)
        assert _enc is None and _pos_table is None, "fused token stage is an inference-only path"
        return F.SlotAttentionFunction.apply(inputs, slots, self.num_iterations, self.epsilon, self.kv_dtype,
                                             *[p[n] for n in F.SA_PARAM_ORDER])


class SlotAttentionEncoder(nn.Module):
    def __init__(self, num_iterations, num_slots, input_channels, slot_size, mlp_hidden_size, pos_channels,
                 num_heads, kv_dtype=None):
        super().__init__()
        self.num_iterations = num_iterations
        self.num_slots = num_slots
        self.input_channels = input_channels
        self.slot_size = slot_size
        self.mlp_hidden_size = mlp_hidden_size
        self.pos_channels = pos_channels

        self.layer_norm = nn.LayerNorm(input_channels)
        self.mlp = nn.Sequential(
            linear(input_channels, input_channels, weight_init="kaiming"),
            nn.ReLU(),
            linear(input_channels, input_channels),
        )
        self.slot_mu = nn.Parameter(torch.zeros(1, 1, slot_size))
        self.slot_log_sigma = nn.Parameter(torch.zeros(1, 1, slot_size))
        nn.init.xavier_uniform_(self.slot_mu)
        nn.init.xavier_uniform_(self.slot_log_sigma)
        self.slot_attention = SlotAttention(num_iterations, num_slots, input_channels, slot_size, mlp_hidden_size,
                                            num_heads, kv_dtype=kv_dtype)

    def _enc_params(self):
        return {"layer_norm.weight": self.layer_norm.weight, "layer_norm.bias": self.layer_norm.bias,
                "mlp.0.weight": self.mlp[0].weight, "mlp.0.bias": self.mlp[0].bias,
                "mlp.2.weight": self.mlp[2].weight, "mlp.2.bias": self.mlp[2].bias}

    def init_slots(self, batch, like):
        # same draw as the reference (slot_attn.py:155): torch's generator on the tensor's device
        # (always in the parameters' dtype: a bf16 feature map must not turn the draw into a bf16 one)
        noise = torch.empty(batch, self.num_slots, self.slot_size, device=like.device, dtype=self.slot_mu.dtype).normal_()
        if torch.is_grad_enabled() and (self.slot_mu.requires_grad or self.slot_log_sigma.requires_grad):
            return self.slot_mu + torch.exp(self.slot_log_sigma) * noise
        # inference: sigma cached until the parameter changes, one fused multiply-add (same values)
        key = (self.slot_log_sigma.data_ptr(), self.slot_log_sigma._version)
        cached = self.__dict__.get("_sigma_cache")
        if cached is None or cached[0] != key:
            cached = (key, torch.exp(self.slot_log_sigma.detach()))
            self.__dict__["_sigma_cache"] = cached
        return torch.addcmul(self.slot_mu.detach(), cached[1], noise)

    def forward(self, x, *, _pos_table=None):
        """x [B,N,C] (or the NCHW feature map when ``_pos_table`` is given) -> (slots, attn)."""
        B = x.shape[0]
        ep = self._enc_params()
        needs_grad = torch.is_grad_enabled() and (
            x.requires_grad or any(t.requires_grad for t in self.parameters()))
        if not needs_grad:
            slots = self.init_slots(B, x)
            return self.slot_attention(x, slots, _enc=ep, _pos_table=_pos_table)
        assert _pos_table is None
        x = self.mlp(self.layer_norm(x))  # token MLP through autograd (library GEMMs) in training
        slots = self.init_slots(B, x)
        return self.slot_attention(x, slots)

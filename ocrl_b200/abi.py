"""ctypes binding of libocrl_sa.so (include/ocrl_sa.h).  No torch types cross this boundary:
only device pointers, sizes and the raw cudaStream_t.

There is no CPU or library fallback: if the shared library is missing, ``lib()`` raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_size_t, c_void_p

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libocrl_sa.so")

DT_F32, DT_BF16 = 0, 1
MATH_FP32, MATH_TENSOR = 0, 1
X_TOKENS_F32, X_NCHW_F32, X_TOKENS_BF16, X_PADDED_BF16 = 0, 1, 2, 3

EXPORTS = [
    "ocrl_version", "ocrl_built_arch", "ocrl_last_error", "ocrl_launch_count", "ocrl_sa_query_workspace",
    "ocrl_kv_proj_fwd_workspace", "ocrl_kv_proj_fwd", "ocrl_kv_proj_bwd_workspace", "ocrl_kv_proj_bwd",
    "ocrl_kv_proj_bwd_lowrank_workspace", "ocrl_kv_proj_bwd_lowrank",
    "ocrl_sa_iter_fwd", "ocrl_sa_iter_fwd_ex", "ocrl_sa_last_kernel", "ocrl_sa_iter_bwd",
    "ocrl_xhat_fwd", "ocrl_sa_iter_fwd_xhat",
    "ocrl_conv_bias_relu_bf16", "ocrl_frames_to_nhwc_bf16", "ocrl_conv_first_relu_bf16", "ocrl_conv_first_relu_u8p",
    "ocrl_conv_padded_bytes", "ocrl_conv5x5_pack_weights", "ocrl_conv5x5_c64_tc", "ocrl_conv_first_relu_bf16p",
    "ocrl_pool_transformer_fwd",
]

# ocrl_sa_launch_opts.variant
SA_AUTO, SA_TCGEN05, SA_PIPE, SA_CLUSTER_TC, SA_FFMA = 0, 1, 2, 3, 4
SA_VARIANTS = {"auto": SA_AUTO, "tcgen05": SA_TCGEN05, "pipe": SA_PIPE, "cluster_tc": SA_CLUSTER_TC, "ffma": SA_FFMA}


class SaDims(Structure):
    _fields_ = [("B", c_int32), ("N", c_int32), ("C_in", c_int32), ("D", c_int32), ("H_mlp", c_int32),
                ("K", c_int32), ("T", c_int32), ("heads", c_int32), ("eps", c_float), ("ln_eps", c_float),
                ("kv_dtype", c_int32), ("math_mode", c_int32), ("x_format", c_int32), ("frame_w", c_int32)]


_SA_W = ["ln_slots_w", "ln_slots_b", "ln_mlp_w", "ln_mlp_b", "wq", "w_ih", "w_hh", "b_ih", "b_hh",
         "w1", "b1", "w2", "b2"]
_TOK_W = ["enc_ln_w", "enc_ln_b", "mlp_w1", "mlp_b1", "mlp_w2", "mlp_b2", "in_ln_w", "in_ln_b", "wk", "wv"]


class SaWeights(Structure):
    _fields_ = [(n, c_void_p) for n in _SA_W]


class LaunchOpts(Structure):
    """ocrl_sa_launch_opts (include/ocrl_sa.h): per-call kernel selection of the iteration loop."""
    _fields_ = [("variant", c_int32), ("max_clusters", c_int32), ("lanes", c_int32), ("strict", c_int32),
                ("trace", c_int32), ("prepared", c_int32)]


def launch_opts(variant="auto", max_clusters=0, lanes=0, strict=False, trace=False, prepared=False) -> LaunchOpts:
    if isinstance(variant, str):
        variant = SA_VARIANTS[variant]
    return LaunchOpts(int(variant), int(max_clusters or 0), int(lanes or 0), int(bool(strict)), int(bool(trace)),
                      int(bool(prepared)))


_POOL_W = ["lin_w", "lin_b", "cls", "in_proj_w", "in_proj_b", "out_proj_w", "out_proj_b", "lin1_w", "lin1_b",
           "lin2_w", "lin2_b", "norm1_w", "norm1_b", "norm2_w", "norm2_b"]


class PoolWeights(Structure):
    """ocrl_pool_weights (include/ocrl_sa.h)."""
    _fields_ = [(n, c_void_p) for n in _POOL_W]


class SaWeightGrads(Structure):
    _fields_ = [(n, c_void_p) for n in _SA_W]


class TokenWeights(Structure):
    _fields_ = [(n, c_void_p) for n in _TOK_W]


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m ocrl_b200.build` "
                "(there is no CPU or PyTorch fallback for the slot-attention path)")
        L = ctypes.CDLL(LIB_PATH)
        L.ocrl_version.restype = c_int
        L.ocrl_built_arch.restype = c_char_p
        L.ocrl_last_error.restype = c_char_p
        L.ocrl_launch_count.restype = ctypes.c_ulonglong
        L.ocrl_sa_query_workspace.argtypes = [POINTER(SaDims), POINTER(c_size_t), POINTER(c_size_t), POINTER(c_size_t)]
        L.ocrl_kv_proj_fwd.argtypes = [POINTER(SaDims), c_void_p, c_void_p, POINTER(TokenWeights), c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p]
        L.ocrl_kv_proj_fwd_workspace.argtypes = [POINTER(SaDims)]
        L.ocrl_kv_proj_fwd_workspace.restype = c_size_t
        L.ocrl_kv_proj_bwd_workspace.argtypes = [POINTER(SaDims)]
        L.ocrl_kv_proj_bwd_workspace.restype = c_size_t
        L.ocrl_kv_proj_bwd.argtypes = [POINTER(SaDims), c_void_p, POINTER(TokenWeights), c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        L.ocrl_kv_proj_bwd_lowrank_workspace.argtypes = [POINTER(SaDims)]
        L.ocrl_kv_proj_bwd_lowrank_workspace.restype = c_size_t
        L.ocrl_kv_proj_bwd_lowrank.argtypes = [POINTER(SaDims), c_void_p, POINTER(TokenWeights), c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
        L.ocrl_kv_proj_bwd_lowrank.restype = c_int
        L.ocrl_sa_iter_fwd.argtypes = [POINTER(SaDims), c_void_p, c_void_p, c_void_p, POINTER(SaWeights), c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p]
        L.ocrl_sa_iter_fwd_ex.argtypes = [POINTER(SaDims), c_void_p, c_void_p, c_void_p, POINTER(SaWeights), c_void_p,
                                          c_void_p, c_void_p, c_void_p, POINTER(LaunchOpts), c_void_p]
        L.ocrl_sa_last_kernel.restype = c_char_p
        L.ocrl_xhat_fwd.argtypes = [POINTER(SaDims), c_void_p, c_void_p, POINTER(TokenWeights), c_void_p, c_void_p,
                                    c_void_p, c_void_p]
        L.ocrl_xhat_fwd.restype = c_int
        L.ocrl_sa_iter_fwd_xhat.argtypes = [POINTER(SaDims), c_void_p, c_void_p, c_void_p, c_void_p, POINTER(SaWeights),
                                            c_void_p, c_void_p, c_void_p, POINTER(LaunchOpts), c_void_p]
        L.ocrl_sa_iter_fwd_xhat.restype = c_int
        L.ocrl_sa_iter_bwd.argtypes = [POINTER(SaDims), c_void_p, c_void_p, c_void_p, POINTER(SaWeights), c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, POINTER(SaWeightGrads), c_void_p,
                                       c_void_p]
        L.ocrl_conv_bias_relu_bf16.argtypes = [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p]
        L.ocrl_frames_to_nhwc_bf16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
        L.ocrl_conv_first_relu_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                                c_int, c_void_p]
        L.ocrl_conv_first_relu_bf16.restype = c_int
        L.ocrl_conv_first_relu_bf16p.argtypes = L.ocrl_conv_first_relu_bf16.argtypes
        L.ocrl_conv_first_relu_bf16p.restype = c_int
        L.ocrl_conv_first_relu_u8p.argtypes = L.ocrl_conv_first_relu_bf16.argtypes
        L.ocrl_conv_first_relu_u8p.restype = c_int
        L.ocrl_conv_padded_bytes.argtypes = [c_int, c_int, c_int]
        L.ocrl_conv_padded_bytes.restype = c_size_t
        L.ocrl_conv5x5_pack_weights.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p]
        L.ocrl_conv5x5_pack_weights.restype = c_int
        L.ocrl_conv5x5_c64_tc.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
        L.ocrl_conv5x5_c64_tc.restype = c_int
        L.ocrl_pool_transformer_fwd.argtypes = [c_void_p, POINTER(PoolWeights), c_void_p, c_int, c_int, c_int, c_int, c_int,
                                                c_int, c_float, c_void_p]
        L.ocrl_pool_transformer_fwd.restype = c_int
        for name in ("ocrl_sa_query_workspace", "ocrl_kv_proj_fwd", "ocrl_kv_proj_bwd", "ocrl_sa_iter_fwd",
                     "ocrl_sa_iter_fwd_ex", "ocrl_sa_iter_bwd", "ocrl_conv_bias_relu_bf16", "ocrl_frames_to_nhwc_bf16"):
            getattr(L, name).restype = c_int
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().ocrl_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t) -> c_void_p:
    if t is None:
        return c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "libocrl_sa takes contiguous CUDA tensors"
    return c_void_p(t.data_ptr())


def stream_ptr() -> c_void_p:
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def make_dims(B, N, C_in, D, H_mlp, K, T, heads=1, eps=1e-8, ln_eps=1e-5, kv_dtype=DT_F32, math_mode=MATH_FP32,
              x_format=X_TOKENS_F32, frame_w=0):
    return SaDims(B, N, C_in, D, H_mlp, K, T, heads, eps, ln_eps, kv_dtype, math_mode, x_format, frame_w)


def query_workspace(dims: SaDims):
    f, b, s = c_size_t(0), c_size_t(0), c_size_t(0)
    check(lib().ocrl_sa_query_workspace(ctypes.byref(dims), ctypes.byref(f), ctypes.byref(b), ctypes.byref(s)),
          "ocrl_sa_query_workspace")
    return f.value, b.value, s.value


def sa_weights(**tensors) -> SaWeights:
    return SaWeights(*[ptr(tensors[n]) for n in _SA_W])


def sa_weight_grads(**tensors) -> SaWeightGrads:
    return SaWeightGrads(*[ptr(tensors[n]) for n in _SA_W])


def token_weights(**tensors) -> TokenWeights:
    return TokenWeights(*[ptr(tensors.get(n)) for n in _TOK_W])

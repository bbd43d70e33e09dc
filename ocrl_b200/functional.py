"""Host-side operators over libocrl_sa.so: plain calls for inference and a
``torch.autograd.Function`` for training.  Every arithmetic step of the slot-attention path runs
in the hand-written sm_100a kernels behind the C ABI; torch is used for device memory and streams.

Reference being replaced: ``SlotAttention.forward`` (ocrs/common/slot_attn.py:47-102) and the
token LayerNorm+MLP of ``SlotAttentionEncoder.forward`` (slot_attn.py:151).
"""
from __future__ import annotations

import contextlib
import contextvars
import ctypes
import os
from typing import Dict, Optional, Tuple

import torch

from . import abi

Tensor = torch.Tensor

SA_PARAM_ORDER = [
    "norm_inputs.weight", "norm_inputs.bias", "norm_slots.weight", "norm_slots.bias",
    "norm_mlp.weight", "norm_mlp.bias", "project_q.weight", "project_k.weight", "project_v.weight",
    "gru.weight_ih", "gru.weight_hh", "gru.bias_ih", "gru.bias_hh",
    "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias",
]

_KV_DTYPES = {"fp32": (abi.DT_F32, torch.float32), "bf16": (abi.DT_BF16, torch.bfloat16)}

# Inference in the bf16 / tensor-core mode streams the normalised tokens x^ instead of k, v wherever the factored kernels
# cover the shape (ocrl_xhat_fwd + ocrl_sa_iter_fwd_xhat, include/ocrl_sa.h); False keeps the k/v form (A/B, cross-checks)
FACTORED = True

# The projection backward from the low-rank coefficients (ocrl_kv_proj_bwd_lowrank); False: dk, dv + ocrl_kv_proj_bwd
LOWRANK_BWD = True

# When set to a list, every C-ABI launch appends (name, start_event, end_event) recorded on the
# launching stream; bench.py uses it to time the kernels live inside the step.
KERNEL_EVENTS = None


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if KERNEL_EVENTS is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if KERNEL_EVENTS is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            KERNEL_EVENTS.append((self.name, self.e0, e1))
        return False


# Launch options of the iteration loop for the calls made inside a ``launch_options(...)`` block (context-local, so
# threads and asyncio tasks do not see each other's setting).  An explicit ``opts=`` argument wins.
_LAUNCH_OPTS: contextvars.ContextVar = contextvars.ContextVar("ocrl_sa_launch_opts", default=None)


@contextlib.contextmanager
def launch_options(**kw):
    """``with launch_options(variant="pipe", max_clusters=8): ...`` -- see ``abi.launch_opts`` / ocrl_sa_launch_opts."""
    tok = _LAUNCH_OPTS.set(abi.launch_opts(**kw))
    try:
        yield
    finally:
        _LAUNCH_OPTS.reset(tok)


def last_kernel() -> str:
    """Which kernel the last ``iterate`` of this thread launched: tcgen05 | pipe | cluster_tc | ffma."""
    return abi.lib().ocrl_sa_last_kernel().decode()


def _math_mode(dt_code: int) -> int:
    """bf16 k/v run the token contractions on the tensor cores unless OCRL_SA_MATH=fp32 forces FFMA."""
    if dt_code == abi.DT_BF16 and os.environ.get("OCRL_SA_MATH", "tensor") != "fp32":
        return abi.MATH_TENSOR
    return abi.MATH_FP32


def _f32c(t: Tensor) -> Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _require_cuda(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: ocrl_b200 runs on CUDA (sm_100a) only; there is no CPU path")


def _sa_weights(p: Dict[str, Tensor]) -> abi.SaWeights:
    return abi.sa_weights(
        ln_slots_w=p["norm_slots.weight"], ln_slots_b=p["norm_slots.bias"],
        ln_mlp_w=p["norm_mlp.weight"], ln_mlp_b=p["norm_mlp.bias"], wq=p["project_q.weight"],
        w_ih=p["gru.weight_ih"], w_hh=p["gru.weight_hh"], b_ih=p["gru.bias_ih"], b_hh=p["gru.bias_hh"],
        w1=p["mlp.0.weight"], b1=p["mlp.0.bias"], w2=p["mlp.2.weight"], b2=p["mlp.2.bias"])


def kv_project(x: Tensor, p: Dict[str, Tensor], *, kv: str = "fp32", enc: Optional[Dict[str, Tensor]] = None,
               pos_table: Optional[Tensor] = None, want_y: bool = False, ln_eps: float = 1e-5, xhat_only: bool = False):
    """Token stage: [pos add + NCHW->tokens] -> [LN+MLP of the encoder] -> norm_inputs -> k, v.
    ``xhat_only`` (bf16 / tensor-core mode): stop after norm_inputs and return (x^ [B,N,C] bf16, None, y) -- the input
    of the factored loop (``iterate_xhat``).

    x: [B,N,C] tokens, or a CNN feature map [B,C,H,W] (NCHW-contiguous fp32, or channels-last fp32 / bf16,
    whose memory already is token-major); ``pos_table`` ([C,H*W]) is added to every image's tokens.
    enc: optional {"layer_norm.weight","layer_norm.bias","mlp.0.weight",...} of SlotAttentionEncoder.
    Returns (k, v, y) with y = token-MLP output [B,N,C] (None unless want_y).
    """
    _require_cuda(x, "kv_project")
    x_format = abi.X_TOKENS_F32
    frame_w = 0
    if hasattr(x, "to_nchw"):  # feature_stage.PaddedMap: padded channels-last bf16 map of the hand-written convolutions
        if not (kv == "bf16" and _math_mode(abi.DT_BF16) == abi.MATH_TENSOR):
            raise RuntimeError("kv_project: a padded bf16 feature map is an input of the bf16 / tensor-core mode only")
        B, C, N, frame_w = x.B, x.data.shape[1], x.H * x.W, x.W
        x = x.data
        x_format = abi.X_PADDED_BF16
    elif x.dim() == 4:  # CNN feature map [B,C,H,W]
        B, C = x.shape[0], x.shape[1]
        N = x.shape[2] * x.shape[3]
        if not x.is_contiguous() and x.permute(0, 2, 3, 1).is_contiguous():
            x = x.permute(0, 2, 3, 1).reshape(B, N, C)  # channels-last memory == token-major [B,N,C], no copy
            if x.dtype == torch.bfloat16 and kv == "bf16" and _math_mode(abi.DT_BF16) == abi.MATH_TENSOR:
                x_format = abi.X_TOKENS_BF16
            else:
                x = x.float()
        else:
            x = _f32c(x)
            x_format = abi.X_NCHW_F32
    else:
        x = _f32c(x)
        B, N, C = x.shape
    if pos_table is not None:
        pos_table = _f32c(pos_table).reshape(C, N)
    D = p["project_k.weight"].shape[0]
    dt_code, dt = _KV_DTYPES[kv]
    dims = abi.make_dims(B, N, C, D, D, 1, 1, kv_dtype=dt_code, ln_eps=ln_eps, math_mode=_math_mode(dt_code),
                         x_format=x_format, frame_w=frame_w)
    ws = None
    if dims.math_mode == abi.MATH_TENSOR:  # bf16 weight copies for the tcgen05 kernel
        ws = torch.empty(abi.lib().ocrl_kv_proj_fwd_workspace(ctypes.byref(dims)), device=x.device, dtype=torch.uint8)
    if xhat_only and dims.math_mode != abi.MATH_TENSOR:
        raise RuntimeError("kv_project(xhat_only): the factored form belongs to the bf16 / tensor-core mode")
    k = torch.empty((B, N, C) if xhat_only else (B, N, D), device=x.device, dtype=dt)
    v = None if xhat_only else torch.empty(B, N, D, device=x.device, dtype=dt)
    y = torch.empty(B, N, C, device=x.device, dtype=torch.float32) if want_y else None
    keep = {n: _f32c(p[n]) for n in ("norm_inputs.weight", "norm_inputs.bias", "project_k.weight", "project_v.weight")}
    tw = dict(in_ln_w=keep["norm_inputs.weight"], in_ln_b=keep["norm_inputs.bias"],
              wk=keep["project_k.weight"], wv=keep["project_v.weight"])
    if enc is not None:
        e = {n: _f32c(t) for n, t in enc.items()}
        keep.update({"e." + n: t for n, t in e.items()})
        tw.update(enc_ln_w=e["layer_norm.weight"], enc_ln_b=e["layer_norm.bias"], mlp_w1=e["mlp.0.weight"],
                  mlp_b1=e["mlp.0.bias"], mlp_w2=e["mlp.2.weight"], mlp_b2=e["mlp.2.bias"])
    w = abi.token_weights(**tw)
    if xhat_only:
        with _timed("xhat_fwd"):
            abi.check(abi.lib().ocrl_xhat_fwd(ctypes.byref(dims), abi.ptr(x), abi.ptr(pos_table), ctypes.byref(w),
                                              abi.ptr(y), abi.ptr(k), abi.ptr(ws), abi.stream_ptr()), "ocrl_xhat_fwd")
        return k, None, y
    with _timed("kv_proj_fwd"):
        abi.check(abi.lib().ocrl_kv_proj_fwd(ctypes.byref(dims), abi.ptr(x), abi.ptr(pos_table), ctypes.byref(w),
                                             abi.ptr(y), abi.ptr(k), abi.ptr(v), abi.ptr(ws), abi.stream_ptr()),
                  "ocrl_kv_proj_fwd")
    return k, v, y


class PreparedWeights:
    """Workspace of the iteration loop that keeps the kernel's prepared weight copies (bf16, LayerNorm-folded) between
    calls: the preparation launch runs again only when a parameter (``data_ptr`` / ``_version``), the shape or the
    launch options change (ocrl_sa_launch_opts.prepared).  One per module; inference only."""

    def __init__(self):
        self._params = None
        self._ws = {}

    def lookup(self, p, shape_key, nbytes, device):
        params = tuple((t.data_ptr(), t._version) for t in p.values())
        if params != self._params:  # a parameter changed: every prepared copy is stale
            self._params = params
            self._ws = {}
        key = (shape_key, str(device))
        ws = self._ws.get(key)
        if ws is not None:
            return ws, True
        # one workspace per shape / option set, kept while the parameters stand: CUDA graphs captured for one batch
        # size keep reading theirs while another batch size (a rollout batch next to a training batch) prepares its own
        if len(self._ws) >= 32:
            self._ws.pop(next(iter(self._ws)))
        ws = self._ws[key] = torch.empty(nbytes, device=device, dtype=torch.uint8)
        return ws, False


def iterate(k: Tensor, v: Tensor, slots0: Tensor, p: Dict[str, Tensor], num_iterations: int, *,
            epsilon: float = 1e-8, ln_eps: float = 1e-5, want_attn: bool = True, save: bool = False,
            _workspace: Optional[Tensor] = None, opts: Optional[abi.LaunchOpts] = None,
            prepared: Optional[PreparedWeights] = None):
    """The fused T-iteration loop.  k, v: [B,N,D] fp32 or bf16; slots0 [B,K,D].
    ``opts``: ``abi.launch_opts(...)`` (kernel variant, cluster cap, lanes, strict); default: the enclosing
    ``launch_options`` block, else the library's own choice.
    ``prepared``: a ``PreparedWeights`` that owns the workspace across calls (frozen-weight inference).
    Returns (slots, attn_vis or None, saved or None)."""
    _require_cuda(k, "iterate")
    B, N, D = k.shape
    K = slots0.shape[1]
    H = p["mlp.0.weight"].shape[0]
    dt_code = abi.DT_BF16 if k.dtype == torch.bfloat16 else abi.DT_F32
    dims = abi.make_dims(B, N, 64, D, H, K, num_iterations, eps=epsilon, ln_eps=ln_eps, kv_dtype=dt_code,
                         math_mode=_math_mode(dt_code))
    pw = {n: _f32c(p[n]) for n in SA_PARAM_ORDER if n in p}
    slots0 = _f32c(slots0)
    slots = torch.empty(B, K, D, device=k.device, dtype=torch.float32)
    attn = torch.empty(B, N, K, device=k.device, dtype=torch.float32) if want_attn else None
    saved = None
    if save:
        _, _, saved_bytes = abi.query_workspace(dims)
        saved = torch.empty(saved_bytes // 4, device=k.device, dtype=torch.float32)
    w = _sa_weights(pw)
    if opts is None:
        opts = _LAUNCH_OPTS.get()
    if _workspace is None and dims.math_mode == abi.MATH_TENSOR:  # bf16 weight copies for the tensor-core slot update
        fwd_ws, _, _ = abi.query_workspace(dims)
        if prepared is not None and not save:
            o = opts if opts is not None else abi.launch_opts()
            shape_key = (B, N, K, num_iterations, D, H, dt_code, o.variant, o.max_clusters, o.lanes, o.strict)
            _workspace, ready = prepared.lookup(pw, shape_key, fwd_ws, k.device)
            opts = abi.LaunchOpts(o.variant, o.max_clusters, o.lanes, o.strict, 0, int(ready))
        else:
            _workspace = torch.empty(fwd_ws, device=k.device, dtype=torch.uint8)
    with _timed("sa_iter_fwd"):
        abi.check(abi.lib().ocrl_sa_iter_fwd_ex(ctypes.byref(dims), abi.ptr(k), abi.ptr(v), abi.ptr(slots0),
                                                ctypes.byref(w), abi.ptr(slots), abi.ptr(attn), abi.ptr(saved),
                                                abi.ptr(_workspace), ctypes.byref(opts) if opts is not None else None,
                                                abi.stream_ptr()), "ocrl_sa_iter_fwd")
    return slots, attn, saved


def factored_covers(C: int, D: int, H: int, K: int, kv: str, opts: Optional[abi.LaunchOpts] = None) -> bool:
    """Shapes of the factored inference kernels (ocrl_sa_iter_fwd_xhat)."""
    if opts is not None and opts.variant not in (abi.SA_AUTO, abi.SA_TCGEN05):
        return False
    return (kv == "bf16" and _math_mode(abi.DT_BF16) == abi.MATH_TENSOR and C == 64 and D == 192 and H == 192
            and K <= 16)


def iterate_xhat(xhat: Tensor, slots0: Tensor, p: Dict[str, Tensor], num_iterations: int, *, epsilon: float = 1e-8,
                 ln_eps: float = 1e-5, want_attn: bool = True, opts: Optional[abi.LaunchOpts] = None,
                 prepared: Optional[PreparedWeights] = None):
    """The fused T-iteration loop in its factored form: streams x^ = norm_inputs(x) [B,N,C] bf16 instead of k, v and
    applies project_k / project_v through folded update weights (include/ocrl_sa.h, ocrl_sa_iter_fwd_xhat).
    Inference only.  Returns (slots, attn_vis or None)."""
    _require_cuda(xhat, "iterate_xhat")
    B, N, C = xhat.shape
    K, D = slots0.shape[1], slots0.shape[2]
    H = p["mlp.0.weight"].shape[0]
    dims = abi.make_dims(B, N, C, D, H, K, num_iterations, eps=epsilon, ln_eps=ln_eps, kv_dtype=abi.DT_BF16,
                         math_mode=abi.MATH_TENSOR)
    pw = {n: _f32c(p[n]) for n in SA_PARAM_ORDER if n in p}
    slots0 = _f32c(slots0)
    slots = torch.empty(B, K, D, device=xhat.device, dtype=torch.float32)
    attn = torch.empty(B, N, K, device=xhat.device, dtype=torch.float32) if want_attn else None
    if opts is None:
        opts = _LAUNCH_OPTS.get()
    fwd_ws, _, _ = abi.query_workspace(dims)
    o = opts if opts is not None else abi.launch_opts()
    if prepared is not None:
        shape_key = ("xhat", B, N, K, num_iterations, D, H, C, o.max_clusters, o.lanes)
        ws, ready = prepared.lookup(pw, shape_key, fwd_ws, xhat.device)
    else:
        ws, ready = torch.empty(fwd_ws, device=xhat.device, dtype=torch.uint8), False
    opts = abi.LaunchOpts(o.variant, o.max_clusters, o.lanes, o.strict, o.trace, int(ready))
    w = _sa_weights(pw)
    with _timed("sa_iter_fwd"):
        abi.check(abi.lib().ocrl_sa_iter_fwd_xhat(ctypes.byref(dims), abi.ptr(xhat), abi.ptr(pw["project_k.weight"]),
                                                  abi.ptr(pw["project_v.weight"]), abi.ptr(slots0), ctypes.byref(w),
                                                  abi.ptr(slots), abi.ptr(attn), abi.ptr(ws), ctypes.byref(opts),
                                                  abi.stream_ptr()), "ocrl_sa_iter_fwd_xhat")
    return slots, attn


def slot_attention(inputs: Tensor, slots0: Tensor, p: Dict[str, Tensor], num_iterations: int, *,
                   epsilon: float = 1e-8, kv: str = "fp32", enc: Optional[Dict[str, Tensor]] = None,
                   pos_table: Optional[Tensor] = None, want_attn: bool = True,
                   opts: Optional[abi.LaunchOpts] = None,
                   prepared: Optional[PreparedWeights] = None,
                   factored: Optional[bool] = None) -> Tuple[Tensor, Optional[Tensor]]:
    """Inference-only SlotAttention.forward (no autograd graph).  ``factored``: None = the module default ``FACTORED``
    wherever the factored kernels cover the shape; False = the k/v form."""
    if factored is None:
        factored = FACTORED
    if opts is None:
        opts = _LAUNCH_OPTS.get()
    C = p["project_k.weight"].shape[1]
    D, H = p["project_q.weight"].shape[0], p["mlp.0.weight"].shape[0]
    if factored and factored_covers(C, D, H, slots0.shape[1], kv, opts):
        xhat, _, _ = kv_project(inputs, p, kv=kv, enc=enc, pos_table=pos_table, xhat_only=True)
        return iterate_xhat(xhat, slots0, p, num_iterations, epsilon=epsilon, want_attn=want_attn, opts=opts,
                            prepared=prepared)
    k, v, _ = kv_project(inputs, p, kv=kv, enc=enc, pos_table=pos_table)
    slots, attn, _ = iterate(k, v, slots0, p, num_iterations, epsilon=epsilon, want_attn=want_attn, opts=opts,
                             prepared=prepared)
    return slots, attn


class SlotAttentionFunction(torch.autograd.Function):
    """SlotAttention.forward(inputs, slots) with a fused CUDA forward and backward.

    The backward recomputes the attention logits from k and the saved per-iteration slots; only
    O(T*K*D) state per image is kept between the passes (plus k, v themselves).
    """

    @staticmethod
    def forward(ctx, inputs, slots0, num_iterations, epsilon, kv, *params):
        p = dict(zip(SA_PARAM_ORDER, params))
        inputs_c = _f32c(inputs)
        k, v, _ = kv_project(inputs_c, p, kv=kv)
        slots, attn, saved = iterate(k, v, slots0, p, num_iterations, epsilon=epsilon, want_attn=True, save=True)
        ctx.save_for_backward(inputs_c, k, v, saved, *[_f32c(t) for t in params])
        ctx.meta = (num_iterations, epsilon, kv, slots0.shape[1])
        return slots, attn

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_slots, d_attn):
        inputs, k, v, saved, *params = ctx.saved_tensors
        T, epsilon, kv, K = ctx.meta
        p = dict(zip(SA_PARAM_ORDER, params))
        B, N, D = k.shape
        C = inputs.shape[-1]
        H = p["mlp.0.weight"].shape[0]
        dt_code = abi.DT_BF16 if k.dtype == torch.bfloat16 else abi.DT_F32
        dims = abi.make_dims(B, N, C, D, H, K, T, eps=epsilon, kv_dtype=dt_code)
        dev = k.device
        _, bwd_ws, _ = abi.query_workspace(dims)
        ws = torch.empty(max(bwd_ws, 16), device=dev, dtype=torch.uint8)
        if B == 0:  # an empty shard (uneven dp.shard): the kernels do not run, every gradient is exactly zero
            return (torch.zeros_like(inputs), torch.zeros(0, K, D, device=dev), None, None, None,
                    *[torch.zeros_like(p[n]) for n in SA_PARAM_ORDER])
        # the projection backward works straight from the rank-(2 T K) coefficients of the iteration backward
        # (dk, dv [B,N,D] fp32 are never written); LOWRANK_BWD = False keeps the two-step form for cross-checks
        lowrank = LOWRANK_BWD and C == 64 and 2 * T * K <= 224
        dk = None if lowrank else torch.empty(B, N, D, device=dev, dtype=torch.float32)
        dv = None if lowrank else torch.empty(B, N, D, device=dev, dtype=torch.float32)
        d_slots0 = torch.empty(B, K, D, device=dev, dtype=torch.float32)
        g = {n: torch.empty_like(p[n]) for n in SA_PARAM_ORDER}
        dw = abi.sa_weight_grads(
            ln_slots_w=g["norm_slots.weight"], ln_slots_b=g["norm_slots.bias"], ln_mlp_w=g["norm_mlp.weight"],
            ln_mlp_b=g["norm_mlp.bias"], wq=g["project_q.weight"], w_ih=g["gru.weight_ih"], w_hh=g["gru.weight_hh"],
            b_ih=g["gru.bias_ih"], b_hh=g["gru.bias_hh"], w1=g["mlp.0.weight"], b1=g["mlp.0.bias"],
            w2=g["mlp.2.weight"], b2=g["mlp.2.bias"])
        w = _sa_weights(p)
        d_slots = _f32c(d_slots)
        d_attn_c = _f32c(d_attn) if d_attn is not None else None
        L = abi.lib()
        with _timed("sa_iter_bwd"):
            abi.check(L.ocrl_sa_iter_bwd(ctypes.byref(dims), abi.ptr(k), abi.ptr(v), abi.ptr(saved), ctypes.byref(w),
                                         abi.ptr(d_slots), abi.ptr(d_attn_c), abi.ptr(dk), abi.ptr(dv),
                                         abi.ptr(d_slots0), ctypes.byref(dw), abi.ptr(ws), abi.stream_ptr()),
                      "ocrl_sa_iter_bwd")
        # norm_inputs + project_k / project_v
        dx = torch.empty_like(inputs)
        tw = abi.token_weights(in_ln_w=p["norm_inputs.weight"], in_ln_b=p["norm_inputs.bias"],
                               wk=p["project_k.weight"], wv=p["project_v.weight"])
        if lowrank:
            ws2 = torch.empty(max(L.ocrl_kv_proj_bwd_lowrank_workspace(ctypes.byref(dims)), 16), device=dev,
                              dtype=torch.uint8)
            with _timed("kv_proj_bwd"):
                abi.check(L.ocrl_kv_proj_bwd_lowrank(ctypes.byref(dims), abi.ptr(inputs), ctypes.byref(tw),
                                                     abi.ptr(saved), abi.ptr(ws), abi.ptr(dx),
                                                     abi.ptr(g["norm_inputs.weight"]), abi.ptr(g["norm_inputs.bias"]),
                                                     abi.ptr(g["project_k.weight"]), abi.ptr(g["project_v.weight"]),
                                                     abi.ptr(ws2), abi.stream_ptr()), "ocrl_kv_proj_bwd_lowrank")
            return (dx, d_slots0, None, None, None, *[g[n] for n in SA_PARAM_ORDER])
        ws2 = torch.empty(max(L.ocrl_kv_proj_bwd_workspace(ctypes.byref(dims)), 16), device=dev, dtype=torch.uint8)
        with _timed("kv_proj_bwd"):
            abi.check(L.ocrl_kv_proj_bwd(ctypes.byref(dims), abi.ptr(inputs), ctypes.byref(tw), abi.ptr(dk),
                                         abi.ptr(dv), abi.ptr(dx), abi.ptr(g["norm_inputs.weight"]),
                                         abi.ptr(g["norm_inputs.bias"]), abi.ptr(g["project_k.weight"]),
                                         abi.ptr(g["project_v.weight"]), abi.ptr(ws2), abi.stream_ptr()),
                      "ocrl_kv_proj_bwd")
        return (dx, d_slots0, None, None, None, *[g[n] for n in SA_PARAM_ORDER])

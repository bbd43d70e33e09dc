"""Timing of the hand-written convolutions against cuDNN (CUDA events); development aid.
    python scripts/quick_conv.py [B] [size]"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import abi  # noqa: E402


def timeit(fn, warm=3, rep=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep


B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = "cuda"
L = abi.lib()
w = torch.randn(64, 64, 5, 5, device=dev) * 0.05
bias = torch.randn(64, device=dev)
pk = torch.empty(25 * 64 * 64, device=dev, dtype=torch.bfloat16)
abi.check(L.ocrl_conv5x5_pack_weights(abi.ptr(w), ctypes.c_void_p(pk.data_ptr()), 64, 64, abi.stream_ptr()), "pack")
n = L.ocrl_conv_padded_bytes(B, S, S) // 128
a = torch.zeros(n, 64, device=dev, dtype=torch.bfloat16)
b = torch.empty_like(a)
a.view(-1, S + 4, 64)[2:].view(B, S + 2, S + 4, 64)[:, :S, 2:S + 2] = torch.randn(B, S, S, 64, device=dev).to(torch.bfloat16)
st = abi.stream_ptr()


def own():
    abi.check(L.ocrl_conv5x5_c64_tc(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(pk.data_ptr()), abi.ptr(bias),
                                    ctypes.c_void_p(b.data_ptr()), B, S, S, 1, st), "conv")


x = torch.randn(B, 64, S, S, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
wb = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
bb = bias.to(torch.bfloat16)


def lib():
    return torch.cudnn_convolution_relu(x, wb, bb, [1, 1], [2, 2], [1, 1], 1)


flops = 2.0 * B * S * S * 64 * 64 * 25
# reference of the padded layout: fp32 convolution of the same bf16 operands (+ bias, ReLU), zeros at the padding positions
ain = a.view(-1, S + 4, 64)[2:].view(B, S + 2, S + 4, 64)[:, :S, 2:S + 2].permute(0, 3, 1, 2).float()
ref = torch.relu(torch.nn.functional.conv2d(ain, w.to(torch.bfloat16).float(), bias, padding=2)).permute(0, 2, 3, 1)
for v in (1, 2, 3, 4):  # 1: three tiles / four stages, 2: two tiles / eight stages, 3 / 4: paired taps (four / three stages)
    L.ocrl_dev_conv_variant(v)
    try:
        b.fill_(7.0)
        own()
        torch.cuda.synchronize()
        got = b.view(-1, S + 4, 64)[2:].view(B, S + 2, S + 4, 64)
        inner = got[:, :S, 2:S + 2].float()
        err = float((inner - ref).norm() / ref.norm())
        pad_zero = float(b.float().abs().sum() - inner.abs().sum())
        tv = timeit(own)
        print(json.dumps({"variant": v, "own_us": round(tv * 1e3, 1), "rel_err": err, "padding_abs_sum": pad_zero}), flush=True)
    except RuntimeError as e:
        print("variant", v, "failed:", str(e)[:100])
L.ocrl_dev_conv_variant(int(os.environ.get("QCV", 0)))
t_own, t_lib = timeit(own), timeit(lib)
print(json.dumps({"B": B, "size": S, "own_us": round(t_own * 1e3, 1), "cudnn_us": round(t_lib * 1e3, 1),
                  "own_tflops": round(flops / t_own / 1e9, 1), "cudnn_tflops": round(flops / t_lib / 1e9, 1),
                  "own_frac_of_1678": round(flops / t_own / 1e9 / 1678.2, 3)}))

trace = torch.zeros(128 + 4 * 148, dtype=torch.int64, device=dev)
L.ocrl_dev_conv_trace(ctypes.c_void_p(trace.data_ptr()))
own(); own()
torch.cuda.synchronize()
L.ocrl_dev_conv_trace(None)
tr = trace.cpu().tolist()
t0 = tr[0]
print(f"CTA 0: kernel start -> last store done: {tr[1] - t0} cycles")
for u in range(8):
    r = tr[16 + u * 8: 24 + u * 8]
    if r[0] == 0:
        break
    print(f" unit {u}: issuer begin {r[0]-t0:7d} slab ready {r[1]-t0:7d} acc free {r[2]-t0:7d} issued {r[3]-t0:7d} (weight waits {r[4]:6d}) | "
          f"epilogue wait {r[5]-t0:7d} acc ready {r[6]-t0:7d} done {r[7]-t0:7d}")

# spread of the persistent CTAs (global timer, ns): start and end relative to the earliest start
ct = [tr[128 + 4 * b: 132 + 4 * b] for b in range(148) if tr[128 + 4 * b]]
if ct:
    t0g = min(c[0] for c in ct)
    starts = sorted(c[0] - t0g for c in ct); ends = sorted(c[1] - t0g for c in ct)
    print(f"CTAs {len(ct)}: start min/median/max {starts[0]} / {starts[len(ct) // 2]} / {starts[-1]} ns; end min/median/max {ends[0]} / {ends[len(ct) // 2]} / {ends[-1]} ns")
    by_tiles = {}
    for c in ct:
        by_tiles.setdefault(c[3], []).append(c[1] - c[0])
    for k in sorted(by_tiles):
        v = sorted(by_tiles[k]); print(f"  {len(v)} CTAs with {k} tiles: duration min/median/max {v[0]} / {v[len(v) // 2]} / {v[-1]} ns")
    slow = sorted(ct, key=lambda c: c[0] - c[1])[:6]
    print("  slowest CTAs (sm, tiles, duration ns):", [(c[2], c[3], c[1] - c[0]) for c in slow])

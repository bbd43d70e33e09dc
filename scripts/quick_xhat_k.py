"""Factored loop at 9 .. 16 slots: time and parity of the dispatcher variants (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import slot_oracle as so
from tests.golden_io import rel_err
from ocrl_b200 import functional as F, abi

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): fn()
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for N, B in ((4096, 64), (16384, 16)):
    for K in (11, 16):
        torch.manual_seed(0)
        p = {k: v.cuda() for k, v in so.random_sa_params(K, 64, 192, 192, seed=3).items()}
        x = torch.randn(B, N, 64, device="cuda"); s0 = torch.randn(B, K, 192, device="cuda")
        xh, _, _ = F.kv_project(x, p, kv="bf16", xhat_only=True)
        pc = {k: v.cpu() for k, v in p.items()}
        sr, ar = so.slot_attention(x[:1].cpu(), s0[:1].cpu(), pc, 3)
        for dv in (6, 0):
            abi.lib().ocrl_dev_iter_variant(dv)
            prep = F.PreparedWeights()
            us = timeit(lambda: F.iterate_xhat(xh, s0, p, 3, prepared=prep))
            s, a = F.iterate_xhat(xh, s0, p, 3)
            print(f"N={N} B={B} K={K} dev_variant={dv}: {us:.1f} us  slots {rel_err(s[:1].cpu(), sr):.2e} attn {rel_err(a[:1].cpu(), ar):.2e}", flush=True)
abi.lib().ocrl_dev_iter_variant(0)

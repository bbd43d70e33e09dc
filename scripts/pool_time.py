"""Rollout pooling: fused kernel vs the module's torch path (development aid).  python scripts/pool_time.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200.pooling import Transformer  # noqa: E402


def timeit(fn, rep=200):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep * 1e3


torch.manual_seed(0)
t = Transformer(192, 128, 8, 1).cuda().eval()
for B in (4, 32, 256):
    x = torch.randn(B, 6, 192, device="cuda")
    with torch.no_grad():
        fused = timeit(lambda: t(x))
        ref = timeit(lambda: t._trans(torch.cat([t._cls_token().repeat(B, 1, 1), t._linear(x)], dim=1).permute(1, 0, 2))[0])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = t(x)
        graphed = timeit(g.replay)
    print(f"B={B}: fused {fused:.1f} us (graph replay {graphed:.1f} us), torch path {ref:.1f} us")

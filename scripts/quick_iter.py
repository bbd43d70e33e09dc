"""Iteration-kernel timing by variant (CUDA events, inputs larger than L2 at B = 64); development aid.

    python scripts/quick_iter.py [B] [N] [K] [T] [D] [H]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import abi, functional as F  # noqa: E402
from oracle import slot_oracle as so  # noqa: E402


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


def main():
    a = [int(x) for x in sys.argv[1:]]
    B, N, K, T, D, H = (a + [64, 4096, 6, 3, 192, 192][len(a):])[:6]
    p = {k: v.cuda() for k, v in so.random_sa_params(K, 64, D, H, seed=3).items()}
    x = torch.randn(B, N, 64, device="cuda")
    s0 = torch.randn(B, K, D, device="cuda")
    k, v, _ = F.kv_project(x, p, kv="bf16")
    bytes_img = 2 * N * D * 2 + N * K * 4 + 2 * K * D * 4
    ref = None
    dev = [int(x) for x in os.environ.get("OCRL_DEV_VARIANTS", "0").split(",")]
    cases = []
    for dv in dev:
        cases += [(f"tcgen05_v{dv}", dict(variant="tcgen05"), dv), (f"tcgen05_2lanes_v{dv}", dict(variant="tcgen05", lanes=2), dv)]
    cases += [("pipe", dict(variant="pipe"), 0)]
    for name, kw, dv in cases:
        abi.lib().ocrl_dev_iter_variant(dv)
        opts = abi.launch_opts(strict=True, **kw)
        try:
            s, at, _ = F.iterate(k, v, s0, p, T, opts=opts)
            torch.cuda.synchronize()
        except RuntimeError as e:
            print(json.dumps({"variant": name, "error": str(e)[:200]}))
            continue
        if ref is None:
            ref = (s, at)
        ms = timeit(lambda: F.iterate(k, v, s0, p, T, opts=opts))
        print(json.dumps({"variant": name, "kernel": F.last_kernel(), "B": B, "N": N, "K": K, "T": T, "D": D,
                          "us": round(ms * 1e3, 1), "GBps": round(B * bytes_img / ms / 1e6, 1),
                          "frac_of_6551": round(B * bytes_img / ms / 1e6 / 6551, 4),
                          "finite": bool(torch.isfinite(s).all()),
                          "slots_vs_first": float((s - ref[0]).norm() / ref[0].norm()),
                          "attn_vs_first": float((at - ref[1]).norm() / ref[1].norm())}))


if __name__ == "__main__":
    main()

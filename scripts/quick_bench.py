"""Quick kernel timing (CUDA events) for development; bench.py is the contract benchmark."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import functional as F  # noqa: E402
from oracle import slot_oracle as so  # noqa: E402


def timeit(fn, warm=3, rep=10):
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
    B = int(os.environ.get("QB_B", 64))
    N = int(os.environ.get("QB_N", 4096))
    K = int(os.environ.get("QB_K", 6))
    T = int(os.environ.get("QB_T", 3))
    D = int(os.environ.get("QB_D", 192))
    H = int(os.environ.get("QB_H", D))
    p = {k: v.cuda() for k, v in so.random_sa_params(K, 64, D, H, seed=3).items()}
    x = torch.randn(B, N, 64, device="cuda")
    s0 = torch.randn(B, K, D, device="cuda")
    out = {}
    for kv, esz in (("fp32", 4), ("bf16", 2)):
        k, v, _ = F.kv_project(x, p, kv=kv)
        t_proj = timeit(lambda: F.kv_project(x, p, kv=kv))
        t_iter = timeit(lambda: F.iterate(k, v, s0, p, T))
        bytes_img = 2 * N * D * esz + N * K * 4 + 2 * K * D * 4
        out[kv] = dict(proj_ms=t_proj, iter_ms=t_iter, iter_img_s=B / t_iter * 1e3,
                       iter_GBps=B * bytes_img / t_iter / 1e6,
                       proj_GBps=B * (N * 64 * 4 + 2 * N * D * esz) / t_proj / 1e6)
    out["cfg"] = dict(B=B, N=N, K=K, T=T, D=D, H=H, cluster=os.environ.get("OCRL_SA_CLUSTER", "auto"))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

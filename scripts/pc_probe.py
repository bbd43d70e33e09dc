"""Development probe of the persistent-cluster iteration kernel (sa_iter_fwd_pc.cu): parity against the
oracle on exact bf16 inputs, then CUDA-event timings of every variant next to the previous kernel.
Run each step under `timeout`: a protocol bug in the cluster exchange shows up as a hang."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import functional as F  # noqa: E402
from oracle import slot_oracle as so  # noqa: E402
from tests.golden_io import load_case, rel_err  # noqa: E402


def cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def set_variant(v):
    """'pipeN' two-engine kernel variant N; 'pcN' single-engine variant N; 'old' the per-image cluster kernel"""
    v = str(v)
    if v.startswith("pipe"):
        os.environ["OCRL_SA_PIPE"] = v[4:] or "0"
    elif v.startswith("pc"):
        os.environ["OCRL_SA_PIPE"] = "-1"
        os.environ["OCRL_SA_PC"] = v[2:] or "0"
    else:
        os.environ["OCRL_SA_PIPE"] = "-1"
        os.environ["OCRL_SA_PC"] = "-1"


def parity(variant):
    set_variant(variant)
    for name in ("sa_slate_grad", "sa_sharp", "sa_small_grad"):
        meta, g = load_case(name)
        if meta["K"] > 8 or g["p"]["project_q.weight"].shape[0] != 192:
            continue
        k_ref, v_ref = so.kv_project(g["in"]["inputs"], g["p"])
        kb, vb = k_ref.bfloat16(), v_ref.bfloat16()
        s_ref, a_ref = so.iterate(kb.float(), vb.float(), g["in"]["slots0"], g["p"], meta["T"], meta["eps"])
        s, a, _ = F.iterate(kb.cuda(), vb.cuda(), g["in"]["slots0"].cuda(), cuda(g["p"]), meta["T"], epsilon=meta["eps"])
        torch.cuda.synchronize()
        print(f"variant {variant} {name}: B={kb.shape[0]} N={kb.shape[1]} K={meta['K']} T={meta['T']} "
              f"slots {rel_err(s.cpu(), s_ref):.2e} attn {rel_err(a.cpu(), a_ref):.2e} "
              f"rowsum {float((a.sum(-1) - 1).abs().max()):.1e}", flush=True)
    # ragged N, odd batch, other K / T
    for (B, N, K, T) in ((3, 100, 5, 2), (5, 1000, 7, 4), (17, 4096, 6, 3), (1, 16, 1, 1)):
        p = so.random_sa_params(K, 64, 192, 192, seed=11)
        gen = torch.Generator().manual_seed(B * 7 + N)
        x = torch.randn(B, N, 64, generator=gen)
        s0 = torch.randn(B, K, 192, generator=gen)
        k_ref, v_ref = so.kv_project(x, p)
        kb, vb = k_ref.bfloat16(), v_ref.bfloat16()
        s_ref, a_ref = so.iterate(kb.float(), vb.float(), s0, p, T, 1e-8)
        s, a, _ = F.iterate(kb.cuda(), vb.cuda(), s0.cuda(), cuda(p), T)
        torch.cuda.synchronize()
        print(f"variant {variant} B={B} N={N} K={K} T={T}: slots {rel_err(s.cpu(), s_ref):.2e} "
              f"attn {rel_err(a.cpu(), a_ref):.2e}", flush=True)


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


def timing(variants, shapes):
    for (B, N, K, T) in shapes:
        p = cuda(so.random_sa_params(K, 64, 192, 192, seed=3))
        x = torch.randn(B, N, 64, device="cuda")
        s0 = torch.randn(B, K, 192, device="cuda")
        k, v, _ = F.kv_project(x, p, kv="bf16")
        ws = torch.empty(1 << 22, device="cuda", dtype=torch.uint8)
        bytes_img = 2 * N * 192 * 2 + N * K * 4 + 2 * K * 192 * 4
        flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
        for var in variants:
            set_variant(var)
            ms = timeit(lambda: F.iterate(k, v, s0, p, T, _workspace=ws))
            # cold: L2 flushed before every launch (k/v come from HBM in the first pass)

            def cold():
                flush.zero_()
                F.iterate(k, v, s0, p, T, _workspace=ws)
            ms_cold = timeit(cold) - timeit(lambda: flush.zero_())
            print(f"B={B} N={N} K={K} T={T} variant {var}: {ms*1e3:.1f} us ({B*bytes_img/ms/1e6:.0f} GB/s, "
                  f"{B*bytes_img/ms/1e6/6551:.3f} of HBM) | L2-flushed {ms_cold*1e3:.1f} us "
                  f"({B*bytes_img/ms_cold/1e6/6551:.3f})", flush=True)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "parity":
        parity(sys.argv[2])
    else:
        variants = sys.argv[2].split(",")
        timing(variants, [(64, 4096, 6, 3)] if mode == "time" else [(64, 4096, 6, 3), (256, 4096, 6, 3), (32, 16384, 6, 3), (64, 4096, 8, 7)])

"""Factored inference path (x^ streaming): parity on the goldens and kernel times at the BASELINE shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import slot_oracle as so
from tests.golden_io import load_case, rel_err
from ocrl_b200 import functional as F, abi

cuda = lambda d: {k: v.cuda() for k, v in d.items()}
for name in ("sa_slate_grad", "sa_sharp", "sa_k1_t1", "sa_k11_t5_ragged", "sa_k16_t7"):
    meta, g = load_case(name)
    if meta["D"] != 192:
        continue
    p = g["p"]
    s_ref, a_ref = so.slot_attention(g["in"]["inputs"], g["in"]["slots0"], p, meta["T"], meta["eps"])
    out = {}
    for fac in (False, True):
        s, a = F.slot_attention(g["in"]["inputs"].cuda(), g["in"]["slots0"].cuda(), cuda(p), meta["T"], epsilon=meta["eps"],
                                kv="bf16", factored=fac)
        torch.cuda.synchronize()
        out[fac] = (s.cpu(), a.cpu())
        print(f"{name:18s} factored={fac!s:5s} kernel={F.last_kernel():13s} slots {rel_err(s.cpu(), s_ref):.2e} attn {rel_err(a.cpu(), a_ref):.2e}", flush=True)

torch.manual_seed(0)
B = int(os.environ.get("QB", 64))
p = cuda(so.random_sa_params(6, 64, 192, 192, seed=3))
x = torch.randn(B, 4096, 64, device="cuda")
s0 = torch.randn(B, 6, 192, device="cuda")
def timeit(fn, n=20):
    """n back-to-back launches replayed from one CUDA graph (no host time between the kernels)"""
    for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn()
        with torch.cuda.graph(g, stream=st):
            for _ in range(n): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
prep = F.PreparedWeights()
xh, _, _ = F.kv_project(x, p, kv="bf16", xhat_only=True)
k, v, _ = F.kv_project(x, p, kv="bf16")
print("token stage  kv   %.1f us" % timeit(lambda: F.kv_project(x, p, kv="bf16")))
print("token stage  xhat %.1f us" % timeit(lambda: F.kv_project(x, p, kv="bf16", xhat_only=True)))
print("iteration    kv   %.1f us" % timeit(lambda: F.iterate(k, v, s0, p, 3, prepared=prep)))
for dv in (0, 5, 2, 4):  # dispatcher variants: 0 = default (256-token steps), 5 = 64-token half tiles, 2 = 3 lanes / 2 streams, 4 = 256-token steps with 4 streams
    abi.lib().ocrl_dev_iter_variant(dv)
    for mc in (0,):
        o = abi.launch_opts(max_clusters=mc)
        print("iteration    xhat dev_variant=%d max_clusters=%d %.1f us" % (dv, mc, timeit(lambda: F.iterate_xhat(xh, s0, p, 3, prepared=prep, opts=o))))
    sa, aa = F.iterate_xhat(xh, s0, p, 3)
    sb, ab, _ = F.iterate(k, v, s0, p, 3)
    print("  xhat vs kv at B=%d: slots %.2e attn %.2e" % (B, rel_err(sa, sb), rel_err(aa, ab)))
abi.lib().ocrl_dev_iter_variant(int(os.environ.get("QV", 0)))
sa, aa = F.iterate_xhat(xh, s0, p, 3)
sb, ab, _ = F.iterate(k, v, s0, p, 3)
print("xhat vs kv at B=%d: slots %.2e attn %.2e" % (B, rel_err(sa, sb), rel_err(aa, ab)))
pc = {k_: v_.cpu() for k_, v_ in p.items()}
sr, ar = so.slot_attention(x[:2].cpu(), s0[:2].cpu(), pc, 3)
print("xhat vs oracle (2 images): slots %.2e attn %.2e" % (rel_err(sa[:2].cpu(), sr), rel_err(aa[:2].cpu(), ar)))

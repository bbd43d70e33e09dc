"""Timeline of the tcgen05 iteration kernel (clock64 of CTA 0).   python scripts/trace_umma.py [B] [lanes] [xhat]
(needs a build with the stamps compiled in:  OCRL_NVCC_FLAGS=-DOCRL_UMMA_TRACE=1 python -m ocrl_b200.build)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import abi, functional as F  # noqa: E402
from oracle import slot_oracle as so  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
XHAT = len(sys.argv) > 3 and sys.argv[3] == "xhat"  # the factored form (ocrl_sa_iter_fwd_xhat)
N, K, T, D = 4096, 6, 3, 192
p = {k: v.cuda() for k, v in so.random_sa_params(K, 64, D, D, seed=3).items()}
x = torch.randn(B, N, 64, device="cuda"); s0 = torch.randn(B, K, D, device="cuda")
k, v, _ = F.kv_project(x, p, kv="bf16")
dims = abi.make_dims(B, N, 64, D, D, K, T, kv_dtype=abi.DT_BF16, math_mode=abi.MATH_TENSOR)
nbytes = abi.query_workspace(dims)[0]
ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
opts = abi.launch_opts(variant="tcgen05", lanes=lanes, strict=True, trace=True)
if XHAT:
    import ctypes
    xh, _, _ = F.kv_project(x, p, kv="bf16", xhat_only=True)
    slots = torch.empty(B, K, D, device="cuda"); attn = torch.empty(B, N, K, device="cuda")
    w = F._sa_weights(p)
    abi.lib().ocrl_dev_iter_variant(int(os.environ.get("QV", 0)))
    for _ in range(3):
        abi.check(abi.lib().ocrl_sa_iter_fwd_xhat(ctypes.byref(dims), abi.ptr(xh), abi.ptr(p["project_k.weight"]),
                                                  abi.ptr(p["project_v.weight"]), abi.ptr(s0), ctypes.byref(w), abi.ptr(slots),
                                                  abi.ptr(attn), abi.ptr(ws), ctypes.byref(opts), abi.stream_ptr()), "xhat")
else:
    for _ in range(3):
        F.iterate(k, v, s0, p, T, _workspace=ws, opts=opts)
torch.cuda.synchronize()
tr = ws[nbytes - 4096:].view(torch.int64).cpu().tolist()
t0 = tr[0]
print(f"B={B} lanes={lanes}: cycles since the post-setup cluster sync")
print(f"first cluster: setup {tr[0]-tr[1]} cycles, main loop {tr[3]-tr[0]}; last cluster: setup {tr[5]-tr[4]}, main loop {tr[6]-tr[5]}; "
      f"last cluster started {tr[4]-tr[1]} cycles after the first (different SM clocks: indicative only)")
for n in range(40):
    b = tr[8 + n * 8: 16 + n * 8]
    if b[0] == 0:
        break
    r = [x - t0 for x in b]
    print(f"op {n:2d}: pass {r[0]:8d} .. {r[1]:8d} ({r[1]-r[0]:6d}) | update {r[2]:8d} R1 +{r[3]-r[2]:5d} GRU +{r[4]-r[3]:5d} "
          f"MLP1 +{r[5]-r[4]:5d} MLP2 +{r[6]-r[5]:5d} end +{r[7]-r[6]:5d}  = {r[7]-r[2]:6d}")
base = tr[340] or tr[346]
print("op 2, per 128-token pair (cycles since the first stamp):")
for pi in range(4):
    b = [x - base if x else -1 for x in tr[340 + pi * 12: 352 + pi * 12]]
    print(f" pair {pi}: softmax wait {b[0]} -> logits ready {b[1]} -> loaded {b[2]} -> softmax done {b[3]} -> w slot free {b[4]} -> w written {b[5]}"
          f" | issuer: logits begin {b[6]} issued {b[7]} ; U begin {b[8]} w ready {b[9]} issued {b[10]}")
print("op 2 end (factored): pairs done %d -> token sums reduced %d -> drain begin %d -> stream's buffers free %d" % tuple(tr[340 + 12 * i + 11] - base for i in range(4)))
print("producers (op 2 half tiles): k [wait begin, slot free] v [wait begin, slot free]")
for j in range(8):
    b = [x - base if x else -1 for x in tr[400 + j * 4: 404 + j * 4]]
    print(f" half {j}: k {b[0]} {b[1]} | v {b[2]} {b[3]}")

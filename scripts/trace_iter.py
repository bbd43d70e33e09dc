"""Phase timeline of the tensor-core iteration kernel (clock64 of CTA 0), development aid.
   OCRL_SA_TRACE=1 python scripts/trace_iter.py"""
import os
import sys

import torch

os.environ["OCRL_SA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import functional as F  # noqa: E402
from oracle import slot_oracle as so  # noqa: E402

B = int(os.environ.get("QB_B", 148)); N = int(os.environ.get("QB_N", 4096)); K, T, D = 6, 3, 192
p = {k: v.cuda() for k, v in so.random_sa_params(K, 64, D, D, seed=3).items()}
x = torch.randn(B, N, 64, device="cuda"); s0 = torch.randn(B, K, D, device="cuda")
k, v, _ = F.kv_project(x, p, kv="bf16")
import ctypes
from ocrl_b200 import abi
dims = abi.make_dims(B, N, 64, D, D, K, T, kv_dtype=abi.DT_BF16, math_mode=abi.MATH_TENSOR)
nbytes = abi.query_workspace(dims)[0]
ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    F.iterate(k, v, s0, p, T, _workspace=ws)
torch.cuda.synchronize()
tr = ws[nbytes - 4096:].view(torch.int64).cpu().tolist()
names = ["start"]
for t in range(T):
    names += [f"t{t} pass-start", f"t{t} pass-done(w0)", f"t{t} all-warps", f"t{t} pushed", f"t{t} sync1", f"t{t} sync2(upd)",
              f"t{t} sync3(gru)", f"t{t} sync4(mlp1)", f"t{t} sync5(mlp2)"]
prev = tr[0]
for i, nme in enumerate(names):
    print(f"{nme:22s} {tr[i]-tr[0]:9d}  (+{tr[i]-prev})")
    prev = tr[i]

print("pass detail (warp 0, first groups): issue+wait | logits | softmax | U-mma | total")
for g in range(6):
    b = tr[128 + g * 5: 128 + g * 5 + 5]
    print(f"  group {g}: wait {b[1]-b[0]:6d}  logits {b[2]-b[1]:6d}  softmax {b[3]-b[2]:6d}  U {b[4]-b[3]:6d}   start-to-start {b[0]-tr[128]:7d}")

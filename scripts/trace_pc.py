"""Phase timeline of the persistent-cluster iteration kernel (clock64 of CTA 0, group 0, thread 0).
   python scripts/trace_pc.py [variant]"""
import os
import sys

import torch

os.environ["OCRL_SA_TRACE"] = "1"
os.environ["OCRL_SA_PIPE"] = "-1"
os.environ["OCRL_SA_PC"] = sys.argv[1] if len(sys.argv) > 1 else "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import abi, functional as F  # noqa: E402
from oracle import slot_oracle as so  # noqa: E402

B = int(os.environ.get("QB_B", 64)); N = int(os.environ.get("QB_N", 4096)); K, T, D = 6, 3, 192
p = {k: v.cuda() for k, v in so.random_sa_params(K, 64, D, D, seed=3).items()}
x = torch.randn(B, N, 64, device="cuda"); s0 = torch.randn(B, K, D, device="cuda")
k, v, _ = F.kv_project(x, p, kv="bf16")
dims = abi.make_dims(B, N, 64, D, D, K, T, kv_dtype=abi.DT_BF16, math_mode=abi.MATH_TENSOR)
nbytes = abi.query_workspace(dims)[0]
ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    F.iterate(k, v, s0, p, T, _workspace=ws)
torch.cuda.synchronize()
tr = ws[nbytes - 4096:].view(torch.int64).cpu().tolist()
print(f"setup (weights, barriers, cluster sync): {tr[1]-tr[0]} cycles")
names = ["pass start", "logit warp 0 done", "pass end (group sync)", "R1 pushed", "R1 arrived", "R2 arrived (updates)",
         "GRU mma done", "R3 arrived (h')", "MLP1 done", "R4 arrived (hidden)", "MLP2 done", "R5 arrived (slots)"]
for m in range(3):
    for t in range(T):
        tb = 2 + (m * T + t) * 12
        if tb + 12 >= 500 or tr[tb] == 0:
            continue
        print(f"image {m} iteration {t}: start at {tr[tb]-tr[0]}")
        prev = tr[tb]
        for i, nme in enumerate(names):
            if i == 11 and t == T - 1:
                continue
            print(f"   {nme:26s} +{tr[tb+i]-prev:7d}   (t={tr[tb+i]-tr[tb]})")
            prev = tr[tb + i]

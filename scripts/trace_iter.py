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
ws = torch.zeros(256, dtype=torch.int64, device="cuda")
for _ in range(3):
    F.iterate(k, v, s0, p, T, _workspace=ws)
torch.cuda.synchronize()
tr = ws.cpu().tolist()
names = ["start"]
for t in range(T):
    names += [f"t{t} pass-start", f"t{t} pass-done(w0)", f"t{t} all-warps", f"t{t} pushed", f"t{t} sync1", f"t{t} sync2(upd)",
              f"t{t} sync3(gru)", f"t{t} sync4(mlp1)", f"t{t} sync5(mlp2)"]
prev = tr[0]
for i, nme in enumerate(names):
    print(f"{nme:22s} {tr[i]-tr[0]:9d}  (+{tr[i]-prev})")
    prev = tr[i]

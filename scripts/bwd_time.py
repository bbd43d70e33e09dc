"""One forward + backward of the fused slot attention at the BASELINE size (development aid: run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel times of the backward).  python scripts/bwd_time.py [kv]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import functional as F  # noqa: E402
from oracle import slot_oracle as so  # noqa: E402

kv = sys.argv[1] if len(sys.argv) > 1 else "bf16"
B, N, K, T, D = 64, 4096, 6, 3, 192
p = {k: v.cuda().requires_grad_(True) for k, v in so.random_sa_params(K, 64, D, D, seed=5).items()}
x = torch.randn(B, N, 64, device="cuda", requires_grad=True)
s0 = torch.randn(B, K, D, device="cuda", requires_grad=True)
for it in range(3):
    slots, attn = F.SlotAttentionFunction.apply(x, s0, T, 1e-8, kv, *[p[n] for n in F.SA_PARAM_ORDER])
    (slots.sum() + (attn * attn).sum()).backward()
torch.cuda.synchronize()
print("ok")

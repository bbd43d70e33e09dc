"""One forward + backward of the fused slot attention at the BASELINE size (development aid: run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel times of the backward).  python scripts/bwd_time.py [kv]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import functional as F  # noqa: E402
from oracle import slot_oracle as so  # noqa: E402

kv = sys.argv[1] if len(sys.argv) > 1 else "bf16"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T = int(sys.argv[3]) if len(sys.argv) > 3 else 3
B, K, D = 64, 6, 192
p = {k: v.cuda().requires_grad_(True) for k, v in so.random_sa_params(K, 64, D, D, seed=5).items()}
x = torch.randn(B, N, 64, device="cuda", requires_grad=True)
s0 = torch.randn(B, K, D, device="cuda", requires_grad=True)
for it in range(3):
    slots, attn = F.SlotAttentionFunction.apply(x, s0, T, 1e-8, kv, *[p[n] for n in F.SA_PARAM_ORDER])
    (slots.sum() + (attn * attn).sum()).backward()
torch.cuda.synchronize()
F.KERNEL_EVENTS = []
for it in range(5):
    slots, attn = F.SlotAttentionFunction.apply(x, s0, T, 1e-8, kv, *[p[n] for n in F.SA_PARAM_ORDER])
    (slots.sum() + (attn * attn).sum()).backward()
torch.cuda.synchronize()
per = {}
for name, e0, e1 in F.KERNEL_EVENTS:
    per.setdefault(name, []).append(e0.elapsed_time(e1))
print(f"kv={kv} N={N} T={T}: " + ", ".join(f"{k} {sum(v) / len(v) * 1e3:.0f} us" for k, v in per.items()))

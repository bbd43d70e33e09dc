"""Run one forward+backward of the fused slot-attention on a golden case (debug helper)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import functional as F  # noqa: E402
from tests.golden_io import load_case, rel_err  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sa_small_grad"
kv = sys.argv[2] if len(sys.argv) > 2 else "fp32"
meta, g = load_case(name)
x = g["in"]["inputs"].cuda().requires_grad_(True)
s0 = g["in"]["slots0"].cuda().requires_grad_(True)
p = {k: v.cuda().requires_grad_(True) for k, v in g["p"].items()}
slots, attn = F.SlotAttentionFunction.apply(x, s0, meta["T"], meta["eps"], kv, *[p[n] for n in F.SA_PARAM_ORDER])
torch.cuda.synchronize()
print("fwd ok", rel_err(slots.detach().cpu(), g["out"]["slots"]))
loss = (slots * g["g.out"]["slots"].cuda()).sum() + (attn * g["g.out"]["attn"].cuda()).sum()
loss.backward()
torch.cuda.synchronize()
print("bwd ok")
print("inputs", rel_err(x.grad.cpu(), g["g.in"]["inputs"]))
print("slots0", rel_err(s0.grad.cpu(), g["g.in"]["slots0"]))
for k, gv in g["g.p"].items():
    print(k, rel_err(p[k].grad.cpu(), gv), float((p[k].grad.cpu() - gv).abs().max()))

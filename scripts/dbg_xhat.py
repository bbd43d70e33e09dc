import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import slot_oracle as so
from tests.golden_io import load_case, rel_err
from ocrl_b200 import functional as F, abi
cuda = lambda d: {k: v.cuda() for k, v in d.items()}
meta, g = load_case("sa_slate_grad")
p = g["p"]; x = g["in"]["inputs"]; s0 = g["in"]["slots0"]
xh_ref = so.layer_norm(x, p["norm_inputs.weight"], p["norm_inputs.bias"])
xh, _, _ = F.kv_project(x.cuda(), cuda(p), kv="bf16", xhat_only=True)
print("xhat vs oracle", rel_err(xh.float().cpu(), xh_ref), xh.shape, xh.dtype)
for T in (1, 2, 3):
    s_ref, a_ref = so.slot_attention(x, s0, p, T, meta["eps"])
    s, a = F.iterate_xhat(xh_ref.bfloat16().cuda(), s0.cuda(), cuda(p), T, epsilon=meta["eps"])
    print("T", T, "slots", rel_err(s.cpu(), s_ref), "attn", rel_err(a.cpu(), a_ref))
# T=1: attention depends only on q'' (initial q phase) and x^
s_ref, a_ref, tr = so.iterate(*so.kv_project(x, p), s0, p, 1, meta["eps"], return_trace=True)
s, a = F.iterate_xhat(xh_ref.bfloat16().cuda(), s0.cuda(), cuda(p), 1, epsilon=meta["eps"])
print("attn row0 ref", a_ref[0, 0], "\n         got", a[0, 0].cpu())
print("slots ref", s_ref[0, 0, :8], "\n      got", s[0, 0, :8].cpu())

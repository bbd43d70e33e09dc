"""Check the tcgen05 token-stage kernel against the oracle on small cases (debug helper)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import functional as F  # noqa: E402
from oracle import slot_oracle as so  # noqa: E402
from tests.golden_io import load_case, rel_err  # noqa: E402

torch.manual_seed(0)
for D in (192, 64, 128):
    p = so.random_sa_params(6, 64, D, D, seed=5)
    enc = {"layer_norm.weight": 1 + 0.1 * torch.randn(64), "layer_norm.bias": 0.1 * torch.randn(64),
           "mlp.0.weight": 0.2 * torch.randn(64, 64), "mlp.0.bias": 0.1 * torch.randn(64),
           "mlp.2.weight": 0.2 * torch.randn(64, 64), "mlp.2.bias": 0.1 * torch.randn(64)}
    pc = {k: v.cuda() for k, v in p.items()}
    ec = {k: v.cuda() for k, v in enc.items()}
    for (B, N) in ((2, 256), (3, 100), (5, 1024)):
        x = torch.randn(B, N, 64) + 0.3
        k_ref, v_ref = so.kv_project(x, p)
        k, v, _ = F.kv_project(x.cuda(), pc, kv="bf16")
        torch.cuda.synchronize()
        print(f"D={D} B={B} N={N} plain : k {rel_err(k.float().cpu(), k_ref):.2e} v {rel_err(v.float().cpu(), v_ref):.2e}", flush=True)
        y_ref = so.token_mlp(x, enc)
        k_ref, v_ref = so.kv_project(y_ref, p)
        k, v, y = F.kv_project(x.cuda(), pc, kv="bf16", enc=ec, want_y=True)
        torch.cuda.synchronize()
        print(f"              mlp   : y {rel_err(y.cpu(), y_ref):.2e} k {rel_err(k.float().cpu(), k_ref):.2e} v {rel_err(v.float().cpu(), v_ref):.2e}", flush=True)
    # NCHW + position table
    S = 16
    fmap = torch.randn(2, 64, S, S)
    pos = torch.randn(64, S * S)
    tok = (fmap.flatten(2) + pos.unsqueeze(0)).permute(0, 2, 1).contiguous()
    y_ref = so.token_mlp(tok, enc)
    k_ref, v_ref = so.kv_project(y_ref, p)
    k, v, _ = F.kv_project(fmap.cuda(), pc, kv="bf16", enc=ec, pos_table=pos.cuda())
    torch.cuda.synchronize()
    print(f"D={D} nchw 16x16   : k {rel_err(k.float().cpu(), k_ref):.2e} v {rel_err(v.float().cpu(), v_ref):.2e}", flush=True)
print("done")

"""Development probe: the D=64 / H=128 ("Slot-Attention small") iteration kernel, cluster-size variants."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import functional as F
from oracle import slot_oracle as so
from tests.golden_io import rel_err

B, N, K, T = int(os.environ.get("QB_B", 64)), 4096, 6, 7
p = so.random_sa_params(K, 64, 64, 128, seed=5)
gen = torch.Generator().manual_seed(1)
x = torch.randn(B, N, 64, generator=gen)
s0 = torch.randn(B, K, 64, generator=gen)
k_ref, v_ref = so.kv_project(x, p)
kb, vb = k_ref.bfloat16(), v_ref.bfloat16()
sel = [0, B // 2, B - 1]
s_ref, a_ref = so.iterate(kb[sel].float(), vb[sel].float(), s0[sel], p, T, 1e-8)
pc = {k_: v_.cuda() for k_, v_ in p.items()}
kc, vc, sc = kb.cuda(), vb.cuda(), s0.cuda()
for var in sys.argv[1:]:
    os.environ["OCRL_SA_PIPE"] = var
    s, a, _ = F.iterate(kc, vc, sc, pc, T)
    torch.cuda.synchronize()
    err = (rel_err(s[sel].cpu(), s_ref), rel_err(a[sel].cpu(), a_ref))
    for _ in range(3):
        F.iterate(kc, vc, sc, pc, T)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        F.iterate(kc, vc, sc, pc, T)
    e1.record(); torch.cuda.synchronize()
    print(f"variant {var}: {e0.elapsed_time(e1) * 100:.1f} us  slots {err[0]:.2e} attn {err[1]:.2e}", flush=True)

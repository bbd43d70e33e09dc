"""Timeline of the two-engine iteration kernel (clock64 of CTA 0): per op n, pass engine (logit warp 0) and
update engine (thread 0) stamps.   python scripts/trace_pipe.py [variant]"""
import os
import sys

import torch

os.environ["OCRL_SA_TRACE"] = "1"
os.environ["OCRL_SA_PIPE"] = sys.argv[1] if len(sys.argv) > 1 else "0"
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
t0 = tr[0]
print("cycles since the post-setup cluster sync; pass: q ready -> logit warp 0 done; update: u_ready seen, R1 arrived, "
      "GRU mma done, MLP1 done, MLP2 done, op end")
for n in range(40):
    b = tr[8 + n * 8: 16 + n * 8]
    if b[0] == 0:
        break
    r = [x - t0 for x in b]
    print(f"op {n:2d}: pass {r[0]:8d} .. {r[1]:8d} ({r[1]-r[0]:6d}) | update {r[2]:8d} R1 +{r[3]-r[2]:5d} GRU +{r[4]-r[3]:5d} "
          f"MLP1 +{r[5]-r[4]:5d} MLP2 +{r[6]-r[5]:5d} end +{r[7]-r[6]:5d}  = {r[7]-r[2]:6d}")

print("chain 0 in op 4 (cycles since the first stamp): logit warp [wait begin, tile landed, logits done, w handed over] | "
      "U warp [wait begin, w ready, U mma done, refill issued]")
base = tr[340]
for i in range(8):
    b = tr[340 + i * 8: 348 + i * 8]
    if b[0] == 0:
        break
    r = [x - base for x in b]
    print(f"tile {4*i:2d}: logit {r[0]:6d} {r[1]:6d} {r[2]:6d} {r[3]:6d}  (wait {r[1]-r[0]:5d} mma {r[2]-r[1]:4d} softmax {r[3]-r[2]:4d}) | "
          f"U {r[4]:6d} {r[5]:6d} {r[6]:6d} {r[7]:6d}  (wait {r[5]-r[4]:5d} mma {r[6]-r[5]:4d} issue {r[7]-r[6]:4d})")

print("softmax segment detail: [logits done -> softmax+attn store done -> movmatrix+sts done -> syncwarp done -> arrived]")
for i in range(8):
    b = tr[340 + i * 8: 348 + i * 8]
    e = tr[420 + i * 4: 424 + i * 4]
    if b[0] == 0:
        break
    print(f"tile {4*i:2d}: softmax+store {e[0]-b[2]:5d}  pack/movmatrix/sts {e[1]-e[0]:5d}  syncwarp {e[2]-e[1]:5d}  arrive {b[3]-e[2]:5d}")

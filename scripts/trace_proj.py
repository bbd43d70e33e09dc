"""Phase timeline of the tcgen05 token-stage kernel (clock64 of CTA 0, row thread 0), development aid."""
import ctypes, os, sys
import torch
os.environ["OCRL_SA_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ocrl_b200 import abi, functional as F
from oracle import slot_oracle as so

p = {k: v.cuda() for k, v in so.random_sa_params(6, 64, 192, 192, seed=3).items()}
g = torch.Generator().manual_seed(1)
enc = {"layer_norm.weight": 1 + 0.1 * torch.randn(64, generator=g), "layer_norm.bias": 0.1 * torch.randn(64, generator=g),
       "mlp.0.weight": 0.2 * torch.randn(64, 64, generator=g), "mlp.0.bias": 0.1 * torch.randn(64, generator=g),
       "mlp.2.weight": 0.2 * torch.randn(64, 64, generator=g), "mlp.2.bias": 0.1 * torch.randn(64, generator=g)}
enc = {k: v.cuda() for k, v in enc.items()}
fmap = torch.randn(64, 64, 64, 64, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
pos = torch.randn(64, 4096, device="cuda")
import ocrl_b200.functional as FF
# reach into kv_project's workspace: run once through the ABI with our own workspace
dims = abi.make_dims(64, 4096, 64, 192, 192, 6, 3, kv_dtype=abi.DT_BF16, math_mode=abi.MATH_TENSOR, x_format=abi.X_TOKENS_BF16)
n = abi.lib().ocrl_kv_proj_fwd_workspace(ctypes.byref(dims))
print("workspace bytes", n)
orig_empty = torch.empty
holder = {}
def patched(*a, **k):
    t = orig_empty(*a, **k)
    if len(a) == 1 and a[0] == n and k.get("dtype") == torch.uint8:
        t.zero_(); holder["ws"] = t
    return t
torch.empty = patched
for _ in range(3):
    F.kv_project(fmap, p, kv="bf16", enc=enc, pos_table=pos)
torch.cuda.synchronize()
torch.empty = orig_empty
tr = holder["ws"][n - 1024:].view(torch.int64).cpu().tolist()
names = ["x tile landed", "x read + position add", "LN1 + A1 stored", "G1 done + relu + A2 stored", "G2 done + LN + A3 stored", "kv MMA done", "kv epilogue done"]
for it in range(8):
    b = tr[it * 8: it * 8 + 8]
    if b[0] == 0: continue
    print(f"tile {it+2}: " + "  ".join(f"{nme} +{b[i+1]-b[i]}" for i, nme in enumerate(names)) + f"   total {b[7]-b[0]}")

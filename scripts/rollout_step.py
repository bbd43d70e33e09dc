"""pooling(ocr(obs)) at the rollout batch, eager (for an ncu launch list).  python scripts/rollout_step.py [batch]"""
import os
import sys
from types import SimpleNamespace as NS

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrl_b200  # noqa: E402
from ocrl_b200 import synth  # noqa: E402
from ocrl_b200.config import slate_config  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
model = ocrl_b200.SLATE(*slate_config(num_slots=6, num_iterations=3, obs_size=64, kv_dtype="bf16"))
model.to("cuda")
model.eval()
pcfg = NS(d_model=128, nhead=8, num_layers=1, pos_emb="None", norm_first=False, use_mlp1=False, use_mlp2=False,
          cw_embedding=False, push_embedding=False)
pool = ocrl_b200.Transformer_Module(model.rep_dim, model.num_slots, pcfg).cuda().eval()
obs = synth.to_obs(torch.from_numpy(synth.random_objs_frames(B, 64, seed=1))).contiguous().cuda()
with torch.no_grad():
    for i in range(4):
        out = pool(model(obs))
torch.cuda.synchronize()
print("ok", tuple(out.shape))

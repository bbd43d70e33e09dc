"""Three eager SLATE encode steps in the bf16 mode at the bench configuration (for ncu: the kernels of the third step
are the ones to profile; inputs larger than the L2).   python scripts/ncu_step.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrl_b200  # noqa: E402
from ocrl_b200 import synth  # noqa: E402
from ocrl_b200.config import slate_config  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
model = ocrl_b200.SLATE(*slate_config(num_slots=6, num_iterations=3, obs_size=64, kv_dtype="bf16"))
model.to("cuda")
model.eval()
pool = synth.to_obs(torch.from_numpy(synth.random_objs_frames(3 * B, 64, seed=1))).contiguous().cuda()
with torch.no_grad():
    for i in range(3):
        out = model(pool[i * B:(i + 1) * B])
torch.cuda.synchronize()
print("ok", tuple(out.shape))

"""Streamed encode throughput for explicit (lanes, max_clusters) of the iteration kernel (development aid)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrl_b200  # noqa: E402
from ocrl_b200 import abi, synth  # noqa: E402
from ocrl_b200.config import slate_config  # noqa: E402

B = 64
torch.manual_seed(0)
model = ocrl_b200.SLATE(*slate_config(num_slots=6, num_iterations=3, obs_size=64, kv_dtype="bf16"))
model.to("cuda")
model.eval()
pool = synth.to_obs(torch.from_numpy(synth.random_objs_frames(8 * B, 64, seed=1))).contiguous().cuda()
outs = [torch.empty(B, 6, 192, device="cuda") for _ in range(3)]
for lanes, ncl in [(0, 0), (3, 6), (3, 7), (2, 0), (2, 8), (2, 10), (2, 11), (2, 13), (2, 15)]:
    model._module._slotattn.slot_attention.launch_opts = abi.launch_opts(lanes=lanes, max_clusters=ncl) if (lanes or ncl) else None
    enc = ocrl_b200.StreamedEncoder(model, pool[:B], iter_clusters=None)
    n = 3000
    for i in range(50):
        enc.submit(pool[(i % 8) * B:(i % 8 + 1) * B], outs[i % 3])
    enc.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        enc.submit(pool[(i % 8) * B:(i % 8 + 1) * B], outs[i % 3])
    enc.synchronize()
    dt = time.perf_counter() - t0
    print(f"lanes {lanes} max_clusters {ncl}: {B * n / dt:.0f} images/s")

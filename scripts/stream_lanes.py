"""Streamed encode throughput over the cluster cap of the iteration kernel and the batches in flight (development aid)."""
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
for buffers, ncl in [(3, "auto"), (3, None), (3, 13), (3, 10), (3, 8), (3, 5), (4, None), (4, 10), (2, None)]:
    enc = ocrl_b200.StreamedEncoder(model, pool[:B], iter_clusters=ncl, buffers=buffers)
    outs = [torch.empty(B, 6, 192, device="cuda") for _ in range(buffers)]
    n = 3000
    for i in range(50):
        enc.submit(pool[(i % 8) * B:(i % 8 + 1) * B], outs[i % buffers])
    enc.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        enc.submit(pool[(i % 8) * B:(i % 8 + 1) * B], outs[i % buffers])
    enc.synchronize()
    dt = time.perf_counter() - t0
    print(f"buffers {buffers} iter_clusters {ncl} (resolved {enc.iter_clusters}): {B * n / dt:.0f} images/s", flush=True)
    del enc

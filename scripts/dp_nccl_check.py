"""Data-parallel OCR training over NCCL (one process per GPU): a few SLATE.update steps on sharded batches.
Checks that every rank holds bit-identical parameters after the steps (gradient all-reduce + replicated clip / Adam)
and reports the training throughput (CUDA events, max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dp_nccl_check.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocrl_b200  # noqa: E402
from ocrl_b200 import dp, synth  # noqa: E402
from ocrl_b200.config import slate_config  # noqa: E402


def main():
    os.environ.setdefault("OCRL_KV_DTYPE", "bf16")
    rank, world, local = dp.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    per_gpu = int(os.environ.get("DP_BATCH", 64))
    steps, warm = int(os.environ.get("DP_STEPS", 6)), 2
    torch.manual_seed(10 + rank)  # different init per rank: make_data_parallel must broadcast rank 0's
    model = ocrl_b200.SLATE(*slate_config(num_slots=6, num_iterations=3, obs_size=64))
    model.to(dev)
    model.train()
    dp.make_data_parallel(model)
    frames = synth.to_obs(torch.from_numpy(synth.random_objs_frames(per_gpu * world, 64, seed=3))).to(dev)
    mine = dp.shard(frames, rank, world)
    torch.manual_seed(1000 + rank)  # slot noise / gumbel noise differ per rank (seed + rank)
    for i in range(warm):
        model.update(mine, None, 100 + i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        m = model.update(mine, None, 200 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    flat = torch.cat([p.detach().reshape(-1).float() for p in model._module.parameters()])
    digest = torch.stack([flat.sum(), flat.abs().sum(), (flat * flat).sum()])
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ref = digest.clone()
        dist.broadcast(ref, src=0)
        ok = torch.tensor([float(torch.equal(ref, digest))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        same = bool(ok.item())
    if rank == 0:
        print(json.dumps({"world": world, "batch_per_gpu": per_gpu, "ms_per_step": float(ms.item()),
                          "train_images_per_s": per_gpu * world / float(ms.item()) * 1e3,
                          "params_identical_across_ranks": same, "loss": float(m["loss"].detach()),
                          "grad_buckets": len(model._grad_reducer.buckets) if world > 1 else 0}), flush=True)
    assert same, "ranks diverged"
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Data parallelism for the OCR wrappers: one process per GPU, batch sharded by image.

Encode / rollout needs no communication (images are independent).  For OCR
training the only exchange is the gradient all-reduce: parameters are bucketed in reverse
construction order (the order autograd finishes them) and every bucket is all-reduced with NCCL
on a side stream as soon as its last gradient has been accumulated, so the exchange of the
decoder / slot-attention gradients runs under the rest of the backward (NVLink 5 / NVSwitch; NVLS
in-switch reduction when NCCL selects it).  ``clip_grad_norm_`` and Adam then run identically on
every rank (ocrs/base.py:60-74 semantics, loss is a per-rank batch mean so the average of the rank
gradients is the global-batch gradient).
"""
from __future__ import annotations

import os
from typing import List

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's environment; returns (rank, world, local_rank)."""
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard(batch: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous batch shard of this rank (mirrors run_sb3s.py's one-process-per-device layout)."""
    per = (batch.shape[0] + world - 1) // world
    return batch[rank * per:(rank + 1) * per]


class GradientReducer:
    """Bucketed, overlapped gradient averaging over the default process group."""

    def __init__(self, params: List[torch.nn.Parameter], bucket_bytes: int = 4 << 20):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in reversed(self.params):  # reverse construction order ~ order gradients become ready
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self._bucket_of = {id(p): i for i, b in enumerate(self.buckets) for p in b}
        self._pending = [0] * len(self.buckets)
        self._flat = [None] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        self._stream = None
        self._hooks = []
        # which parameters of a bucket produce a gradient is a property of the model (use_bcdec leaves the dVAE and the
        # transformer decoder untouched), not of the step: the per-bucket flag vectors and the all-rank verdict are
        # computed once and reused while this rank's pattern stays the same (re-checked every `recheck` steps), so a
        # steady-state step has no host <-> device synchronisation between backward and the optimizer
        self._pattern = [None] * len(self.buckets)
        self._flags = [None] * len(self.buckets)
        self._touched = [None] * len(self.buckets)
        self._steps = 0
        self.recheck = 256
        if self.world > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.reset()

    def reset(self):
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    def _on_grad(self, p):
        i = self._bucket_of[id(p)]
        if self._work[i] is not None:
            # the bucket is already in flight with the first backward's gradients: a second backward before finish()
            # (gradient accumulation) would be reduced partially
            raise RuntimeError("GradientReducer: backward() ran twice before finish(); call finish() (or update()) "
                               "after every backward, or accumulate micro-batches before the hooks fire")
        self._pending[i] -= 1
        if self._pending[i] == 0:
            self._launch(i)

    def _launch(self, i):
        bucket = self.buckets[i]
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in bucket]
        dev = grads[0].device
        # one flag per parameter rides at the end of the bucket: did THIS rank produce a gradient?  Parameters no rank
        # touched (e.g. the dVAE / transformer decoder when use_bcdec is on) keep grad = None, as on a single GPU, so
        # the optimizer state matches a 1-GPU run
        pattern = tuple(p.grad is not None for p in bucket)
        if pattern != self._pattern[i] or self._flags[i] is None or self._steps % self.recheck == 0:
            self._pattern[i] = pattern
            self._flags[i] = torch.tensor([1.0 if hit else 0.0 for hit in pattern], device=dev, dtype=grads[0].dtype)
            self._touched[i] = None  # decided again from the reduced flags
        flags = self._flags[i]
        if dev.type == "cuda":
            if self._stream is None:
                self._stream = torch.cuda.Stream(device=dev)
            self._stream.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self._stream):
                flat = torch.cat([g.reshape(-1) for g in grads] + [flags])
                self._work[i] = dist.all_reduce(flat, async_op=True)
        else:
            flat = torch.cat([g.reshape(-1) for g in grads] + [flags])
            self._work[i] = dist.all_reduce(flat, async_op=True)
        self._flat[i] = flat

    def finish(self):
        """Wait for every bucket, write the averaged gradients back (call after backward)."""
        if self.world == 1:
            return
        for i, bucket in enumerate(self.buckets):
            if self._work[i] is None:  # parameters that received no gradient this step
                self._launch(i)
        for i, bucket in enumerate(self.buckets):
            self._work[i].wait()
            flat = self._flat[i]
            if flat.is_cuda:
                torch.cuda.current_stream(flat.device).wait_stream(self._stream)
            if self._touched[i] is None:  # some rank produced a gradient for the parameter (host sync, first step only)
                self._touched[i] = (flat[-len(bucket):] > 0).tolist()
            flat.div_(self.world)
            off, dst, src = 0, [], []
            for p, hit in zip(bucket, self._touched[i]):
                n = p.numel()
                if hit:
                    if p.grad is None:
                        p.grad = torch.empty_like(p)
                    dst.append(p.grad)
                    src.append(flat[off:off + n].view_as(p))
                off += n
            if dst:
                torch._foreach_copy_(dst, src)  # one fused copy per bucket
        self._steps += 1
        self.reset()


def make_data_parallel(model, bucket_bytes: int = 4 << 20):
    """Turn an OCR wrapper (``ocrl_b200.SLATE``-style object with ``_module`` and ``update``) into
    its data-parallel version in place: parameters are broadcast from rank 0 and ``update`` averages
    gradients across ranks between backward and the gradient clip."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return model
    for t in list(model._module.parameters()) + list(model._module.buffers()):
        dist.broadcast(t.data, src=0)
    reducer = GradientReducer(list(model._module.parameters()), bucket_bytes)
    model._grad_reducer = reducer
    model._after_backward = reducer.finish
    return model

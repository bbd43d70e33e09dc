"""World-size-2 test of the data-parallel gradient exchange on CPU (gloo): the averaged rank
gradients equal the single-process gradient of the full batch, for the OCR wrapper's update()."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ocrl_b200 import dp
from ocrl_b200.config import to_namespace
from ocrl_b200.slate import Base


class _ToyModule(torch.nn.Module):
    rep_dim, num_slots = 8, 2

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(12, 16)
        self.b = torch.nn.Linear(16, 8)
        self.unused = torch.nn.Parameter(torch.zeros(3))

    def get_loss(self, obs, masks):
        out = self.b(torch.tanh(self.a(obs.flatten(1))))
        return {"loss": (out ** 2).sum() / obs.shape[0]}


class _ToyOCR(Base):
    def __init__(self):
        self._module = _ToyModule()
        cfg = to_namespace(dict(name="Toy", learning=dict(lr=1e-2, clip=0.05)))
        super().__init__(cfg, to_namespace(dict(obs_size=2, obs_channels=3)))

    def get_loss(self, obs, masks):
        return self._module.get_loss(obs, masks)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, obs, expect, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dp.init_from_env("gloo")
    torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
    model = _ToyOCR()
    dp.make_data_parallel(model, bucket_bytes=256)
    assert len(model._grad_reducer.buckets) > 1
    metrics = model.update(dp.shard(obs, rank, world), None, 0)
    got = {k: v.detach().clone() for k, v in model._module.state_dict().items()}
    ok = all(torch.allclose(got[k], expect[k], atol=1e-6) for k in expect)
    # a parameter no rank touched keeps grad None (and no Adam state), exactly as in the single-process run
    ok = ok and model._module.unused.grad is None and model._module.unused not in model._opt.state
    out[rank] = (ok, float(metrics["norm"]))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_update_equals_full_batch_update():
    torch.manual_seed(100)  # rank 0's seed: its weights are broadcast
    ref = _ToyOCR()
    obs = torch.rand(8, 3, 2, 2)
    ref.update(obs, None, 0)
    expect = {k: v.detach().clone() for k, v in ref._module.state_dict().items()}
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, obs, expect, out), nprocs=2, join=True)
    assert out[0][0] and out[1][0]
    assert out[0][1] == pytest.approx(out[1][1])


def test_shard_covers_batch():
    x = torch.arange(10).view(10, 1)
    parts = [dp.shard(x, r, 4) for r in range(4)]
    assert torch.equal(torch.cat(parts), x)

"""Host-side multi-process logic with the gloo backend, world_size 2 (no GPU): scene sharding, gradient
all-reduce through DDP, metric all-reduce -- the N>1 plumbing of run/distill.py on this repository's helpers."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from openscene_b200 import distill


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, _, w = distill.init_distributed('gloo')
    assert (r, w) == (rank, world)
    idx = distill.shard_indices(7, rank, world, epoch=3)
    torch.manual_seed(0)
    model = distill.wrap_ddp(torch.nn.Linear(4, 3))
    x = torch.full((5, 4), float(rank + 1))
    loss = distill.distill_loss(model(x), torch.ones(5, 3))
    loss.backward()
    g = model.module.weight.grad.clone()
    inter = torch.tensor([1.0 + rank, 2.0])
    distill.allreduce_sum(inter)
    ret[rank] = (idx, g, inter)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_world2_sharding_ddp_and_metric_allreduce():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    (i0, g0, m0), (i1, g1, m1) = ret[0], ret[1]
    assert len(i0) == len(i1) == 4 and sorted(set(i0 + i1)) == list(range(7))     # padded by wrapping, disjoint otherwise
    assert torch.allclose(g0, g1)                                                  # DDP averaged the gradients
    assert torch.equal(m0, torch.tensor([3.0, 4.0])) and torch.equal(m0, m1)
    # single process: averaged gradient equals the mean of the two per-rank gradients
    torch.manual_seed(0)
    lin = torch.nn.Linear(4, 3)
    gs = []
    for r in range(2):
        lin.zero_grad()
        distill.distill_loss(lin(torch.full((5, 4), float(r + 1))), torch.ones(5, 3)).backward()
        gs.append(lin.weight.grad.clone())
    assert torch.allclose(g0, (gs[0] + gs[1]) / 2, atol=1e-6)


def test_shard_indices_matches_distributed_sampler():
    from torch.utils.data.distributed import DistributedSampler
    ds = list(range(11))
    for rank in range(4):
        s = DistributedSampler(ds, num_replicas=4, rank=rank, shuffle=True, seed=5)
        s.set_epoch(2)
        assert list(iter(s)) == distill.shard_indices(11, rank, 4, epoch=2, shuffle=True, seed=5)


def test_shard_indices_with_fewer_items_than_half_the_ranks():
    """3 scenes on 8 ranks: the padded list must repeat (every rank gets one index, or DDP's all-reduce deadlocks)."""
    from torch.utils.data.distributed import DistributedSampler
    ds = list(range(3))
    for rank in range(8):
        s = DistributedSampler(ds, num_replicas=8, rank=rank, shuffle=True, seed=1)
        s.set_epoch(4)
        got = distill.shard_indices(3, rank, 8, epoch=4, shuffle=True, seed=1)
        assert len(got) == 1 and got == list(iter(s))


def test_poly_lr():
    assert abs(distill.poly_learning_rate(1e-4, 50, 100) - 1e-4 * 0.5 ** 0.9) < 1e-12

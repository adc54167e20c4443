"""CPU, world_size 2, gloo: sharded-gradient AdamW (prfl_b200.sharding) == torch.optim.AdamW on the rank-averaged
gradients, for a model with a `blocks` ModuleList (one shard unit per block + a root unit)."""
import os
import sys

import torch
import torch.multiprocessing as mp
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.blocks = nn.ModuleList([nn.Sequential(nn.Linear(7, 5), nn.Linear(5, 7)) for _ in range(3)])
        self.head = nn.Linear(7, 3)

    def forward(self, x):
        for b in self.blocks:
            x = x + b(x)
        return self.head(x)


def _data(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(11, 7, generator=g), torch.randn(11, 3, generator=g)


def _worker(rank, world, port, q, hooks=False):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from prfl_b200.sharding import ShardedAdamW, fsdp_units
    m = Toy()
    assert len(fsdp_units(m)) == 4
    opt = ShardedAdamW(m, lr=1e-2, weight_decay=0.01)
    if hooks:
        opt.attach_hooks()          # streaming mode: each unit is reduce-scattered and freed inside backward
    norms = []
    for _ in range(3):
        x, y = _data(rank)
        ((m(x) - y) ** 2).mean().backward()
        norms.append(float(opt.step(max_norm=1.0)))
        assert all(p.grad is None for p in m.parameters())
    q.put((rank, [p.detach().numpy().copy() for p in m.parameters()], norms))    # numpy: no fd passing through the queue
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("hooks", [False, True])
def test_sharded_adamw_matches_dense_adamw(hooks):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29647 + int(hooks), q, hooks)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda r: r[0])
    [p.join(60) for p in procs]
    # dense reference: average of the two ranks' gradients, global-norm clip, AdamW
    m = Toy()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-2, weight_decay=0.01)
    ref_norms = []
    for _ in range(3):
        opt.zero_grad()
        for r in range(world):
            x, y = _data(r)
            (((m(x) - y) ** 2).mean() / world).backward()
        ref_norms.append(float(torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)))
        opt.step()
    for r in range(world):
        for a, b in zip(res[r][1], m.parameters()):
            torch.testing.assert_close(torch.from_numpy(a), b.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(torch.tensor(res[r][2]), torch.tensor(ref_norms), rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------
# gradient accumulation: two backward() per step, a parameter that never receives a gradient inside a unit,
# clip_grad_norm_ on the accumulated gradients after EVERY micro-step, optimizer step at the boundary
# (train_prfl.py:822-830 with gradient_accumulation_steps = 2)
# ---------------------------------------------------------------------------------------------------
class ToyUnused(Toy):
    def __init__(self):
        super().__init__()
        self.blocks[1].register_parameter("unused", nn.Parameter(torch.randn(4, 4)))   # part of unit 1, never used in forward


def _worker_accum(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from prfl_b200.sharding import ShardedAdamW
    m = ToyUnused()
    opt = ShardedAdamW(m, lr=1e-2, weight_decay=0.01).attach_hooks()
    norms = []
    for step in range(3):
        for micro in range(2):
            g = torch.Generator().manual_seed(1000 * step + 10 * micro + rank)
            x, y = torch.randn(11, 7, generator=g), torch.randn(11, 3, generator=g)
            ((m(x) - y) ** 2).mean().backward()
            norms.append(float(opt.clip_grad_norm_(0.5)))
        opt.step()
    q.put((rank, [p.detach().numpy().copy() for p in m.parameters()], norms))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_accumulation_with_unused_parameter():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_accum, args=(r, world, 29651, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda r: r[0])
    [p.join(60) for p in procs]
    # dense reference with FSDP flat-parameter semantics: a parameter of a unit that received no gradient has a ZERO
    # gradient (it still decays), gradients are rank-averaged, clipping acts on the accumulated gradient every micro-step
    m = ToyUnused()
    opt = torch.optim.AdamW(m.parameters(), lr=1e-2, weight_decay=0.01)
    ref_norms = []
    for step in range(3):
        for p in m.parameters():
            p.grad = torch.zeros_like(p)
        for micro in range(2):
            for r in range(world):
                g = torch.Generator().manual_seed(1000 * step + 10 * micro + r)
                x, y = torch.randn(11, 7, generator=g), torch.randn(11, 3, generator=g)
                (((m(x) - y) ** 2).mean() / world).backward()
            ref_norms.append(float(torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)))
        opt.step()
    for r in range(world):
        for a, b in zip(res[r][1], m.parameters()):
            torch.testing.assert_close(torch.from_numpy(a), b.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(torch.tensor(res[r][2]), torch.tensor(ref_norms), rtol=1e-5, atol=1e-6)

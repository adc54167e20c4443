"""CPU, world_size 2, gloo: the Ulysses host logic of prfl_b200.parallel (layout algebra, autograd mirror,
all_gather with slice-backward, group bookkeeping) against the REFERENCE's own gloo run (tests/golden/a2a_gloo2.pt,
produced by diffusers_lite/utils/communication.py under gloo)."""
import os
import sys

import torch
import torch.multiprocessing as mp

from conftest import golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from prfl_b200 import parallel as P
    P.initialize_sequence_parallel_state(world)
    assert P.get_sequence_parallel_state() and P.nccl_info.sp_size == world and P.nccl_info.rank_within_group == rank
    g = torch.Generator().manual_seed(1234)
    full = torch.randn(1, 12 * world, 2 * world, 8, generator=g)
    mine = full.chunk(world, dim=1)[rank].clone().requires_grad_(True)
    out = P.all_to_all_4D(mine, scatter_dim=2, gather_dim=1)
    back = P.all_to_all_4D(out, scatter_dim=1, gather_dim=2)
    (out * (rank + 1)).sum().backward()
    x = mine.detach().clone().requires_grad_(True)
    gathered = P.all_gather(x, dim=1)
    (gathered * torch.arange(gathered.shape[1]).view(1, -1, 1, 1)).sum().backward()
    t = torch.full((3,), float(rank))
    P.broadcast(t)
    q.put((rank,) + tuple(v.detach().numpy().copy() for v in (out, back, mine.grad, gathered, x.grad, t)))   # numpy: no fd passing
    dist.barrier()
    dist.destroy_process_group()


def test_ulysses_gloo_world2_matches_reference():
    fx = golden("a2a_gloo2")
    world = fx["world"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29641, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in range(world)], key=lambda r: r[0])
    [p.join(60) for p in procs]
    for r in range(world):
        _, out, back, grad, gathered, xg, t = (res[r][0],) + tuple(torch.from_numpy(v) for v in res[r][1:])
        assert torch.equal(out, fx["out"][r])            # forward exchange == reference's
        assert torch.equal(back, fx["back"][r])          # round trip
        assert torch.equal(grad, fx["grad"][r])          # backward = swapped exchange (communication.py:143-152)
        assert torch.equal(gathered, fx["gathered"][r])  # all_gather + cat
        n = gathered.shape[1] // world                   # all_gather backward = own slice, no reduction (:248-260)
        exp = torch.arange(r * n, (r + 1) * n, dtype=torch.float32).view(1, -1, 1, 1).expand_as(xg)
        assert torch.equal(xg, exp)
        assert torch.equal(t, torch.zeros(3))            # broadcast from the group's first rank


def test_merge_partial_poolings_equals_global_softmax_pooling():
    """QueryAttention(sp_local=True): softmax pooling over the union of P token chunks from per-chunk (pooled, max, sum)."""
    import torch
    from prfl_b200.network import merge_partial_poolings
    g = torch.Generator().manual_seed(4)
    P, Ll, nh, C = 4, 37, 8, 48
    x = torch.randn(P * Ll, C, generator=g, dtype=torch.float64)
    wk = torch.randn(nh, C, generator=g, dtype=torch.float64)
    scores = x @ wk.t() * 3.0                                            # [L, nh], spread so the chunks' maxima differ
    want = torch.softmax(scores, 0).t() @ x                               # [nh, C]
    parts = []
    for r in range(P):
        s_r, x_r = scores[r * Ll:(r + 1) * Ll], x[r * Ll:(r + 1) * Ll]
        m = s_r.max(0).values
        e = torch.exp(s_r - m)
        parts.append(torch.cat([(e / e.sum(0)).t() @ x_r, m[:, None], e.sum(0)[:, None]], dim=1))
    torch.testing.assert_close(merge_partial_poolings(torch.stack(parts)), want, rtol=1e-12, atol=1e-12)

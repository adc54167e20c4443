"""CPU, gloo, world_size 4 and 2: the Ulysses x Ring schedule of `parallel.usp_attention` (SURVEY.md §8f row 4;
xdit_context_parallel.py:190-233 + xfuser's LongContextAttention) — group layout, all-to-all on the Ulysses sub-group,
K/V blocks travelling round the ring, LSE merge, ragged key length — against plain softmax attention over the gathered
sequence.  The attention of one (query block, key block) pair is injected as a torch function (the product uses the
tcgen05 kernel there; it has no CPU path), so this test covers exactly the host logic and the merge algebra."""
import math
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _attn_ref(q, k, v):
    """[Lq, H, d] x [Lk, H, d] -> (o [Lq, H, d] bf16, lse [H, Lq] fp32, natural log) — the contract of ops.attn_fwd(need_lse=True)."""
    s = torch.einsum("qhd,khd->hqk", q.float(), k.float()) / math.sqrt(q.shape[-1])
    lse = torch.logsumexp(s, dim=-1)
    o = torch.einsum("hqk,khd->qhd", torch.softmax(s, dim=-1), v.float())
    return o.to(q.dtype), lse


def _worker(rank, world, port, U, R, klen, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from prfl_b200 import parallel
    parallel.initialize_usp_state(U, R)
    ni = parallel.nccl_info
    assert (ni.sp_size, ni.ulysses_degree, ni.ring_degree) == (U * R, U, R)
    assert ni.rank_within_group == ni.ring_rank * U + ni.ulysses_rank            # chunk index = sp rank (use_ulysses_low layout)
    L, H, d = 96, 4, 128
    g = torch.Generator().manual_seed(5)
    Q, K, V = (torch.randn(L, H, d, generator=g).bfloat16() for _ in range(3))
    Lc = L // world
    sl = slice(rank * Lc, (rank + 1) * Lc)
    out = parallel.usp_attention(Q[sl].clone(), K[sl].clone(), V[sl].clone(), klen, attn_fn=_attn_ref)
    want, _ = _attn_ref(Q, K[:klen], V[:klen])
    err = float((out.float() - want[sl].float()).abs().max())
    q.put((rank, err, tuple(out.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,U,R,klen", [(4, 2, 2, 70), (4, 1, 4, 50), (2, 1, 2, 96)])
def test_usp_attention_schedule(world, U, R, klen):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29670 + world * 3 + U + R + (klen % 7)
    procs = [ctx.Process(target=_worker, args=(r, world, port, U, R, klen, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=180) for _ in range(world))
    [p.join(60) for p in procs]
    for rank, err, shape in res:
        assert shape == (96 // world, 4, 128)
        assert err <= 2e-2, (rank, err)          # bf16 rounding of the per-block outputs before the fp32 merge

"""Sequence-parallel (DeepSpeed-Ulysses) state and collectives with the reference's API
(diffusers_lite/utils/parallel_states.py:10-74, diffusers_lite/utils/communication.py:17-20,40-160,224-273).

Same names (`nccl_info`, `initialize_sequence_parallel_state`, `get_sequence_parallel_state`,
`all_to_all_4D`, `all_gather`, `broadcast`), same tensor contracts.  Differences, all on purpose:
  * payloads are bf16 (the reference ships q, k and the attention output as fp32);
  * one staging copy per exchange (a CUDA kernel writing/reading the [P, L/P, H/P, 128] wire layout)
    instead of reshape/transpose/contiguous on both sides;
  * no torch.cuda.synchronize() after the collective (communication.py:80,113).
The collective itself is NCCL `all_to_all_single` through torch.distributed (gloo on CPU in the tests).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


class COMM_INFO:
    def __init__(self):
        self.group = None
        self.sp_size = 1
        self.global_rank = 0
        self.rank_within_group = 0
        self.group_id = 0
        # Ulysses x Ring (USP): sp_size = ulysses_degree * ring_degree; ring_degree == 1 => plain Ulysses
        self.ulysses_group = None
        self.ring_group = None
        self.ulysses_degree = 1
        self.ring_degree = 1
        self.ulysses_rank = 0
        self.ring_rank = 0
        self.ring_ranks = ()          # global ranks of this rank's ring, in ring order


nccl_info = COMM_INFO()
_SEQUENCE_PARALLEL_STATE = False
_adopted = [False]            # True while the state above mirrors the reference's own bookkeeping (adopt_reference_state)


def initialize_sequence_parallel_state(sequence_parallel_size: int):
    """parallel_states.py:34-44."""
    global _SEQUENCE_PARALLEL_STATE
    _adopted[0] = False                                           # set natively from here on: not a mirror of the reference's module
    if sequence_parallel_size > 1:
        _SEQUENCE_PARALLEL_STATE = True
        initialize_sequence_parallel_group(sequence_parallel_size)
    else:
        _SEQUENCE_PARALLEL_STATE = False
        nccl_info.group = None
        nccl_info.sp_size = 1
        nccl_info.global_rank = int(os.getenv("RANK", "0"))
        nccl_info.rank_within_group = 0
        nccl_info.group_id = int(os.getenv("RANK", "0"))
    nccl_info.ulysses_group, nccl_info.ring_group = nccl_info.group, None
    nccl_info.ulysses_degree, nccl_info.ring_degree = nccl_info.sp_size, 1
    nccl_info.ulysses_rank, nccl_info.ring_rank, nccl_info.ring_ranks = nccl_info.rank_within_group, 0, ()


def initialize_usp_state(ulysses_degree: int, ring_degree: int):
    """xfuser's unified sequence parallelism as the reference's inference entry point sets it up
    (scripts/prfl/inference_prfl.py:71-82: `initialize_model_parallel(sequence_parallel_degree=world, ring_degree,
    ulysses_degree)`): the sp group of U x R ranks holds U*R contiguous token chunks (chunk index = rank in the sp
    group = ring_rank * U + ulysses_rank, xfuser's `use_ulysses_low` layout); Ulysses groups are the R runs of U
    consecutive ranks, ring groups the U strided sets {u, u + U, ...}.  Inside self-attention (xdit_context_parallel.py:
    190-233) an all-to-all over the Ulysses group turns [L/(UR), H] into [L/R, H/U], then the K/V blocks travel round
    the ring while every rank attends its L/R queries to each block and merges the partial results by their LSEs."""
    sp = ulysses_degree * ring_degree
    initialize_sequence_parallel_state(sp)
    nccl_info.ulysses_degree, nccl_info.ring_degree = ulysses_degree, ring_degree
    nccl_info.ulysses_group, nccl_info.ring_group = nccl_info.group, None
    nccl_info.ulysses_rank, nccl_info.ring_rank = nccl_info.rank_within_group, 0
    if sp <= 1 or ring_degree == 1:
        return
    rank = dist.get_rank()
    world = dist.get_world_size()
    for base in range(0, world, sp):
        for r in range(ring_degree):
            ranks = [base + r * ulysses_degree + u for u in range(ulysses_degree)]
            g = dist.new_group(ranks)
            if rank in ranks:
                nccl_info.ulysses_group, nccl_info.ulysses_rank = g, ranks.index(rank)
        for u in range(ulysses_degree):
            ranks = [base + u + r * ulysses_degree for r in range(ring_degree)]
            g = dist.new_group(ranks)
            if rank in ranks:
                nccl_info.ring_group, nccl_info.ring_rank, nccl_info.ring_ranks = g, ranks.index(rank), tuple(ranks)


def set_sequence_parallel_state(state: bool):
    global _SEQUENCE_PARALLEL_STATE
    _adopted[0] = False
    _SEQUENCE_PARALLEL_STATE = state


def get_sequence_parallel_state() -> bool:
    return _SEQUENCE_PARALLEL_STATE


def initialize_sequence_parallel_group(sequence_parallel_size: int):
    """parallel_states.py:56-74: contiguous ranks form one SP group."""
    rank = dist.get_rank() if dist.is_initialized() else int(os.getenv("RANK", "0"))
    world_size = dist.get_world_size() if dist.is_initialized() else int(os.getenv("WORLD_SIZE", "1"))
    assert world_size % sequence_parallel_size == 0, (
        f"world_size must be divisible by sequence_parallel_size, but got world_size: {world_size}, "
        f"sequence_parallel_size: {sequence_parallel_size}")
    nccl_info.sp_size = sequence_parallel_size
    nccl_info.global_rank = rank
    for i in range(world_size // sequence_parallel_size):
        ranks = range(i * sequence_parallel_size, (i + 1) * sequence_parallel_size)
        group = dist.new_group(ranks)
        if rank in ranks:
            nccl_info.group = group
            nccl_info.rank_within_group = rank - i * sequence_parallel_size
            nccl_info.group_id = i


def adopt_reference_state(ref_states=None) -> bool:
    """`plugin.install()` patches the blocks of a REFERENCE model, whose trainer initialised the reference's own bookkeeping
    (`diffusers_lite.utils.parallel_states`: `initialize_sequence_parallel_state`, train_prfl.py:119) — not this module's.  The
    patched blocks read THIS module's state, so mirror the reference's (flag, SP group, sizes, ranks) into it whenever the two
    differ; without this a patched model would silently run its self-attention without the Ulysses exchange.  Returns True if
    something was adopted.  No-op when the reference module is not loaded in this process."""
    global _SEQUENCE_PARALLEL_STATE
    import sys
    mod = ref_states if ref_states is not None else sys.modules.get("diffusers_lite.utils.parallel_states")
    if mod is None or mod is sys.modules.get(__name__):
        return False
    info = getattr(mod, "nccl_info", None)
    if info is None or not hasattr(mod, "get_sequence_parallel_state"):
        return False
    on = bool(mod.get_sequence_parallel_state()) and int(getattr(info, "sp_size", 1)) > 1
    if not on and not _adopted[0]:
        return False                                             # this module's state was set natively (class-swap route): not ours to undo
    same = (on == _SEQUENCE_PARALLEL_STATE) and (not on or (info.group is nccl_info.group and info.sp_size == nccl_info.sp_size and
                                                           info.rank_within_group == nccl_info.rank_within_group))
    if same:
        return False
    _adopted[0] = on
    _p2p_cache.clear()                                           # peer buffers belong to the previous group
    _SEQUENCE_PARALLEL_STATE = on
    nccl_info.group = info.group if on else None
    nccl_info.sp_size = int(info.sp_size) if on else 1
    nccl_info.global_rank = int(info.global_rank)
    nccl_info.rank_within_group = int(info.rank_within_group) if on else 0
    nccl_info.group_id = int(info.group_id)
    nccl_info.ulysses_group, nccl_info.ring_group = nccl_info.group, None
    nccl_info.ulysses_degree, nccl_info.ring_degree = nccl_info.sp_size, 1
    nccl_info.ulysses_rank, nccl_info.ring_rank, nccl_info.ring_ranks = nccl_info.rank_within_group, 0, ()
    return True


def destroy_sequence_parallel_group():
    """parallel_states.py:127-129 (+ the peer-mapped exchange buffers of this group are released first)."""
    global _SEQUENCE_PARALLEL_STATE
    _p2p_cache.clear()
    _SEQUENCE_PARALLEL_STATE = False
    dist.destroy_process_group()


def broadcast(input_: torch.Tensor):
    """communication.py:17-19."""
    src = nccl_info.group_id * nccl_info.sp_size
    dist.broadcast(input_, src=src, group=nccl_info.group)


# ------------------------------------------------------------------------------------------------
# layout algebra shared by the CUDA path and the CPU (gloo) tests
# ------------------------------------------------------------------------------------------------
def _pack_heads(x: torch.Tensor, P: int) -> torch.Tensor:
    """[L_loc, H, d] -> wire layout [P, L_loc, H/P, d]  (destination rank major)."""
    L, H, d = x.shape
    if x.is_cuda:
        from . import ops
        packed = torch.empty(P, L, H // P, d, dtype=x.dtype, device=x.device)
        return ops.a2a_pack(x, packed, P)
    return x.reshape(L, P, H // P, d).permute(1, 0, 2, 3).contiguous()


def _unpack_heads(packed: torch.Tensor, P: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """wire layout [P, L_loc, H/P, d] (source rank major) -> [L_loc, H, d] (`out`: a strided [L_loc, H, d] destination)."""
    _, L, Hl, d = packed.shape
    if packed.is_cuda:
        from . import ops
        if out is None:
            out = torch.empty(L, P * Hl, d, dtype=packed.dtype, device=packed.device)
        return ops.a2a_pack(out, packed, P, unpack=True)
    res = packed.permute(1, 0, 2, 3).reshape(L, P * Hl, d).contiguous()
    if out is not None:
        out.copy_(res)
        return out
    return res


def _a2a(buf: torch.Tensor, group=None) -> torch.Tensor:
    out = torch.empty_like(buf)
    dist.all_to_all_single(out, buf, group=nccl_info.group if group is None else group)
    return out


def ulysses_scatter_tokens(x: torch.Tensor, P: int, group=None) -> torch.Tensor:
    """scatter heads / gather tokens: local [L/P, H, d] -> [L, H/P, d] (communication.py:60-89).
    The receive buffer [P, L/P, H/P, d] already is [L, H/P, d] in global token order."""
    L, H, d = x.shape
    return _a2a(_pack_heads(x, P), group).view(P * L, H // P, d)


def ulysses_gather_tokens(x: torch.Tensor, P: int, out: Optional[torch.Tensor] = None, group=None) -> torch.Tensor:
    """scatter tokens / gather heads: [L, H/P, d] -> local [L/P, H, d] (communication.py:91-123).
    The send buffer is x itself ([P, L/P, H/P, d] by token chunk).  `out`: optional strided destination view."""
    L, Hl, d = x.shape
    recv = _a2a(x.contiguous().view(P, L // P, Hl, d), group)
    return _unpack_heads(recv, P, out)


# ------------------------------------------------------------------------------------------------
# Ulysses x Ring attention (no-grad; the reference uses it for inference only, text2video.py:137-148)
# ------------------------------------------------------------------------------------------------
def _merge_host(o_acc, lse_acc, o_new, lse_new, first):
    """Host-logic twin of prfl_attn_merge for CPU tensors (gloo tests): same algebra in torch ops."""
    if first:
        o_acc.copy_(o_new.float())
        lse_acc.copy_(lse_new)
        return
    m = torch.maximum(lse_acc, lse_new)
    wa, wn = torch.exp(lse_acc - m), torch.exp(lse_new - m)
    inv = 1.0 / (wa + wn)
    o_acc.mul_((wa * inv).t().unsqueeze(-1)).add_(o_new.float() * (wn * inv).t().unsqueeze(-1))
    lse_acc.copy_(m + torch.log(wa + wn))


def usp_attention(q3: torch.Tensor, k3: torch.Tensor, v3: torch.Tensor, klen: int, attn_fn=None) -> torch.Tensor:
    """Self-attention of one sample under Ulysses x Ring.  q3 / k3 / v3: this rank's [L/(UR), H, 128] bf16 token chunk
    (q, k already normalised + rotated at their global positions); klen: number of valid keys of the sample.  Returns
    this rank's [L/(UR), H, 128] attention output.

      1. all-to-all over the Ulysses group: [L/(UR), H] -> [L/R, H/U]            (communication.py:60-89 on the sub-group)
      2. R ring steps: attend the local L/R queries to the K/V block currently held (`prfl_attn_fwd` with LSE), fold it
         into the running result (`prfl_attn_merge`), pass the block to the next rank of the ring while the attention
         over it runs (isend / irecv on the ring group; blocks past `klen` are skipped, a partially valid block is cut)
      3. all-to-all back: [L/R, H/U] -> [L/(UR), H]
    `attn_fn(q, k, v) -> (o, lse)` is injectable for the CPU tests of the schedule; on the GPU it is ops.attn_fwd."""
    U, R = nccl_info.ulysses_degree, nccl_info.ring_degree
    rr = nccl_info.ring_rank
    if attn_fn is None:
        from . import ops
        attn_fn = lambda q, k, v: ops.attn_fwd(q, k, v, need_lse=True)
    ug = nccl_info.ulysses_group
    qg, kg, vg = ((ulysses_scatter_tokens(t, U, ug) if U > 1 else t.contiguous()) for t in (q3, k3, v3))     # [L/R, H/U, d]
    Lr, Hl, d = qg.shape
    o_acc = torch.empty(Lr, Hl, d, dtype=torch.float32, device=qg.device)
    lse_acc = torch.empty(Hl, Lr, dtype=torch.float32, device=qg.device)
    out = torch.empty(Lr, Hl, d, dtype=qg.dtype, device=qg.device)
    kv_cur = torch.stack([kg, vg])                                     # one message per ring step
    nxt, prv = (nccl_info.ring_ranks[(rr + 1) % R], nccl_info.ring_ranks[(rr - 1) % R]) if R > 1 else (None, None)
    first = True
    for s in range(R):
        reqs, kv_next = [], None
        if s + 1 < R:
            kv_next = torch.empty_like(kv_cur)
            ops_ = [dist.P2POp(dist.isend, kv_cur, nxt, group=nccl_info.ring_group),
                    dist.P2POp(dist.irecv, kv_next, prv, group=nccl_info.ring_group)]
            reqs = dist.batch_isend_irecv(ops_)
        blk = (rr - s) % R                                            # which token block the held K/V belongs to
        valid = min(max(klen - blk * Lr, 0), Lr)
        if valid > 0:
            o_s, lse_s = attn_fn(qg, kv_cur[0][:valid], kv_cur[1][:valid])
            last = all(min(max(klen - ((rr - t) % R) * Lr, 0), Lr) == 0 for t in range(s + 1, R))
            if qg.is_cuda:
                from . import ops
                ops.attn_merge_(o_acc, lse_acc, o_s, lse_s, first, out if last else None)
            else:
                _merge_host(o_acc, lse_acc, o_s, lse_s, first)
                if last:
                    out.copy_(o_acc.to(out.dtype))
            first = False
        for r_ in reqs:
            r_.wait()
        if kv_next is not None:
            kv_cur = kv_next
    assert not first, "no valid keys"
    return ulysses_gather_tokens(out, U, group=ug) if U > 1 else out


class SeqAllToAll4D(torch.autograd.Function):
    """communication.py:128-152: backward = the exchange with scatter/gather swapped."""

    @staticmethod
    def forward(ctx, group, input_, scatter_idx, gather_idx):
        ctx.scatter_idx, ctx.gather_idx = scatter_idx, gather_idx
        return _all_to_all_4D(input_, scatter_idx, gather_idx)

    @staticmethod
    def backward(ctx, grad_output):
        return None, _all_to_all_4D(grad_output, ctx.gather_idx, ctx.scatter_idx), None, None


def _all_to_all_4D(input_: torch.Tensor, scatter_idx: int, gather_idx: int) -> torch.Tensor:
    assert input_.dim() == 4, f"input must be 4D tensor, got {input_.dim()} and shape {input_.shape}"
    P = nccl_info.sp_size
    if scatter_idx == 2 and gather_idx == 1:
        return torch.stack([ulysses_scatter_tokens(input_[b], P) for b in range(input_.shape[0])])
    if scatter_idx == 1 and gather_idx == 2:
        return torch.stack([ulysses_gather_tokens(input_[b], P) for b in range(input_.shape[0])])
    raise RuntimeError("scatter_idx must be 1 or 2 and gather_idx must be 1 or 2")


def all_to_all_4D(input_: torch.Tensor, scatter_dim: int = 2, gather_dim: int = 1):
    """communication.py:155-160.  [b, L/P, H, d] <-> [b, L, H/P, d]."""
    return SeqAllToAll4D.apply(nccl_info.group, input_, scatter_dim, gather_dim)


class _AllGather(torch.autograd.Function):
    """communication.py:224-260: forward all_gather + cat, backward = own slice (no reduction)."""

    @staticmethod
    def forward(ctx, input_, dim):
        ctx.dim = dim
        ctx.input_size = input_.size(dim)
        world = nccl_info.sp_size
        input_ = input_.contiguous()
        if dim == 0 or all(s == 1 for s in input_.shape[:dim]):
            out = torch.empty((world,) + tuple(input_.shape), dtype=input_.dtype, device=input_.device)
            dist.all_gather_into_tensor(out.view(-1), input_.view(-1), group=nccl_info.group)
            shp = list(input_.shape)
            shp[dim] *= world
            return out.reshape(shp)     # leading dims are all 1 => the rank-major buffer IS the concatenation
        parts = [torch.empty_like(input_) for _ in range(world)]
        dist.all_gather(parts, input_, group=nccl_info.group)
        return torch.cat(parts, dim=dim)

    @staticmethod
    def backward(ctx, grad_output):
        rank = nccl_info.rank_within_group
        return grad_output.narrow(ctx.dim, rank * ctx.input_size, ctx.input_size).contiguous(), None


def all_gather(input_: torch.Tensor, dim: int = 1):
    """communication.py:263-273."""
    return _AllGather.apply(input_, dim)


# ------------------------------------------------------------------------------------------------
# NCCL-free Ulysses exchange for the no-grad forward: peer stores over NVLink into symmetric buffers
# ------------------------------------------------------------------------------------------------
class P2PUlysses:
    """Symmetric (peer-mapped) receive buffers + the two fused exchanges of one self-attention:
         q/k/v : prfl_a2a_scatter_p2p — every rank stores its token chunk's heads straight into the owners' buffers
         out   : prfl_attn_fwd_p2p    — the attention epilogue stores each query row straight into its home rank
    with one cross-rank barrier after each (torch symmetric-memory signal pads).  Replaces 4 NCCL all-to-alls, the pack
    and the unpack copies per block (communication.py:40-160 does 4 all_to_all_single + 8 transposes + 4 device syncs).
    The training path (recompute-forward + backward of a checkpointed block) uses the same buffers un-fused:
    scatter_qkv / gather_out / scatter_grad / gather_grads below."""

    def __init__(self, L: int, H: int, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.P, self.rank = nccl_info.sp_size, nccl_info.rank_within_group
        self.L, self.H = L, H
        # Set-up happens with the device DRAINED on every rank and is followed by a group-wide barrier (the unanimity
        # all-reduce of get_p2p_ulysses): a peer's first signal must not be able to reach this rank's signal pads before their
        # (stream-ordered) initialisation has executed here.  The one hang of round 2 (profiles/r02_bench_n4_deadlock.log)
        # froze the devices right where the gradient buffer used to be created lazily in the middle of a backward — ~1000
        # queued launches behind the host, and the only symmetric buffer whose creation was NOT followed by that barrier.
        torch.cuda.synchronize(device)
        self.L_loc, self.Hl = L // self.P, H // self.P
        self.qkv = symm_mem.empty(4, L, self.Hl, 128, dtype=torch.bfloat16, device=device)      # slab 3: dO in the backward
        self.o = symm_mem.empty(self.L_loc, H, 128, dtype=torch.bfloat16, device=device)
        self.h_qkv = symm_mem.rendezvous(self.qkv, nccl_info.group)
        self.h_o = symm_mem.rendezvous(self.o, nccl_info.group)
        self.qkv_ptrs = [int(p) for p in self.h_qkv.buffer_ptrs]
        self.o_ptrs = [int(p) for p in self.h_o.buffer_ptrs]
        self.slab = L * self.Hl * 128 * 2
        # [L_loc, 3*H*128] fused gradient buffer of the training path.  Created here, with the other buffers, while every rank
        # is at the same quiescent point (the rendezvous is a host-side collective): not lazily in the middle of a backward.
        self.dqkv = symm_mem.empty(self.L_loc, 3 * H * 128, dtype=torch.bfloat16, device=device)
        self.h_dqkv = symm_mem.rendezvous(self.dqkv, nccl_info.group)
        self.dqkv_ptrs = [int(p) for p in self.h_dqkv.buffer_ptrs]
        # every initialisation enqueued above has executed here before this rank joins the group-wide agreement in
        # get_p2p_ulysses (an all-reduce + host read = a barrier): no rank can launch its first exchange earlier
        torch.cuda.synchronize(device)

    def attention(self, q3: torch.Tensor, k3: torch.Tensor, v3: torch.Tensor, klen: int) -> torch.Tensor:
        """q3/k3/v3: local [L/P, H, 128] bf16 views.  Returns this rank's [L/P, H, 128] attention output (a view of the
        symmetric buffer: consume it before the next call)."""
        from . import ops
        for j, t in enumerate((q3, k3, v3)):
            ops.a2a_scatter_p2p(t, [p + j * self.slab for p in self.qkv_ptrs], self.P, self.rank)
        self.h_qkv.barrier(channel=0)
        ops.attn_fwd_p2p(self.qkv[0], self.qkv[1][:klen], self.qkv[2][:klen], self.o_ptrs, self.L_loc, self.rank * self.Hl, self.H)
        self.h_o.barrier(channel=1)
        return self.o


    # ---- training path (recompute-forward + backward of a checkpointed block): same peer stores, nothing fused into the
    #      attention kernels because the backward needs q / k / v / o in the attention layout on this rank ----------------
    def scatter_qkv(self, q3: torch.Tensor, k3: torch.Tensor, v3: torch.Tensor):
        """local [L/P, H, 128] views -> this rank's [L, H/P, 128] q, k, v (views of the symmetric buffer, valid until the next
        scatter; the backward of the same block runs before that)."""
        from . import ops
        for j, t in enumerate((q3, k3, v3)):
            ops.a2a_scatter_p2p(t, [p + j * self.slab for p in self.qkv_ptrs], self.P, self.rank)
        self.h_qkv.barrier(channel=0)
        return self.qkv[0], self.qkv[1], self.qkv[2]

    def gather_out(self, og: torch.Tensor) -> torch.Tensor:
        """attention output [L, H/P, 128] of this rank's heads -> this rank's tokens [L/P, H, 128] (the symmetric buffer)."""
        from . import ops
        ops.a2a_gather_p2p(og, self.o_ptrs, self.H * 128, 128, self.P, self.rank)
        self.h_o.barrier(channel=1)
        return self.o

    def scatter_grad(self, do3: torch.Tensor) -> torch.Tensor:
        """dL/d(attention output) of this rank's tokens [L/P, H, 128] -> [L, H/P, 128] (slab 3 of the symmetric buffer)."""
        from . import ops
        ops.a2a_scatter_p2p(do3, [p + 3 * self.slab for p in self.qkv_ptrs], self.P, self.rank)
        self.h_qkv.barrier(channel=0)
        return self.qkv[3]

    def gather_grads(self, dqg: torch.Tensor, dkg: torch.Tensor, dvg: torch.Tensor) -> torch.Tensor:
        """dq, dk, dv [L, H/P, 128] of this rank's heads -> the fused [L/P, 3*H*128] gradient of this rank's tokens, written by
        the peers straight into the column blocks of a symmetric buffer (no unpack, no concatenation)."""
        from . import ops
        C = self.H * 128
        for j, t in enumerate((dqg, dkg, dvg)):
            ops.a2a_gather_p2p(t, [p + j * C * 2 for p in self.dqkv_ptrs], 3 * C, 128, self.P, self.rank)
        self.h_dqkv.barrier(channel=0)
        return self.dqkv


_p2p_cache = {}
_p2p_disabled = os.environ.get("PRFL_ULYSSES", "p2p").lower() == "nccl"


def get_p2p_ulysses(L: int, H: int, device) -> Optional[P2PUlysses]:
    """The peer-store exchange if symmetric memory can be set up on this box, else None (=> NCCL all-to-all path)."""
    global _p2p_disabled
    if _p2p_disabled or not get_sequence_parallel_state() or nccl_info.sp_size > 8:
        return None
    key = (L, H, str(device))
    if key not in _p2p_cache:
        obj, err = None, None
        try:
            obj = P2PUlysses(L, H, device)
        except Exception as e:  # no fabric / no peer access / out of memory on THIS rank
            err = e
        # the choice must be unanimous: a rank that fell back to NCCL while its peers wait in a symmetric-memory barrier
        # would deadlock the group, so agree on the outcome (MIN over the SP group) before committing to either path.  The
        # all-reduce + host read is also the barrier that ends the set-up: every rank has finished (and synchronised) its
        # P2PUlysses.__init__ before any rank can launch its first exchange into a peer's buffers / signal pads
        ok = torch.tensor([0 if obj is None else 1], dtype=torch.int32, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=nccl_info.group)
        if int(ok) == 0:
            import warnings
            why = f"{type(err).__name__}: {err}" if err is not None else "a peer rank could not set it up"
            warnings.warn(f"prfl_b200: symmetric-memory Ulysses unavailable ({why}); using NCCL all-to-all on every rank")
            del obj                                   # frees this rank's partially / fully created symmetric buffers
            _p2p_disabled = True
            return None
        _p2p_cache[key] = obj
    return _p2p_cache[key]

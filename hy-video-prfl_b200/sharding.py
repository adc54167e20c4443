"""Sharded training state for the trainable Wan-DiT (the FSDP FULL_SHARD role in the reference,
diffusers_lite/utils/fsdp_utils.py:66-122 + train_prfl.py:346-362, 482-491, 822-830), re-cut for 180 GB B200s.

The reference shards fp32 parameters, gradients and AdamW state over ALL ranks (SP ranks included), all-gathers every
block's fp32 parameters twice per step (1.41 GB per block per direction) and lets autocast re-cast them to bf16 on
every use.  Here, per FSDP unit (one per WanAttentionBlock, `_no_split_modules`, + one root unit):

  * the block's MATRICES live in one flat **bf16** buffer that stays resident on every rank (28 GB for the 14B model):
    the nn.Parameters are views into it, the GEMM operands (fused [3C, C] QKV, [2C, C] context K/V, ...) are views of
    the same bytes — no operand caches, no fp32 replica, no per-block parameter all-gather in forward / backward;
  * fp32 MASTER weights and both AdamW moments exist only as 1/W shards;
  * weight gradients are written by the wgrad GEMMs straight into a flat fp32 staging buffer (two buffers, shared by
    all units) through a *gradient sink* (`engine.BlockFn` bypasses autograd's per-parameter .grad), REDUCE-SCATTERED
    (fp32, averaged over the group exactly as FSDP does — under Ulysses every rank holds the partial gradient of its
    token chunk) on a side stream while the next block's backward runs, and accumulated
    into a 1/W fp32 gradient shard (gradient accumulation over micro-batches = more adds into the same shard);
  * `step()` = one `prfl_adamw_step` launch per unit on the shards (clip coefficient read from device memory, the bf16
    compute copy of the updated slice written by the same kernel) + an in-place **bf16** all-gather into the resident
    buffer on the side stream;
  * everything that is not a block matrix (biases, norm weights, modulation, embeddings, head: < 2 % of the
    parameters) forms the ROOT unit: a flat replicated **fp32** buffer (those parameters are consumed in fp32 by the
    reference too), fed by autograd hooks and by the sink, reduce-scattered once per flush, updated in place on each
    rank's slice and all-gathered in fp32.

`clip_grad_norm_()` and `step()` are separate calls so the reference's gradient-accumulation pattern — clip the
accumulated gradients on EVERY micro-step, optimizer.step() at accumulation boundaries (train_prfl.py:822-830) — is
reproducible; `step(max_norm=...)` fuses the two for accumulation == 1.

Models that are not resident (CPU toys in the gloo tests, `resident_bf16=False`) use flat fp32 units only: one per
block + root.  Collectives are NCCL through torch.distributed; on backends without reduce-scatter (gloo) the same
result is produced with all_reduce + slice.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

# order of a block's matrices inside its resident buffer: operands that are consumed fused are adjacent
MATRIX_ORDER = ("self_attn.q.weight", "self_attn.k.weight", "self_attn.v.weight", "self_attn.o.weight",
                "cross_attn.q.weight", "cross_attn.k.weight", "cross_attn.v.weight",
                "cross_attn.k_img.weight", "cross_attn.v_img.weight", "cross_attn.o.weight",
                "ffn.0.weight", "ffn.2.weight")


# Test hook (tests/test_host_logic_emulated_cpu.py): lets the resident layout be built from CPU tensors so that its host logic
# (views, gradient sink, staging) can run over the emulated kernels.  It opens no CPU path: every kernel wrapper in ops.py
# still refuses CPU tensors, so with the real library a CPU-resident model fails at its first op.
_ALLOW_CPU_UNITS = False


def _dist_info(group) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def fsdp_units(model: nn.Module) -> List[List[nn.Parameter]]:
    """One unit per block (load.py:6-12 wraps WanAttentionBlock) + a root unit with everything else."""
    units, seen = [], set()
    for blk in getattr(model, "blocks", []):
        ps = [p for p in blk.parameters() if p.requires_grad]
        seen.update(id(p) for p in ps)
        if ps:
            units.append(ps)
    root = [p for p in model.parameters() if p.requires_grad and id(p) not in seen]
    if root:
        units.append(root)
    return units


# ------------------------------------------------------------------------------------------------
# units
# ------------------------------------------------------------------------------------------------
class ResidentUnit:
    """The matrices of one WanAttentionBlock in a flat bf16 buffer (parameters re-pointed to views of it) and, if
    trainable, this rank's 1/W fp32 master shard."""
    kind = "resident"

    def __init__(self, blk: nn.Module, world: int = 1, rank: int = 0, trainable: bool = True):
        named = dict(blk.named_parameters())
        mats = [(n, named[n]) for n in MATRIX_ORDER if n in named]
        assert mats, "not a WanAttentionBlock"
        dev = mats[0][1].device
        assert dev.type == "cuda" or _ALLOW_CPU_UNITS, "resident bf16 units live on the GPU (there is no CPU path)"
        self.world, self.rank = world, rank
        self.offsets: Dict[str, Tuple[int, int, torch.Size]] = {}
        o = 0
        for n, p in mats:
            assert p.dim() == 2 and p.numel() % 8 == 0
            self.offsets[n] = (o, p.numel(), p.shape)
            o += p.numel()
        self.n = o
        self.pad = (-o) % (8 * world)
        self.shard = (self.n + self.pad) // world
        self.wflat = torch.empty(self.n + self.pad, dtype=torch.bfloat16, device=dev)
        if self.pad:
            self.wflat[self.n:].zero_()
        self.master = torch.zeros(self.shard, dtype=torch.float32, device=dev) if trainable else None
        lo, hi = rank * self.shard, (rank + 1) * self.shard
        self.params = []
        for n, p in mats:                                  # one matrix at a time: its fp32 storage is freed as soon as re-pointed
            o, cnt, shp = self.offsets[n]
            src = p.data.reshape(-1)
            self.wflat[o:o + cnt].copy_(src)
            if trainable:
                a, b = max(o, lo), min(o + cnt, hi)
                if a < b:
                    self.master[a - lo:b - lo].copy_(src[a - o:b - o])
            p.data = self.wflat[o:o + cnt].view(shp)
            self.params.append(p)
        self.m = self.v = self.gshard = None
        self.t, self.has_grad = 0, False
        self.sink = None
        blk.__dict__["_prfl_unit"] = self

    def my_slice(self, t: torch.Tensor) -> torch.Tensor:
        return t[self.rank * self.shard:(self.rank + 1) * self.shard]

    def view(self, buf: torch.Tensor, names: Sequence[str]) -> torch.Tensor:
        """[sum rows, K] view of `buf` (same layout as wflat) covering the adjacent matrices `names`."""
        o0, _, shp0 = self.offsets[names[0]]
        rows, o = 0, o0
        for n in names:
            oo, cnt, shp = self.offsets[n]
            assert oo == o and shp[1] == shp0[1], "matrices are not adjacent in the resident layout"
            rows, o = rows + shp[0], o + cnt
        return buf[o0:o].view(rows, shp0[1])


class FlatUnit:
    """A set of fp32 parameters re-pointed into one flat replicated fp32 buffer; each rank's slice of it is that rank's
    master shard (updated in place, all-gathered in place)."""
    kind = "flat"

    def __init__(self, named: Sequence[Tuple[str, nn.Parameter]], world: int = 1, rank: int = 0):
        assert named
        dev = named[0][1].device
        self.world, self.rank = world, rank
        self.offsets: Dict[str, Tuple[int, int, torch.Size]] = {}
        o = 0
        for n, p in named:
            self.offsets[n] = (o, p.numel(), p.shape)
            o += p.numel() + ((-p.numel()) % 4)             # keep every parameter 16-byte aligned
        self.n = o
        self.pad = (-o) % (4 * world)
        self.shard = (self.n + self.pad) // world
        self.flat = torch.zeros(self.n + self.pad, dtype=torch.float32, device=dev)
        self.params, self.by_id = [], {}
        for n, p in named:
            o, cnt, shp = self.offsets[n]
            self.flat[o:o + cnt].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + cnt].view(shp)
            self.params.append(p)
            self.by_id[id(p)] = (o, cnt)
        self.master = self.my_slice(self.flat)
        self.m = self.v = self.gshard = self.gflat = None
        self.t, self.has_grad, self.dirty = 0, False, False
        self.sink = None

    def my_slice(self, t: torch.Tensor) -> torch.Tensor:
        return t[self.rank * self.shard:(self.rank + 1) * self.shard]

    def gbuf(self) -> torch.Tensor:
        if self.gflat is None:
            self.gflat = torch.zeros_like(self.flat)
        return self.gflat

    def add_grad(self, key, grad: torch.Tensor):
        o, cnt = self.by_id[key] if not isinstance(key, str) else self.offsets[key][:2]
        self.gbuf()[o:o + cnt].add_(grad.reshape(-1))
        self.dirty = True


class _BlockSink:
    """What `engine.BlockFn.backward` writes a resident block's weight gradients into (instead of returning them to
    autograd): fp32 [rows, K] views of a flat staging buffer for the matrices, adds into the root unit for the rest."""

    def __init__(self, opt: "ShardedAdamW", unit: ResidentUnit, prefix: str):
        self.opt, self.unit, self.prefix = opt, unit, prefix
        self.buf = None
        self.written = set()
        self.slot = 0

    def begin(self):
        self.buf, self.slot = self.opt._stage_acquire(self.unit)
        if self.unit.world > 1 or not self.unit.has_grad:
            self.written = set()

    def matrix_out(self, names: Sequence[str]) -> Tuple[torch.Tensor, bool]:
        """(fp32 [rows, K] view to write dW into, beta): beta = accumulate onto what is already there."""
        beta = names[0] in self.written
        self.written.update(names)
        return self.unit.view(self.buf, names), beta

    def small(self, name: str, grad: torch.Tensor):
        self.opt.root.add_grad(self.prefix + name, grad)

    def end(self):
        missing = [n for n in self.unit.offsets if n not in self.written]
        for n in missing:                                     # never expected on the Wan path; keeps stale staging data out
            self.unit.view(self.buf, [n]).zero_()
            self.written.add(n)
        self.opt._stage_release(self.unit, self.buf, self.slot)
        self.buf = None


# ------------------------------------------------------------------------------------------------
# residency helpers
# ------------------------------------------------------------------------------------------------
def make_resident(model: nn.Module, trainable: bool = False, group=None) -> nn.Module:
    """Convert every WanAttentionBlock of `model` (already on its CUDA device) to the resident bf16 layout.  With
    trainable=False (a frozen model, e.g. PRFL's reward transformer) no master shard is kept: weights = bf16(fp32
    checkpoint), which is what the reference's autocast feeds its GEMMs."""
    from .model import WanAttentionBlock, bump_weight_epoch
    world, rank = _dist_info(group)
    for blk in getattr(model, "blocks", []):
        if isinstance(blk, WanAttentionBlock) and "_prfl_unit" not in blk.__dict__:
            ResidentUnit(blk, world, rank, trainable)
    bump_weight_epoch()
    return model


def build_wan(model_type: str, in_dim: int, num_layers: int, device, *, dim: int = 5120, ffn_dim: int = 13824,
              num_heads: int = 40, resident_bf16: bool = True, head: bool = True, frozen: bool = False, group=None,
              **kw) -> nn.Module:
    """Random-init WanModel (model.py:413-729 initialisation) built block by block, so that with resident_bf16 the fp32
    initial values never occupy more than one block (1.6 GB) beyond the resident copies: a 40-block 14B model comes up
    in 28 GB (+ 1/W fp32 master shards if trainable) instead of 56 GB fp32 + 28 GB of operand caches."""
    from .model import WanAttentionBlock, WanModel, bump_weight_epoch
    world, rank = _dist_info(group)
    with torch.device(device):
        m = WanModel(model_type=model_type, in_dim=in_dim, dim=dim, ffn_dim=ffn_dim, num_heads=num_heads, num_layers=0, **kw)
        if not head:
            m.head = None
        ca = "t2v_cross_attn" if model_type == "t2v" else "i2v_cross_attn"
        for _ in range(num_layers):
            blk = WanAttentionBlock(ca, dim, ffn_dim, num_heads, m.window_size, m.qk_norm, m.cross_attn_norm, m.eps)
            for mod in blk.modules():                          # WanModel.init_weights (model.py:707-716) for this block
                if isinstance(mod, nn.Linear):
                    nn.init.xavier_uniform_(mod.weight)
                    if mod.bias is not None:
                        nn.init.zeros_(mod.bias)
            if resident_bf16:
                ResidentUnit(blk, world, rank, trainable=not frozen)
            m.blocks.append(blk)
    m.num_layers = m.config["num_layers"] = num_layers
    bump_weight_epoch()
    return m


# ------------------------------------------------------------------------------------------------
# the optimizer
# ------------------------------------------------------------------------------------------------
class ShardedAdamW:
    """AdamW over reduce-scattered gradient shards with sharded fp32 masters and moments; see the module docstring."""

    def __init__(self, model: nn.Module, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, group=None,
                 average: bool = True, resident_bf16: Optional[bool] = None):
        self.group = group
        self.world, self.rank = _dist_info(group)
        self.lr, self.betas, self.eps, self.wd, self.average = lr, betas, eps, weight_decay, average
        self.model = model
        blocks = list(getattr(model, "blocks", []))
        try:
            first = next(model.parameters())
        except StopIteration:
            raise ValueError("model has no parameters")
        from .model import WanAttentionBlock
        wan = len(blocks) > 0 and all(isinstance(b, WanAttentionBlock) for b in blocks) and (first.is_cuda or _ALLOW_CPU_UNITS)
        self.resident = wan if resident_bf16 is None else bool(resident_bf16)
        assert not self.resident or wan, "resident_bf16 needs WanAttentionBlock blocks on a CUDA device"
        self.units: List = []
        self.root: Optional[FlatUnit] = None
        if self.resident:
            taken = set()
            for i, blk in enumerate(blocks):
                if not any(p.requires_grad for p in blk.parameters()):
                    continue
                u = blk.__dict__.get("_prfl_unit")
                if u is None:
                    u = ResidentUnit(blk, self.world, self.rank, trainable=True)
                assert u.master is not None, f"blocks.{i} was made resident as frozen (no fp32 master) but requires grad"
                assert (u.world, u.rank) == (self.world, self.rank), "resident unit was built for another process group"
                u.sink = _BlockSink(self, u, f"blocks.{i}.")
                taken.update(id(p) for p in u.params)
                self.units.append(u)
            rest = [(n, p) for n, p in model.named_parameters() if p.requires_grad and id(p) not in taken]
            if rest:
                self.root = FlatUnit(rest, self.world, self.rank)
                self.units.append(self.root)
            from .model import bump_weight_epoch
            bump_weight_epoch()
        else:
            names = {id(p): n for n, p in model.named_parameters()}
            for ps in fsdp_units(model):
                self.units.append(FlatUnit([(names[id(p)], p) for p in ps], self.world, self.rank))
            self.root = self.units[-1] if self.units else None
        for u in self.units:
            u.m = torch.zeros_like(u.master)
            u.v = torch.zeros_like(u.master)
        self.state = self.units                                # (kept name: one entry per unit)
        self._hooks = []
        self._stage: List[Optional[torch.Tensor]] = [None, None]
        self._stage_free: List[Optional[torch.cuda.Event]] = [None, None]
        self._stage_next = 0
        self._comm = None
        # PRFL_RS=serial keeps every collective of the gradient / parameter path on the compute stream (no side stream, no
        # cross-stream events): nothing overlaps, and the order of work on each rank's single stream is the program order on
        # every rank.  The conservative mode `bench.py` retries the training-step leg in if the overlapped mode hangs.
        self.overlap = os.environ.get("PRFL_RS", "overlap").lower() != "serial"
        if first.is_cuda and self.world > 1 and self.overlap:
            self._comm = torch.cuda.Stream(device=first.device)

    @torch.no_grad()
    def warmup_collectives(self):
        """Allocate every large buffer of the streaming gradient path (both staging buffers, the gradient shards) and run one
        reduce-scatter and one bf16 all-gather of the REAL sizes on the side stream, then synchronise.  Call it once, at a
        quiescent point (all ranks aligned, no peer-memory kernels in flight): NCCL sets up the connections / algorithm of
        a message size on first use, and the training path interleaves its collectives with spinning symmetric-memory
        barrier kernels — first-use setup is kept away from that.  No-op for a single rank."""
        if self.world == 1:
            return
        res = [u for u in self.units if u.kind == "resident"]
        if not res:
            return
        big = max(u.n + u.pad for u in res)
        dev = res[0].wflat.device
        for k in range(2):
            if self._stage[k] is None or self._stage[k].numel() < big:
                self._stage[k] = torch.zeros(big, dtype=torch.float32, device=dev)
        for u in res:
            if u.gshard is None:
                u.gshard = torch.empty(u.shard, dtype=torch.float32, device=dev)
        u = max(res, key=lambda v: v.n)
        torch.cuda.synchronize()
        with torch.cuda.stream(self._comm if self._comm is not None else torch.cuda.current_stream()):
            tmp = torch.empty(u.shard, dtype=torch.float32, device=dev)
            op = dist.ReduceOp.AVG if (self.average and self._nccl()) else dist.ReduceOp.SUM
            dist.reduce_scatter_tensor(tmp, self._stage[0][:u.n + u.pad], op=op, group=self.group)
            full16 = torch.empty(u.n + u.pad, dtype=torch.bfloat16, device=dev)
            dist.all_gather_into_tensor(full16, u.my_slice(u.wflat).clone(), group=self.group)
        torch.cuda.synchronize()
        del tmp, full16

    # -- gradient intake ----------------------------------------------------------------------------
    def attach_hooks(self):
        """Move every flat unit's gradients into its flat fp32 gradient buffer as soon as autograd has accumulated them
        (post-accumulate-grad hooks) and free the per-parameter .grad.  Hooks ADD, so several backward() calls before a
        step() accumulate (micro-batches), whichever parameters each of them touches."""
        if self._hooks:
            return self
        for u in self.units:
            if u.kind != "flat":
                continue
            for p in u.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(u)))
        return self

    @staticmethod
    def _make_hook(u: FlatUnit):
        def hook(p):
            if p.grad is not None:          # the hook also fires when a Function returned None for this input (gradient sink)
                u.add_grad(id(p), p.grad)
                p.grad = None
        return hook

    # staging buffers for resident units (two, shared): wgrad GEMMs of unit i write one while unit i+1's is in flight
    def _stage_acquire(self, u: ResidentUnit):
        if self.world == 1:                                    # no collective: the "staging buffer" is the gradient shard itself
            if u.gshard is None:
                u.gshard = torch.zeros(u.n + u.pad, dtype=torch.float32, device=u.wflat.device)
            return u.gshard, 0
        if u.gshard is None:                                   # allocated on the compute stream, used on both
            u.gshard = torch.empty(u.shard, dtype=torch.float32, device=u.wflat.device)
        k = self._stage_next
        self._stage_next ^= 1
        if self._stage[k] is None or self._stage[k].numel() < u.n + u.pad:
            self._stage[k] = torch.zeros(u.n + u.pad, dtype=torch.float32, device=u.wflat.device)
        if self._stage_free[k] is not None:
            torch.cuda.current_stream().wait_event(self._stage_free[k])
        return self._stage[k][:u.n + u.pad], k

    def _stage_release(self, u: ResidentUnit, buf: torch.Tensor, k: int):
        if self.world == 1:
            u.has_grad = True
            return
        if self._comm is None:                                 # PRFL_RS=serial: in stream order, nothing to wait for later
            self._reduce_into_shard(u, buf)
            return
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ready)
            self._reduce_into_shard(u, buf)
            done = torch.cuda.Event()
            done.record()
        self._stage_free[k] = done

    # -- collectives --------------------------------------------------------------------------------
    def _nccl(self) -> bool:
        return self.world > 1 and dist.get_backend(self.group) == "nccl"

    def _reduce_into_shard(self, u, flat: torch.Tensor):
        """u.gshard (+)= this rank's slice of the group-reduced `flat`."""
        if self.world == 1:
            if u.gshard is None:
                u.gshard = flat.clone()
            elif u.has_grad:
                u.gshard.add_(flat)
            else:
                u.gshard.copy_(flat)
            u.has_grad = True
            return
        if self._nccl():
            op = dist.ReduceOp.AVG if self.average else dist.ReduceOp.SUM
            if u.gshard is None:
                u.gshard = torch.empty(u.shard, dtype=torch.float32, device=flat.device)
            if u.has_grad:
                tmp = torch.empty_like(u.gshard)
                dist.reduce_scatter_tensor(tmp, flat, op=op, group=self.group)
                u.gshard.add_(tmp)
            else:
                dist.reduce_scatter_tensor(u.gshard, flat, op=op, group=self.group)
        else:                                                   # gloo (CPU tests): same numbers through all_reduce + slice
            red = flat.clone()
            dist.all_reduce(red, op=dist.ReduceOp.SUM, group=self.group)
            part = u.my_slice(red)
            if self.average:
                part = part / self.world
            if u.gshard is None or not u.has_grad:
                u.gshard = part.clone()
            else:
                u.gshard.add_(part)
        u.has_grad = True

    def _publish(self, u):
        """All-gather the updated slices into the resident buffer (bf16 for block matrices, fp32 for flat units), in place."""
        if self.world == 1:
            return
        full = u.wflat if u.kind == "resident" else u.flat
        mine = u.my_slice(full)
        if not self._nccl():
            mine = mine.clone()
        dist.all_gather_into_tensor(full, mine, group=self.group)

    # -- API ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def flush(self):
        """Bring every gradient produced so far into the 1/W gradient shards: parameters whose .grad autograd left in
        place (no hooks attached) are collected, dirty flat gradient buffers are reduce-scattered and cleared, and the
        side stream's reduce-scatters of the resident units are joined."""
        for u in self.units:
            if u.kind != "flat":
                continue
            for p in u.params:
                if p.grad is not None:
                    u.add_grad(id(p), p.grad)
                    p.grad = None
            if u.dirty:
                self._reduce_into_shard(u, u.gflat)
                u.gflat.zero_()
                u.dirty = False
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)

    def reduce_gradients(self) -> List[torch.Tensor]:
        """Flush and return the gradient shards (one fp32 tensor per unit; zeros for units that received nothing)."""
        self.flush()
        return [u.gshard if u.has_grad else torch.zeros_like(u.master) for u in self.units]

    def _grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the sharded (accumulated) gradient, computed on the device (no host sync)."""
        live = [u for u in self.units if u.has_grad]
        dev = self.units[0].master.device
        if dev.type == "cuda":
            from . import ops
            acc = torch.zeros((), dtype=torch.float64, device=dev)
            for u in live:
                ops.sumsq_(u.gshard, acc)                       # one pass per shard, double accumulation
        else:                                                   # CPU tensors: only the world_size-2 gloo tests get here
            acc = torch.zeros((), dtype=torch.float64)
            for u in live:
                acc = acc + u.gshard.double().pow(2).sum()
        if self.world > 1:
            dist.all_reduce(acc, group=self.group)
        return acc.sqrt().float()

    @torch.no_grad()
    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """FSDP.clip_grad_norm_ on the ACCUMULATED gradient shards (train_prfl.py:825 calls it on every micro-step):
        scales them in place, returns the pre-clip global norm."""
        self.flush()
        norm = self._grad_norm()
        coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        for u in self.units:
            if u.has_grad:
                u.gshard.mul_(coef)
        return norm

    @torch.no_grad()
    def zero_grad(self):
        self.flush()
        for u in self.units:
            u.has_grad = False
            if u.sink is not None:
                u.sink.written = set()

    @torch.no_grad()
    def step(self, max_norm: Optional[float] = None):
        """[clip_grad_norm_(max_norm) +] AdamW on the shards + all-gather of the updated parameters (train_prfl.py:825-830).
        On the GPU every unit is one `prfl_adamw_step` launch with the clip coefficient read from device memory.  Units
        that received no gradient since the last step are left untouched (torch.optim skips parameters without .grad).
        Returns the pre-clip gradient norm if max_norm is given."""
        self.flush()
        norm = coef = None
        if max_norm is not None:
            norm = self._grad_norm()
            coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        b1, b2 = self.betas
        for u in self.units:
            if not u.has_grad:
                continue
            u.t += 1
            t = u.t
            w, m, v = u.master, u.m, u.v
            g = u.gshard
            if w.is_cuda:
                from . import ops
                out16 = u.my_slice(u.wflat) if u.kind == "resident" else None
                ops.adamw_step_(g, w, m, v, t, self.lr, self.betas, self.eps, self.wd, coef, bf16_out=out16)
                if self._comm is not None:
                    ready = torch.cuda.Event()
                    ready.record()
                    with torch.cuda.stream(self._comm):
                        self._comm.wait_event(ready)
                        self._publish(u)
                else:
                    self._publish(u)
            else:                                               # host-logic path of the gloo tests (torch.optim.AdamW math)
                if coef is not None:
                    g = g * coef
                w.mul_(1 - self.lr * self.wd)
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(self.eps)
                w.addcdiv_(m, denom, value=-self.lr / (1 - b1 ** t))
                if u.kind == "resident":                            # (test hook _ALLOW_CPU_UNITS) bf16(master), as prfl_adamw_step writes it
                    u.my_slice(u.wflat).copy_(w)
                self._publish(u)
            u.has_grad = False
            if u.sink is not None:
                u.sink.written = set()
        if self._comm is not None:
            torch.cuda.current_stream().wait_stream(self._comm)
        from .model import bump_weight_epoch
        bump_weight_epoch()                                     # derived operands (fp32 bias copies, split head weights, K/V caches) refresh
        return norm

    # -- full-precision weights for checkpoints ---------------------------------------------------------
    @torch.no_grad()
    def full_state_dict(self, to_cpu: bool = True) -> Dict[str, torch.Tensor]:
        """model.state_dict() with every resident (bf16) matrix replaced by its fp32 MASTER, gathered unit by unit from the
        1/W shards (what FSDP's FULL_STATE_DICT yields in the reference, model_utils.py:70-86).  Collective: call on all ranks."""
        sd = {k: (v.detach().cpu() if to_cpu else v.detach().clone()) for k, v in self.model.state_dict().items()}
        if not self.resident:
            return sd
        for u in self.units:
            if u.kind != "resident":
                continue
            if self.world > 1:
                full = torch.empty(u.n + u.pad, dtype=torch.float32, device=u.master.device)
                dist.all_gather_into_tensor(full, u.master, group=self.group)
            else:
                full = u.master
            for n, (o, cnt, shp) in u.offsets.items():
                t = full[o:o + cnt].view(shp)
                sd[u.sink.prefix + n] = t.cpu() if to_cpu else t.clone()
        return sd

    # -- optimizer state on disk (the reference saves none, train_prfl.py:485-491) ---------------
    def state_dict(self) -> dict:
        """This rank's shard of the optimizer state, flat fp32 tensors keyed by unit index (safetensors-friendly)."""
        out = {}
        for ui, u in enumerate(self.units):
            out[f"unit{ui:03d}.master"], out[f"unit{ui:03d}.m"], out[f"unit{ui:03d}.v"] = u.master, u.m, u.v
            out[f"unit{ui:03d}.t"] = torch.tensor([u.t], dtype=torch.int64)
        return out

    @torch.no_grad()
    def load_state_dict(self, sd: dict):
        for ui, u in enumerate(self.units):
            for k, dst in (("master", u.master), ("m", u.m), ("v", u.v)):
                src = sd[f"unit{ui:03d}.{k}"]
                assert src.shape == dst.shape, (ui, k, src.shape, dst.shape)
                dst.copy_(src)
            u.t = int(sd[f"unit{ui:03d}.t"][0])
            if u.kind == "resident":
                u.my_slice(u.wflat).copy_(u.master)
            self._publish(u)
        from .model import bump_weight_epoch
        bump_weight_epoch()

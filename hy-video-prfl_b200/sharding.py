"""Sharded gradients + optimizer state for the trainable Wan-DiT (the FSDP FULL_SHARD role in the reference,
diffusers_lite/utils/fsdp_utils.py:66-122 + train_prfl.py:346-362, 482-491), re-cut for 180 GB B200s.

The reference shards parameters, gradients and AdamW state over ALL ranks (SP ranks included) and all-gathers every
block's fp32 parameters twice per step (1.41 GB per block per direction).  With 180 GB per GPU the 14B bf16 operands
(28 GB) fit replicated, so this build keeps the *compute copies* resident and shards only what is large and cold:

  * per FSDP unit (one per WanAttentionBlock + one root unit, `_no_split_modules`), gradients are flattened and
    REDUCE-SCATTERED (fp32, averaged over the world group exactly as FSDP does — under Ulysses every rank holds the
    partial gradient of its token chunk) so each rank keeps 1/W of them;
  * fp32 master weights and AdamW moments exist only as 1/W shards;
  * after the update the new parameters are ALL-GATHERED once per step (not per block per pass).

Collectives are NCCL through torch.distributed (`reduce_scatter_tensor` / `all_gather_into_tensor`); on backends
without reduce-scatter (gloo, used by the CPU tests) the same result is produced with all_reduce + slice.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn


def fsdp_units(model: nn.Module) -> List[List[nn.Parameter]]:
    """One unit per block (load.py:6-12 wraps WanAttentionBlock) + a root unit with everything else."""
    units, seen = [], set()
    blocks = getattr(model, "blocks", [])
    for blk in blocks:
        ps = [p for p in blk.parameters() if p.requires_grad]
        seen.update(id(p) for p in ps)
        if ps:
            units.append(ps)
    root = [p for p in model.parameters() if p.requires_grad and id(p) not in seen]
    if root:
        units.append(root)
    return units


class ShardedAdamW:
    """AdamW over reduce-scattered gradient shards with sharded fp32 masters and moments (ZeRO-2 layout)."""

    def __init__(self, model: nn.Module, lr=1e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, group=None,
                 average: bool = True):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lr, self.betas, self.eps, self.wd, self.average = lr, betas, eps, weight_decay, average
        self.units = fsdp_units(model)
        self.state = []
        for ps in self.units:
            n = sum(p.numel() for p in ps)
            pad = (-n) % self.world
            shard = (n + pad) // self.world
            flat = torch.cat([p.detach().reshape(-1).float() for p in ps])
            if pad:
                flat = torch.cat([flat, flat.new_zeros(pad)])
            master = flat[self.rank * shard:(self.rank + 1) * shard].clone()
            self.state.append(dict(n=n, pad=pad, shard=shard, master=master, m=torch.zeros_like(master),
                                   v=torch.zeros_like(master), t=0))

        self._pending = None        # unit index -> [flat buffer, params still missing] while a backward is running
        self._shards = None         # reduce-scattered gradient shards accumulated by the hooks
        self._hooks = []

    # -- streaming mode: reduce-scatter each unit as soon as its last gradient lands ---------------------
    def attach_hooks(self):
        """Register post-accumulate-grad hooks so that every unit's gradients are flattened, reduce-scattered and FREED
        during backward, block by block (what FSDP does after each block's backward, fsdp_utils.py:86-109).  Peak gradient
        memory is one unit (1.4 GB for a 14B block) + the 1/W shards instead of the full 56 GB."""
        if self._hooks:
            return self
        self._pending, self._shards = {}, [None] * len(self.units)
        for ui, ps in enumerate(self.units):
            offs, o = {}, 0
            for p in ps:
                offs[id(p)] = o
                o += p.numel()
            for p in ps:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(ui, offs)))
        return self

    def _make_hook(self, ui, offs):
        def hook(p):
            st, ps = self.state[ui], self.units[ui]
            ent = self._pending.get(ui)
            if ent is None:
                flat = torch.zeros(st["n"] + st["pad"], dtype=torch.float32, device=p.device)
                ent = self._pending[ui] = [flat, len(ps)]
            o = offs[id(p)]
            ent[0][o:o + p.numel()].copy_(p.grad.reshape(-1))
            p.grad = None
            ent[1] -= 1
            if ent[1] == 0:
                shard = self._reduce_scatter(ent[0], st["shard"])
                self._shards[ui] = shard if self._shards[ui] is None else self._shards[ui] + shard
                del self._pending[ui]
        return hook

    def _take_hook_shards(self):
        # units whose parameters did not all receive a gradient this step (e.g. unused img_emb) are flushed here
        for ui in list(self._pending):
            flat, _ = self._pending.pop(ui)
            shard = self._reduce_scatter(flat, self.state[ui]["shard"])
            self._shards[ui] = shard if self._shards[ui] is None else self._shards[ui] + shard
        out = []
        for ui, st in enumerate(self.state):
            sh = self._shards[ui]
            out.append(sh if sh is not None else torch.zeros_like(st["master"]))
            self._shards[ui] = None
        return out

    # -- collectives --------------------------------------------------------------------------------
    def _reduce_scatter(self, flat: torch.Tensor, shard: int) -> torch.Tensor:
        if self.world == 1:
            return flat
        op = dist.ReduceOp.AVG if (self.average and dist.get_backend(self.group) == "nccl") else dist.ReduceOp.SUM
        if dist.get_backend(self.group) == "nccl":
            out = torch.empty(shard, dtype=flat.dtype, device=flat.device)
            dist.reduce_scatter_tensor(out, flat, op=op, group=self.group)
            return out
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)       # gloo path (tests): same numbers
        out = flat[self.rank * shard:(self.rank + 1) * shard].clone()
        return out / self.world if self.average else out

    def _all_gather(self, shard_t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return shard_t
        out = torch.empty(self.world * shard_t.numel(), dtype=shard_t.dtype, device=shard_t.device)
        dist.all_gather_into_tensor(out, shard_t, group=self.group)
        return out

    # -- API ------------------------------------------------------------------------------------------
    def reduce_gradients(self) -> List[torch.Tensor]:
        """Flatten + reduce-scatter every unit's gradients; frees the full-size .grad tensors. Returns the shards."""
        shards = []
        for ps, st in zip(self.units, self.state):
            flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in ps])
            if st["pad"]:
                flat = torch.cat([flat, flat.new_zeros(st["pad"])])
            for p in ps:
                p.grad = None
            shards.append(self._reduce_scatter(flat, st["shard"]))
        return shards

    def clip_grad_norm_(self, shards: List[torch.Tensor], max_norm: float) -> torch.Tensor:
        """Global L2 norm over all shards of all ranks (FSDP.clip_grad_norm_, train_prfl.py:825)."""
        sq = torch.stack([s.pow(2).sum() for s in shards]).sum()
        if self.world > 1:
            dist.all_reduce(sq, group=self.group)
        norm = sq.sqrt()
        coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        for s in shards:
            s.mul_(coef)
        return norm

    def _grad_norm(self, shards: List[torch.Tensor]) -> torch.Tensor:
        """Global L2 norm of the sharded gradient, on the device (no host sync)."""
        if shards[0].is_cuda:
            from . import ops
            acc = torch.zeros((), dtype=torch.float64, device=shards[0].device)
            for s in shards:
                ops.sumsq_(s, acc)                                   # one pass per shard, double accumulation
        else:                                                        # CPU tensors: only the world_size-2 gloo tests get here
            acc = torch.stack([s.double().pow(2).sum() for s in shards]).sum()
        if self.world > 1:
            dist.all_reduce(acc, group=self.group)
        return acc.sqrt().float()

    @torch.no_grad()
    def step(self, shards: Optional[List[torch.Tensor]] = None, max_norm: Optional[float] = None):
        """clip_grad_norm_(max_norm) + AdamW on the shards + all-gather of the updated parameters (train_prfl.py:825-830).
        On the GPU every unit is one `prfl_adamw_step` launch with the clip coefficient read from device memory."""
        if shards is None:
            shards = self._take_hook_shards() if self._hooks else self.reduce_gradients()
        norm, coef = None, None
        if max_norm is not None:
            norm = self._grad_norm(shards)
            coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0)
        b1, b2 = self.betas
        for ps, st, g in zip(self.units, self.state, shards):
            st["t"] += 1
            t = st["t"]
            w, m, v = st["master"], st["m"], st["v"]
            if w.is_cuda:
                from . import ops
                ops.adamw_step_(g.contiguous(), w, m, v, t, self.lr, self.betas, self.eps, self.wd, coef)
            else:                                                    # host-logic path of the gloo tests (torch.optim.AdamW math)
                if coef is not None:
                    g = g * coef
                w.mul_(1 - self.lr * self.wd)
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (v.sqrt() / math.sqrt(1 - b2 ** t)).add_(self.eps)
                w.addcdiv_(m, denom, value=-self.lr / (1 - b1 ** t))
            full = self._all_gather(w)
            off = 0
            for p in ps:
                p.data.copy_(full[off:off + p.numel()].view_as(p))      # bumps p._version => bf16 operand caches refresh
                off += p.numel()
        return norm

    # -- optimizer state on disk (the reference saves none, train_prfl.py:485-491) ---------------
    def state_dict(self) -> dict:
        """This rank's shard of the optimizer state, flat fp32 tensors keyed by unit index (safetensors-friendly)."""
        out = {}
        for ui, st in enumerate(self.state):
            for k in ("master", "m", "v"):
                out[f"unit{ui:03d}.{k}"] = st[k]
            out[f"unit{ui:03d}.t"] = torch.tensor([st["t"]], dtype=torch.int64)
        return out

    def load_state_dict(self, sd: dict):
        for ui, st in enumerate(self.state):
            for k in ("master", "m", "v"):
                src = sd[f"unit{ui:03d}.{k}"]
                assert src.shape == st[k].shape, (ui, k, src.shape, st[k].shape)
                st[k].copy_(src)
            st["t"] = int(sd[f"unit{ui:03d}.t"][0])

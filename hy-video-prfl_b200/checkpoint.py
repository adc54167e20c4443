"""Checkpoint wire formats — the on-disk layout the reference trainers write and read
(diffusers_lite/utils/model_utils.py:70-141), so checkpoints move freely between the reference and this path:

  checkpoint-{step}[-ema]/
      diffusion_pytorch_model.safetensors                      (total <= 5 GiB)              model_utils.py:88-90
      diffusion_pytorch_model-{i:05}-of-{n:05}.safetensors     + diffusion_pytorch_model.safetensors.index.json
                                                               (keys sorted, greedy 5 GiB shards)      :91-118
      config.json                                              (transformer.config minus "dtype")       :120-126
      optimizer-rank{r:05}-of-{w:05}.safetensors               (NOT in the reference, which saves no optimizer state,
                                                               train_prfl.py:485-491: this rank's ShardedAdamW shard)

Parameter names are the reference's (the drop-in modules keep its state-dict keys).  In this build the bf16 compute weights
are replicated and the fp32 masters are 1/W shards: `ShardedAdamW.full_state_dict()` gathers them for
`save_checkpoint(..., state_dict=...)` (the FULL_STATE_DICT role).
Cold path: plain Python + safetensors, no kernels.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import torch

__all__ = ["save_checkpoint", "write_model_dir", "load_model_dir", "load_state_dict", "update_ema_model", "save_optimizer", "load_optimizer"]

MAX_SHARD_BYTES = 5 * 1024 ** 3


def _config_dict(transformer) -> dict:
    cfg = getattr(transformer, "config", {})
    cfg = dict(cfg) if isinstance(cfg, dict) else {k: v for k, v in vars(cfg).items() if not k.startswith("_")}
    cfg.pop("dtype", None)
    return cfg


def save_checkpoint(transformer, rank: int, output_dir: str, step: int, ema: bool = False, max_bytes: int = MAX_SHARD_BYTES,
                    state_dict: Optional[Dict[str, torch.Tensor]] = None) -> Optional[str]:
    """model_utils.py:70-126.  Returns the directory written (rank <= 0) or None.
    `state_dict`: what to write instead of `transformer.state_dict()` — with resident bf16 weights pass
    `ShardedAdamW.full_state_dict()` (the fp32 masters gathered from the 1/W shards, a collective every rank calls), which
    is what the reference's FSDP FULL_STATE_DICT gather yields (model_utils.py:75-86); without it the bf16 compute copies
    are written."""
    if rank > 0:
        return None
    src = transformer.state_dict() if state_dict is None else state_dict
    save_dir = os.path.join(output_dir, f"checkpoint-{step}-ema" if ema else f"checkpoint-{step}")
    write_model_dir(save_dir, src, _config_dict(transformer), max_bytes)
    return save_dir


def write_model_dir(save_dir: str, state: Dict[str, torch.Tensor], config: dict, max_bytes: int = MAX_SHARD_BYTES) -> None:
    """The body of model_utils.py:88-126: one `diffusion_pytorch_model.safetensors` if the state is <= max_bytes, else greedy
    shards over the sorted keys + index JSON; then `config.json`.  Shared by `save_checkpoint` and `WanModel.save_pretrained`."""
    from safetensors.torch import save_file
    cpu_state = {k: v.detach().to("cpu").contiguous() for k, v in state.items()}
    os.makedirs(save_dir, exist_ok=True)
    total_bytes = sum(v.numel() * v.element_size() for v in cpu_state.values())
    if total_bytes <= max_bytes:
        save_file(cpu_state, os.path.join(save_dir, "diffusion_pytorch_model.safetensors"))
    else:
        shard, shards, current = {}, [], 0
        for k, v in sorted(cpu_state.items()):
            size = v.numel() * v.element_size()
            if current + size > max_bytes and shard:
                shards.append(shard)
                shard, current = {}, 0
            shard[k], current = v, current + size
        if shard:
            shards.append(shard)
        index = {"metadata": {"total_size": total_bytes}, "weight_map": {}}
        for i, sh in enumerate(shards, start=1):
            name = f"diffusion_pytorch_model-{i:05}-of-{len(shards):05}.safetensors"
            save_file(sh, os.path.join(save_dir, name))
            for key in sh:
                index["weight_map"][key] = name
        with open(os.path.join(save_dir, "diffusion_pytorch_model.safetensors.index.json"), "w") as f:
            json.dump(index, f, indent=2)
    with open(os.path.join(save_dir, "config.json"), "w") as f:
        json.dump(config, f, indent=4)


def load_model_dir(model_dir: str) -> Dict[str, torch.Tensor]:
    """The weights of a `from_pretrained`-style directory (what diffusers' `ModelMixin.from_pretrained`, the loader behind the
    reference's `WanModel.from_pretrained(...)` calls — train_prfl.py:182-217 — reads): the index JSON's shard list if there is
    one, else `diffusion_pytorch_model.safetensors`, else the legacy `diffusion_pytorch_model.bin`.  Unlike `load_state_dict`
    (the reference's merge-every-file helper) other `.safetensors` files in the directory are not touched."""
    from safetensors.torch import load_file
    idx = os.path.join(model_dir, "diffusion_pytorch_model.safetensors.index.json")
    one = os.path.join(model_dir, "diffusion_pytorch_model.safetensors")
    legacy = os.path.join(model_dir, "diffusion_pytorch_model.bin")
    if os.path.exists(idx):
        with open(idx) as f:
            weight_map = json.load(f)["weight_map"]
        state = {}
        for name in sorted(set(weight_map.values())):
            chunk = load_file(os.path.join(model_dir, name), device="cpu")
            state.update(chunk)
        missing = sorted(set(weight_map) - set(state))
        if missing:
            raise KeyError(f"{idx} lists {len(missing)} tensors its shards do not hold (first: {missing[0]})")
        return state
    if os.path.exists(one):
        return load_file(one, device="cpu")
    if os.path.exists(legacy):
        return torch.load(legacy, map_location="cpu", weights_only=True)
    raise FileNotFoundError(f"no diffusion_pytorch_model.safetensors[.index.json] / .bin under {model_dir}")


def load_state_dict(model_dir: str, postfix: str = ".safetensors") -> Dict[str, torch.Tensor]:
    """model_utils.py:128-141: merge every `*{postfix}` file of the directory (optimizer shards excluded)."""
    from safetensors.torch import load_file
    state = {}
    for name in sorted(os.listdir(model_dir)):
        if not name.endswith(postfix) or name.startswith("optimizer-"):
            continue
        path = os.path.join(model_dir, name)
        chunk = load_file(path, device="cpu") if postfix == ".safetensors" else torch.load(path, map_location="cpu")
        if "module" in chunk:
            chunk = chunk["module"]
        state.update(chunk)
    return state


@torch.no_grad()
def update_ema_model(transformer, ema_transformer, ema_decay: float) -> None:
    """model_utils.py:172-175: p_ema <- decay * p_ema + (1 - decay) * p for every trainable parameter.  The rule writes through
    `.data`, which PyTorch's version counters do not see, so the derived GEMM operands of the averaged model are invalidated
    here.  Resident bf16 weights are refused: increments of (1 - decay) * p fall below one bf16 ulp and would be lost — average
    the fp32 masters (`ShardedAdamW.full_state_dict()`) instead."""
    from .model import bump_weight_epoch
    for p_averaged, p_model in zip(ema_transformer.parameters(), transformer.parameters()):
        if p_model.requires_grad:
            if p_averaged.dtype != torch.float32:
                raise NotImplementedError(f"EMA over {p_averaged.dtype} parameters is lossy; keep the averaged model in fp32")
            p_averaged.data.mul_(ema_decay).add_(p_model.data.to(p_averaged.dtype), alpha=1 - ema_decay)
    bump_weight_epoch()


def _opt_name(rank: int, world: int) -> str:
    return f"optimizer-rank{rank:05}-of-{world:05}.safetensors"


def save_optimizer(opt, output_dir: str, step: int, ema: bool = False) -> str:
    """Every rank writes its own shard of the ShardedAdamW state (fp32 masters, both moments, step counts)."""
    from safetensors.torch import save_file
    save_dir = os.path.join(output_dir, f"checkpoint-{step}-ema" if ema else f"checkpoint-{step}")
    os.makedirs(save_dir, exist_ok=True)
    sd = {k: v.detach().to("cpu").contiguous() for k, v in opt.state_dict().items()}
    path = os.path.join(save_dir, _opt_name(opt.rank, opt.world))
    save_file(sd, path, metadata={"rank": str(opt.rank), "world": str(opt.world), "lr": repr(opt.lr), "betas": repr(tuple(opt.betas)),
                                  "eps": repr(opt.eps), "weight_decay": repr(opt.wd)})
    return path


def load_optimizer(opt, model_dir: str) -> None:
    """Restore this rank's shard; the world size must match the one the checkpoint was written with."""
    from safetensors.torch import load_file
    path = os.path.join(model_dir, _opt_name(opt.rank, opt.world))
    if not os.path.exists(path):
        have = [n for n in os.listdir(model_dir) if n.startswith("optimizer-")]
        raise FileNotFoundError(f"{path} not found (checkpoint holds {have}); resharding optimizer state across a different "
                                "world size is not supported")
    opt.load_state_dict(load_file(path, device=str(opt.units[0].master.device)))

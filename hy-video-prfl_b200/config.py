"""The `configs/*.yaml` surface of the reference (SURVEY.md §5.6, §8b "Config surface"), consumed by this package's objects.

The reference trainers read their YAML with OmegaConf (train_prfl.py:1199) and turn a handful of its entries into the objects
on the hot path.  This module does the same mapping onto the drop-in classes, so a shipped `configs/train_prfl_*.yaml` /
`train_pavrm_*.yaml` drives the B200 path without edits:

    cfg = load_config("configs/train_prfl_i2v_720.yaml")
    init_parallel_from_config(cfg)                                   # dataset.sp_size            train_prfl.py:119
    lrm = truncate_reward_transformer(WanModel.from_pretrained(cfg.model.base_path), cfg)         # :217-258
    query_attention, mlp = reward_head_from_config(cfg)              # lrm.mlp_dim / query_attention  :268-305
    noise_scheduler = scheduler_from_config(cfg)                     # extra_model.scheduler      :411-413
    optimizer = optimizer_from_config(cfg, transformer)              # optimizer.*                :482-491
    loss, reward = refl_chain(transformer, lrm, query_attention, mlp, noise_scheduler, ..., **refl_kwargs(cfg))

Host-side cold path: plain Python.  OmegaConf is not required: PyYAML parses the files, and the two things OmegaConf does that
the trainers rely on are restated here — attribute access on nested mappings, and YAML-1.2 number parsing (PyYAML's YAML-1.1
resolver reads `5e-6` — `optimizer.learning_rate` in every shipped config — as a *string*).
"""
from __future__ import annotations

import re
from typing import Any, Dict, Optional, Tuple

__all__ = ["ConfigNode", "load_config", "WAN_ARCH", "arch_from_task", "init_parallel_from_config", "reward_head_from_config",
           "truncate_reward_transformer", "scheduler_from_config", "optimizer_from_config", "refl_kwargs", "accumulation_steps"]

_FLOAT = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$")


class ConfigNode(dict):
    """Nested mapping with attribute access, `hasattr` / `getattr(..., default)` / `.get` as the trainers use them."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value


def _wrap(v: Any) -> Any:
    if isinstance(v, dict):
        return ConfigNode({k: _wrap(u) for k, u in v.items()})
    if isinstance(v, list):
        return [_wrap(u) for u in v]
    if isinstance(v, str) and _FLOAT.match(v.strip()) and not v.strip().lstrip("+-").isdigit():
        return float(v)                                             # `5e-6` and friends (YAML 1.2 floats PyYAML leaves as strings)
    return v


def load_config(path: str) -> ConfigNode:
    import yaml
    with open(path) as f:
        return _wrap(yaml.safe_load(f))


# Architectures by `task` (the keys of NAME_MAPPING, train_prfl.py:86-93; values of diffusers_lite/wan/configs/wan_*.py:20-36 and
# shared_config.py:11).  With real weights the architecture comes from the checkpoint's config.json through
# `WanModel.from_pretrained`; this table is for random-init builds (benchmarks, tests).
_B14 = dict(dim=5120, ffn_dim=13824, num_heads=40, num_layers=40, freq_dim=256, text_len=512, eps=1e-6)
_B1_3 = dict(dim=1536, ffn_dim=8960, num_heads=12, num_layers=30, freq_dim=256, text_len=512, eps=1e-6)
WAN_ARCH: Dict[str, Dict[str, Any]] = {
    "t2v-1.3b": dict(model_type="t2v", in_dim=16, **_B1_3),
    "i2v-1.3b": dict(model_type="i2v", in_dim=36, **_B1_3),           # train_prfl.py:199-205: the T2V 1.3B config with in_dim 36
    "t2v-14b": dict(model_type="t2v", in_dim=16, **_B14),
    "i2v-14b-480p": dict(model_type="i2v", in_dim=36, **_B14),
    "i2v-14b-720p": dict(model_type="i2v", in_dim=36, **_B14),
    "flf2v-14b-720p": dict(model_type="flf2v", in_dim=36, **_B14),
}


def arch_from_task(task: str) -> Dict[str, Any]:
    if task not in WAN_ARCH:
        raise KeyError(f"task {task!r}: expected one of {sorted(WAN_ARCH)}")
    return dict(WAN_ARCH[task])


def init_parallel_from_config(cfg) -> int:
    """dataset.sp_size -> Ulysses groups of that many consecutive ranks (train_prfl.py:119).  Returns the SP degree."""
    from .parallel import initialize_sequence_parallel_state
    sp = int(cfg.dataset.sp_size)
    initialize_sequence_parallel_state(sp)
    return sp


def reward_head_from_config(cfg, device=None) -> Tuple["QueryAttention", "MLP"]:
    """lrm.mlp_dim + lrm.query_attention.{num_queries,num_heads,dropout,return_type,product_text,text_dim} -> (QueryAttention,
    MLP), fp32, eval, with the trainer's defaults for absent keys (train_prfl.py:268-314).  Weights are the caller's to load
    (`model.lrm_mlp_path`, `model.lrm_query_attention_path`)."""
    import torch
    from .network import MLP, QueryAttention
    dim = int(cfg.lrm.mlp_dim)
    qc = getattr(cfg.lrm, "query_attention", None) or {}
    qa = QueryAttention(feature_dim=dim, num_queries=int(qc.get("num_queries", 1)), num_heads=int(qc.get("num_heads", 8)),
                        dropout=float(qc.get("dropout", 0.)), layer_norm=bool(qc.get("layer_norm", False)),
                        return_type=qc.get("return_type", None), product_text=bool(qc.get("product_text", False)),
                        text_dim=int(qc.get("text_dim", 4096)))
    mlp = MLP(dim)
    if device is not None:
        qa, mlp = qa.to(device=device, dtype=torch.float32), mlp.to(device)
    return qa.eval(), mlp.eval()


def truncate_reward_transformer(model, cfg, trainable: bool = False):
    """The latent reward model's transformer as the trainers cut it (train_prfl.py:219-258, train_pavrm.py:195-235): the
    embeddings are frozen, only `lrm.trainable_blocks` are kept (as a new ModuleList), the head is removed;
    `lrm.feature_layer` defaults to [6, 7] when the YAML has none.  `trainable=False` (PRFL: the reward model is not in the
    optimizer) freezes the kept blocks too, which selects the dgrad-only backward."""
    import torch.nn as nn
    for name in ("patch_embedding", "text_embedding", "time_embedding", "time_projection", "img_emb"):
        mod = getattr(model, name, None)
        if mod is not None:
            for p in mod.parameters():
                p.requires_grad = False
    keep = [int(i) for i in cfg.lrm.trainable_blocks]
    if not hasattr(cfg.lrm, "feature_layer"):
        cfg.lrm.feature_layer = [6, 7]
    kept = []
    for i, blk in enumerate(model.blocks):
        on = i in keep
        for p in blk.parameters():
            p.requires_grad = bool(on and trainable)
        if on:
            kept.append(blk)
    model.blocks = nn.ModuleList(kept)
    model.head = None
    if hasattr(model, "num_layers"):
        model.num_layers = len(kept)
    return model


def scheduler_from_config(cfg):
    """train_prfl.py:411-413: FlowUniPCMultistepScheduler(num_train_timesteps, shift=1, use_dynamic_shifting=False); the flow
    shift is applied per chain by `set_timesteps(..., shift=extra_model.scheduler.flow_shift)` (`refl_kwargs`)."""
    from .scheduler import FlowUniPCMultistepScheduler
    return FlowUniPCMultistepScheduler(num_train_timesteps=int(cfg.extra_model.scheduler.num_train_timesteps), shift=1,
                                       use_dynamic_shifting=False)


def optimizer_from_config(cfg, transformer, group=None):
    """train_prfl.py:482-491 (AdamW over the transformer's trainable parameters, eps 1e-8) on the sharded training state that
    replaces the FSDP wrap (`model.fsdp.fsdp_sharding_startegy: full`)."""
    from .sharding import ShardedAdamW
    o = cfg.optimizer
    return ShardedAdamW(transformer, lr=float(o.learning_rate), betas=(float(o.adam_beta1), float(o.adam_beta2)), eps=1e-8,
                        weight_decay=float(o.weight_decay), group=group)


def refl_kwargs(cfg) -> Dict[str, Any]:
    """Keyword arguments of `prfl.refl_chain` that come from the YAML (train_prfl.py:632-633, 764-770)."""
    fl = getattr(cfg.lrm, "feature_layer", None) or [6, 7]
    return dict(flow_shift=float(cfg.extra_model.scheduler.flow_shift), feature_layer=[int(v) for v in fl])


def accumulation_steps(cfg) -> int:
    """train.gradient_accumulation_steps is written as a float (`5.`) in the shipped files."""
    return max(1, int(round(float(getattr(cfg.train, "gradient_accumulation_steps", 1)))))

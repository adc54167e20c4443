"""In-place installation of the B200 path into an existing REFERENCE model, in the reference's own
plugin style: `diffusers_lite/wan/modules/context_parallel/plugins.py:26-37` keeps `module.old_forward` and swaps
`module.forward` behind an enable flag; `text2video.py:145-148` patches methods with `types.MethodType`.

    from prfl_b200.plugin import install, uninstall
    install(ref_model)          # ref_model: diffusers_lite.wan.modules.model.WanModel, already on its CUDA device
    ...                         # trainers / pipelines call ref_model(...) exactly as before
    ref_model.prfl_b200_enable(False)   # back to the reference kernels (A/B), True to re-enable
    uninstall(ref_model)

Each reference `WanAttentionBlock` gets a shadow `prfl_b200.model.WanAttentionBlock` holding the SAME parameters (the
shadow's parameters are re-pointed at the reference block's tensors, so optimizers / checkpoints / FSDP-free training keep
working on the reference's own nn.Parameters and no weights are duplicated); its forward replaces the block's forward.
Everything outside the blocks (embeddings, head, unpatchify) stays the reference's code: the blocks are >= 99 % of the
work.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .model import WanAttentionBlock
from .parallel import adopt_reference_state

__all__ = ["install", "uninstall"]


def _shadow(blk: nn.Module, cross_attn_type: str) -> WanAttentionBlock:
    with torch.device("meta"):
        fast = WanAttentionBlock(cross_attn_type, blk.dim, blk.ffn_dim, blk.num_heads, tuple(blk.window_size), blk.qk_norm,
                                 blk.cross_attn_norm, blk.eps)
    ref = dict(blk.named_parameters())
    mine = dict(fast.named_parameters())
    missing, extra = sorted(set(mine) - set(ref)), sorted(set(ref) - set(mine))
    if missing or extra:
        raise RuntimeError(f"reference block and prfl_b200 block disagree on parameter names: missing {missing}, unexpected {extra}")
    for name, p in ref.items():                               # share the reference's own Parameters (no copy)
        mod, _, leaf = name.rpartition(".")
        owner = fast.get_submodule(mod) if mod else fast
        if mine[name].shape != p.shape:
            raise RuntimeError(f"{name}: shape {tuple(p.shape)} != {tuple(mine[name].shape)}")
        owner._parameters[leaf] = p
    return fast


def install(ref_model: nn.Module) -> nn.Module:
    """Patch every block of a reference WanModel to run on the prfl_b200 kernels.  Idempotent."""
    cross = "t2v_cross_attn" if getattr(ref_model, "model_type", "t2v") == "t2v" else "i2v_cross_attn"
    for i, blk in enumerate(ref_model.blocks):                  # checked up front: either every block is patched or none
        if hasattr(blk, "_fsdp_wrapped_module") or hasattr(blk, "_checkpoint_wrapped_module"):
            # FSDP swaps a wrapped block's parameters for views of its flat parameter on every forward: a shadow holding the
            # original nn.Parameter objects would read stale storage.  Refuse instead of computing with old weights.
            raise RuntimeError(f"blocks.{i} is wrapped ({type(blk).__name__}): install() before / instead of the FSDP and "
                               "activation-checkpoint wrap — sharded training uses prfl_b200.sharding (INTEGRATION.md B')")
    for i, blk in enumerate(ref_model.blocks):
        if hasattr(blk, "_prfl_b200_fast"):
            continue
        fast = _shadow(blk, cross)
        fast.train(blk.training)
        object.__setattr__(blk, "_prfl_b200_fast", fast)      # not registered as a submodule: no duplicate state-dict keys
        if not hasattr(blk, "old_forward"):
            blk.old_forward = blk.forward                     # ModulePlugin convention (plugins.py:28-29)
        blk._prfl_b200_on = True

        def forward(x, e, seq_lens, grid_sizes, freqs, context, context_lens, _blk=blk, _fast=fast, _first=(i == 0)):
            if not _blk._prfl_b200_on:
                return _blk.old_forward(x, e, seq_lens, grid_sizes, freqs, context, context_lens)
            if _first:
                # the reference model chunks the tokens by ITS sequence-parallel state (model.py:618-619); the patched blocks must
                # exchange by the same one (checked once per forward, at block 0: a dict lookup and a few comparisons)
                adopt_reference_state()
            # the reference feeds block 0 the bf16 patch embedding (model.py:345 with x.dtype == bf16): keep that rounding
            return _fast(x.float().contiguous(), e, seq_lens, grid_sizes, freqs, context, context_lens,
                         first_block_bf16_input=_first and x.dtype == torch.bfloat16)
        blk.forward = forward

    def enable(flag: bool = True, _m=ref_model):
        for b in _m.blocks:
            if hasattr(b, "_prfl_b200_fast"):
                b._prfl_b200_on = bool(flag)
    ref_model.prfl_b200_enable = enable
    adopt_reference_state()
    return ref_model


def uninstall(ref_model: nn.Module) -> nn.Module:
    for blk in ref_model.blocks:
        if hasattr(blk, "_prfl_b200_fast"):
            del blk.forward                                   # instance attribute: the class's forward shows through again
            del blk.old_forward
            object.__delattr__(blk, "_prfl_b200_fast")
            del blk._prfl_b200_on
    if hasattr(ref_model, "prfl_b200_enable"):
        del ref_model.prfl_b200_enable
    return ref_model

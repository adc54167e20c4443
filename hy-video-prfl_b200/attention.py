"""`flash_attention` with the reference's signature (diffusers_lite/wan/modules/attention.py:24-130), on the tcgen05
kernels (prfl_attn_fwd / prfl_attn_bwd).  Differentiable (autograd.Function with the hand-written backward).

Only the configuration the Wan-DiT path uses is implemented — non-causal, no dropout, default or explicit softmax
scale, full-length queries, optional per-sample key lengths — everything else raises instead of silently differing.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import ops

__all__ = ["flash_attention"]


class _FlashAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, scale):
        # q: [Lq, N, 128], k/v: [Lk, N, 128] bf16
        o, lse = ops.attn_fwd(q, k, v, scale=scale, need_lse=True)
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        do = do.contiguous() if do.stride(2) != 1 else do
        dq, dk, dv = ops.attn_bwd(q, k, v, o, do.to(torch.bfloat16), lse, scale=ctx.scale)
        return dq, dk, dv, None


def flash_attention(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None, causal=False,
                    window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16, version=None):
    """q: [B, Lq, N, 128]; k, v: [B, Lk, N, 128]; k_lens: [B] or None.  Returns [B, Lq, N, 128] in q's dtype
    (attention.py:57,130).  Inputs that are not 16-bit are rounded to `dtype` as attention.py:59-82 does."""
    if q_lens is not None or causal or dropout_p != 0. or tuple(window_size) != (-1, -1):
        raise NotImplementedError("prfl_b200.flash_attention: only q_lens=None, non-causal, no dropout, global window "
                                  "(the configuration WanModel uses) is implemented")
    assert dtype == torch.bfloat16, "the tcgen05 kernels are bf16"
    assert q.size(-1) == 128, "head_dim 128 only"
    out_dtype = q.dtype
    if q_scale is not None:
        q = q * q_scale
    scale = float(softmax_scale) if softmax_scale is not None else 1.0 / math.sqrt(q.size(-1))
    outs = []
    for i in range(q.size(0)):
        lk = k.size(1) if k_lens is None else int(k_lens[i])
        qi, ki, vi = (t[i].to(torch.bfloat16) for t in (q, k[:, :lk], v[:, :lk]))
        qi, ki, vi = (t if t.stride(2) == 1 else t.contiguous() for t in (qi, ki, vi))
        if torch.is_grad_enabled() and (q.requires_grad or k.requires_grad or v.requires_grad):
            outs.append(_FlashAttnFn.apply(qi, ki, vi, scale))
        else:
            outs.append(ops.attn_fwd(qi, ki, vi, scale=scale))
    return torch.stack(outs).type(out_dtype)

"""PAVRM reward head with the reference's API (diffusers_lite/utils/network.py:8-152): `QueryAttention`,
`MLP`, `forward_mlp`, `forward_siamese`; same parameter names (`multihead_attn.in_proj_weight`, `queries`,
`fc1..3`).

QueryAttention with its single learnable query is evaluated as two streaming passes over the features
(prfl_sq_pool_fwd) instead of the [L, C] x [C, 2C] key/value in-projection GEMM and an hd-640 attention:
    scores[l, h] = x_l . (Wk_h^T q_h) / sqrt(hd)          (q_h . bk_h is constant in l and cancels)
    out_h        = Wv_h (sum_l softmax(scores)_l x_l) + bv_h
which is algebraically identical to nn.MultiheadAttention with one query (network.py:80-85).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .model import no_autocast


class _SqPool(torch.autograd.Function):
    """pooled[h] = sum_l softmax_l(x_l . wk_eff[h]) x_l   with hand-written forward and backward kernels."""

    @staticmethod
    def forward(ctx, x, wk_eff):
        pooled, scores, stats = ops.sq_pool(x, wk_eff)
        ctx.save_for_backward(x, wk_eff, scores, stats, pooled)
        return pooled

    @staticmethod
    def backward(ctx, dpooled):
        x, wk_eff, scores, stats, pooled = ctx.saved_tensors
        need_w = ctx.needs_input_grad[1]
        dx, ds = ops.sq_pool_bwd(x, wk_eff, scores, stats, pooled, dpooled.float().contiguous(), need_ds=need_w)
        dwk = None
        if need_w:
            dwk = ds.t().contiguous() @ x           # [8, L] x [L, C]: 16*L*C flops, negligible
        return (dx if ctx.needs_input_grad[0] else None), dwk


def _sp_pool(x_local: torch.Tensor, wk_eff: torch.Tensor) -> torch.Tensor:
    """Softmax pooling over the tokens of ALL sequence-parallel ranks from per-rank partial poolings:
    pooled = sum_r w_r pooled_r / sum_r w_r with w_r = sum_r * exp(max_r - max)."""
    import torch.distributed as dist
    from .parallel import get_sequence_parallel_state, nccl_info
    pooled, _, stats = ops.sq_pool(x_local, wk_eff)                  # pooled [nh, C] (local softmax), stats [max | sum]
    if not get_sequence_parallel_state():
        return pooled
    nh, C = pooled.shape
    packed = torch.cat([pooled, stats.view(2, nh).t()], dim=1).contiguous()        # [nh, C + 2]
    allp = torch.empty((nccl_info.sp_size,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
    dist.all_gather_into_tensor(allp.view(-1), packed.view(-1), group=nccl_info.group)
    return merge_partial_poolings(allp)


def merge_partial_poolings(allp: torch.Tensor) -> torch.Tensor:
    """allp [P, nh, C + 2]: per rank the locally normalised pooling [nh, C], the local score max and the local sum of
    exp(score - max).  Returns the pooling under the softmax over the union of all ranks' tokens, [nh, C]."""
    C = allp.shape[2] - 2
    mx, sm = allp[:, :, C], allp[:, :, C + 1]                        # [P, nh]
    w = sm * torch.exp(mx - mx.max(dim=0, keepdim=True).values)
    return (allp[:, :, :C] * w.unsqueeze(-1)).sum(0) / w.sum(0).unsqueeze(-1)


class QueryAttention(nn.Module):
    """network.py:8-110."""

    def __init__(self, feature_dim, num_queries=1, num_heads=8, dropout=0.1, layer_norm=False, return_type=None,
                 product_text=False, text_dim=768):
        super().__init__()
        assert not layer_norm and not product_text, "only the configuration the trainers use is on the path"
        assert num_queries == 1, "the shipped configs use one learnable query (configs/*.yaml lrm.query_attention)"
        self.feature_dim = feature_dim
        self.num_queries = num_queries
        self.num_heads = num_heads
        self.layer_norm = layer_norm
        self.return_type = return_type
        self.product_text = product_text
        self.multihead_attn = nn.MultiheadAttention(embed_dim=feature_dim, num_heads=num_heads, dropout=dropout, batch_first=True)
        self.queries = nn.Parameter(torch.randn(num_queries, feature_dim))
        nn.init.xavier_uniform_(self.queries)

    @no_autocast
    def forward(self, x, e=None, text=None, sp_local: bool = False):
        """x: [B, L, C] or [n_sel, B, L, C] fp32 features (network.py:44-110).
        `sp_local=True` (no-grad scoring under Ulysses sequence parallelism): x holds only THIS rank's token chunk
        [.., L/P, C]; every rank pools its chunk and the per-rank (max, sum, pooled) triples are merged exactly
        (softmax over the union of the chunks) through one tiny all-gather — instead of all-gathering the fp32 features
        (671 MB at 480P) and repeating the identical pooling on every rank as the reference does."""
        assert not (self.training and self.multihead_attn.dropout > 0), "dropout in the pooling is not supported"
        shape = x.shape
        if x.dim() == 2:
            x = x.unsqueeze(1)
        elif x.dim() == 4:
            x = x.reshape(shape[0] * shape[1], shape[2], shape[3])
        bsz, L, C = x.shape
        nh, hd = self.num_heads, C // self.num_heads
        mha = self.multihead_attn
        w, b = mha.in_proj_weight.float(), mha.in_proj_bias.float()
        queries = self.queries.float()
        if e is not None:
            queries = queries + e
        q = F.linear(queries, w[:C], b[:C]).view(nh, hd)                                   # [nh, hd]
        wk_eff = torch.einsum("hd,hdc->hc", q, w[C:2 * C].view(nh, hd, C)) / math.sqrt(hd)  # [nh, C]
        wv, bv = w[2 * C:].view(nh, hd, C), b[2 * C:]
        outs = []
        for i in range(bsz):
            if sp_local:
                assert not torch.is_grad_enabled(), "sp_local pooling is the no-grad scoring path"
                pooled = _sp_pool(x[i].float().contiguous(), wk_eff.contiguous())
            else:
                pooled = _SqPool.apply(x[i].float().contiguous(), wk_eff.contiguous())     # [nh, C]
            o = torch.einsum("hc,hdc->hd", pooled, wv).reshape(1, C) + bv
            outs.append(F.linear(o, mha.out_proj.weight.float(), mha.out_proj.bias.float()))
        out = torch.cat(outs)                                                              # [bsz, C]
        if len(shape) == 4:
            out = out.view(shape[0], bsz // shape[0], -1).mean(dim=0)
        if self.return_type == "query":
            out = out + queries.unsqueeze(0).expand(bsz, -1, -1)                           # [B, 1, C] (network.py:73-75,103-104: incl. + e)
        return out


class MLP(nn.Module):
    """network.py:112-134 — 3 tiny fp32 Linears on a [B, 1, C] vector; kept in PyTorch."""

    def __init__(self, input_dim):
        super().__init__()
        self.fc1 = nn.Linear(input_dim, 1024)
        self.fc2 = nn.Linear(1024, 512)
        self.fc3 = nn.Linear(512, 1)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        x = torch.relu(self.fc1(x))
        x = torch.relu(self.fc2(x))
        return self.fc3(x)


def forward_mlp(model, input):
    return torch.sigmoid(model(input))


def forward_siamese(model, input1, input2):
    return torch.sigmoid(model(input1) - model(input2))

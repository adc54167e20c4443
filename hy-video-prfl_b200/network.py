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

    def forward(self, x, e=None, text=None):
        """x: [B, L, C] or [n_sel, B, L, C] fp32 features (network.py:44-110)."""
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
            pooled = _SqPool.apply(x[i].float().contiguous(), wk_eff.contiguous())         # [nh, C]
            o = torch.einsum("hc,hdc->hd", pooled, wv).reshape(1, C) + bv
            outs.append(F.linear(o, mha.out_proj.weight.float(), mha.out_proj.bias.float()))
        out = torch.cat(outs)                                                              # [bsz, C]
        if len(shape) == 4:
            out = out.view(shape[0], bsz // shape[0], -1).mean(dim=0)
        if self.return_type == "query":
            out = out + self.queries.float().unsqueeze(0).expand(bsz, -1, -1)              # [B, 1, C] (network.py:103-104)
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

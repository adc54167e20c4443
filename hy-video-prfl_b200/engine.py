"""Block engine: the sequence of C-ABI calls that is one WanAttentionBlock forward (model.py:320-359 of the
reference) and its hand-written backward, plus the autograd glue.

Training keeps only each block's fp32 input (per-block activation checkpointing, as the reference's
`apply_fsdp_checkpointing`, fsdp_utils.py:23-50); the backward re-runs the forward with a stash of the
intermediates and then walks the chain backwards with the backward kernels:

  FFN      : gate_bwd -> GEMM(dgrad, dGELU epilogue) + GEMM(wgrad) -> GEMM(dgrad) + GEMM(wgrad) -> ln_mod_bwd
  cross    : cast -> GEMM(dgrad o) + wgrad -> attn_bwd -> rmsnorm_bwd(q), rmsnorm_bwd(k) -> GEMMs -> ln_mod_bwd (affine)
  self     : gate_bwd -> GEMM(dgrad o) + wgrad -> [a2a] attn_bwd [a2a] -> rmsnorm_rope_bwd(q,k) -> GEMM(dgrad QKV) + wgrad
             -> ln_mod_bwd
Weight gradients are fp32 (GEMM EPI_F32), activations' gradients bf16, the residual-stream gradient fp32 (in place).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .parallel import (get_p2p_ulysses, get_sequence_parallel_state, nccl_info, ulysses_gather_tokens,
                       ulysses_scatter_tokens)
from .rope import rope_tables

T5_CONTEXT_TOKEN_NUMBER = 512

# Activation checkpointing granularity of the training path.  The reference recomputes the WHOLE block in backward
# (fsdp_utils.py:23-50).  On a 180 GB B200 the self-attention output (bf16, [L, H/P, 128] in the attention layout) and its LSE
# are cheap to keep — 98 MB per block at 720P on 8 GPUs, 3.9 GB for 40 blocks — and they let the recompute skip the one
# kernel that is 2/3 of a block forward: "selective" (default) saves them in the forward that records the graph and reuses
# them in backward (bit-identical gradients: the recompute would produce the same bytes); PRFL_CKPT=full restores the
# reference's full recompute.
import os as _os
SAVE_ATTENTION = _os.environ.get("PRFL_CKPT", "selective").lower() != "full"


def _f(p):
    return p.detach().float()


# ------------------------------------------------------------------------------------------------
# forward
# ------------------------------------------------------------------------------------------------
def block_forward(blk, x: torch.Tensor, em: torch.Tensor, ctx: torch.Tensor, seq_len_i: int, grid: Tuple[int, int, int],
                  first_block: bool, stash: Optional[Dict] = None, sample: int = 0, save_attn: Optional[list] = None,
                  reuse_attn: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
    """One sample through one block.  x: [M, C] fp32 (updated in place), em: [6, C] fp32 (modulation + e),
    ctx: [Lc, C] bf16.  With `stash` (a dict) every intermediate the backward needs is kept; x is then READ ONLY: the two
    intermediate residual states are written to fresh buffers by the residual-epilogue GEMMs (`resid=`), no clones.
    `save_attn` (a list, graph-recording forward): receives (self-attention output in the attention layout, LSE) so that the
    backward's recompute can skip the attention kernel; `reuse_attn` (recompute): that pair."""
    sa, ca = blk.self_attn, blk.cross_attn
    M, C = x.shape
    n, d = sa.num_heads, sa.head_dim
    keep = stash is not None
    sp = get_sequence_parallel_state()
    P = nccl_info.sp_size if sp else 1
    rank = nccl_info.rank_within_group if sp else 0

    # ---- self-attention ----
    if keep:
        h1, mean1, rstd1 = ops.ln_mod(x, em[0], em[1], None, None, blk.eps, round_bf16=first_block, save_stats=True)
    else:
        h1 = ops.ln_mod(x, em[0], em[1], None, None, blk.eps, round_bf16=first_block)
    wqkv, bqkv = sa._qkv_operands()
    qkv = ops.gemm(h1, wqkv, bias=bqkv, epi=ops.EPI_BF16)                                   # [M, 3C]
    cos, sin = rope_tables(grid, x.device, d, pad_to=M * P)
    n_rot = M if sp else min(seq_len_i, M)
    pos0 = rank * M
    qk_raw = qkv[:, :2 * C].clone() if keep else None
    wq_n, wk_n = _f(sa.norm_q.weight), _f(sa.norm_k.weight)
    if keep:
        _, rstd_q = ops.rmsnorm_rope_(qkv[:, :C], wq_n, cos, sin, sa.eps, n_rot, pos0, save_rstd=True)
        _, rstd_k = ops.rmsnorm_rope_(qkv[:, C:2 * C], wk_n, cos, sin, sa.eps, n_rot, pos0, save_rstd=True)
    else:
        ops.rmsnorm_rope_(qkv[:, :C], wq_n, cos, sin, sa.eps, n_rot, pos0)
        ops.rmsnorm_rope_(qkv[:, C:2 * C], wk_n, cos, sin, sa.eps, n_rot, pos0)
    q3, k3, v3 = (qkv[:, j * C:(j + 1) * C].unflatten(1, (n, d)) for j in range(3))
    klen = seq_len_i
    lse1 = None
    want_lse = keep or save_attn is not None
    if not sp:
        if reuse_attn is not None:
            a1, lse1 = reuse_attn
        elif want_lse:
            a1, lse1 = ops.attn_fwd(q3, k3[:klen], v3[:klen], need_lse=True)
        else:
            a1 = ops.attn_fwd(q3, k3[:klen], v3[:klen])
        if save_attn is not None:
            save_attn.append((a1, lse1))
        qg = kg = vg = og = None
    elif nccl_info.ring_degree > 1:
        # Ulysses x Ring (xdit_context_parallel.py:190-233): inference only in the reference, no-grad only here
        if keep:
            raise RuntimeError("Ulysses x Ring sequence parallelism is a no-grad (inference) path; train with ring_degree == 1")
        from .parallel import usp_attention
        a1 = usp_attention(q3, k3, v3, klen)
        qg = kg = vg = og = None
    else:
        p2p = get_p2p_ulysses(M * P, n, x.device)
        if p2p is not None and not want_lse:
            # no-grad forward: both exchanges are peer stores fused into our own kernels (no NCCL, no staging copies)
            a1 = p2p.attention(q3, k3, v3, klen)                                             # [M, n, d] (symmetric buffer)
            qg = kg = vg = og = None
        elif p2p is not None:
            # graph-recording forward / recompute of a checkpointed block: the same peer stores, but q / k / v / o stay on this
            # rank in the attention layout for the backward (views of the symmetric buffers: valid until the next exchange)
            qg, kg, vg = p2p.scatter_qkv(q3, k3, v3)                                         # [L, n/P, d]
            if reuse_attn is not None:
                og, lse1 = reuse_attn
            else:
                og, lse1 = ops.attn_fwd(qg, kg[:klen], vg[:klen], need_lse=True)
            a1 = p2p.gather_out(og)                                                          # [M, n, d] (symmetric buffer)
        else:
            qg, kg, vg = (ulysses_scatter_tokens(t, P) for t in (q3, k3, v3))                # [L, n/P, d]
            if reuse_attn is not None:
                og, lse1 = reuse_attn
            elif want_lse:
                og, lse1 = ops.attn_fwd(qg, kg[:klen], vg[:klen], need_lse=True)
            else:
                og = ops.attn_fwd(qg, kg[:klen], vg[:klen])
            a1 = ulysses_gather_tokens(og, P)                                                # [M, n, d]
        if save_attn is not None:
            save_attn.append((og, lse1))
    a1 = a1.reshape(M, C)
    wo, bo = sa.o.operands()
    y1 = torch.empty(M, C, dtype=torch.bfloat16, device=x.device) if keep else None
    if keep:
        x_in = x
        x = ops.gemm(a1, wo, bias=bo, epi=ops.EPI_RESIDUAL, resid=x_in, gate=em[2], aux=y1)    # x1 = x_in + e2 * o(attn)
    else:
        x_in = None
        ops.gemm(a1, wo, bias=bo, epi=ops.EPI_RESIDUAL, out=x, gate=em[2], aux=y1)           # x += e2 * o(attn)

    # ---- cross-attention ----
    x1 = x if keep else None
    if blk.cross_attn_norm:
        g3, b3 = _f(blk.norm3.weight), _f(blk.norm3.bias)
        if keep:
            h3, mean3, rstd3 = ops.ln_mod(x, None, None, g3, b3, blk.eps, save_stats=True)
        else:
            h3 = ops.ln_mod(x, None, None, g3, b3, blk.eps)
    else:
        h3, mean3, rstd3 = ops.cast_bf16(x), None, None
    wcq, bcq = ca.q.operands()
    q2 = ops.gemm(h3, wcq, bias=bcq, epi=ops.EPI_BF16)
    q2_raw = q2.clone() if keep else None
    wq2_n = _f(ca.norm_q.weight)
    rstd_q2 = None
    if keep:
        _, rstd_q2 = ops.rmsnorm_rope_(q2, wq2_n, None, None, ca.eps, save_rstd=True)
    else:
        ops.rmsnorm_rope_(q2, wq2_n, None, None, ca.eps)
    q2_3 = q2.unflatten(1, (n, d))
    groups = []          # (names, norm module, ctx slice)
    if hasattr(ca, "k_img"):
        n_img = ctx.shape[0] - T5_CONTEXT_TOKEN_NUMBER
        groups.append((("k_img", "v_img"), ca.norm_k_img, ctx[:n_img].contiguous()))
        groups.append((("k", "v"), ca.norm_k, ctx[n_img:].contiguous()))
    else:
        groups.append((("k", "v"), ca.norm_k, ctx))
    a2 = None
    cross = []
    kv_cache = None if keep else ca.__dict__.get("_kv_cache")    # PreparedContext (model.py): K/V of a constant prompt
    for names, norm, c_in in groups:
        kv = None if kv_cache is None else kv_cache.get((names, sample))
        k_raw = rstd_kc = None
        if kv is None:
            wkv, bkv = ca._kv_operands(names)
            kv = ops.gemm(c_in, wkv, bias=bkv, epi=ops.EPI_BF16)                             # [Lc, 2C]
            k_raw = kv[:, :C].clone() if keep else None
            if keep:
                _, rstd_kc = ops.rmsnorm_rope_(kv[:, :C], _f(norm.weight), None, None, ca.eps, save_rstd=True)
            else:
                ops.rmsnorm_rope_(kv[:, :C], _f(norm.weight), None, None, ca.eps)
            if kv_cache is not None:
                kv_cache[(names, sample)] = kv
        kc, vc = kv[:, :C].unflatten(1, (n, d)), kv[:, C:].unflatten(1, (n, d))
        if keep:
            o_g, lse_g = ops.attn_fwd(q2_3, kc, vc, need_lse=True)
            cross.append(dict(names=names, norm=norm, c_in=c_in, kv=kv, k_raw=k_raw, rstd=rstd_kc, o=o_g, lse=lse_g))
        else:
            o_g = ops.attn_fwd(q2_3, kc, vc)
        a2 = o_g if a2 is None else a2 + o_g                                                 # model.py:269 (bf16 add)
    a2 = a2.reshape(M, C)
    wco, bco = ca.o.operands()
    if keep:
        x = ops.gemm(a2, wco, bias=bco, epi=ops.EPI_RESIDUAL, resid=x1)                      # x2 = x1 + o(cross)
    else:
        ops.gemm(a2, wco, bias=bco, epi=ops.EPI_RESIDUAL, out=x)                             # x += o(cross)

    # ---- FFN ----
    x2 = x if keep else None
    if keep:
        h2, mean2, rstd2 = ops.ln_mod(x, em[3], em[4], None, None, blk.eps, save_stats=True)
    else:
        h2 = ops.ln_mod(x, em[3], em[4], None, None, blk.eps)
    w1, b1 = blk.ffn[0].operands()
    w2, b2 = blk.ffn[2].operands()
    u = torch.empty(M, blk.ffn_dim, dtype=torch.bfloat16, device=x.device) if keep else None
    f = ops.gemm(h2, w1, bias=b1, epi=ops.EPI_BF16_GELU, aux=u)
    if keep:
        # the block output itself is not needed by the backward: stop here (ffn.2 is re-run there for the gate grad)
        stash.update(x_in=x_in, h1=h1, mean1=mean1, rstd1=rstd1, qkv=qkv, qk_raw=qk_raw, rstd_q=rstd_q, rstd_k=rstd_k,
                     cos=cos, sin=sin, n_rot=n_rot, pos0=pos0, klen=klen, a1=a1, lse1=lse1, qg=qg, kg=kg, vg=vg, og=og,
                     y1=y1, x1=x1, h3=h3, mean3=mean3, rstd3=rstd3, q2=q2, q2_raw=q2_raw, rstd_q2=rstd_q2, cross=cross,
                     a2=a2, x2=x2, h2=h2, mean2=mean2, rstd2=rstd2, u=u, f=f, em=em, P=P)
        return x
    ops.gemm(f, w2, bias=b2, epi=ops.EPI_RESIDUAL, out=x, gate=em[5])                        # x += e5 * ffn
    return x


# ------------------------------------------------------------------------------------------------
# backward
# ------------------------------------------------------------------------------------------------
def _wgrad(dy: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """dW[n, k] = sum_m dy[m, n] x[m, k]  (fp32)."""
    return ops.gemm(dy, x, a_trans=True, b_trans=True, epi=ops.EPI_F32)


def block_backward(blk, st: Dict, dx: torch.Tensor, need_ctx_grad: bool, need_w: bool = True, need_e: bool = True, sink=None):
    """dx: [M, C] fp32 gradient w.r.t. the block output; overwritten with the gradient w.r.t. the block input.
    Returns (grads: name -> fp32 tensor, dem [6, C], dctx [Lc, C] fp32 or None).  With a gradient `sink`
    (sharding._BlockSink) the weight-gradient GEMMs write (or accumulate) straight into the sink's flat fp32 buffer —
    fused operands (QKV, context K/V) as ONE [3C, C] / [2C, C] GEMM output — and only the small gradients are returned."""
    sa, ca = blk.self_attn, blk.cross_attn
    M, C = dx.shape
    n, d = sa.num_heads, sa.head_dim
    em = st["em"]
    g: Dict[str, torch.Tensor] = {}
    # frozen blocks (PRFL's reward model, Appendix B item 10): dgrad only, every weight-gradient kernel is skipped
    colsum = ops.colsum if need_w else (lambda a: None)

    def wgrad(names, dy, x):
        """dW of the (fused) weight whose row blocks are `names`."""
        if not need_w:
            return
        if sink is not None:
            out, beta = sink.matrix_out(names)
            ops.gemm(dy, x, a_trans=True, b_trans=True, epi=ops.EPI_F32, out=out, beta=beta)
            return
        dw = _wgrad(dy, x)
        if len(names) == 1:
            g[names[0]] = dw
        else:
            r = dw.shape[0] // len(names)
            for j, nm in enumerate(names):
                g[nm] = dw[j * r:(j + 1) * r]

    # ---- FFN ----
    w1, b1 = blk.ffn[0].operands()
    w2, b2 = blk.ffn[2].operands()
    # the gate / shift / scale gradients feed `modulation` and the time embedding: skipped for frozen blocks whose
    # conditioning needs no gradient (the reward model in PRFL) — that also saves re-running ffn.2 for the gate gradient
    need_mod = need_w or need_e
    y2 = ops.gemm(st["f"], w2, bias=b2, epi=ops.EPI_BF16) if need_mod else None
    dy2, de5 = ops.gate_bwd(dx, y2, em[5])
    del y2
    du = ops.gemm(dy2, w2, b_trans=True, epi=ops.EPI_BF16_DGELU, aux=st["u"])                # [M, ffn]
    wgrad(("ffn.2.weight",), dy2, st["f"])
    g["ffn.2.bias"] = colsum(dy2)
    del dy2
    dh2 = ops.gemm(du, w1, b_trans=True, epi=ops.EPI_BF16)                                   # [M, C]
    wgrad(("ffn.0.weight",), du, st["h2"])
    g["ffn.0.bias"] = colsum(du)
    del du
    dsh2, dsc2 = ops.ln_mod_bwd(st["x2"], dh2, em[4], None, st["mean2"], st["rstd2"], dx, need_mod)
    del dh2

    # ---- cross-attention ----
    dyc, _ = ops.gate_bwd(dx, None, None)                                                     # bf16 cast of dx
    wco, _ = ca.o.operands()
    da2 = ops.gemm(dyc, wco, b_trans=True, epi=ops.EPI_BF16)
    wgrad(("cross_attn.o.weight",), dyc, st["a2"])
    g["cross_attn.o.bias"] = colsum(dyc)
    del dyc
    q2_3 = st["q2"].unflatten(1, (n, d))
    da2_3 = da2.unflatten(1, (n, d))
    dq2 = None
    dctx_parts = []
    for cg in st["cross"]:
        kv = cg["kv"]
        Lc = kv.shape[0]
        kc, vc = kv[:, :C].unflatten(1, (n, d)), kv[:, C:].unflatten(1, (n, d))
        dkv = torch.empty(Lc, 2 * C, dtype=torch.bfloat16, device=dx.device)
        dq_g, _, _ = ops.attn_bwd(q2_3, kc, vc, cg["o"], da2_3, cg["lse"], dk=dkv[:, :C].unflatten(1, (n, d)),
                                  dv=dkv[:, C:].unflatten(1, (n, d)))
        dq2 = dq_g if dq2 is None else dq2 + dq_g
        kn, vn = cg["names"]
        nname = "cross_attn.norm_k_img.weight" if kn == "k_img" else "cross_attn.norm_k.weight"
        g[nname] = ops.rmsnorm_rope_bwd_(cg["k_raw"], _f(cg["norm"].weight), None, None, dkv[:, :C], cg["rstd"], need_dw=need_w)
        wkv, _ = ca._kv_operands(cg["names"])
        if need_w:
            wgrad((f"cross_attn.{kn}.weight", f"cross_attn.{vn}.weight"), dkv, cg["c_in"])   # [2C, C]
            dbkv = colsum(dkv)
            g[f"cross_attn.{kn}.bias"], g[f"cross_attn.{vn}.bias"] = dbkv[:C], dbkv[C:]
        if need_ctx_grad:
            dctx_parts.append(ops.gemm(dkv, wkv, b_trans=True, epi=ops.EPI_F32))              # [Lc, C]
    dq2 = dq2.reshape(M, C)
    g["cross_attn.norm_q.weight"] = ops.rmsnorm_rope_bwd_(st["q2_raw"], _f(ca.norm_q.weight), None, None, dq2, st["rstd_q2"], need_dw=need_w)
    wcq, _ = ca.q.operands()
    dh3 = ops.gemm(dq2, wcq, b_trans=True, epi=ops.EPI_BF16)
    wgrad(("cross_attn.q.weight",), dq2, st["h3"])
    g["cross_attn.q.bias"] = colsum(dq2)
    del dq2, da2
    if blk.cross_attn_norm:
        db3, dg3 = ops.ln_mod_bwd(st["x1"], dh3, None, _f(blk.norm3.weight), st["mean3"], st["rstd3"], dx, need_w)
        g["norm3.weight"], g["norm3.bias"] = dg3, db3
    else:
        dx += dh3.float()
    del dh3

    # ---- self-attention ----
    dy1, de2 = ops.gate_bwd(dx, st["y1"] if need_mod else None, em[2])
    wo, _ = sa.o.operands()
    da1 = ops.gemm(dy1, wo, b_trans=True, epi=ops.EPI_BF16)
    wgrad(("self_attn.o.weight",), dy1, st["a1"])
    g["self_attn.o.bias"] = colsum(dy1)
    del dy1
    qkv = st["qkv"]
    klen, P = st["klen"], st["P"]
    q3, k3, v3 = (qkv[:, j * C:(j + 1) * C].unflatten(1, (n, d)) for j in range(3))
    if P == 1:
        dqkv = (torch.zeros if klen < M else torch.empty)(M, 3 * C, dtype=torch.bfloat16, device=dx.device)
        dq3, dk3, dv3 = (dqkv[:, j * C:(j + 1) * C].unflatten(1, (n, d)) for j in range(3))
        ops.attn_bwd(q3, k3[:klen], v3[:klen], st["a1"].unflatten(1, (n, d)), da1.unflatten(1, (n, d)), st["lse1"],
                     dq=dq3, dk=dk3[:klen], dv=dv3[:klen])
    else:
        p2p = get_p2p_ulysses(M * P, n, dx.device)
        da1_3 = da1.unflatten(1, (n, d))
        dog = p2p.scatter_grad(da1_3) if p2p is not None else ulysses_scatter_tokens(da1_3, P)   # [L, n/P, d]
        L = dog.shape[0]
        dkg = (torch.zeros if klen < L else torch.empty)(L, n // P, d, dtype=torch.bfloat16, device=dx.device)
        dvg = torch.zeros_like(dkg) if klen < L else torch.empty_like(dkg)
        dqg, _, _ = ops.attn_bwd(st["qg"], st["kg"][:klen], st["vg"][:klen], st["og"], dog, st["lse1"], dk=dkg[:klen], dv=dvg[:klen])
        if p2p is not None:
            dqkv = p2p.gather_grads(dqg, dkg, dvg)                                            # [M, 3C] symmetric buffer, written by the peers
        else:
            dqkv = torch.empty(M, 3 * C, dtype=torch.bfloat16, device=dx.device)
            for j, t in enumerate((dqg, dkg, dvg)):                                           # unpacked straight into the fused buffer
                ulysses_gather_tokens(t, P, out=dqkv[:, j * C:(j + 1) * C].unflatten(1, (n, d)))
    del da1
    qk_raw = st["qk_raw"]
    g["self_attn.norm_q.weight"] = ops.rmsnorm_rope_bwd_(qk_raw[:, :C], _f(sa.norm_q.weight), st["cos"], st["sin"], dqkv[:, :C],
                                                        st["rstd_q"], st["n_rot"], st["pos0"], need_dw=need_w)
    g["self_attn.norm_k.weight"] = ops.rmsnorm_rope_bwd_(qk_raw[:, C:], _f(sa.norm_k.weight), st["cos"], st["sin"], dqkv[:, C:2 * C],
                                                        st["rstd_k"], st["n_rot"], st["pos0"], need_dw=need_w)
    wqkv, _ = sa._qkv_operands()
    dh1 = ops.gemm(dqkv, wqkv, b_trans=True, epi=ops.EPI_BF16)                               # [M, C]
    if need_w:
        wgrad(("self_attn.q.weight", "self_attn.k.weight", "self_attn.v.weight"), dqkv, st["h1"])   # [3C, C]
        dbqkv = colsum(dqkv)
        for j, nm in enumerate(("q", "k", "v")):
            g[f"self_attn.{nm}.bias"] = dbqkv[j * C:(j + 1) * C]
    del dqkv
    dsh1, dsc1 = ops.ln_mod_bwd(st["x_in"], dh1, em[1], None, st["mean1"], st["rstd1"], dx, need_mod)
    dem = torch.stack([dsh1, dsc1, de2, dsh2, dsc2, de5]) if need_mod else None              # [6, C]
    dctx = None
    if need_ctx_grad:
        dctx = dctx_parts[0] if len(dctx_parts) == 1 else torch.cat(dctx_parts, dim=0)
    return g, dem, dctx


# ------------------------------------------------------------------------------------------------
# autograd glue
# ------------------------------------------------------------------------------------------------
def block_param_names(blk) -> List[str]:
    return [n for n, _ in blk.named_parameters()]


class BlockFn(torch.autograd.Function):
    """x_out = block(x); saves only x (per-block activation checkpointing) and recomputes in backward."""

    @staticmethod
    def forward(ctx, x, e, context, blk, seq_lens, grids, first_block, names, *params):
        ctx.blk, ctx.seq_lens, ctx.grids, ctx.first, ctx.names = blk, seq_lens, grids, first_block, names
        out = x.detach().clone()
        em = (_f(blk.modulation) + e.detach()).contiguous()
        cb = context.detach()
        cb = cb if cb.dtype == torch.bfloat16 else cb.to(torch.bfloat16)
        saved = [] if (SAVE_ATTENTION and nccl_info.ring_degree == 1) else None
        for i in range(x.shape[0]):
            block_forward(blk, out[i], em[i], cb[i].contiguous(), int(seq_lens[i]), grids[i], first_block, save_attn=saved)
        ctx.n_saved = 0 if saved is None else len(saved)
        flat = [] if saved is None else [t for pair in saved for t in pair]
        ctx.save_for_backward(x, e, context, *flat)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, e, context = ctx.saved_tensors[:3]
        attn_saved = ctx.saved_tensors[3:]
        blk = ctx.blk
        B = x.shape[0]
        dx = dout.detach().float().contiguous().clone()
        em = (_f(blk.modulation) + e.detach()).contiguous()
        cb = context.detach()
        cb = cb if cb.dtype == torch.bfloat16 else cb.to(torch.bfloat16)
        need_ctx = ctx.needs_input_grad[2]
        need_w = any(ctx.needs_input_grad[8:])
        need_e = ctx.needs_input_grad[1]
        unit = blk.__dict__.get("_prfl_unit")
        sink = unit.sink if (unit is not None and need_w) else None     # resident bf16 block owned by a ShardedAdamW
        if need_w and unit is not None and sink is None:
            raise RuntimeError("a resident bf16 block has no fp32 gradient path through autograd: attach a sharding.ShardedAdamW "
                               "(its gradient sink receives the weight gradients) or freeze the block")
        if sink is not None:
            sink.begin()
        tot: Dict[str, torch.Tensor] = {}
        dems, dctxs = [], []
        for i in range(B):
            st: Dict = {}
            reuse = (attn_saved[2 * i], attn_saved[2 * i + 1]) if ctx.n_saved else None
            block_forward(blk, x[i].detach(), em[i], cb[i].contiguous(), int(ctx.seq_lens[i]), ctx.grids[i], ctx.first, st, reuse_attn=reuse)
            gi, dem, dctx = block_backward(blk, st, dx[i], need_ctx, need_w, need_e, sink)
            del st
            for k, v in gi.items():
                if v is not None:
                    tot[k] = v if k not in tot else tot[k] + v
            dems.append(dem)
            dctxs.append(dctx)
        de = torch.stack(dems) if dems[0] is not None else None                               # [B, 6, C]
        if need_w:
            tot["modulation"] = de.sum(0, keepdim=True)
        if not need_e:
            de = None
        dctx = torch.stack(dctxs).to(context.dtype) if need_ctx else None
        pg = []
        if sink is not None:
            for nm, gr in tot.items():                                                        # biases, norm weights, modulation -> root unit
                if gr is not None:
                    sink.small(nm, gr)
            sink.end()                                                                        # reduce-scatter on the side stream
            pg = [None] * len(ctx.names)
        else:
            for nm, p in zip(ctx.names, ctx.blk.parameters()):
                gr = tot.get(nm)
                pg.append(None if gr is None else gr.reshape(p.shape).to(p.dtype))
        return (dx, de, dctx, None, None, None, None, None, *pg)


class HeadFn(torch.autograd.Function):
    """Head of one sample (model.py:379-389: LayerNorm -> * (1 + e1) + e0 -> Linear(dim -> prod(patch) * out_dim), fp32 in the
    reference) with fp32-grade products on the bf16 tensor pipe: h = hi + lo and W = Whi + Wlo are split into bf16 pairs and
    hi Whi + hi Wlo + lo Whi is accumulated in fp32 (relative error ~1e-5); the same split gives the weight gradient
    dW = dy^T h.  The input gradient goes through bf16 (dh = bf16(dy) W) into the LayerNorm backward kernel, as every other
    activation gradient on the path does.  Backward recomputes the normalised activations instead of storing them."""

    @staticmethod
    def forward(ctx, x, shift, scale, weight, bias, head):
        xf, sh, sc = (t.detach().float().contiguous() for t in (x, shift, scale))
        w_hi, w_lo = head._split_operands()
        hi, lo = ops.ln_mod_split(xf, sh, sc, head.eps)
        o = ops.gemm(hi, w_hi, bias=bias.detach().float().contiguous(), epi=ops.EPI_F32)
        ops.gemm(hi, w_lo, epi=ops.EPI_F32, out=o, beta=True)
        ops.gemm(lo, w_hi, epi=ops.EPI_F32, out=o, beta=True)
        ctx.save_for_backward(xf, sh, sc)
        ctx.head = head
        return o

    @staticmethod
    def backward(ctx, do):
        xf, sh, sc = ctx.saved_tensors
        head = ctx.head
        need_x, need_sh, need_sc, need_w, need_b = ctx.needs_input_grad[:5]
        do = do.float().contiguous()                                     # [L, 64] fp32
        d_hi = ops.cast_bf16(do)
        dW = db = dx = dsh = dsc = None
        if need_w or need_b:
            d_lo = ops.cast_bf16(do - d_hi.float())
            if need_w:
                hi, lo = ops.ln_mod_split(xf, sh, sc, head.eps)
                dW = _wgrad(d_hi, hi)
                ops.gemm(d_hi, lo, a_trans=True, b_trans=True, epi=ops.EPI_F32, out=dW, beta=True)
                ops.gemm(d_lo, hi, a_trans=True, b_trans=True, epi=ops.EPI_F32, out=dW, beta=True)
                del hi, lo
            if need_b:
                db = ops.colsum(d_hi) + ops.colsum(d_lo)
        if need_x or need_sh or need_sc:
            w_hi, _ = head._split_operands()
            dh = ops.gemm(d_hi, w_hi, b_trans=True, epi=ops.EPI_BF16)    # [L, C] = dy . W   (W as the [K = 64, N = C] operand)
            _, mean, rstd = ops.ln_mod(xf, sh, sc, eps=head.eps, save_stats=True)
            dx = torch.zeros_like(xf)
            dsh, dsc = ops.ln_mod_bwd(xf, dh, sc, None, mean, rstd, dx, need_sh or need_sc)
        return dx, dsh, dsc, dW, db, None


class LinearFn(torch.autograd.Function):
    """y = bf16(x W^T + b) through the tcgen05 GEMM, with dgrad / wgrad / bias-grad kernels in backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, lin):
        w, b = lin.operands()
        x2 = x.detach().reshape(-1, x.shape[-1])
        x2 = (x2 if x2.dtype == torch.bfloat16 else x2.to(torch.bfloat16)).contiguous()
        ctx.save_for_backward(x2)
        ctx.lin, ctx.shape, ctx.xdtype = lin, x.shape, x.dtype
        return ops.gemm(x2, w, bias=b, epi=ops.EPI_BF16).view(*x.shape[:-1], -1)

    @staticmethod
    def backward(ctx, dy):
        (x2,) = ctx.saved_tensors
        w, _ = ctx.lin.operands()
        dy2 = dy.detach().reshape(-1, dy.shape[-1])
        dy2 = (dy2 if dy2.dtype == torch.bfloat16 else dy2.to(torch.bfloat16)).contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gemm(dy2, w, b_trans=True, epi=ops.EPI_BF16).view(ctx.shape).to(ctx.xdtype)
        dw = _wgrad(dy2, x2) if ctx.needs_input_grad[1] else None
        db = ops.colsum(dy2) if (ctx.lin.bias is not None and ctx.needs_input_grad[2]) else None
        return dx, dw, db, None


class PatchEmbedFn(torch.autograd.Function):
    """Conv3d(k = s = (1,2,2)) as gather + GEMM (model.py:578-581); backward = wgrad / bias reductions + scatter."""

    @staticmethod
    def forward(ctx, x, y, weight, bias, model):
        wp, bp = model._patch_operands()
        xf = x.detach().float().contiguous()
        yf = None if y is None else y.detach().float().contiguous()
        patches = ops.patchify(xf, yf)
        ctx.save_for_backward(patches)
        ctx.model, ctx.xshape, ctx.wshape, ctx.xdtype = model, x.shape, weight.shape, x.dtype
        return ops.gemm(patches, wp, bias=bp, epi=ops.EPI_BF16)

    @staticmethod
    def backward(ctx, dy):
        (patches,) = ctx.saved_tensors
        wp, _ = ctx.model._patch_operands()
        dy2 = dy.detach()
        dy2 = (dy2 if dy2.dtype == torch.bfloat16 else dy2.to(torch.bfloat16)).contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dp = ops.gemm(dy2, wp, b_trans=True, epi=ops.EPI_F32)                             # [L, Ctot*4]
            Cx, F, H, W = ctx.xshape
            dx = ops.patchify_bwd(dp, Cx, F, H, W).to(ctx.xdtype)
        dw = _wgrad(dy2, patches).reshape(ctx.wshape) if ctx.needs_input_grad[2] else None
        db = ops.colsum(dy2) if ctx.needs_input_grad[3] else None
        return dx, None, dw, db, None

"""CPU oracle for the Wan-DiT hot path of HY-Video-PRFL.  TEST INFRASTRUCTURE ONLY.

This file is a plain-PyTorch (CPU, fp32 by default) *restatement* of the algorithm the
reference executes on the path BASELINE.json names: WanModel.forward -> WanAttentionBlock
-> self/cross attention -> (features | head -> unpatchify) and the PAVRM reward head
(QueryAttention + MLP).  It is functional (state-dict in, tensors out) rather than a module
tree, so it shares no structure with the reference; every function cites the reference
file:line it follows (paths relative to the reference checkout).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module; the product (`prfl_b200`) never does and fails loudly when its
CUDA library is missing.

Pinning: `tests/golden/make_golden.py` imports the real reference modules (with shims, in the
build container where /root/reference exists), runs them on seeded weights/inputs and commits
the outputs under `tests/golden/*.pt`; `tests/test_oracle_golden.py` checks this oracle against
those fixtures.  The reference ships no tests or golden vectors of its own, so
parity is pinned by "outputs of the reference itself run here".  In the build container
`tests/test_oracle_live_14b_cpu.py` additionally runs the unmodified reference LIVE at the 14B
dimensions BASELINE.json is quoted on (one block + embeddings + head, forward and backward; the
reward head at dim 5120) and holds this oracle to the same bounds.

Precision choreography reproduced: fp32 residual stream, RMSNorm over
the full channel dim with a rounding to the input dtype before the weight multiply, RoPE in
float64, fp32 head and time-embedding path, unmasked cross-attention over 512 padded tokens.
`autocast_dtype=torch.bfloat16` emulates the reference's `torch.autocast(bf16)` Linear / attention
rounding on CPU (Linear inputs+weights rounded to bf16, fp32 accumulate, bf16 output).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

T5_CONTEXT_TOKEN_NUMBER = 512  # model.py:18


@dataclass
class WanConfig:
    """Constructor surface of WanModel (model.py:424-439)."""
    model_type: str = "t2v"
    patch_size: Tuple[int, int, int] = (1, 2, 2)
    text_len: int = 512
    in_dim: int = 16
    dim: int = 2048
    ffn_dim: int = 8192
    freq_dim: int = 256
    text_dim: int = 4096
    out_dim: int = 16
    num_heads: int = 16
    num_layers: int = 32
    window_size: Tuple[int, int] = (-1, -1)
    qk_norm: bool = True
    cross_attn_norm: bool = True
    eps: float = 1e-6

    def kwargs(self) -> dict:
        return dict(self.__dict__)


# ----------------------------------------------------------------------------------------------
# autocast emulation helpers
# ----------------------------------------------------------------------------------------------
class _Prec:
    """Emulates what `torch.autocast(dtype)` does to F.linear on the reference path."""

    def __init__(self, autocast_dtype: Optional[torch.dtype], native: bool = False):
        self.dt = autocast_dtype
        # native=True (bench.py's gpu_baseline leg only): run the low-precision ops for real on the tensors' device —
        # F.linear in `dt` (cuBLAS on CUDA) and flash-attn 2 — instead of emulating their rounding in fp32.  This is
        # the reference's own kernel stack (attention.py:113-127, nn.Linear under autocast) driven by this restatement.
        self.native = native and autocast_dtype is not None

    def linear(self, x, w, b=None):
        if self.native:
            return F.linear(x.to(self.dt), w.to(self.dt), None if b is None else b.to(self.dt))
        if self.dt is None:
            return F.linear(x.to(w.dtype) if x.dtype != w.dtype else x, w, b)
        # autocast: inputs and params cast to the low dtype, fp32 accumulate, low-dtype output
        y = F.linear(x.to(self.dt).float(), w.to(self.dt).float(),
                     None if b is None else b.to(self.dt).float())
        return y.to(self.dt)

    def linear_fp32(self, x, w, b=None):
        """Linear inside `amp.autocast(dtype=torch.float32)` regions (model.py:339,386,590)."""
        return F.linear(x.float(), w.float(), None if b is None else b.float())

    def half(self, x):
        """flash_attention's `half()` (attention.py:59-60): non-16-bit inputs go to bf16."""
        if self.dt is None:
            return x
        return x.to(self.dt)


# ----------------------------------------------------------------------------------------------
# embeddings / rope   (model.py:22-103)
# ----------------------------------------------------------------------------------------------
def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    """model.py:22-32 — float64 sinusoid, [cos | sin] concatenated."""
    assert dim % 2 == 0
    half = dim // 2
    pos = position.to(torch.float64)
    inv = torch.pow(torch.tensor(10000.0, dtype=torch.float64),
                    -torch.arange(half, dtype=torch.float64) / half).to(pos.device)
    ang = pos[:, None] * inv[None, :]
    return torch.cat([ang.cos(), ang.sin()], dim=1)


def rope_angles(max_seq_len: int, dim: int, theta: float = 10000.0) -> torch.Tensor:
    """model.py:35-43 — returns the *angles* (float64) whose polar form is the reference table."""
    assert dim % 2 == 0
    inv = 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float64),
                          torch.arange(0, dim, 2, dtype=torch.float64) / dim)
    return torch.arange(max_seq_len, dtype=torch.float64)[:, None] * inv[None, :]


def rope_axis_split(head_dim: int) -> Tuple[int, int, int]:
    """model.py:65 + 521-526 — complex pairs per head given to (frame, height, width)."""
    c = head_dim // 2
    return c - 2 * (c // 3), c // 3, c // 3


def rope_table(grid: Tuple[int, int, int], head_dim: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """cos/sin [F*H*W, head_dim/2] float64 for one sample — model.py:78-83 with the table of
    model.py:518-526 (three `rope_params(1024, ·)` concatenated on dim 1)."""
    f, h, w = grid
    d = head_dim
    cf, ch, cw = rope_axis_split(d)
    a_f = rope_angles(1024, d - 4 * (d // 6))[:f]      # [f, cf]
    a_h = rope_angles(1024, 2 * (d // 6))[:h]          # [h, ch]
    a_w = rope_angles(1024, 2 * (d // 6))[:w]          # [w, cw]
    assert a_f.shape[1] == cf and a_h.shape[1] == ch and a_w.shape[1] == cw
    ang = torch.cat([
        a_f.view(f, 1, 1, cf).expand(f, h, w, cf),
        a_h.view(1, h, 1, ch).expand(f, h, w, ch),
        a_w.view(1, 1, w, cw).expand(f, h, w, cw),
    ], dim=-1).reshape(f * h * w, d // 2)
    return ang.cos(), ang.sin()


def rope_apply(x: torch.Tensor, grids: Sequence[Tuple[int, int, int]],
               sp_rank: int = 0, sp_size: int = 1) -> torch.Tensor:
    """model.py:60-103.  x: [B, s, n, d]; pairs (2j, 2j+1) rotate by the position's angle, in
    float64; tokens beyond the sample's grid pass through; result is fp32.  With sequence
    parallelism the table row is offset by sp_rank*s (model.py:89-96; ones-padding = identity)."""
    b, s, n, d = x.shape
    out = []
    for i, g in enumerate(grids):
        seq_len = g[0] * g[1] * g[2]
        cos, sin = rope_table(g, d)
        cos, sin = cos.to(x.device), sin.to(x.device)
        if sp_size > 1:
            total = s * sp_size
            if total > seq_len:  # pad_freqs: multiply by 1+0j
                cos = torch.cat([cos, torch.ones(total - seq_len, d // 2, dtype=cos.dtype, device=cos.device)])
                sin = torch.cat([sin, torch.zeros(total - seq_len, d // 2, dtype=sin.dtype, device=sin.device)])
            cos = cos[sp_rank * s:(sp_rank + 1) * s]
            sin = sin[sp_rank * s:(sp_rank + 1) * s]
            nrot = s
        else:
            nrot = seq_len
        xi = x[i, :nrot].to(torch.float64).reshape(nrot, n, d // 2, 2)
        xr, xim = xi[..., 0], xi[..., 1]
        c, sn = cos[:nrot, None, :], sin[:nrot, None, :]
        yr = xr * c - xim * sn
        yi = xr * sn + xim * c
        y = torch.stack([yr, yi], dim=-1).reshape(nrot, n, d)
        y = torch.cat([y, x[i, nrot:].to(torch.float64)])
        out.append(y)
    return torch.stack(out).float()


# ----------------------------------------------------------------------------------------------
# norms   (model.py:106-135)
# ----------------------------------------------------------------------------------------------
def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """model.py:114-122 — fp32 normalise over the whole channel dim, round to x.dtype, times w."""
    xf = x.float()
    y = xf * torch.rsqrt(xf.pow(2).mean(dim=-1, keepdim=True) + eps)
    return y.type_as(x) * weight


def layer_norm(x: torch.Tensor, eps: float, weight=None, bias=None) -> torch.Tensor:
    """model.py:130-135 — LayerNorm in fp32, cast back to x.dtype."""
    return F.layer_norm(x.float(), (x.shape[-1],),
                        None if weight is None else weight.float(),
                        None if bias is None else bias.float(), eps).type_as(x)


# ----------------------------------------------------------------------------------------------
# attention   (attention.py:24-130 semantics on this path)
# ----------------------------------------------------------------------------------------------
def flash_attention(q, k, v, prec: _Prec, k_len: Optional[int] = None) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) v, non-causal, no dropout; q:[B,Lq,N,d], k/v:[B,Lk,N,d].
    Inputs are rounded the way attention.py:59-82 does (to bf16 unless already 16-bit) and the
    result is returned in q's *original* dtype (attention.py:57,130)."""
    out_dtype = q.dtype
    qh, kh, vh = prec.half(q), prec.half(k), prec.half(v)
    if k_len is not None:
        kh, vh = kh[:, :k_len], vh[:, :k_len]
    if prec.native and qh.is_cuda:
        from flash_attn import flash_attn_func                # attention.py:113-127 (FA2), [B, L, N, d] layout
        return flash_attn_func(qh.contiguous(), kh.contiguous(), vh.contiguous()).to(out_dtype)
    o = F.scaled_dot_product_attention(qh.float().transpose(1, 2), kh.float().transpose(1, 2),
                                       vh.float().transpose(1, 2))
    o = o.transpose(1, 2)
    if prec.dt is not None:
        o = o.to(prec.dt)
    return o.to(out_dtype)


# ----------------------------------------------------------------------------------------------
# Ulysses layout algebra (communication.py:40-160) — single-process emulation over a list of
# per-rank tensors; used by the SP tests and by the multi-rank oracle below.
# ----------------------------------------------------------------------------------------------
def all_to_all_4d_emulated(shards: List[torch.Tensor], scatter_dim: int, gather_dim: int):
    """shards[r]: rank r's tensor.  scatter 2 / gather 1: [b, L/P, H, d] -> [b, L, H/P, d]
    (communication.py:60-89); scatter 1 / gather 2: the inverse (communication.py:91-123)."""
    p = len(shards)
    if scatter_dim == 2 and gather_dim == 1:
        full = torch.cat(shards, dim=1)                       # [b, L, H, d]
        return list(full.chunk(p, dim=2))
    if scatter_dim == 1 and gather_dim == 2:
        full = torch.cat(shards, dim=2)                       # [b, L, H, d]
        return list(full.chunk(p, dim=1))
    raise RuntimeError("scatter_idx must be 1 or 2 and gather_idx must be 1 or 2")


# ----------------------------------------------------------------------------------------------
# blocks   (model.py:138-389)
# ----------------------------------------------------------------------------------------------
def self_attention(sd, pfx, x, grids, seq_lens, cfg: WanConfig, prec: _Prec,
                   sp_rank=0, sp_size=1, sp_exchange=None):
    """model.py:163-201.  `sp_exchange(tensor, scatter_dim, gather_dim)` performs the Ulysses
    all-to-all when sp_size > 1 (None => single rank)."""
    b, s, n, d = x.shape[0], x.shape[1], cfg.num_heads, cfg.dim // cfg.num_heads
    q = prec.linear(x, sd[pfx + "q.weight"], sd[pfx + "q.bias"])
    k = prec.linear(x, sd[pfx + "k.weight"], sd[pfx + "k.bias"])
    v = prec.linear(x, sd[pfx + "v.weight"], sd[pfx + "v.bias"]).view(b, s, n, d)
    if cfg.qk_norm:
        q = rms_norm(q, sd[pfx + "norm_q.weight"], cfg.eps)
        k = rms_norm(k, sd[pfx + "norm_k.weight"], cfg.eps)
    q = rope_apply(q.view(b, s, n, d), grids, sp_rank, sp_size)
    k = rope_apply(k.view(b, s, n, d), grids, sp_rank, sp_size)
    if sp_size > 1:
        q = sp_exchange(q, 2, 1)
        k = sp_exchange(k, 2, 1)
        v = sp_exchange(v, 2, 1)
    # k_lens = seq_lens (model.py:188-193): keys beyond the sample's length are dropped
    outs = []
    for i in range(b):
        outs.append(flash_attention(q[i:i + 1], k[i:i + 1], v[i:i + 1], prec, int(seq_lens[i])))
    o = torch.cat(outs)
    if sp_size > 1:
        o = sp_exchange(o, 1, 2)
    return prec.linear(o.flatten(2), sd[pfx + "o.weight"], sd[pfx + "o.bias"])


def cross_attention(sd, pfx, x, context, cfg: WanConfig, prec: _Prec):
    """model.py:206-226 (t2v) and 244-271 (i2v: CLIP tokens first, text = last 512)."""
    b, n, d = x.shape[0], cfg.num_heads, cfg.dim // cfg.num_heads
    q = prec.linear(x, sd[pfx + "q.weight"], sd[pfx + "q.bias"])
    if cfg.qk_norm:
        q = rms_norm(q, sd[pfx + "norm_q.weight"], cfg.eps)
    q = q.view(b, -1, n, d)
    if cfg.model_type == "t2v":
        ctx_txt, ctx_img = context, None
    else:
        n_img = context.shape[1] - T5_CONTEXT_TOKEN_NUMBER
        ctx_img, ctx_txt = context[:, :n_img], context[:, n_img:]
    k = prec.linear(ctx_txt, sd[pfx + "k.weight"], sd[pfx + "k.bias"])
    if cfg.qk_norm:
        k = rms_norm(k, sd[pfx + "norm_k.weight"], cfg.eps)
    v = prec.linear(ctx_txt, sd[pfx + "v.weight"], sd[pfx + "v.bias"])
    o = flash_attention(q, k.view(b, -1, n, d), v.view(b, -1, n, d), prec).flatten(2)
    if ctx_img is not None:
        ki = prec.linear(ctx_img, sd[pfx + "k_img.weight"], sd[pfx + "k_img.bias"])
        if cfg.qk_norm:
            ki = rms_norm(ki, sd[pfx + "norm_k_img.weight"], cfg.eps)
        vi = prec.linear(ctx_img, sd[pfx + "v_img.weight"], sd[pfx + "v_img.bias"])
        oi = flash_attention(q, ki.view(b, -1, n, d), vi.view(b, -1, n, d), prec).flatten(2)
        o = o + oi
    return prec.linear(o, sd[pfx + "o.weight"], sd[pfx + "o.bias"])


def attention_block(sd, pfx, x, e0, grids, seq_lens, context, cfg: WanConfig, prec: _Prec,
                    sp_rank=0, sp_size=1, sp_exchange=None):
    """model.py:320-359.  e0: [B, 6, C] fp32."""
    assert e0.dtype == torch.float32
    e = (sd[pfx + "modulation"].float() + e0).chunk(6, dim=1)
    h = layer_norm(x, cfg.eps).float() * (1 + e[1]) + e[0]
    y = self_attention(sd, pfx + "self_attn.", h, grids, seq_lens, cfg, prec,
                       sp_rank, sp_size, sp_exchange)
    x = x + y * e[2]                                           # fp32 from here on (model.py:348)
    if cfg.cross_attn_norm:
        h = layer_norm(x, cfg.eps, sd[pfx + "norm3.weight"], sd[pfx + "norm3.bias"])
    else:
        h = x
    x = x + cross_attention(sd, pfx + "cross_attn.", h, context, cfg, prec)
    h = layer_norm(x, cfg.eps).float() * (1 + e[4]) + e[3]
    y = prec.linear(h, sd[pfx + "ffn.0.weight"], sd[pfx + "ffn.0.bias"])
    y = F.gelu(y.float(), approximate="tanh").to(y.dtype)
    y = prec.linear(y, sd[pfx + "ffn.2.weight"], sd[pfx + "ffn.2.bias"])
    x = x + y * e[5]
    return x


def head_forward(sd, x, e, cfg: WanConfig, prec: _Prec):
    """model.py:379-389 — everything in fp32."""
    m = (sd["head.modulation"].float() + e.unsqueeze(1)).chunk(2, dim=1)
    h = layer_norm(x, cfg.eps) * (1 + m[1]) + m[0]
    return prec.linear_fp32(h, sd["head.head.weight"], sd["head.head.bias"])


def unpatchify(x, grids, cfg: WanConfig):
    """model.py:683-705 — [L, prod(patch)*c] -> [c, F*pt, H*ph, W*pw]."""
    c = cfg.out_dim
    pt, ph, pw = cfg.patch_size
    out = []
    for u, (f, h, w) in zip(x, grids):
        u = u[:f * h * w].view(f, h, w, pt, ph, pw, c)
        u = u.permute(6, 0, 3, 1, 4, 2, 5).reshape(c, f * pt, h * ph, w * pw)
        out.append(u)
    return out


def img_emb(sd, clip_fea, prec: _Prec):
    """MLPProj, model.py:392-410 (i2v; no flf positional embedding on this path)."""
    h = F.layer_norm(clip_fea.float(), (clip_fea.shape[-1],), sd["img_emb.proj.0.weight"].float(),
                     sd["img_emb.proj.0.bias"].float(), 1e-5)
    if prec.dt is not None:
        h = h  # autocast runs layer_norm in fp32 and keeps fp32 output
    h = prec.linear(h, sd["img_emb.proj.1.weight"], sd["img_emb.proj.1.bias"])
    h = F.gelu(h.float()).to(h.dtype)
    h = prec.linear(h, sd["img_emb.proj.3.weight"], sd["img_emb.proj.3.bias"])
    h = F.layer_norm(h.float(), (h.shape[-1],), sd["img_emb.proj.4.weight"].float(),
                     sd["img_emb.proj.4.bias"].float(), 1e-5)
    return h


# ----------------------------------------------------------------------------------------------
# WanModel.forward   (model.py:534-681)
# ----------------------------------------------------------------------------------------------
def patch_embed(sd, u, cfg: WanConfig, prec: _Prec):
    """Conv3d with kernel == stride == patch_size (model.py:497-498,578-581) restated as a
    gather of patches + Linear; returns [1, L, C] and the (F, H, W) token grid."""
    pt, ph, pw = cfg.patch_size
    c_in, fr, hh, ww = u.shape
    f, h, w = fr // pt, hh // ph, ww // pw
    p = u.view(c_in, f, pt, h, ph, w, pw).permute(1, 3, 5, 0, 2, 4, 6).reshape(f * h * w, -1)
    wgt = sd["patch_embedding.weight"].reshape(cfg.dim, -1)
    y = prec.linear(p, wgt, sd["patch_embedding.bias"])
    return y.unsqueeze(0), (f, h, w)


def wan_forward(sd: Dict[str, torch.Tensor], cfg: WanConfig, x: List[torch.Tensor], t: torch.Tensor,
                context: List[torch.Tensor], seq_len: int, clip_fea=None, y=None,
                output_features: bool = False, selected_layers: Sequence[int] = (20, 30, 40),
                autocast_dtype: Optional[torch.dtype] = None, num_blocks: Optional[int] = None,
                sp_size: int = 1, return_block_outputs: bool = False, native: bool = False):
    """WanModel.forward (model.py:534-681).  With sp_size > 1 the P sequence-parallel ranks are
    emulated in this one process: tokens are chunked (model.py:618-619), the Ulysses exchanges
    of model.py:183-196 are performed on the list of per-rank tensors, and features / head output
    are concatenated (all_gather, model.py:663-664,675-676)."""
    prec = _Prec(autocast_dtype, native)
    if cfg.model_type in ("i2v", "flf2v"):
        assert clip_fea is not None and y is not None
    if y is not None:
        x = [torch.cat([u, v], dim=0) for u, v in zip(x, y)]
    emb, grids = [], []
    for u in x:
        e_, g_ = patch_embed(sd, u, cfg, prec)
        emb.append(e_)
        grids.append(g_)
    seq_lens = [u.shape[1] for u in emb]
    assert max(seq_lens) <= seq_len
    xs = torch.cat([torch.cat([u, u.new_zeros(1, seq_len - u.shape[1], u.shape[2])], dim=1) for u in emb])

    # time embeddings, fp32 (model.py:589-594)
    e = prec.linear_fp32(sinusoidal_embedding_1d(cfg.freq_dim, t).float(),
                         sd["time_embedding.0.weight"], sd["time_embedding.0.bias"])
    e = prec.linear_fp32(F.silu(e), sd["time_embedding.2.weight"], sd["time_embedding.2.bias"])
    e0 = prec.linear_fp32(F.silu(e), sd["time_projection.1.weight"],
                          sd["time_projection.1.bias"]).unflatten(1, (6, cfg.dim))

    # context (model.py:596-607); context_lens=None => unmasked over text_len
    ctx = torch.stack([torch.cat([u, u.new_zeros(cfg.text_len - u.shape[0], u.shape[1])]) for u in context])
    ctx = prec.linear(ctx, sd["text_embedding.0.weight"], sd["text_embedding.0.bias"])
    ctx = F.gelu(ctx.float(), approximate="tanh").to(ctx.dtype)
    ctx = prec.linear(ctx, sd["text_embedding.2.weight"], sd["text_embedding.2.bias"])
    if clip_fea is not None:
        ctx = torch.cat([img_emb(sd, clip_fea, prec).to(ctx.dtype), ctx], dim=1)

    nb = cfg.num_layers if num_blocks is None else num_blocks
    feats, block_outs = [], []
    if sp_size == 1:
        for i in range(nb):
            xs = attention_block(sd, f"blocks.{i}.", xs, e0, grids, seq_lens, ctx, cfg, prec)
            if return_block_outputs:
                block_outs.append(xs)
            if output_features and (i + 1) in selected_layers:
                feats.append(xs)
    else:
        assert seq_len % sp_size == 0 and cfg.num_heads % sp_size == 0
        shards = list(torch.chunk(xs, sp_size, dim=1))
        for i in range(nb):
            shards = _sp_block(sd, f"blocks.{i}.", shards, e0, grids, seq_lens, ctx, cfg, prec)
            if return_block_outputs:
                block_outs.append(torch.cat(shards, dim=1))
            if output_features and (i + 1) in selected_layers:
                feats.append(torch.cat(shards, dim=1))
        xs = torch.cat(shards, dim=1)
    if output_features:
        return (feats, block_outs) if return_block_outputs else feats
    out = head_forward(sd, xs, e, cfg, prec)
    out = [u.float() for u in unpatchify(out, grids, cfg)]
    return (out, block_outs) if return_block_outputs else out


def _sp_block(sd, pfx, shards, e0, grids, seq_lens, ctx, cfg, prec):
    """One attention block executed 'on P ranks' in lock-step so the exchange can be emulated."""
    p = len(shards)
    e = (sd[pfx + "modulation"].float() + e0).chunk(6, dim=1)
    n, d = cfg.num_heads, cfg.dim // cfg.num_heads
    qs, ks, vs = [], [], []
    for r, xr in enumerate(shards):
        b, s = xr.shape[:2]
        h = layer_norm(xr, cfg.eps).float() * (1 + e[1]) + e[0]
        sp = pfx + "self_attn."
        q = prec.linear(h, sd[sp + "q.weight"], sd[sp + "q.bias"])
        k = prec.linear(h, sd[sp + "k.weight"], sd[sp + "k.bias"])
        v = prec.linear(h, sd[sp + "v.weight"], sd[sp + "v.bias"]).view(b, s, n, d)
        if cfg.qk_norm:
            q = rms_norm(q, sd[sp + "norm_q.weight"], cfg.eps)
            k = rms_norm(k, sd[sp + "norm_k.weight"], cfg.eps)
        qs.append(rope_apply(q.view(b, s, n, d), grids, r, p))
        ks.append(rope_apply(k.view(b, s, n, d), grids, r, p))
        vs.append(v)
    qh = all_to_all_4d_emulated(qs, 2, 1)
    kh = all_to_all_4d_emulated(ks, 2, 1)
    vh = all_to_all_4d_emulated(vs, 2, 1)
    oh = []
    for r in range(p):
        outs = [flash_attention(qh[r][i:i + 1], kh[r][i:i + 1], vh[r][i:i + 1], prec, int(seq_lens[i]))
                for i in range(qh[r].shape[0])]
        oh.append(torch.cat(outs))
    os_ = all_to_all_4d_emulated(oh, 1, 2)
    new = []
    for r, xr in enumerate(shards):
        sp = pfx + "self_attn."
        yv = prec.linear(os_[r].flatten(2), sd[sp + "o.weight"], sd[sp + "o.bias"])
        xr = xr + yv * e[2]
        h = layer_norm(xr, cfg.eps, sd[pfx + "norm3.weight"], sd[pfx + "norm3.bias"]) if cfg.cross_attn_norm else xr
        xr = xr + cross_attention(sd, pfx + "cross_attn.", h, ctx, cfg, prec)
        h = layer_norm(xr, cfg.eps).float() * (1 + e[4]) + e[3]
        yv = prec.linear(h, sd[pfx + "ffn.0.weight"], sd[pfx + "ffn.0.bias"])
        yv = F.gelu(yv.float(), approximate="tanh").to(yv.dtype)
        yv = prec.linear(yv, sd[pfx + "ffn.2.weight"], sd[pfx + "ffn.2.bias"])
        new.append(xr + yv * e[5])
    return new


# ----------------------------------------------------------------------------------------------
# PAVRM reward head   (network.py:8-152)
# ----------------------------------------------------------------------------------------------
def query_attention(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_heads: int = 8,
                    return_type: Optional[str] = "query", autocast_dtype=None, native: bool = False) -> torch.Tensor:
    """QueryAttention.forward (network.py:44-110) for layer_norm=False, product_text=False, eval
    mode (dropout inactive).  x: [n_sel, B, L, C] | [B, L, C] | [B, C].  nn.MultiheadAttention
    semantics: q = queries Wq^T + bq, k = x Wk^T + bk, v = x Wv^T + bv (packed in_proj), per-head
    softmax(q k^T / sqrt(hd)) v, out_proj."""
    prec = _Prec(autocast_dtype, native)
    shape = x.shape
    if x.dim() == 2:
        x = x.unsqueeze(1)
    elif x.dim() == 4:
        x = x.reshape(shape[0] * shape[1], shape[2], shape[3])
    bsz, L, C = x.shape
    hd = C // num_heads
    queries = sd["queries"].unsqueeze(0).expand(bsz, -1, -1)          # [B', nq, C]
    w, bqkv = sd["multihead_attn.in_proj_weight"], sd["multihead_attn.in_proj_bias"]
    q = prec.linear(queries, w[:C], bqkv[:C])
    k = prec.linear(x, w[C:2 * C], bqkv[C:2 * C])
    v = prec.linear(x, w[2 * C:], bqkv[2 * C:])
    nq = q.shape[1]
    qh = q.reshape(bsz, nq, num_heads, hd).transpose(1, 2).float()
    kh = k.reshape(bsz, L, num_heads, hd).transpose(1, 2).float()
    vh = v.reshape(bsz, L, num_heads, hd).transpose(1, 2).float()
    att = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    o = (att @ vh).transpose(1, 2).reshape(bsz, nq, C)
    if prec.dt is not None:
        o = o.to(prec.dt)
    o = prec.linear(o, sd["multihead_attn.out_proj.weight"], sd["multihead_attn.out_proj.bias"])
    out = o.mean(dim=1) if nq > 1 else o.squeeze(1)                   # network.py:90-93
    if len(shape) == 4:                                               # network.py:96-98
        out = out.view(shape[0], bsz // shape[0], -1).mean(dim=0)
    if return_type == "query":                                        # network.py:103-104
        out = out + sd["queries"].unsqueeze(0).expand(bsz, -1, -1)
    return out


def reward_mlp(sd: Dict[str, torch.Tensor], x: torch.Tensor, autocast_dtype=None, native: bool = False) -> torch.Tensor:
    """MLP.forward (network.py:130-134): the pre-sigmoid reward logit."""
    prec = _Prec(autocast_dtype, native)
    h = torch.relu(prec.linear(x, sd["fc1.weight"], sd["fc1.bias"]))
    h = torch.relu(prec.linear(h, sd["fc2.weight"], sd["fc2.bias"]))
    return prec.linear(h, sd["fc3.weight"], sd["fc3.bias"])


def forward_mlp(sd, x, autocast_dtype=None):
    """network.py:151-152."""
    return torch.sigmoid(reward_mlp(sd, x, autocast_dtype))


def pavrm_reward(sd_model, cfg, sd_qa, sd_mlp, x, t, context, seq_len, clip_fea=None, y=None,
                 selected_layers=(8,), num_blocks=8, qa_heads=8, autocast_dtype=None, sp_size=1, native=False):
    """PAVRM scoring chain (train_pavrm.py:792-845, train_prfl.py:764-796): features of the
    selected block(s) -> list2batch -> QueryAttention -> MLP.  Returns (logit, features)."""
    feats = wan_forward(sd_model, cfg, x, t, context, seq_len, clip_fea, y, output_features=True,
                        selected_layers=selected_layers, autocast_dtype=autocast_dtype,
                        num_blocks=num_blocks, sp_size=sp_size, native=native)
    stacked = torch.stack(feats)                                      # [n_sel, B, L, C]
    pooled = query_attention(sd_qa, stacked, qa_heads, "query", autocast_dtype, native)
    return reward_mlp(sd_mlp, pooled, autocast_dtype, native).float(), stacked

"""B200-native Wan-DiT with the reference's module API (diffusers_lite/wan/modules/model.py).

Same class names, constructor signatures, parameter names (state-dict keys) and forward signatures as
the reference `WanModel` / `WanAttentionBlock` / `WanSelfAttention` / `Wan{T2V,I2V}CrossAttention` /
`WanRMSNorm` / `WanLayerNorm` / `Head` / `MLPProj`, so reference checkpoints load and the reference
trainers can call it unchanged; but every forward runs hand-written sm_100a kernels through the C ABI
(prfl_b200.ops) instead of ATen / cuBLAS / flash-attn.  There is no CPU or PyTorch fallback: on a
machine without the CUDA library these forwards raise.

Precision choreography kept from the reference: fp32 residual stream, bf16
GEMM/attention operands with fp32 accumulation, RMSNorm over the full channel dim with the bf16 rounding
before the weight multiply, RoPE angles from float64, fp32 time-embedding path and fp32 head.

Per block (sequence of C-ABI calls, B = 1 sample at a time, M = local tokens):
  ln_mod -> gemm(QKV fused, N=3C) -> rmsnorm_rope(q), rmsnorm_rope(k) -> [a2a] -> attn_fwd -> [a2a]
  -> gemm(o, gated residual epilogue) -> ln_mod(affine) -> gemm(cross q) -> rmsnorm -> gemm(ctx KV)
  -> rmsnorm -> attn_fwd (x2 for i2v) -> gemm(cross o, residual epilogue) -> ln_mod
  -> gemm(ffn.0, GELU epilogue) -> gemm(ffn.2, gated residual epilogue)
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .parallel import get_sequence_parallel_state, nccl_info, ulysses_gather_tokens, ulysses_scatter_tokens, all_gather
from .rope import rope_tables

__all__ = ["WanModel", "ModelConfig", "PreparedContext", "WanAttentionBlock", "WanSelfAttention", "WanT2VCrossAttention", "WanI2VCrossAttention",
           "WanRMSNorm", "WanLayerNorm", "Head", "MLPProj", "sinusoidal_embedding_1d", "rope_params", "rope_apply"]

T5_CONTEXT_TOKEN_NUMBER = 512


def no_autocast(fn):
    """The reference trainers call the model inside `torch.autocast("cuda", dtype=bf16)` (train_prfl.py:669, 723, 748, 960) and
    keep selected parts in fp32 with nested `amp.autocast(dtype=torch.float32)` blocks (model.py:339-354, 386, 590: time
    embedding, modulation, head).  Here the precision of every step is explicit (bf16 GEMM / attention operands, fp32
    accumulation, fp32 residual stream / time embedding / head), so the caller's autocast state must not re-cast the few
    PyTorch ops on the path: the decorated forward runs with CUDA autocast off, whatever the caller's context."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        with torch.autocast(device_type="cuda", enabled=False):
            return fn(*args, **kwargs)
    return wrapper


# ------------------------------------------------------------------------------------------------
# bf16 operand cache: fp32 master parameters -> bf16 copies, refreshed when the parameter changes
# (what torch.autocast re-does on every Linear call in the reference).
# ------------------------------------------------------------------------------------------------
_WEIGHT_EPOCH = [0]


def bump_weight_epoch():
    """Invalidate every derived operand (concatenated / split / fp32-bias copies, PreparedContext K/V are the caller's):
    called by `sharding.ShardedAdamW.step()`, whose updates land in the flat buffers the parameters are views of and
    therefore do not bump the parameters' own version counters."""
    _WEIGHT_EPOCH[0] += 1


class _OperandCache:
    def __init__(self):
        self._store = {}

    def get(self, key, params: Sequence[torch.Tensor], build):
        ver = (_WEIGHT_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in params)
        hit = self._store.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        val = build()
        self._store[key] = (ver, val)
        return val

    def clear(self):
        self._store.clear()


def _cat_bf16(ws: Sequence[torch.Tensor]) -> torch.Tensor:
    """The [sum rows, K] bf16 operand of one or several weights.  Resident bf16 parameters (sharding.ResidentUnit) that are
    adjacent in their flat buffer ARE the operand: a view, no copy."""
    if all(p.dtype == torch.bfloat16 and p.is_contiguous() for p in ws):
        if len(ws) == 1:
            return ws[0].detach().reshape(ws[0].shape[0], -1)
        k = ws[0][0].numel()
        adjacent = all(a.data_ptr() + a.numel() * 2 == b.data_ptr() and b[0].numel() == k and
                       a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr() for a, b in zip(ws, ws[1:]))
        if adjacent:
            rows = sum(p.shape[0] for p in ws)
            return torch.as_strided(ws[0].detach(), (rows, k), (k, 1))
        return torch.cat([p.detach().reshape(p.shape[0], -1) for p in ws], dim=0)
    w = torch.cat([p.detach().reshape(p.shape[0], -1) for p in ws], dim=0).float().contiguous()
    return ops.cast_bf16(w)


def _cat_f32(bs: Sequence[torch.Tensor]) -> torch.Tensor:
    # autocast casts the bias to bf16 before the addmm; keep that rounding, store as fp32 for the epilogue
    return torch.cat([b.detach().float() for b in bs]).bfloat16().float().contiguous()


class _Linear(nn.Linear):
    """nn.Linear whose forward is the tcgen05 GEMM (bf16 operands, fp32 accumulate, bf16 out)."""

    def operands(self):
        cache = self.__dict__.setdefault("_prfl_cache", _OperandCache())
        w = cache.get("w", [self.weight], lambda: _cat_bf16([self.weight]))
        b = None if self.bias is None else cache.get("b", [self.bias], lambda: _cat_f32([self.bias]))
        return w, b

    def forward(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            from .engine import LinearFn
            return LinearFn.apply(x, self.weight, self.bias, self)
        w, b = self.operands()
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        if x2.dtype != torch.bfloat16:
            x2 = x2.to(torch.bfloat16)
        y = ops.gemm(x2.contiguous(), w, bias=b, epi=ops.EPI_BF16)
        return y.view(*shp[:-1], -1)


def sinusoidal_embedding_1d(dim, position):
    """model.py:22-32 (float64)."""
    assert dim % 2 == 0
    half = dim // 2
    position = position.type(torch.float64)
    sinusoid = torch.outer(position, torch.pow(10000, -torch.arange(half).to(position).div(half)))
    return torch.cat([torch.cos(sinusoid), torch.sin(sinusoid)], dim=1)


def rope_params(max_seq_len, dim, theta=10000):
    """model.py:35-43 — kept for API parity (`WanModel.freqs`); the kernels use rope.rope_tables."""
    assert dim % 2 == 0
    freqs = torch.outer(torch.arange(max_seq_len),
                        1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
    return torch.polar(torch.ones_like(freqs), freqs)


def rope_apply(x, grid_sizes, freqs=None):
    """API-compatibility helper with the reference's signature (model.py:60-103): x [B, s, n, d] -> fp32, pairs
    (2j, 2j+1) of every head rotated by the (frame, row, column) angle of the token, tokens past the grid passed through,
    Ulysses rank offset applied under sequence parallelism.  The model itself never calls this: RoPE is fused into
    prfl_rmsnorm_rope_fwd; this exists for callers that use the free function.  `freqs` is ignored (tables come from
    rope.rope_tables, float64-derived)."""
    b, s, n, d = x.shape
    sp = get_sequence_parallel_state()
    P = nccl_info.sp_size if sp else 1
    rank = nccl_info.rank_within_group if sp else 0
    out = []
    for i, g in enumerate(_grid_list(grid_sizes)):
        cos, sin = rope_tables(g, x.device, d, pad_to=s * P)
        nrot = s if sp else min(g[0] * g[1] * g[2], s)
        c, sn = cos[rank * s:rank * s + nrot, None, :], sin[rank * s:rank * s + nrot, None, :]
        xi = x[i, :nrot].float().reshape(nrot, n, d // 2, 2)
        y = torch.stack([xi[..., 0] * c - xi[..., 1] * sn, xi[..., 0] * sn + xi[..., 1] * c], dim=-1).reshape(nrot, n, d)
        out.append(torch.cat([y, x[i, nrot:].float()]))
    return torch.stack(out)


class WanRMSNorm(nn.Module):
    """model.py:106-122.  forward() accepts a bf16 [.., C] tensor and normalises over all C channels."""

    def __init__(self, dim, eps=1e-5):
        super().__init__()
        self.dim = dim
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).to(torch.bfloat16).contiguous()
        out = torch.empty_like(x2)
        ops.rmsnorm_rope_(x2, self.weight.detach().float(), None, None, self.eps, out=out)
        return out.view(shp)


class WanLayerNorm(nn.LayerNorm):
    """model.py:125-135.  fp32 in -> bf16 out (the consumer is always a bf16 GEMM)."""

    def __init__(self, dim, eps=1e-6, elementwise_affine=False):
        super().__init__(dim, elementwise_affine=elementwise_affine, eps=eps)

    def forward(self, x, shift=None, scale=None, round_bf16=False):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).float().contiguous()
        g = self.weight.detach().float() if self.elementwise_affine else None
        b = self.bias.detach().float() if self.elementwise_affine else None
        return ops.ln_mod(x2, shift, scale, g, b, self.eps, round_bf16).view(shp)


def _grid_list(grid_sizes):
    return [tuple(int(v) for v in g) for g in (grid_sizes.tolist() if torch.is_tensor(grid_sizes) else grid_sizes)]


class WanSelfAttention(nn.Module):
    """model.py:138-201."""

    def __init__(self, dim, num_heads, window_size=(-1, -1), qk_norm=True, eps=1e-6):
        assert dim % num_heads == 0
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        assert self.head_dim == 128, "prfl_b200 attention kernels are head_dim 128 only"
        if tuple(window_size) != (-1, -1):      # every shipped config (wan/configs/*.py) uses the global window; refuse rather than differ
            raise NotImplementedError(f"window_size={tuple(window_size)}: the attention kernels implement global attention only")
        self.window_size = window_size
        self.qk_norm = qk_norm
        self.eps = eps
        self.q = _Linear(dim, dim)
        self.k = _Linear(dim, dim)
        self.v = _Linear(dim, dim)
        self.o = _Linear(dim, dim)
        self.norm_q = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()
        self.norm_k = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()

    def _qkv_operands(self):
        cache = self.__dict__.setdefault("_prfl_cache", _OperandCache())
        ws = [self.q.weight, self.k.weight, self.v.weight]
        bs = [self.q.bias, self.k.bias, self.v.bias]
        return cache.get("wqkv", ws, lambda: _cat_bf16(ws)), cache.get("bqkv", bs, lambda: _cat_f32(bs))

    def attend(self, h, seq_lens, grid_sizes):
        """h: [B, s, C] bf16 (already normalised + modulated) -> attention output [B, s, C] bf16,
        BEFORE the output projection (the caller fuses `o` with the gated residual)."""
        assert self.qk_norm, "qk_norm=False is not on the reference path"
        b, s, C = h.shape
        n, d = self.num_heads, self.head_dim
        wqkv, bqkv = self._qkv_operands()
        sp = get_sequence_parallel_state()
        P = nccl_info.sp_size if sp else 1
        rank = nccl_info.rank_within_group if sp else 0
        grids = _grid_list(grid_sizes)
        outs = []
        for i in range(b):
            qkv = ops.gemm(h[i], wqkv, bias=bqkv, epi=ops.EPI_BF16)            # [s, 3C]
            f, hh, ww = grids[i]
            seq_len = f * hh * ww
            cos, sin = rope_tables(grids[i], h.device, d, pad_to=s * P)
            n_rot = s if sp else min(seq_len, s)
            ops.rmsnorm_rope_(qkv[:, :C], self.norm_q.weight.detach().float(), cos, sin, self.eps, n_rot, rank * s)
            ops.rmsnorm_rope_(qkv[:, C:2 * C], self.norm_k.weight.detach().float(), cos, sin, self.eps, n_rot, rank * s)
            q3, k3, v3 = (qkv[:, j * C:(j + 1) * C].unflatten(1, (n, d)) for j in range(3))
            klen = int(seq_lens[i])
            if not sp:
                o = ops.attn_fwd(q3, k3[:klen], v3[:klen])                        # [s, n, d]
            else:
                qg, kg, vg = (ulysses_scatter_tokens(t, P) for t in (q3, k3, v3))  # [s*P, n/P, d]
                og = ops.attn_fwd(qg, kg[:klen], vg[:klen])
                o = ulysses_gather_tokens(og, P)                                 # [s, n, d]
            outs.append(o.reshape(s, C))
        return torch.stack(outs)

    def forward(self, x, seq_lens, grid_sizes, freqs=None):
        """Reference signature (model.py:163).  x: [B, L, C]; returns o(attention) as bf16 [B, L, C]."""
        h = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
        return self.o(self.attend(h.contiguous(), seq_lens, grid_sizes))


class WanT2VCrossAttention(WanSelfAttention):
    """model.py:204-226."""

    def _kv_operands(self, names=("k", "v")):
        cache = self.__dict__.setdefault("_prfl_cache", _OperandCache())
        ws = [getattr(self, nm).weight for nm in names]
        bs = [getattr(self, nm).bias for nm in names]
        return (cache.get("w" + "".join(names), ws, lambda: _cat_bf16(ws)),
                cache.get("b" + "".join(names), bs, lambda: _cat_f32(bs)))

    def _attend_ctx(self, q3, ctx, names, norm, sample=0):
        C, n, d = self.dim, self.num_heads, self.head_dim
        cache = self.__dict__.get("_kv_cache")           # set by WanModel.forward for a PreparedContext (no-grad only)
        kv = None if cache is None else cache.get((names, sample))
        if kv is None:
            wkv, bkv = self._kv_operands(names)
            kv = ops.gemm(ctx, wkv, bias=bkv, epi=ops.EPI_BF16)                   # [Lc, 2C]
            ops.rmsnorm_rope_(kv[:, :C], norm.weight.detach().float(), None, None, self.eps)
            if cache is not None:
                cache[(names, sample)] = kv
        return ops.attn_fwd(q3, kv[:, :C].unflatten(1, (n, d)), kv[:, C:].unflatten(1, (n, d)))

    def attend(self, h, context, context_lens=None):
        assert context_lens is None, "the reference path passes context_lens=None (model.py:597)"
        b, s, C = h.shape
        n, d = self.num_heads, self.head_dim
        wq, bq = self.q.operands()
        outs = []
        for i in range(b):
            q = ops.gemm(h[i], wq, bias=bq, epi=ops.EPI_BF16)
            ops.rmsnorm_rope_(q, self.norm_q.weight.detach().float(), None, None, self.eps)
            o = self._attend_ctx(q.unflatten(1, (n, d)), context[i], ("k", "v"), self.norm_k, i)
            outs.append(o.reshape(s, C))
        return torch.stack(outs)

    def forward(self, x, context, context_lens):
        h = x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)
        return self.o(self.attend(h.contiguous(), context.to(torch.bfloat16).contiguous(), context_lens))


class WanI2VCrossAttention(WanT2VCrossAttention):
    """model.py:229-271: CLIP image tokens first, text = last 512 tokens; two attentions, summed."""

    def __init__(self, dim, num_heads, window_size=(-1, -1), qk_norm=True, eps=1e-6):
        super().__init__(dim, num_heads, window_size, qk_norm, eps)
        self.k_img = _Linear(dim, dim)
        self.v_img = _Linear(dim, dim)
        self.norm_k_img = WanRMSNorm(dim, eps=eps) if qk_norm else nn.Identity()

    def attend(self, h, context, context_lens=None):
        assert context_lens is None
        b, s, C = h.shape
        n, d = self.num_heads, self.head_dim
        n_img = context.shape[1] - T5_CONTEXT_TOKEN_NUMBER
        wq, bq = self.q.operands()
        outs = []
        for i in range(b):
            q = ops.gemm(h[i], wq, bias=bq, epi=ops.EPI_BF16)
            ops.rmsnorm_rope_(q, self.norm_q.weight.detach().float(), None, None, self.eps)
            q3 = q.unflatten(1, (n, d))
            o_img = self._attend_ctx(q3, context[i, :n_img].contiguous(), ("k_img", "v_img"), self.norm_k_img, i)
            o_txt = self._attend_ctx(q3, context[i, n_img:].contiguous(), ("k", "v"), self.norm_k, i)
            outs.append((o_txt + o_img).reshape(s, C))                            # model.py:269 (bf16 add)
        return torch.stack(outs)


class PreparedContext:
    """Opaque result of `WanModel.prepare_context`: embedded context [B, n_ctx, dim] bf16 + per-block K/V cache."""

    def __init__(self, embedded: torch.Tensor):
        self.embedded = embedded
        self.kv = {}


WAN_CROSSATTENTION_CLASSES = {"t2v_cross_attn": WanT2VCrossAttention, "i2v_cross_attn": WanI2VCrossAttention}


class WanAttentionBlock(nn.Module):
    """model.py:280-359."""

    def __init__(self, cross_attn_type, dim, ffn_dim, num_heads, window_size=(-1, -1), qk_norm=True,
                 cross_attn_norm=False, eps=1e-6):
        super().__init__()
        self.dim = dim
        self.ffn_dim = ffn_dim
        self.num_heads = num_heads
        self.window_size = window_size
        self.qk_norm = qk_norm
        self.cross_attn_norm = cross_attn_norm
        self.eps = eps
        self.norm1 = WanLayerNorm(dim, eps)
        self.self_attn = WanSelfAttention(dim, num_heads, window_size, qk_norm, eps)
        self.norm3 = WanLayerNorm(dim, eps, elementwise_affine=True) if cross_attn_norm else nn.Identity()
        self.cross_attn = WAN_CROSSATTENTION_CLASSES[cross_attn_type](dim, num_heads, (-1, -1), qk_norm, eps)
        self.norm2 = WanLayerNorm(dim, eps)
        self.ffn = nn.Sequential(_Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), _Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.randn(1, 6, dim) / dim ** 0.5)

    def forward(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens, first_block_bf16_input=False):
        """x: [B, L, C] fp32 residual stream, e: [B, 6, C] fp32, context: [B, Lc, C] bf16.
        Without autograd the stream is updated IN PLACE and returned.  With autograd the block is one
        autograd.Function that saves only its input and recomputes in backward (engine.BlockFn).
        `first_block_bf16_input` reproduces the reference's extra bf16 rounding of norm1's output in block 0,
        whose input is bf16 (model.py:345 with x.dtype == bf16)."""
        assert e.dtype == torch.float32 and context_lens is None
        from . import engine
        grids = _grid_list(grid_sizes)
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        if torch.is_grad_enabled() and (x.requires_grad or e.requires_grad or context.requires_grad
                                        or any(p.requires_grad for p in self.parameters())):
            names = tuple(engine.block_param_names(self))
            return engine.BlockFn.apply(x, e, context, self, [int(v) for v in seq_lens], grids, first_block_bf16_input, names,
                                        *self.parameters())
        em = (self.modulation.detach().float() + e.detach()).contiguous()                    # [B, 6, C]
        ctx = context.detach()
        ctx = ctx if ctx.dtype == torch.bfloat16 else ctx.to(torch.bfloat16)
        x = x.detach()
        for i in range(x.shape[0]):
            engine.block_forward(self, x[i], em[i], ctx[i].contiguous(), int(seq_lens[i]), grids[i], first_block_bf16_input, sample=i)
        return x


class Head(nn.Module):
    """model.py:362-389 — fp32 throughout."""

    def __init__(self, dim, out_dim, patch_size, eps=1e-6):
        super().__init__()
        self.dim = dim
        self.out_dim = out_dim
        self.patch_size = patch_size
        self.eps = eps
        out_dim = math.prod(patch_size) * out_dim
        self.norm = WanLayerNorm(dim, eps)
        self.head = nn.Linear(dim, out_dim)
        self.modulation = nn.Parameter(torch.randn(1, 2, dim) / dim ** 0.5)

    def _split_operands(self):
        cache = self.__dict__.setdefault("_prfl_cache", _OperandCache())

        def build():
            w = self.head.weight.detach().float()
            hi = w.bfloat16()
            lo = (w - hi.float()).bfloat16()
            return hi.contiguous(), lo.contiguous()
        return cache.get("w_split", [self.head.weight], build)

    def forward(self, x, e):
        """x: [B, L, C] fp32, e: [B, C] fp32 -> [B, L, prod(patch)*out_dim] fp32."""
        assert e.dtype == torch.float32
        m = (self.modulation.float() + e.unsqueeze(1)).chunk(2, dim=1)                      # 2 x [B, 1, C]
        if torch.is_grad_enabled() and (x.requires_grad or e.requires_grad or self.head.weight.requires_grad):
            from .engine import HeadFn                                   # same kernels as below + their backward
            return torch.stack([HeadFn.apply(x[i], m[0][i, 0], m[1][i, 0], self.head.weight, self.head.bias, self)
                                for i in range(x.shape[0])])
        w_hi, w_lo = self._split_operands()
        bias = self.head.bias.detach().float().contiguous()
        outs = []
        for i in range(x.shape[0]):
            hi, lo = ops.ln_mod_split(x[i].float().contiguous(), m[0][i, 0].contiguous(), m[1][i, 0].contiguous(), self.eps)
            o = ops.gemm(hi, w_hi, bias=bias, epi=ops.EPI_F32)
            ops.gemm(hi, w_lo, epi=ops.EPI_F32, out=o, beta=True)
            ops.gemm(lo, w_hi, epi=ops.EPI_F32, out=o, beta=True)
            outs.append(o)
        return torch.stack(outs)


class MLPProj(nn.Module):
    """model.py:392-410.  [B, 257, 1280] CLIP features -> [B, 257, dim]; ~0.01 % of the FLOPs, kept in PyTorch
     except the two Linears which go through the tcgen05 GEMM."""

    def __init__(self, in_dim, out_dim, flf_pos_emb=False):
        super().__init__()
        self.proj = nn.Sequential(nn.LayerNorm(in_dim), _Linear(in_dim, in_dim), nn.GELU(), _Linear(in_dim, out_dim),
                                  nn.LayerNorm(out_dim))
        if flf_pos_emb:
            self.emb_pos = nn.Parameter(torch.zeros(1, 257 * 2, 1280))

    def forward(self, image_embeds):
        if hasattr(self, "emb_pos"):
            bs, n, d = image_embeds.shape
            image_embeds = image_embeds.view(-1, 2 * n, d) + self.emb_pos
        p = self.proj
        h = F.layer_norm(image_embeds.float(), p[0].normalized_shape, p[0].weight.float(), p[0].bias.float(), p[0].eps)
        h = p[1](h)
        h = F.gelu(h.float()).to(torch.bfloat16)
        h = p[3](h)
        return F.layer_norm(h.float(), p[4].normalized_shape, p[4].weight.float(), p[4].bias.float(), p[4].eps)


class _UnpatchifyFn(torch.autograd.Function):
    """'fhwpqrc->cfphqwr' scatter (model.py:700-703) and its transpose."""

    @staticmethod
    def forward(ctx, tokens, c, grid):
        ctx.rows = tokens.shape[0]
        return ops.unpatchify(tokens.detach(), c, grid)

    @staticmethod
    def backward(ctx, dvid):
        return ops.unpatchify_bwd(dvid.detach().float().contiguous(), ctx.rows), None, None


class ModelConfig(dict):
    """`model.config` as the callers use it: a mapping (`dict(transformer.config)`, model_utils.py:119) whose entries are also
    attributes, readable and assignable (`transformer.config.lora_rank = ...`, train_prfl.py:355-357)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name) from None

    def __setattr__(self, name, value):
        self[name] = value


class WanModel(nn.Module):
    """model.py:413-729.  Drop-in: same __init__ / forward signature, attributes (`blocks`, `head`, `freqs`,
    `enable_teacache`, `_no_split_modules`, `config`) and state-dict keys."""

    ignore_for_config = ["patch_size", "cross_attn_norm", "qk_norm", "text_dim", "window_size"]
    _no_split_modules = ["WanAttentionBlock"]
    enable_teacache = False          # class attribute, as the trainers set it (train_pavrm.py:237)

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048, ffn_dim=8192,
                 freq_dim=256, text_dim=4096, out_dim=16, num_heads=16, num_layers=32, window_size=(-1, -1),
                 qk_norm=True, cross_attn_norm=True, eps=1e-6):
        super().__init__()
        assert model_type in ["t2v", "i2v", "flf2v"]
        assert tuple(patch_size) == (1, 2, 2), "prfl_b200 patchify kernels implement patch_size (1, 2, 2)"
        self.config = ModelConfig(model_type=model_type, patch_size=tuple(patch_size), text_len=text_len, in_dim=in_dim, dim=dim,
                                  ffn_dim=ffn_dim, freq_dim=freq_dim, text_dim=text_dim, out_dim=out_dim, num_heads=num_heads,
                                  num_layers=num_layers, window_size=tuple(window_size), qk_norm=qk_norm,
                                  cross_attn_norm=cross_attn_norm, eps=eps)
        self.model_type = model_type
        self.patch_size = tuple(patch_size)
        self.text_len = text_len
        self.in_dim = in_dim
        self.dim = dim
        self.ffn_dim = ffn_dim
        self.freq_dim = freq_dim
        self.text_dim = text_dim
        self.out_dim = out_dim
        self.num_heads = num_heads
        self.num_layers = num_layers
        self.window_size = window_size
        self.qk_norm = qk_norm
        self.cross_attn_norm = cross_attn_norm
        self.eps = eps

        self.patch_embedding = nn.Conv3d(in_dim, dim, kernel_size=patch_size, stride=patch_size)
        self.text_embedding = nn.Sequential(_Linear(text_dim, dim), nn.GELU(approximate="tanh"), _Linear(dim, dim))
        self.time_embedding = nn.Sequential(nn.Linear(freq_dim, dim), nn.SiLU(), nn.Linear(dim, dim))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(dim, dim * 6))
        cross_attn_type = "t2v_cross_attn" if model_type == "t2v" else "i2v_cross_attn"
        self.blocks = nn.ModuleList([
            WanAttentionBlock(cross_attn_type, dim, ffn_dim, num_heads, window_size, qk_norm, cross_attn_norm, eps)
            for _ in range(num_layers)])
        self.head = Head(dim, out_dim, patch_size, eps)
        assert (dim % num_heads) == 0 and (dim // num_heads) % 2 == 0
        d = dim // num_heads
        self.freqs = torch.cat([rope_params(1024, d - 4 * (d // 6)), rope_params(1024, 2 * (d // 6)),
                                rope_params(1024, 2 * (d // 6))], dim=1)
        if model_type in ("i2v", "flf2v"):
            self.img_emb = MLPProj(1280, dim, flf_pos_emb=model_type == "flf2v")
        self.init_weights()

    # -- embeddings ---------------------------------------------------------------------------
    def _patch_operands(self):
        cache = self.__dict__.setdefault("_prfl_cache", _OperandCache())
        w = cache.get("wp", [self.patch_embedding.weight], lambda: _cat_bf16([self.patch_embedding.weight]))
        b = cache.get("bp", [self.patch_embedding.bias], lambda: _cat_f32([self.patch_embedding.bias]))
        return w, b

    @torch.no_grad()
    def prepare_context(self, context, clip_fea=None) -> "PreparedContext":
        """Embed a prompt ONCE for many no-grad forwards: the text / CLIP embeddings and, filled lazily
        by the first forward, every block's cross-attention K / V of the context are constants of the prompt and the
        weights — the reference recomputes them on every rank, block and call (model.py:206-226, 590-607), i.e. 80-100
        times per sampled video and m times per PRFL step.  Pass the result as `context=`; results are bit-identical.
        Valid until the weights change; never used when autograd is recording."""
        return PreparedContext(self._embed_context(context, clip_fea))

    def _embed_context(self, context, clip_fea):
        dev = self.patch_embedding.weight.device
        ctx = torch.stack([torch.cat([u, u.new_zeros(self.text_len - u.size(0), u.size(1))]) for u in context])
        ctx = ctx.to(device=dev, dtype=torch.bfloat16)
        B = ctx.shape[0]
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.text_embedding.parameters()):
            h = self.text_embedding[0](ctx)                              # LinearFn (tcgen05 fwd, dgrad, wgrad)
            h = F.gelu(h.float(), approximate="tanh").to(torch.bfloat16)
            ctx = self.text_embedding[2](h)
        else:
            w0, b0 = self.text_embedding[0].operands()
            w2, b2 = self.text_embedding[2].operands()
            h = ops.gemm(ctx.view(B * self.text_len, -1), w0, bias=b0, epi=ops.EPI_BF16_GELU)
            ctx = ops.gemm(h, w2, bias=b2, epi=ops.EPI_BF16).view(B, self.text_len, self.dim)
        if clip_fea is not None:
            ctx = torch.cat([self.img_emb(clip_fea.to(dev)).to(torch.bfloat16), ctx], dim=1)
        return ctx

    @no_autocast
    def forward(self, x, t, context, seq_len, clip_fea=None, y=None, cond_flag=False, output_features=False,
                selected_layers=[20, 30, 40], gather_features=True):
        """Same contract as the reference (model.py:534-681): x list of [C_in, F, H, W], t [B], context list of
        [L, C]; returns a list of [C_out, F, H, W] fp32 tensors, or with output_features the list of
        [B, L, dim] fp32 features after the selected (1-based) blocks.  `gather_features=False` (extension, sequence
        parallelism only) returns each rank's own [B, L/P, dim] token chunk instead of all-gathering it (for
        `QueryAttention(..., sp_local=True)`)."""
        prepared = context if isinstance(context, PreparedContext) else None
        if self.model_type in ("i2v", "flf2v"):
            assert (clip_fea is not None or prepared is not None) and y is not None
        assert not self.enable_teacache, "teacache is an inference-time heuristic outside this path"
        dev = self.patch_embedding.weight.device
        wp, bp = self._patch_operands()
        train = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or any(u.requires_grad for u in x))
        sp_on = get_sequence_parallel_state()
        sp_P = nccl_info.sp_size if sp_on else 1
        sp_r = nccl_info.rank_within_group if sp_on else 0
        local_only = sp_on and not train                 # no-grad under Ulysses: embed only this rank's token chunk
        if local_only:
            assert seq_len % sp_P == 0
            M_loc = seq_len // sp_P
            xs = torch.empty(len(x), M_loc, self.dim, dtype=torch.float32, device=dev)
        embs, grids, lens = [], [], []
        for i, u in enumerate(x):
            yi = None if y is None else y[i].to(device=dev)
            grids.append((u.shape[1] // self.patch_size[0], u.shape[2] // self.patch_size[1], u.shape[3] // self.patch_size[2]))
            lens.append(grids[-1][0] * grids[-1][1] * grids[-1][2])
            if train:
                from .engine import PatchEmbedFn
                embs.append(PatchEmbedFn.apply(u.to(dev), yi, self.patch_embedding.weight, self.patch_embedding.bias, self))
            else:
                yf = None if yi is None else yi.to(torch.float32).contiguous()
                patches = ops.patchify(u.to(device=dev, dtype=torch.float32).contiguous(), yf)
                if local_only:
                    # rows [r*M, (r+1)*M) of the padded sequence: the patch-embedding GEMM and the fp32 stream exist only for
                    # them (the reference embeds and zero-fills all L tokens on every rank, then chunks: model.py:578-619)
                    lo, hi = sp_r * M_loc, min((sp_r + 1) * M_loc, lens[-1])
                    if hi > lo:
                        xs[i, :hi - lo] = ops.gemm(patches[lo:hi], wp, bias=bp, epi=ops.EPI_BF16)
                    if hi - lo < M_loc:
                        xs[i, max(hi - lo, 0):].zero_()
                else:
                    embs.append(ops.gemm(patches, wp, bias=bp, epi=ops.EPI_BF16))    # [L_i, dim] bf16, like the autocast conv
        seq_lens = torch.tensor(lens, dtype=torch.long)
        grid_sizes = torch.tensor(grids, dtype=torch.long)
        assert int(seq_lens.max()) <= seq_len
        if train:
            xs = torch.stack([torch.cat([e_.float(), e_.new_zeros(seq_len - e_.size(0), self.dim, dtype=torch.float32)]) for e_ in embs])
        elif not local_only:
            xs = torch.zeros(len(embs), seq_len, self.dim, dtype=torch.float32, device=dev)
            for i, e_ in enumerate(embs):
                xs[i, :e_.size(0)] = e_

        # time embeddings — fp32 (model.py:589-594); tiny, PyTorch
        tt = t.to(dev)
        te = self.time_embedding
        e = F.linear(sinusoidal_embedding_1d(self.freq_dim, tt).float(), te[0].weight.float(), te[0].bias.float())
        e = F.linear(F.silu(e), te[2].weight.float(), te[2].bias.float())
        tp = self.time_projection[1]
        e0 = F.linear(F.silu(e), tp.weight.float(), tp.bias.float()).unflatten(1, (6, self.dim))

        use_kv_cache = prepared is not None and not torch.is_grad_enabled()
        if prepared is not None and torch.is_grad_enabled() and any(p.requires_grad for p in self.text_embedding.parameters()):
            raise RuntimeError("a PreparedContext is detached from the text / CLIP embedding weights: use it for no-grad forwards "
                               "only, or pass the raw context when those weights are being trained")
        ctx = prepared.embedded if prepared is not None else self._embed_context(context, clip_fea)

        if sp_on and not local_only:
            xs = torch.chunk(xs, nccl_info.sp_size, dim=1)[nccl_info.rank_within_group].contiguous()

        features_list = []
        for index, block in enumerate(self.blocks):
            if use_kv_cache:
                block.cross_attn.__dict__["_kv_cache"] = prepared.kv.setdefault(index, {})
            try:
                xs = block(xs, e=e0, seq_lens=seq_lens, grid_sizes=grid_sizes, freqs=self.freqs, context=ctx,
                           context_lens=None, first_block_bf16_input=(index == 0))
            finally:
                block.cross_attn.__dict__.pop("_kv_cache", None)
            if output_features and index + 1 in selected_layers:
                if get_sequence_parallel_state() and gather_features:
                    features_list.append(all_gather(xs, dim=1))
                else:
                    features_list.append(xs if xs.requires_grad else xs.clone())
        if output_features:
            return features_list

        out = self.head(xs, e)
        if get_sequence_parallel_state():
            out = all_gather(out, dim=1)
        return [u.float() for u in self.unpatchify(out, grid_sizes, self.out_dim)]

    def unpatchify(self, x, grid_sizes, c):
        """model.py:683-705."""
        return [_UnpatchifyFn.apply(u[:math.prod(g)].float().contiguous(), c, g) for u, g in zip(x, _grid_list(grid_sizes))]

    def init_weights(self):
        """model.py:707-729."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.xavier_uniform_(self.patch_embedding.weight.flatten(1))
        for m in self.text_embedding.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=.02)
        for m in self.time_embedding.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, std=.02)
        nn.init.zeros_(self.head.head.weight)

    _CONFIG_KEYS = ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
                    "num_heads", "num_layers", "window_size", "qk_norm", "cross_attn_norm", "eps")

    @classmethod
    def from_config(cls, config):
        """`ModelMixin.from_config` as the trainers use it (train_prfl.py:205): a dict (a parsed config.json) -> a randomly
        initialised model.  Keys that are not constructor arguments (`_class_name`, `_diffusers_version`, `dtype`, LoRA
        annotations) are ignored; constructor arguments the file omits (the reference's `ignore_for_config` list) take
        their defaults."""
        return cls(**{k: config[k] for k in cls._CONFIG_KEYS if k in config})

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder=None, torch_dtype=None, strict=True, **unused):
        """`WanModel.from_pretrained(dir)` — how every reference trainer and pipeline obtains the model (train_prfl.py:182-217,
        train_pavrm.py:170-181, inference_pavrm.py:171-182, text2video.py; diffusers' `ModelMixin.from_pretrained` in the
        reference): `config.json` -> `from_config`, weights from `diffusion_pytorch_model.safetensors` or its sharded form
        (the layout `checkpoint.save_checkpoint` and the reference's `save_checkpoint` write), strict key match, `eval()`.
        A local directory only (no hub); `torch_dtype` casts the floating-point parameters after loading.  With
        `strict=False` returns `(model, missing_keys, unexpected_keys)`."""
        import json
        import os
        from .checkpoint import load_model_dir
        path = str(pretrained_model_name_or_path)
        if subfolder:
            path = os.path.join(path, subfolder)
        cfg_path = os.path.join(path, "config.json")
        if not os.path.isfile(cfg_path):
            raise FileNotFoundError(f"{cfg_path} not found: from_pretrained takes a local checkpoint directory")
        with open(cfg_path) as f:
            config = json.load(f)
        model = cls.from_config(config)
        extra = {k: v for k, v in config.items() if k not in cls._CONFIG_KEYS and not k.startswith("_")}
        model.config.update(extra)                                   # e.g. lora_rank annotations survive a save / load cycle
        state = load_model_dir(path)
        missing, unexpected = model.load_state_dict(state, strict=False)
        if strict and (missing or unexpected):
            raise RuntimeError(f"{path}: state dict does not match {cls.__name__}({ {k: config.get(k) for k in ('model_type', 'dim', 'num_layers')} }): "
                               f"{len(missing)} missing (first: {missing[:3]}), {len(unexpected)} unexpected (first: {unexpected[:3]})")
        if torch_dtype is not None:
            model = model.to(torch_dtype)
        model.eval()
        return model if strict else (model, list(missing), list(unexpected))

    def save_pretrained(self, save_directory, max_shard_size: int = 5 * 1024 ** 3, state_dict=None):
        """Write `save_directory/{config.json, diffusion_pytorch_model*.safetensors[, index]}` — the directory
        `from_pretrained` (this one and the reference's) reads.  `state_dict`: see `checkpoint.save_checkpoint`."""
        from .checkpoint import _config_dict, write_model_dir
        cfg = {"_class_name": type(self).__name__, **_config_dict(self)}
        write_model_dir(str(save_directory), self.state_dict() if state_dict is None else state_dict, cfg, max_shard_size)

"""Tensor-level wrappers over the C ABI (include/prfl_b200.h).  PyTorch is used only to own device
memory and to name the current stream; every FLOP / byte moved here happens in libprfl_b200.so."""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import EPI_BF16, EPI_BF16_DGELU, EPI_BF16_GELU, EPI_F32, EPI_RESIDUAL, check, lib

bf16, f32 = torch.bfloat16, torch.float32


def _stream():
    return torch.cuda.current_stream().cuda_stream


class KernelTimer:
    """Optional CUDA-event timing of individual launches on the launching (current) stream; bench.py
    switches it on for the timed region to get the dominant kernel's live duration for the roofline."""

    def __init__(self, names):
        self.names = set(names)
        self.events = {n: [] for n in names}

    def elapsed_ms(self, name):
        return [a.elapsed_time(b) for a, b in self.events[name]]


TIMER: Optional[KernelTimer] = None


class _timed:
    def __init__(self, name):
        self.on = TIMER is not None and name in TIMER.names
        self.name = name

    def __enter__(self):
        if self.on:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()

    def __exit__(self, *exc):
        if self.on:
            self.b.record()
            TIMER.events[self.name].append((self.a, self.b))


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str):
    if not t.is_cuda:
        raise _lib.PrflError(f"{name}: tensor must live on a CUDA device (prfl_b200 has no CPU path)")
    if t.dtype != dtype:
        raise _lib.PrflError(f"{name}: expected {dtype}, got {t.dtype}")


# ------------------------------------------------------------------------------------------------
def ln_mod(x: torch.Tensor, shift=None, scale=None, gamma=None, beta=None, eps: float = 1e-6,
           round_bf16: bool = False, save_stats: bool = False):
    """LayerNorm(+affine)(+modulate): fp32 [rows, C] -> bf16 [rows, C]  (model.py:345,352,353)."""
    _req(x, f32, "ln_mod.x")
    assert x.is_contiguous()
    C = x.shape[-1]
    rows = x.numel() // C
    out = torch.empty(x.shape, dtype=bf16, device=x.device)
    mean = torch.empty(rows, dtype=f32, device=x.device) if save_stats else None
    rstd = torch.empty(rows, dtype=f32, device=x.device) if save_stats else None
    for t, n in ((shift, "shift"), (scale, "scale"), (gamma, "gamma"), (beta, "beta")):
        if t is not None:
            _req(t, f32, "ln_mod." + n)
            assert t.numel() == C and t.is_contiguous()
    check(lib().prfl_ln_mod_fwd(_p(x), _p(shift), _p(scale), _p(gamma), _p(beta), _p(out), _p(mean), _p(rstd), rows, C,
                                float(eps), int(round_bf16), _stream()), "prfl_ln_mod_fwd")
    return (out, mean, rstd) if save_stats else out


def ln_mod_split(x: torch.Tensor, shift, scale, eps: float = 1e-6):
    """LayerNorm + modulate with the result split into a bf16 (hi, lo) pair: hi + lo == fp32 result to ~16 bits."""
    _req(x, f32, "ln_mod_split.x")
    assert x.is_contiguous()
    C = x.shape[-1]
    rows = x.numel() // C
    hi = torch.empty(x.shape, dtype=bf16, device=x.device)
    lo = torch.empty(x.shape, dtype=bf16, device=x.device)
    check(lib().prfl_ln_mod_split_fwd(_p(x), _p(shift), _p(scale), _p(hi), _p(lo), rows, C, float(eps), _stream()),
          "prfl_ln_mod_split_fwd")
    return hi, lo


def rmsnorm_rope_(x: torch.Tensor, w: torch.Tensor, cos: Optional[torch.Tensor], sin: Optional[torch.Tensor],
                  eps: float, n_rot: int = 0, pos0: int = 0, out: Optional[torch.Tensor] = None,
                  save_rstd: bool = False):
    """RMSNorm over the whole channel dim (+RoPE) on a bf16 [rows, C] view whose rows may be strided
    (e.g. the q or k slice of a fused QKV buffer).  In place unless `out` is given."""
    _req(x, bf16, "rmsnorm_rope.x")
    _req(w, f32, "rmsnorm_rope.w")
    assert x.dim() == 2 and x.stride(1) == 1
    rows, C = x.shape
    o = x if out is None else out
    assert o.shape == x.shape and o.stride(1) == 1 and o.dtype == bf16
    rstd = torch.empty(rows, dtype=f32, device=x.device) if save_rstd else None
    if cos is not None:
        _req(cos, f32, "rmsnorm_rope.cos")
        assert cos.shape[-1] == 64 and cos.is_contiguous() and sin.is_contiguous()
        assert pos0 + min(n_rot, rows) <= cos.shape[0]
    check(lib().prfl_rmsnorm_rope_fwd(_p(x), x.stride(0), _p(w), _p(cos), _p(sin), _p(o), o.stride(0), _p(rstd), rows, C,
                                      int(n_rot), int(pos0), float(eps), _stream()), "prfl_rmsnorm_rope_fwd")
    return (o, rstd) if save_rstd else o


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_trans: bool = False, b_trans: bool = False,
         bias: Optional[torch.Tensor] = None, epi: int = EPI_BF16, out: Optional[torch.Tensor] = None,
         gate: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None, beta: bool = False,
         resid: Optional[torch.Tensor] = None) -> torch.Tensor:
    """acc[m, n] = sum_k A(m, k) B(n, k) on tcgen05 with a fused epilogue.
    a: [M, K] (or [K, M] if a_trans), b: [N, K] (or [K, N] if b_trans); inner stride 1, row stride free.
    EPI_RESIDUAL: out = resid + gate * bf16(acc + bias); resid defaults to out (in place)."""
    _req(a, bf16, "gemm.a")
    _req(b, bf16, "gemm.b")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    M, K = (a.shape[1], a.shape[0]) if a_trans else a.shape
    N, Kb = (b.shape[1], b.shape[0]) if b_trans else b.shape
    assert K == Kb, (a.shape, b.shape, a_trans, b_trans)
    f32_out = epi in (EPI_F32, EPI_RESIDUAL)
    if out is None:
        assert (epi != EPI_RESIDUAL or resid is not None) and not beta
        out = torch.empty(M, N, dtype=f32 if f32_out else bf16, device=a.device)
    if resid is not None:
        _req(resid, f32, "gemm.resid")
        assert epi == EPI_RESIDUAL and resid.shape == (M, N) and resid.stride(1) == 1 and resid.stride(0) == out.stride(0)
    _req(out, f32 if f32_out else bf16, "gemm.out")
    assert out.shape == (M, N) and out.stride(1) == 1
    if bias is not None:
        _req(bias, f32, "gemm.bias")
        assert bias.numel() == N
    if gate is not None:
        _req(gate, f32, "gemm.gate")
        assert gate.numel() == N
    ldaux = 0
    if aux is not None:          # input for DGELU; extra output (pre-activation / un-gated branch) for GELU / RESIDUAL
        _req(aux, bf16, "gemm.aux")
        assert aux.shape == (M, N) and aux.stride(1) == 1
        ldaux = aux.stride(0)
    with _timed("gemm"):
        check(lib().prfl_gemm_bf16(_p(a), a.stride(0), int(a_trans), _p(b), b.stride(0), int(b_trans), _p(out), out.stride(0),
                                   _p(bias), _p(gate), _p(resid), _p(aux), ldaux, M, N, K, epi, int(beta), _stream()), "prfl_gemm_bf16")
    return out


def _attn_ws(Lq: int, Lk: int, H: int, device):
    """Workspace for the key-split tail of the forward kernel (wave quantisation, csrc/attention_fwd.cu); None if unused."""
    n = int(lib().prfl_attn_fwd_ws_bytes(Lq, Lk, H))
    return torch.empty(n // 4, dtype=f32, device=device) if n > 0 else None


def attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: Optional[float] = None,
             out: Optional[torch.Tensor] = None, need_lse: bool = False):
    """q: [Lq, H, 128], k/v: [Lk, H, 128] bf16 views (last stride 1; token / head strides free)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, bf16, "attn." + n)
        assert t.dim() == 3 and t.shape[2] == 128 and t.stride(2) == 1, (n, t.shape, t.stride())
    Lq, H, _ = q.shape
    Lk = k.shape[0]
    assert k.shape == v.shape and k.shape[1] == H
    if out is None:
        out = torch.empty(Lq, H, 128, dtype=bf16, device=q.device)
    lse = torch.empty(H, Lq, dtype=f32, device=q.device) if need_lse else None
    if scale is None:
        scale = 1.0 / math.sqrt(128)
    ws = _attn_ws(Lq, Lk, H, q.device)
    with _timed("attn_fwd_self" if Lk > 1024 else "attn_fwd_cross"):
        check(lib().prfl_attn_fwd(_p(q), q.stride(0), q.stride(1), _p(k), k.stride(0), k.stride(1), _p(v), v.stride(0), v.stride(1),
                                  _p(out), out.stride(0), out.stride(1), _p(lse), Lq, Lk, H, float(scale), _p(ws), _stream()),
              "prfl_attn_fwd")
    return (out, lse) if need_lse else out


def attn_bwd(q, k, v, o, dout, lse, dq=None, dk=None, dv=None, scale: Optional[float] = None):
    """Gradients of attn_fwd.  All tensors [L, H, 128] bf16 views (last stride 1); lse [H, Lq] fp32 from the forward."""
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (o, "o"), (dout, "dout")):
        _req(t, bf16, "attn_bwd." + n)
        assert t.dim() == 3 and t.shape[2] == 128 and t.stride(2) == 1, (n, t.shape, t.stride())
    Lq, H, _ = q.shape
    Lk = k.shape[0]
    dq = torch.empty(Lq, H, 128, dtype=bf16, device=q.device) if dq is None else dq
    dk = torch.empty(Lk, H, 128, dtype=bf16, device=q.device) if dk is None else dk
    dv = torch.empty(Lk, H, 128, dtype=bf16, device=q.device) if dv is None else dv
    ws = torch.empty(int(lib().prfl_attn_bwd_ws_floats(Lq, H)), dtype=f32, device=q.device)
    _req(lse, f32, "attn_bwd.lse")
    assert lse.shape == (H, Lq) and lse.is_contiguous()
    if scale is None:
        scale = 1.0 / math.sqrt(128)
    args = []
    for t in (q, k, v, o, dout):
        args += [_p(t), t.stride(0), t.stride(1)]
    args += [_p(lse), _p(ws)]
    for t in (dq, dk, dv):
        assert t.dtype == bf16 and t.stride(2) == 1
        args += [_p(t), t.stride(0), t.stride(1)]
    with _timed("attn_bwd_self" if Lk > 1024 else "attn_bwd_cross"):
        check(lib().prfl_attn_bwd(*args, Lq, Lk, H, float(scale), _stream()), "prfl_attn_bwd")
    return dq, dk, dv


def attn_merge_(o_acc: torch.Tensor, lse_acc: torch.Tensor, o_new: torch.Tensor, lse_new: torch.Tensor, first: bool,
                out: Optional[torch.Tensor] = None):
    """Ring-attention merge: fold (o_new bf16 [L, H, 128] view, lse_new [H, L]) into the running fp32 (o_acc [L, H, 128]
    contiguous, lse_acc [H, L]); `out` (bf16 [L, H, 128] view) also receives bf16(o_acc)."""
    _req(o_acc, f32, "attn_merge.o_acc")
    _req(lse_acc, f32, "attn_merge.lse_acc")
    _req(o_new, bf16, "attn_merge.o_new")
    _req(lse_new, f32, "attn_merge.lse_new")
    L, H, d = o_new.shape
    assert d == 128 and o_new.stride(2) == 1 and o_acc.is_contiguous() and o_acc.shape == (L, H, 128)
    assert lse_acc.shape == (H, L) and lse_new.shape == (H, L) and lse_acc.is_contiguous() and lse_new.is_contiguous()
    if out is not None:
        _req(out, bf16, "attn_merge.out")
        assert out.shape == (L, H, 128) and out.stride(2) == 1
    check(lib().prfl_attn_merge(_p(o_acc), _p(lse_acc), _p(o_new), o_new.stride(0), o_new.stride(1), _p(lse_new), int(first),
                                _p(out), 0 if out is None else out.stride(0), 0 if out is None else out.stride(1), L, H, _stream()),
          "prfl_attn_merge")
    return o_acc


def colsum_parts(rows: int) -> int:
    return (rows + 255) // 256


def ln_mod_bwd(x, dy, scale, gamma, mean, rstd, dx_accum, need_param_grads: bool):
    """dx_accum += dL/dx; returns (sum dy, sum dy*xhat) as [C] fp32 (or (None, None))."""
    _req(x, f32, "ln_mod_bwd.x")
    _req(dy, bf16, "ln_mod_bwd.dy")
    _req(dx_accum, f32, "ln_mod_bwd.dx")
    assert x.is_contiguous() and dy.is_contiguous() and dx_accum.is_contiguous()
    rows, C = x.shape
    p1 = p2 = None
    if need_param_grads:
        p1 = torch.empty(colsum_parts(rows), C, dtype=f32, device=x.device)
        p2 = torch.empty_like(p1)
    check(lib().prfl_ln_mod_bwd(_p(x), _p(dy), _p(scale), _p(gamma), _p(mean), _p(rstd), _p(dx_accum), _p(p1), _p(p2), rows, C,
                                _stream()), "prfl_ln_mod_bwd")
    if need_param_grads:
        return p1.sum(0), p2.sum(0)
    return None, None


def rmsnorm_rope_bwd_(x, w, cos, sin, dy, rstd, n_rot: int = 0, pos0: int = 0, need_dw: bool = True):
    """dy (bf16 [rows, C] view) is overwritten with dL/dx; returns dL/dw [C] fp32 (or None)."""
    _req(x, bf16, "rmsnorm_rope_bwd.x")
    _req(dy, bf16, "rmsnorm_rope_bwd.dy")
    rows, C = x.shape
    gw = torch.empty(rows, C, dtype=bf16, device=x.device) if need_dw else None
    check(lib().prfl_rmsnorm_rope_bwd(_p(x), x.stride(0), _p(w), _p(cos), _p(sin), _p(dy), dy.stride(0), _p(rstd), _p(dy),
                                      dy.stride(0), _p(gw), C if need_dw else 0, rows, C, int(n_rot), int(pos0), _stream()),
          "prfl_rmsnorm_rope_bwd")
    return colsum(gw) if need_dw else None


def colsum(a: torch.Tensor) -> torch.Tensor:
    """Column sums of a bf16 [rows, N] view (row stride free) -> [N] fp32."""
    _req(a, bf16, "colsum.a")
    assert a.dim() == 2 and a.stride(1) == 1
    rows, N = a.shape
    part = torch.empty(colsum_parts(rows), N, dtype=f32, device=a.device)
    check(lib().prfl_colsum_bf16(_p(a), a.stride(0), _p(part), rows, N, _stream()), "prfl_colsum_bf16")
    return part.sum(0)


def gate_bwd(dx: torch.Tensor, y: Optional[torch.Tensor], gate: Optional[torch.Tensor]):
    """dy = bf16(dx * gate) and dgate = sum_rows dx * y (None if y is None)."""
    _req(dx, f32, "gate_bwd.dx")
    assert dx.is_contiguous() and dx.dim() == 2
    rows, N = dx.shape
    dy = torch.empty(rows, N, dtype=bf16, device=dx.device)
    part = torch.empty(colsum_parts(rows), N, dtype=f32, device=dx.device) if y is not None else None
    if y is not None:
        _req(y, bf16, "gate_bwd.y")
        assert y.is_contiguous()
    check(lib().prfl_gate_bwd(_p(dx), _p(y), _p(gate), _p(dy), _p(part), rows, N, _stream()), "prfl_gate_bwd")
    return dy, (part.sum(0) if part is not None else None)


def _ptr_array(ptrs):
    import ctypes
    return (ctypes.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def a2a_scatter_p2p(strided: torch.Tensor, peer_recv_ptrs, P: int, rank: int):
    """q/k/v exchange as direct NVLink stores: peer[p][(rank*L_loc + t), hl, :] = strided[t, p*H/P + hl, :]."""
    _req(strided, bf16, "a2a_scatter_p2p.src")
    L_loc, H, d = strided.shape
    assert d == 128 and strided.stride(2) == 1 and len(peer_recv_ptrs) == P
    check(lib().prfl_a2a_scatter_p2p(_p(strided), strided.stride(0), strided.stride(1), _ptr_array(peer_recv_ptrs), L_loc, H, P,
                                     rank, _stream()), "prfl_a2a_scatter_p2p")


def a2a_gather_p2p(src: torch.Tensor, peer_dst_ptrs, dst_ld_tok: int, dst_ld_head: int, P: int, rank: int):
    """The reverse exchange as direct NVLink stores: src [P*L_loc, Hl, 128] (this rank's heads, all tokens);
    peer[p][t*dst_ld_tok + (rank*Hl + hl)*dst_ld_head + :] = src[p*L_loc + t, hl, :]."""
    _req(src, bf16, "a2a_gather_p2p.src")
    L, Hl, d = src.shape
    assert d == 128 and src.stride(2) == 1 and L % P == 0 and len(peer_dst_ptrs) == P
    check(lib().prfl_a2a_gather_p2p(_p(src), src.stride(0), src.stride(1), _ptr_array(peer_dst_ptrs), int(dst_ld_tok), int(dst_ld_head),
                                    L // P, Hl, P, rank, _stream()), "prfl_a2a_gather_p2p")


def attn_fwd_p2p(q, k, v, o_peer_ptrs, L_loc: int, head_off: int, H_total: int, scale: Optional[float] = None):
    """attn_fwd whose epilogue stores row i into rank (i // L_loc)'s [L_loc, H_total, 128] buffer (peer pointers)."""
    for t, n in ((q, "q"), (k, "k"), (v, "v")):
        _req(t, bf16, "attn_p2p." + n)
        assert t.dim() == 3 and t.shape[2] == 128 and t.stride(2) == 1
    Lq, H, _ = q.shape
    Lk = k.shape[0]
    if scale is None:
        scale = 1.0 / math.sqrt(128)
    ws = _attn_ws(Lq, Lk, H, q.device)
    with _timed("attn_fwd_self" if Lk > 1024 else "attn_fwd_cross"):
        check(lib().prfl_attn_fwd_p2p(_p(q), q.stride(0), q.stride(1), _p(k), k.stride(0), k.stride(1), _p(v), v.stride(0), v.stride(1),
                                      _ptr_array(o_peer_ptrs), len(o_peer_ptrs), L_loc, head_off, H_total * 128, 128, None, Lq, Lk, H,
                                      float(scale), _p(ws), _stream()), "prfl_attn_fwd_p2p")


def cast_bf16(src: torch.Tensor) -> torch.Tensor:
    _req(src, f32, "cast.src")
    assert src.is_contiguous()
    dst = torch.empty(src.shape, dtype=bf16, device=src.device)
    check(lib().prfl_cast_f32_bf16(_p(src), _p(dst), src.numel(), _stream()), "prfl_cast_f32_bf16")
    return dst


def patchify(x: torch.Tensor, y: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[Cx, F, H, W] (+ [Cy, F, H, W]) fp32 -> [F*(H/2)*(W/2), (Cx+Cy)*4] bf16 (model.py:574-581)."""
    _req(x, f32, "patchify.x")
    assert x.is_contiguous() and x.dim() == 4
    Cx, F, H, W = x.shape
    Cy = 0
    if y is not None:
        _req(y, f32, "patchify.y")
        assert y.is_contiguous() and y.shape[1:] == x.shape[1:]
        Cy = y.shape[0]
    out = torch.empty(F * (H // 2) * (W // 2), (Cx + Cy) * 4, dtype=bf16, device=x.device)
    check(lib().prfl_patchify(_p(x), Cx, _p(y), Cy, _p(out), F, H, W, _stream()), "prfl_patchify")
    return out


def patchify_bwd(dpatches: torch.Tensor, Cx: int, F: int, H: int, W: int) -> torch.Tensor:
    _req(dpatches, f32, "patchify_bwd.dp")
    assert dpatches.is_contiguous()
    Ct = dpatches.shape[1] // 4
    dx = torch.empty(Cx, F, H, W, dtype=f32, device=dpatches.device)
    check(lib().prfl_patchify_bwd(_p(dpatches), Cx, Ct, _p(dx), F, H, W, _stream()), "prfl_patchify_bwd")
    return dx


def unpatchify(tokens: torch.Tensor, c: int, grid: Tuple[int, int, int]) -> torch.Tensor:
    """[F*h*w, 4c] fp32 -> [c, F, 2h, 2w] fp32 (model.py:683-705)."""
    _req(tokens, f32, "unpatchify.tokens")
    assert tokens.is_contiguous()
    F, h, w = grid
    assert tokens.shape[0] >= F * h * w and tokens.shape[1] == 4 * c
    vid = torch.empty(c, F, 2 * h, 2 * w, dtype=f32, device=tokens.device)
    check(lib().prfl_unpatchify(_p(tokens), _p(vid), c, F, h, w, 0, _stream()), "prfl_unpatchify")
    return vid


def unpatchify_bwd(dvid: torch.Tensor, rows: int) -> torch.Tensor:
    _req(dvid, f32, "unpatchify_bwd.dvid")
    assert dvid.is_contiguous()
    c, F, H2, W2 = dvid.shape
    tok = torch.zeros(rows, 4 * c, dtype=f32, device=dvid.device)
    check(lib().prfl_unpatchify(_p(tok), _p(dvid), c, F, H2 // 2, W2 // 2, 1, _stream()), "prfl_unpatchify(inverse)")
    return tok


def sq_pool(x: torch.Tensor, wk_eff: torch.Tensor):
    """Single-query attention pooling: x [L, C] fp32, wk_eff [8, C] fp32 -> pooled [8, C], (scores, stats)."""
    _req(x, f32, "sq_pool.x")
    _req(wk_eff, f32, "sq_pool.wk_eff")
    assert x.is_contiguous() and wk_eff.is_contiguous() and x.dim() == 2
    L, Cc = x.shape
    NH = wk_eff.shape[0]
    scores = torch.empty(L, NH, dtype=f32, device=x.device)
    stats = torch.empty(2 * NH, dtype=f32, device=x.device)
    pooled = torch.empty(NH, Cc, dtype=f32, device=x.device)
    check(lib().prfl_sq_pool_fwd(_p(x), _p(wk_eff), _p(scores), _p(stats), _p(pooled), L, Cc, NH, _stream()), "prfl_sq_pool_fwd")
    return pooled, scores, stats


def sq_pool_bwd(x, wk_eff, scores, stats, pooled, dpooled, dx: Optional[torch.Tensor] = None, need_ds: bool = False):
    L, Cc = x.shape
    NH = wk_eff.shape[0]
    acc = dx is not None
    if dx is None:
        dx = torch.empty_like(x)
    ds = torch.empty(L, NH, dtype=f32, device=x.device) if need_ds else None
    _req(dpooled, f32, "sq_pool_bwd.dpooled")
    check(lib().prfl_sq_pool_bwd(_p(x), _p(wk_eff), _p(scores), _p(stats), _p(pooled), _p(dpooled.contiguous()), _p(dx), _p(ds),
                                 L, Cc, NH, int(acc), _stream()), "prfl_sq_pool_bwd")
    return dx, ds


def a2a_pack(strided: torch.Tensor, packed: torch.Tensor, P: int, unpack: bool = False):
    """strided: [L_loc, H, 128] bf16 view; packed: contiguous [P, L_loc, H/P, 128] bf16."""
    _req(strided, bf16, "a2a.strided")
    _req(packed, bf16, "a2a.packed")
    L_loc, H, d = strided.shape
    assert d == 128 and strided.stride(2) == 1 and packed.is_contiguous() and packed.numel() == strided.numel()
    check(lib().prfl_a2a_pack(_p(strided), strided.stride(0), strided.stride(1), _p(packed), L_loc, H, P, int(unpack), _stream()),
          "prfl_a2a_pack")
    return strided if unpack else packed


# ------------------------------------------------------------------------------------------------
# PRFL chain glue (scheduler step)
# ------------------------------------------------------------------------------------------------
def unipc_step(sample, model_output, last_sample, hist, sigma: float, corr_coef, pred_coef, model_output_uncond=None,
               guide_scale: float = 1.0):
    """One fused FlowUniPC update (fm_solvers_unipc.py:655-739).  All tensors fp32, same shape, contiguous.
    hist: previous x0 predictions, newest first (up to 3).  corr_coef: 5 floats or None; pred_coef: 5 floats.
    With model_output_uncond the velocity is uncond + guide_scale * (model_output - uncond) (text2video.py:295-296).
    Returns (x0, corrected | None, prev)."""
    import ctypes as C
    for t, n in ((sample, "sample"), (model_output, "model_output"), (model_output_uncond, "model_output_uncond")):
        if t is None:
            continue
        _req(t, f32, "unipc_step." + n)
        assert t.is_contiguous() and t.shape == sample.shape
    hs = [None, None, None]
    for i, h in enumerate(hist[:3]):
        if h is not None:
            _req(h, f32, "unipc_step.hist")
            assert h.is_contiguous() and h.shape == sample.shape
            hs[i] = h
    x0 = torch.empty_like(sample)
    prev = torch.empty_like(sample)
    corrected = None
    cc = None
    if corr_coef is not None:
        _req(last_sample, f32, "unipc_step.last_sample")
        assert last_sample.is_contiguous() and last_sample.shape == sample.shape
        corrected = torch.empty_like(sample)
        cc = (C.c_float * 5)(*[float(c) for c in corr_coef])
    pc = (C.c_float * 5)(*[float(c) for c in pred_coef])
    check(lib().prfl_unipc_step(_p(sample), _p(model_output), _p(model_output_uncond), float(guide_scale),
                                _p(last_sample) if cc is not None else None, _p(hs[0]), _p(hs[1]), _p(hs[2]), float(sigma), cc, pc,
                                _p(x0), _p(corrected), _p(prev), sample.numel(), _stream()),
          "prfl_unipc_step")
    return x0, corrected, prev


def scale2(g, a: float, b: Optional[float] = None):
    """(a * g, b * g) in one pass (b optional) — backward of unipc_step."""
    _req(g, f32, "scale2.g")
    g = g.contiguous()
    ya = torch.empty_like(g)
    yb = torch.empty_like(g) if b is not None else None
    check(lib().prfl_scale2_f32(_p(g), float(a), _p(ya), float(b or 0.0), _p(yb), g.numel(), _stream()), "prfl_scale2_f32")
    return ya, yb


# ------------------------------------------------------------------------------------------------
# sharded optimizer
# ------------------------------------------------------------------------------------------------
def adamw_step_(grad, master, exp_avg, exp_avg_sq, step: int, lr: float, betas, eps: float, weight_decay: float, clip_coef=None,
                bf16_out: Optional[torch.Tensor] = None):
    """In-place AdamW on one fp32 shard (all four tensors flat, same length).  clip_coef: 0-dim CUDA fp32 tensor or None.
    bf16_out: flat bf16 tensor of the same length receiving bf16(master) in the same pass (the resident compute copy)."""
    for t, n in ((grad, "grad"), (master, "master"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _req(t, f32, "adamw_step." + n)
        assert t.is_contiguous() and t.numel() == master.numel()
    if clip_coef is not None:
        _req(clip_coef, f32, "adamw_step.clip_coef")
    if bf16_out is not None:
        _req(bf16_out, bf16, "adamw_step.bf16_out")
        assert bf16_out.is_contiguous() and bf16_out.numel() == master.numel()
    check(lib().prfl_adamw_step(_p(grad), _p(master), _p(exp_avg), _p(exp_avg_sq), _p(clip_coef), _p(bf16_out), master.numel(), float(lr),
                                float(betas[0]), float(betas[1]), float(eps), float(weight_decay), int(step), _stream()), "prfl_adamw_step")


def sumsq_(x, acc):
    """acc (0-dim CUDA float64) += sum(x^2)."""
    _req(x, f32, "sumsq.x")
    _req(acc, torch.float64, "sumsq.acc")
    check(lib().prfl_sumsq_f32(_p(x.contiguous()), x.numel(), _p(acc), _stream()), "prfl_sumsq_f32")
    return acc


# ------------------------------------------------------------------------------------------------
# NVTX ranges per kernel family: PRFL_NVTX=1 wraps every op of this module in a range named
# "prfl/<family>/<op>" so that nsys / ncu --nvtx timelines group launches by family; off by default (two extra
# host calls per launch).
# ------------------------------------------------------------------------------------------------
NVTX_FAMILIES = {
    "norm": ("ln_mod", "ln_mod_split", "rmsnorm_rope_", "ln_mod_bwd", "rmsnorm_rope_bwd_", "colsum", "gate_bwd", "cast_bf16"),
    "gemm": ("gemm",),
    "attention": ("attn_fwd", "attn_bwd", "attn_fwd_p2p", "attn_merge_"),
    "exchange": ("a2a_pack", "a2a_scatter_p2p", "a2a_gather_p2p"),
    "patch": ("patchify", "patchify_bwd", "unpatchify", "unpatchify_bwd"),
    "reward": ("sq_pool", "sq_pool_bwd"),
    "scheduler": ("unipc_step", "scale2"),
    "optimizer": ("adamw_step_", "sumsq_"),
}


def install_nvtx(enable: Optional[bool] = None) -> bool:
    import functools
    import os
    if enable is None:
        enable = os.environ.get("PRFL_NVTX", "0") == "1"
    if not enable or globals().get("_NVTX_ON"):
        return bool(globals().get("_NVTX_ON"))
    g = globals()
    for fam, names in NVTX_FAMILIES.items():
        for n in names:
            fn = g[n]

            def wrap(fn=fn, label=f"prfl/{fam}/{n.rstrip('_')}"):
                @functools.wraps(fn)
                def inner(*a, **k):
                    torch.cuda.nvtx.range_push(label)
                    try:
                        return fn(*a, **k)
                    finally:
                        torch.cuda.nvtx.range_pop()
                return inner
            g[n] = wrap()
    g["_NVTX_ON"] = True
    return True


install_nvtx()

"""TEST INFRASTRUCTURE — torch-CPU stand-ins for the C-ABI wrappers of `prfl_b200.ops`, so that the package's HOST logic (the
sequence of kernel calls in engine.py / model.py / network.py / scheduler.py / prfl.py / sampling.py, the hand-written block
backward, the per-sample loops, buffer reuse, in-place epilogues) can be exercised on a machine without a GPU.

Each function computes what `tests/test_kernels_gpu.py` holds the corresponding CUDA kernel to (the same plain-PyTorch fp32
references, including the bf16 rounding points of the precision choreography), with the wrapper's exact calling convention:
in-place updates of strided views, optional outputs, returned tuples.  It is NOT a fallback: nothing under
`hy-video-prfl_b200/` imports it, and the product raises on CPU tensors; tests install it with `monkeypatch`
(`install(monkeypatch)`), which also proves the host code reaches the device only through `prfl_b200.ops`.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

bf16, f32 = torch.bfloat16, torch.float32
EPI_BF16, EPI_BF16_GELU, EPI_F32, EPI_RESIDUAL, EPI_BF16_DGELU = 0, 1, 2, 3, 4          # include/prfl_b200.h:117-121
CALLS = []          # names of the emulated ops in call order (tests read / clear it)


def _log(name):
    CALLS.append(name)


def _ln(x, eps):
    mean = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + eps)
    return (x - mean) * rstd, mean.squeeze(-1), rstd.squeeze(-1)


def ln_mod(x, shift=None, scale=None, gamma=None, beta=None, eps=1e-6, round_bf16=False, save_stats=False):
    _log("ln_mod")
    assert x.dtype == f32 and x.is_contiguous()
    C = x.shape[-1]
    assert C % 256 == 0 and _al16(x, shift, scale, gamma, beta) and (shift is None) == (scale is None) and (gamma is None) == (beta is None)
    assert all(t is None or (t.dtype == f32 and t.numel() == C and t.is_contiguous()) for t in (shift, scale, gamma, beta))
    xh, mean, rstd = _ln(x.reshape(-1, C), eps)
    y = xh
    if gamma is not None:
        y = y * gamma + beta
    if round_bf16:
        y = y.bfloat16().float()
    if shift is not None:
        y = y * (1 + scale) + shift
    out = y.bfloat16().view(x.shape)
    return (out, mean, rstd) if save_stats else out


def ln_mod_split(x, shift, scale, eps=1e-6):
    _log("ln_mod_split")
    xh, _, _ = _ln(x, eps)
    v = xh * (1 + scale) + shift
    hi = v.bfloat16()
    return hi, (v - hi.float()).bfloat16()


def _rope(t, cos, sin, n_rot, pos0, inverse=False):
    """t: [rows, C] fp32; rotate pairs (2j, 2j+1) of every 128-wide head of the first n_rot rows by the angle of row pos0 + r."""
    rows, C = t.shape
    n_rot = min(int(n_rot), rows)
    if cos is None or n_rot <= 0:
        return t
    H = C // 128
    c, s = cos[pos0:pos0 + n_rot, None, :].to(t.dtype), sin[pos0:pos0 + n_rot, None, :].to(t.dtype)
    if inverse:
        s = -s
    r = t[:n_rot].reshape(n_rot, H, 64, 2)
    rot = torch.stack([r[..., 0] * c - r[..., 1] * s, r[..., 0] * s + r[..., 1] * c], -1).reshape(n_rot, C)
    return torch.cat([rot, t[n_rot:]])


def rmsnorm_rope_(x, w, cos, sin, eps, n_rot=0, pos0=0, out=None, save_rstd=False):
    _log("rmsnorm_rope_")
    assert x.dtype == bf16 and w.dtype == f32 and x.dim() == 2 and x.stride(1) == 1
    assert x.shape[1] % 256 == 0 and x.stride(0) % 8 == 0 and _al16(x, w, out, cos, sin), (x.shape, x.stride())
    assert cos is None or (cos.shape[-1] == 64 and pos0 + min(int(n_rot), x.shape[0]) <= cos.shape[0])
    xf = x.float()
    rstd = torch.rsqrt(xf.pow(2).mean(-1) + eps)
    t = (xf * rstd[:, None]).bfloat16().float() * w           # model.py:119: the bf16 rounding sits before the weight
    t = _rope(t, cos, sin, n_rot, pos0)
    o = x if out is None else out
    o.copy_(t.bfloat16())
    return (o, rstd) if save_rstd else o


def rmsnorm_rope_bwd_(x, w, cos, sin, dy, rstd, n_rot=0, pos0=0, need_dw=True):
    _log("rmsnorm_rope_bwd_")
    xf = x.float()
    dt = _rope(dy.float(), cos, sin, n_rot, pos0, inverse=True)        # back through the rotation
    xn = xf * rstd[:, None]
    dw = (dt * xn.bfloat16().float()).sum(0) if need_dw else None
    g = dt * w                                                          # d / d(x * rstd)
    dx = rstd[:, None] * (g - xn * (g * xn).mean(-1, keepdim=True))
    dy.copy_(dx.bfloat16())
    return dw


def _gelu_tanh_grad(x):
    k = math.sqrt(2.0 / math.pi)
    u = k * (x + 0.044715 * x ** 3)
    t = torch.tanh(u)
    return 0.5 * (1 + t) + 0.5 * x * (1 - t * t) * k * (1 + 3 * 0.044715 * x * x)


def _al16(*ts):
    return all(t is None or t.data_ptr() % 16 == 0 for t in ts)


def gemm(a, b, *, a_trans=False, b_trans=False, bias=None, epi=EPI_BF16, out=None, gate=None, aux=None, beta=False, resid=None):
    _log("gemm")
    assert a.dtype == bf16 and b.dtype == bf16 and a.dim() == 2 and b.dim() == 2
    # the argument contract of prfl_gemm_bf16 (csrc/gemm.cu:376-386): a host path that violates it fails on the GPU with PRFL_E_*
    M, K = (a.shape[1], a.shape[0]) if a_trans else a.shape
    N = b.shape[1] if b_trans else b.shape[0]
    assert M > 0 and N > 0 and K > 0 and N % 8 == 0, (M, N, K)
    assert K % 8 == 0 or (a_trans and b_trans), f"K={K} must be a multiple of 8 unless both operands are transposed"
    assert not (a_trans and M % 8), f"transposed A needs M%8==0 (M={M})"
    assert a.stride(1) == 1 and b.stride(1) == 1 and a.stride(0) % 8 == 0 and b.stride(0) % 8 == 0, (a.stride(), b.stride())
    assert _al16(a, b, out, bias, gate, aux, resid), "operands must be 16-byte aligned"
    assert out is None or (out.stride(1) == 1 and out.stride(0) % 4 == 0), out.stride()
    assert aux is None or (aux.stride(1) == 1 and aux.stride(0) % 8 == 0 and aux.shape == (M, N)), aux.stride()
    A = a.float().t() if a_trans else a.float()
    B = b.float() if b_trans else b.float().t()
    acc = A @ B
    if bias is not None:
        acc = acc + bias.float()
    if epi == EPI_BF16:
        res = acc.bfloat16()
    elif epi == EPI_F32:
        res = acc + out if beta else acc
    elif epi == EPI_BF16_GELU:
        pre = acc.bfloat16()
        if aux is not None:
            aux.copy_(pre)
        res = F.gelu(pre.float(), approximate="tanh").bfloat16()
    elif epi == EPI_RESIDUAL:
        y = acc.bfloat16()
        if aux is not None:
            aux.copy_(y)
        base = resid if resid is not None else out
        assert base is not None and base.dtype == f32
        res = base + (y.float() * gate.float() if gate is not None else y.float())
    elif epi == EPI_BF16_DGELU:
        res = (acc * _gelu_tanh_grad(aux.float())).bfloat16()
    else:
        raise ValueError(epi)
    if out is None:
        assert not beta
        return res.contiguous()
    assert out.shape == res.shape and out.dtype == res.dtype, (out.shape, res.shape, out.dtype, res.dtype)
    out.copy_(res)
    return out


def _attn(q, k, v, scale):
    qf, kf, vf = q.float().transpose(0, 1), k.float().transpose(0, 1), v.float().transpose(0, 1)
    s = (qf @ kf.transpose(1, 2)) * scale
    return (torch.softmax(s, -1) @ vf).transpose(0, 1), torch.logsumexp(s, -1)


def attn_fwd(q, k, v, scale=None, out=None, need_lse=False):
    _log("attn_fwd")
    for t in (q, k, v):
        assert t.dtype == bf16 and t.dim() == 3 and t.shape[2] == 128 and t.stride(2) == 1
        assert t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0 and _al16(t), (t.stride(), t.data_ptr() % 16)     # csrc/attention_fwd.cu:668-671
    assert k.shape == v.shape and k.shape[1] == q.shape[1] and k.shape[0] > 0 and q.shape[0] > 0 and q.shape[1] > 0
    scale = 1.0 / math.sqrt(128) if scale is None else scale
    o, lse = _attn(q, k, v, scale)
    if out is None:
        out = torch.empty(q.shape, dtype=bf16)
    out.copy_(o.bfloat16())
    return (out, lse.contiguous()) if need_lse else out


def attn_bwd(q, k, v, o, dout, lse, dq=None, dk=None, dv=None, scale=None):
    _log("attn_bwd")
    scale = 1.0 / math.sqrt(128) if scale is None else scale
    assert lse.shape == (q.shape[1], q.shape[0]) and lse.is_contiguous()
    for t in (q, k, v, o, dout, dq, dk, dv):                         # csrc/attention_bwd.cu:412-418
        assert t is None or (t.dtype == bf16 and t.stride(2) == 1 and t.stride(0) % 8 == 0 and t.stride(1) % 8 == 0 and _al16(t)), t.stride()
    assert q.shape[0] > 0 and k.shape[0] > 0
    with torch.enable_grad():
        qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
        ref, _ = _attn(qf, kf, vf, scale)
        gq, gk, gv = torch.autograd.grad(ref, (qf, kf, vf), dout.float())
    dq = torch.empty(q.shape, dtype=bf16) if dq is None else dq
    dk = torch.empty(k.shape, dtype=bf16) if dk is None else dk
    dv = torch.empty(v.shape, dtype=bf16) if dv is None else dv
    dq.copy_(gq.bfloat16())
    dk.copy_(gk.bfloat16())
    dv.copy_(gv.bfloat16())
    return dq, dk, dv


def attn_merge_(o_acc, lse_acc, o_new, lse_new, first, out=None):
    _log("attn_merge_")
    if first:
        o_acc.copy_(o_new.float())
        lse_acc.copy_(lse_new)
    else:
        m = torch.maximum(lse_acc, lse_new)
        wa, wn = torch.exp(lse_acc - m), torch.exp(lse_new - m)
        tot = wa + wn
        o_acc.copy_(o_acc * (wa / tot).t()[:, :, None] + o_new.float() * (wn / tot).t()[:, :, None])
        lse_acc.copy_(m + torch.log(tot))
    if out is not None:
        out.copy_(o_acc.bfloat16())
    return o_acc


def ln_mod_bwd(x, dy, scale, gamma, mean, rstd, dx_accum, need_param_grads):
    _log("ln_mod_bwd")
    assert x.dtype == f32 and dy.dtype == bf16 and dx_accum.dtype == f32
    xh = (x - mean[:, None]) * rstd[:, None]
    dyf = dy.float()
    g = dyf * (1 + scale) if scale is not None else (dyf * gamma if gamma is not None else dyf)
    dx_accum.add_(rstd[:, None] * (g - g.mean(-1, keepdim=True) - xh * (g * xh).mean(-1, keepdim=True)))
    if need_param_grads:
        return dyf.sum(0), (dyf * xh).sum(0)
    return None, None


def colsum(a):
    _log("colsum")
    assert a.dtype == bf16 and a.dim() == 2
    return a.float().sum(0)


def gate_bwd(dx, y, gate):
    _log("gate_bwd")
    assert dx.dtype == f32 and dx.dim() == 2
    dy = (dx * gate if gate is not None else dx).bfloat16()
    return dy, ((dx * y.float()).sum(0) if y is not None else None)


def cast_bf16(src):
    _log("cast_bf16")
    assert src.dtype == f32
    return src.bfloat16()


def patchify(x, y=None):
    _log("patchify")
    u = x if y is None else torch.cat([x, y])
    Ct, Fr, H, W = u.shape
    return u.view(Ct, Fr, 1, H // 2, 2, W // 2, 2).permute(1, 3, 5, 0, 2, 4, 6).reshape(Fr * (H // 2) * (W // 2), -1).bfloat16()


def patchify_bwd(dp, Cx, Fr, H, W):
    _log("patchify_bwd")
    Ct = dp.shape[1] // 4
    return dp.view(Fr, H // 2, W // 2, Ct, 1, 2, 2).permute(3, 0, 4, 1, 5, 2, 6).reshape(Ct, Fr, H, W)[:Cx].contiguous()


def unpatchify(tokens, c, grid):
    _log("unpatchify")
    Fr, h, w = grid
    return tokens[:Fr * h * w].view(Fr, h, w, 1, 2, 2, c).permute(6, 0, 3, 1, 4, 2, 5).reshape(c, Fr, 2 * h, 2 * w).contiguous()


def unpatchify_bwd(dvid, rows):
    _log("unpatchify_bwd")
    c, Fr, H2, W2 = dvid.shape
    h, w = H2 // 2, W2 // 2
    tok = torch.zeros(rows, 4 * c, dtype=f32)
    tok[:Fr * h * w] = dvid.view(c, Fr, 1, h, 2, w, 2).permute(1, 3, 5, 2, 4, 6, 0).reshape(Fr * h * w, 4 * c)
    return tok


def sq_pool(x, wk_eff):
    _log("sq_pool")
    scores = x @ wk_eff.t()
    mx = scores.max(0).values
    e = torch.exp(scores - mx)
    sm = e.sum(0)
    return (e / sm).t() @ x, scores, torch.cat([mx, sm])


def sq_pool_bwd(x, wk_eff, scores, stats, pooled, dpooled, dx=None, need_ds=False):
    _log("sq_pool_bwd")
    NH = wk_eff.shape[0]
    p = torch.exp(scores - stats[:NH]) / stats[NH:]
    ds = p * (x @ dpooled.t() - (dpooled * pooled).sum(1))
    new = p @ dpooled + ds @ wk_eff
    if dx is None:
        dx = new
    else:
        dx.add_(new)
    return dx, (ds if need_ds else None)


def unipc_step(sample, model_output, last_sample, hist, sigma, corr_coef, pred_coef, model_output_uncond=None, guide_scale=1.0):
    _log("unipc_step")
    v = model_output if model_output_uncond is None else model_output_uncond + guide_scale * (model_output - model_output_uncond)
    hs = [h for h in hist[:3]] + [None] * 3
    x0 = sample - sigma * v

    def comb(c, first):
        acc = c[0] * first + c[1] * x0
        for k in range(3):
            if hs[k] is not None:
                acc = acc + c[2 + k] * hs[k]
            else:
                assert c[2 + k] == 0.0, "a missing history tensor must carry a zero coefficient"
        return acc
    corrected = comb(corr_coef, last_sample) if corr_coef is not None else None
    prev = comb(pred_coef, corrected if corrected is not None else sample)
    return x0, corrected, prev


def scale2(g, a, b=None):
    _log("scale2")
    return a * g, (b * g if b is not None else None)


EMULATED = ("ln_mod", "ln_mod_split", "rmsnorm_rope_", "rmsnorm_rope_bwd_", "gemm", "attn_fwd", "attn_bwd", "attn_merge_", "ln_mod_bwd",
            "colsum", "gate_bwd", "cast_bf16", "patchify", "patchify_bwd", "unpatchify", "unpatchify_bwd", "sq_pool", "sq_pool_bwd",
            "unipc_step", "scale2")


def install(monkeypatch):
    """Point `prfl_b200.ops.<name>` at the emulation for the duration of a test; everything else in ops (the peer-memory
    exchanges, the fused optimizer) keeps raising on CPU tensors."""
    from prfl_b200 import ops
    g = globals()
    for n in EMULATED:
        monkeypatch.setattr(ops, n, g[n])
    del CALLS[:]
    return ops

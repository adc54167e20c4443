"""GPU: each CUDA kernel (called through the C ABI) against a plain PyTorch fp32 reference of the same op
on the same device.  Tolerances are bf16-level and written next to each check."""
import math

import pytest
import torch

from conftest import cos_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from prfl_b200 import ops as o
    return o


def _rand(*shape, dtype=torch.float32, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("C,rows", [(256, 7), (1536, 130), (5120, 257)])
@pytest.mark.parametrize("variant", ["plain", "mod", "affine", "mod_round"])
def test_ln_mod(ops, C, rows, variant):
    x = _rand(rows, C, seed=1) * 3 + 0.5
    shift = scale = gamma = beta = None
    if variant in ("mod", "mod_round"):
        shift, scale = _rand(C, seed=2), _rand(C, seed=3) * 0.3
    if variant == "affine":
        gamma, beta = 1 + 0.1 * _rand(C, seed=4), 0.1 * _rand(C, seed=5)
    out, mean, rstd = ops.ln_mod(x, shift, scale, gamma, beta, 1e-6, round_bf16=(variant == "mod_round"), save_stats=True)
    ref = torch.nn.functional.layer_norm(x, (C,), gamma, beta, 1e-6)
    if variant == "mod_round":
        ref = ref.bfloat16().float()
    if shift is not None:
        ref = ref * (1 + scale) + shift
    assert out.dtype == torch.bfloat16
    # bf16 output rounding: rel 2^-8 of the value
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(mean, x.mean(-1), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(rstd, torch.rsqrt(x.var(-1, unbiased=False) + 1e-6), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("C,rows", [(256, 9), (1536, 77), (5120, 200)])
@pytest.mark.parametrize("rope", [False, True])
def test_rmsnorm_rope(ops, C, rows, rope):
    from prfl_b200.rope import rope_tables
    H = C // 128
    qkv = _rand(rows, 3 * C, dtype=torch.bfloat16, seed=6)
    x = qkv[:, C:2 * C]                      # strided view, like k inside fused QKV
    w = 1 + 0.1 * _rand(C, seed=7)
    n_rot = rows - 3 if rope else 0
    cos = sin = None
    ref = x.float() * torch.rsqrt(x.float().pow(2).mean(-1, keepdim=True) + 1e-6)
    ref = ref.bfloat16().float() * w
    if rope:
        cos, sin = rope_tables((2, 5, (n_rot + 9) // 10 + 1), torch.device("cuda"))
        pos0 = 4
        c, s = cos[pos0:pos0 + n_rot, None, :].double(), sin[pos0:pos0 + n_rot, None, :].double()
        r = ref[:n_rot].double().view(n_rot, H, 64, 2)
        rot = torch.stack([r[..., 0] * c - r[..., 1] * s, r[..., 0] * s + r[..., 1] * c], -1).view(n_rot, C)
        ref = torch.cat([rot.float(), ref[n_rot:]])
    else:
        pos0 = 0
    expected_rest = qkv.clone()
    out, rstd = ops.rmsnorm_rope_(x, w, cos, sin, 1e-6, n_rot, pos0, save_rstd=True)
    torch.testing.assert_close(out.float(), ref, rtol=1.6e-2, atol=1.6e-2)
    # the neighbouring q / v slices of the fused buffer are untouched
    assert torch.equal(qkv[:, :C], expected_rest[:, :C]) and torch.equal(qkv[:, 2 * C:], expected_rest[:, 2 * C:])


# ---------------------------------------------------------------------------------------------
GEMM_SHAPES = [(128, 256, 64), (300, 512, 256), (1000, 768, 144), (257, 264, 1536), (4096, 5120, 5120)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("a_t,b_t", [(False, False), (False, True), (True, False), (True, True)])
def test_gemm_layouts(ops, M, N, K, a_t, b_t):
    if a_t and M % 8:
        M = (M // 8) * 8
    a = _rand(M, K, dtype=torch.bfloat16, seed=10, scale=0.5)
    b = _rand(N, K, dtype=torch.bfloat16, seed=11, scale=0.5)
    bias = _rand(N, seed=12)
    ref = a.float() @ b.float().t() + bias
    A = a.t().contiguous() if a_t else a
    B = b.t().contiguous() if b_t else b
    out = ops.gemm(A, B, a_trans=a_t, b_trans=b_t, bias=bias, epi=ops.EPI_F32)
    cos, rel = cos_rel(out, ref)
    assert cos > 0.99999 and rel < 2e-3, (cos, rel)       # fp32 accumulate of bf16 products, K <= 5120


def test_gemm_epilogues(ops):
    M, N, K = 515, 1024, 512
    a = _rand(M, K, dtype=torch.bfloat16, seed=20, scale=0.5)
    b = _rand(N, K, dtype=torch.bfloat16, seed=21, scale=0.1)
    bias = _rand(N, seed=22)
    acc = a.float() @ b.float().t() + bias
    out = ops.gemm(a, b, bias=bias, epi=ops.EPI_BF16)
    torch.testing.assert_close(out.float(), acc, rtol=1e-2, atol=1e-2)
    out = ops.gemm(a, b, bias=bias, epi=ops.EPI_BF16_GELU)
    ref = torch.nn.functional.gelu(acc.bfloat16().float(), approximate="tanh")
    torch.testing.assert_close(out.float(), ref, rtol=1.6e-2, atol=1.6e-2)
    gate = _rand(N, seed=23)
    x0 = _rand(M, N, seed=24)
    x = x0.clone()
    ops.gemm(a, b, bias=bias, epi=ops.EPI_RESIDUAL, out=x, gate=gate)
    torch.testing.assert_close(x, x0 + gate * acc.bfloat16().float(), rtol=1e-2, atol=2e-2)
    x = x0.clone()
    ops.gemm(a, b, bias=None, epi=ops.EPI_RESIDUAL, out=x)
    torch.testing.assert_close(x, x0 + (acc - bias).bfloat16().float(), rtol=1e-2, atol=2e-2)
    x = x0.clone()
    ops.gemm(a, b, bias=None, epi=ops.EPI_F32, out=x, beta=True)
    torch.testing.assert_close(x, x0 + acc - bias, rtol=1e-3, atol=1e-2)
    aux = _rand(M, N, dtype=torch.bfloat16, seed=25)
    out = ops.gemm(a, b, epi=ops.EPI_BF16_DGELU, aux=aux)
    xa = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(xa, approximate="tanh").backward(acc - bias)
    torch.testing.assert_close(out.float(), xa.grad, rtol=2e-2, atol=2e-2)
    # strided output / operand views (q slice of fused QKV, etc.)
    big = torch.zeros(M, 3 * N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, b, bias=bias, epi=ops.EPI_BF16, out=big[:, N:2 * N])
    torch.testing.assert_close(big[:, N:2 * N].float(), acc, rtol=1e-2, atol=1e-2)
    assert float(big[:, :N].abs().max()) == 0 and float(big[:, 2 * N:].abs().max()) == 0


# ---------------------------------------------------------------------------------------------
def _attn_ref(q, k, v, scale):
    qf, kf, vf = q.float().transpose(0, 1), k.float().transpose(0, 1), v.float().transpose(0, 1)
    s = (qf @ kf.transpose(1, 2)) * scale
    lse = torch.logsumexp(s, -1)
    return (torch.softmax(s, -1) @ vf).transpose(0, 1), lse


@pytest.mark.parametrize("Lq,Lk,H", [(128, 128, 1), (256, 256, 2), (300, 300, 2), (1950, 1950, 3), (300, 512, 2),
                                      (100, 769, 2), (4095, 4095, 2), (1000, 50, 1)])
def test_attn_fwd(ops, Lq, Lk, H):
    q = _rand(Lq, H, 128, dtype=torch.bfloat16, seed=30)
    k = _rand(Lk, H, 128, dtype=torch.bfloat16, seed=31)
    v = _rand(Lk, H, 128, dtype=torch.bfloat16, seed=32)
    scale = 1 / math.sqrt(128)
    out, lse = ops.attn_fwd(q, k, v, need_lse=True)
    ref, lse_ref = _attn_ref(q, k, v, scale)
    cos, rel = cos_rel(out, ref)
    assert cos > 0.9999 and rel < 2e-2, (cos, rel)          # bf16 P and bf16 output
    torch.testing.assert_close(lse, lse_ref, rtol=1e-3, atol=2e-3)


@pytest.mark.parametrize("Lq,Lk,H", [(4096, 4096, 10), (4000, 4100, 10), (2100, 8200, 17)])
def test_attn_fwd_key_split_tail(ops, Lq, Lk, H):
    """Shapes whose unit count leaves a small remainder modulo the SM count (160 / 153 units on 148 SMs): the tail units run
    split along the key axis + combine (csrc/attention_fwd.cu, "Wave quantisation").  Same tolerance as the plain kernel."""
    from prfl_b200 import _lib
    assert _lib.lib().prfl_attn_fwd_ws_bytes(Lq, Lk, H) > 0, "this shape is meant to take the split path"
    q = _rand(Lq, H, 128, dtype=torch.bfloat16, seed=40)
    k = _rand(Lk, H, 128, dtype=torch.bfloat16, seed=41)
    v = _rand(Lk, H, 128, dtype=torch.bfloat16, seed=42)
    out, lse = ops.attn_fwd(q, k, v, need_lse=True)
    ref, lse_ref = _attn_ref(q, k, v, 1 / math.sqrt(128))
    for h0 in range(0, H, 4):                               # tail units are the LAST heads: check every head group
        cos, rel = cos_rel(out[:, h0:h0 + 4], ref[:, h0:h0 + 4])
        assert cos > 0.9999 and rel < 2e-2, (h0, cos, rel)
    torch.testing.assert_close(lse, lse_ref, rtol=1e-3, atol=2e-3)
    # gradients through the split forward's LSE / O
    do = _rand(Lq, H, 128, dtype=torch.bfloat16, seed=43)
    dq, dk, dv = ops.attn_bwd(q, k, v, out, do, lse)
    assert torch.isfinite(dq.float()).all() and torch.isfinite(dk.float()).all() and torch.isfinite(dv.float()).all()


@pytest.mark.parametrize("Lq,blocks,H", [(300, (200, 137, 64), 2), (1000, (512, 512), 3), (96, (96,), 2)])
def test_attn_merge_ring_blocks(ops, Lq, blocks, H):
    """prfl_attn_merge (ring attention, xdit_context_parallel.py:214-219 / xfuser): attention over key blocks folded in one at
    a time == attention over the concatenated keys; the last fold also emits bf16."""
    q = _rand(Lq, H, 128, dtype=torch.bfloat16, seed=50)
    ks = [_rand(n, H, 128, dtype=torch.bfloat16, seed=51 + i) for i, n in enumerate(blocks)]
    vs = [_rand(n, H, 128, dtype=torch.bfloat16, seed=61 + i) for i, n in enumerate(blocks)]
    acc = torch.empty(Lq, H, 128, dtype=torch.float32, device="cuda")
    lse = torch.empty(H, Lq, dtype=torch.float32, device="cuda")
    out = torch.empty(Lq, H, 128, dtype=torch.bfloat16, device="cuda")
    for i, (k, v) in enumerate(zip(ks, vs)):
        o_i, lse_i = ops.attn_fwd(q, k, v, need_lse=True)
        ops.attn_merge_(acc, lse, o_i, lse_i, i == 0, out if i == len(blocks) - 1 else None)
    ref, lse_ref = _attn_ref(q, torch.cat(ks), torch.cat(vs), 1 / math.sqrt(128))
    cos, rel = cos_rel(out, ref)
    assert cos > 0.9999 and rel < 2e-2, (cos, rel)
    assert torch.equal(out, acc.bfloat16())
    assert float((lse - lse_ref).abs().max()) < 2e-3


def test_attn_fwd_strided_and_peaked(ops):
    """q/k/v as slices of one fused [L, 3, H, 128] buffer; large-magnitude scores exercise the lazy rescale."""
    L, H = 700, 2
    qkv = _rand(L, 3, H, 128, dtype=torch.bfloat16, seed=33, scale=4.0)
    q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
    out = ops.attn_fwd(q, k, v)
    ref, _ = _attn_ref(q, k, v, 1 / math.sqrt(128))
    cos, rel = cos_rel(out, ref)
    assert cos > 0.9995 and rel < 3e-2, (cos, rel)


# ---------------------------------------------------------------------------------------------
def test_patchify_unpatchify_cast(ops):
    x, y = _rand(16, 3, 10, 14, seed=40), _rand(20, 3, 10, 14, seed=41)
    p = ops.patchify(x, y)
    u = torch.cat([x, y])
    ref = u.view(36, 3, 1, 5, 2, 7, 2).permute(1, 3, 5, 0, 2, 4, 6).reshape(3 * 5 * 7, -1)
    assert torch.equal(p, ref.bfloat16())
    p1 = ops.patchify(x)
    assert torch.equal(p1, ref[:, :64].bfloat16())
    dp = _rand(105, 144, seed=42)
    dx = ops.patchify_bwd(dp, 16, 3, 10, 14)
    ref_dx = dp.view(3, 5, 7, 36, 1, 2, 2).permute(3, 0, 4, 1, 5, 2, 6).reshape(36, 3, 10, 14)[:16]
    assert torch.equal(dx, ref_dx)
    tok = _rand(105, 64, seed=43)
    vid = ops.unpatchify(tok, 16, (3, 5, 7))
    ref_v = tok.view(3, 5, 7, 1, 2, 2, 16).permute(6, 0, 3, 1, 4, 2, 5).reshape(16, 3, 10, 14)
    assert torch.equal(vid, ref_v)
    assert torch.equal(ops.unpatchify_bwd(vid, 105), tok)
    w = _rand(1000003, seed=44)
    assert torch.equal(ops.cast_bf16(w), w.bfloat16())


def test_a2a_pack(ops):
    L_loc, H, P = 37, 8, 4
    x = _rand(L_loc, 3, H, 128, dtype=torch.bfloat16, seed=50)[:, 1]          # strided
    packed = torch.empty(P, L_loc, H // P, 128, dtype=torch.bfloat16, device="cuda")
    ops.a2a_pack(x, packed, P)
    ref = x.reshape(L_loc, P, H // P, 128).permute(1, 0, 2, 3)
    assert torch.equal(packed, ref)
    back = torch.zeros(L_loc, H, 128, dtype=torch.bfloat16, device="cuda")
    ops.a2a_pack(back, packed, P, unpack=True)
    assert torch.equal(back, x)


@pytest.mark.parametrize("L,C", [(300, 256), (1950, 1536), (4000, 5120)])
def test_sq_pool(ops, L, C):
    x = _rand(L, C, seed=60)
    wk = _rand(8, C, seed=61) * 0.05
    pooled, scores, stats = ops.sq_pool(x, wk)
    s_ref = x @ wk.t()
    torch.testing.assert_close(scores, s_ref, rtol=1e-4, atol=1e-4)
    p_ref = torch.softmax(s_ref, 0)
    torch.testing.assert_close(pooled, p_ref.t() @ x, rtol=1e-3, atol=1e-4)
    dpooled = _rand(8, C, seed=62)
    xr = x.clone().requires_grad_(True)
    wr = wk.clone().requires_grad_(True)
    (torch.softmax(xr @ wr.t(), 0).t() @ xr * dpooled).sum().backward()
    dx, ds = ops.sq_pool_bwd(x, wk, scores, stats, pooled, dpooled, need_ds=True)
    cos, rel = cos_rel(dx, xr.grad)
    assert cos > 0.999999 and rel < 1e-4, (cos, rel)
    cos, rel = cos_rel(ds.t() @ x, wr.grad)
    assert cos > 0.99999 and rel < 1e-3, (cos, rel)


# ---------------------------------------------------------------------------------------------
# backward kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Lq,Lk,H", [(128, 128, 1), (256, 192, 2), (300, 300, 2), (1950, 1950, 2), (300, 512, 2), (700, 769, 1)])
def test_attn_bwd(ops, Lq, Lk, H):
    q = _rand(Lq, H, 128, dtype=torch.bfloat16, seed=70)
    k = _rand(Lk, H, 128, dtype=torch.bfloat16, seed=71)
    v = _rand(Lk, H, 128, dtype=torch.bfloat16, seed=72)
    do = _rand(Lq, H, 128, dtype=torch.bfloat16, seed=73)
    o, lse = ops.attn_fwd(q, k, v, need_lse=True)
    dq, dk, dv = ops.attn_bwd(q, k, v, o, do, lse)
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref, _ = _attn_ref(qf, kf, vf, 1 / math.sqrt(128))
    ref.backward(do.float())
    for name, a, b in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        cos, rel = cos_rel(a, b)
        assert cos > 0.9995 and rel < 2e-2, (name, cos, rel)     # bf16 P / dS operands, bf16 outputs


@pytest.mark.parametrize("C,rows", [(256, 300), (1536, 515), (5120, 300)])
@pytest.mark.parametrize("variant", ["mod", "affine", "plain"])
def test_ln_mod_bwd(ops, C, rows, variant):
    x = _rand(rows, C, seed=80) * 2 + 0.3
    dy = _rand(rows, C, dtype=torch.bfloat16, seed=81)
    shift = scale = gamma = beta = None
    if variant == "mod":
        shift, scale = _rand(C, seed=82), _rand(C, seed=83) * 0.3
    if variant == "affine":
        gamma, beta = 1 + 0.1 * _rand(C, seed=84), 0.1 * _rand(C, seed=85)
    _, mean, rstd = ops.ln_mod(x, shift, scale, gamma, beta, 1e-6, save_stats=True)
    dx0 = _rand(rows, C, seed=86)
    dx = dx0.clone()
    g1, g2 = ops.ln_mod_bwd(x, dy, scale, gamma, mean, rstd, dx, need_param_grads=(variant != "plain"))
    xr = x.clone().requires_grad_(True)
    leaves = [t.clone().requires_grad_(True) if t is not None else None for t in (shift, scale, gamma, beta)]
    y = torch.nn.functional.layer_norm(xr, (C,), leaves[2], leaves[3], 1e-6)
    if variant == "mod":
        y = y * (1 + leaves[1]) + leaves[0]
    y.backward(dy.float())
    torch.testing.assert_close(dx - dx0, xr.grad, rtol=2e-3, atol=2e-3)
    if variant == "mod":
        torch.testing.assert_close(g1, leaves[0].grad, rtol=2e-3, atol=5e-2)
        torch.testing.assert_close(g2, leaves[1].grad, rtol=2e-3, atol=5e-2)
    if variant == "affine":
        torch.testing.assert_close(g1, leaves[3].grad, rtol=2e-3, atol=5e-2)
        torch.testing.assert_close(g2, leaves[2].grad, rtol=2e-3, atol=5e-2)


@pytest.mark.parametrize("C,rows", [(256, 40), (1536, 300), (5120, 130)])
@pytest.mark.parametrize("rope", [False, True])
def test_rmsnorm_rope_bwd(ops, C, rows, rope):
    from prfl_b200.rope import rope_tables
    H = C // 128
    x = _rand(rows, C, dtype=torch.bfloat16, seed=90)
    w = 1 + 0.1 * _rand(C, seed=91)
    dy = _rand(rows, C, dtype=torch.bfloat16, seed=92)
    n_rot, pos0 = (rows - 2, 3) if rope else (0, 0)
    cos = sin = None
    if rope:
        cos, sin = rope_tables((2, 5, (rows + 9) // 10 + 1), torch.device("cuda"))
    _, rstd = ops.rmsnorm_rope_(x.clone(), w, cos, sin, 1e-6, n_rot, pos0, save_rstd=True)
    xr = x.float().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    t = xr * torch.rsqrt(xr.pow(2).mean(-1, keepdim=True) + 1e-6) * wr
    if rope:
        c, s = cos[pos0:pos0 + n_rot, None, :], sin[pos0:pos0 + n_rot, None, :]
        r = t[:n_rot].view(n_rot, H, 64, 2)
        rot = torch.stack([r[..., 0] * c - r[..., 1] * s, r[..., 0] * s + r[..., 1] * c], -1).view(n_rot, C)
        t = torch.cat([rot, t[n_rot:]])
    t.backward(dy.float())
    g = dy.clone()
    dw = ops.rmsnorm_rope_bwd_(x, w, cos, sin, g, rstd, n_rot, pos0)
    cosv, rel = cos_rel(g, xr.grad)
    assert cosv > 0.9999 and rel < 1.5e-2, (cosv, rel)
    cosv, rel = cos_rel(dw, wr.grad)
    assert cosv > 0.9999 and rel < 1.5e-2, (cosv, rel)


def test_colsum_gate_bwd(ops):
    rows, N = 1000, 1536
    a = _rand(rows, 2 * N, dtype=torch.bfloat16, seed=95)[:, N:]
    torch.testing.assert_close(ops.colsum(a), a.float().sum(0), rtol=1e-4, atol=1e-3)
    dx, y, gate = _rand(rows, N, seed=96), _rand(rows, N, dtype=torch.bfloat16, seed=97), _rand(N, seed=98)
    dy, dg = ops.gate_bwd(dx, y, gate)
    torch.testing.assert_close(dy.float(), dx * gate, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(dg, (dx * y.float()).sum(0), rtol=1e-4, atol=1e-3)
    dy, dg = ops.gate_bwd(dx, None, None)
    assert dg is None and torch.equal(dy, dx.bfloat16())


def test_gemm_aux_outputs(ops):
    M, N, K = 300, 512, 256
    a = _rand(M, K, dtype=torch.bfloat16, seed=100, scale=0.5)
    b = _rand(N, K, dtype=torch.bfloat16, seed=101, scale=0.2)
    bias = _rand(N, seed=102)
    acc = a.float() @ b.float().t() + bias
    pre = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    out = ops.gemm(a, b, bias=bias, epi=ops.EPI_BF16_GELU, aux=pre)
    torch.testing.assert_close(pre.float(), acc, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(out.float(), torch.nn.functional.gelu(pre.float(), approximate="tanh"), rtol=1.6e-2, atol=1.6e-2)
    x0 = _rand(M, N, seed=103)
    x = x0.clone()
    y = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    gate = _rand(N, seed=104)
    ops.gemm(a, b, bias=bias, epi=ops.EPI_RESIDUAL, out=x, gate=gate, aux=y)
    torch.testing.assert_close(y.float(), acc, rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(x, x0 + gate * y.float(), rtol=1e-4, atol=1e-4)


def test_head_split_gemm_is_fp32_grade(ops):
    """LayerNorm+modulate emitted as a bf16 (hi, lo) pair + three tensor-core GEMMs == fp32 Linear to ~1e-5."""
    M, C, N = 777, 1536, 64
    x = _rand(M, C, seed=110) * 2
    sh, sc = _rand(C, seed=111), _rand(C, seed=112) * 0.3
    w = _rand(N, C, seed=113) * 0.05
    bias = _rand(N, seed=114)
    hi, lo = ops.ln_mod_split(x, sh, sc)
    ref_h = torch.nn.functional.layer_norm(x, (C,), None, None, 1e-6) * (1 + sc) + sh
    assert float((hi.float() + lo.float() - ref_h).abs().max()) < 2e-4
    w_hi = w.bfloat16()
    w_lo = (w - w_hi.float()).bfloat16()
    o = ops.gemm(hi, w_hi, bias=bias, epi=ops.EPI_F32)
    ops.gemm(hi, w_lo, epi=ops.EPI_F32, out=o, beta=True)
    ops.gemm(lo, w_hi, epi=ops.EPI_F32, out=o, beta=True)
    ref = (ref_h.double() @ w.double().t() + bias.double()).float()
    cos, rel = cos_rel(o, ref)
    assert cos > 0.9999999 and rel < 5e-5, (cos, rel)


@pytest.mark.parametrize("n", [1024, 1_000_003])
def test_adamw_step_and_sumsq(ops, n):
    """prfl_adamw_step against torch.optim.AdamW (same hyper-parameters, 3 steps, with a clip coefficient on the device)
    and prfl_sumsq_f32 against a float64 torch reduction.  fp32 elementwise: 1e-6 relative."""
    g = torch.Generator(device="cuda").manual_seed(n)
    w0 = torch.randn(n, generator=g, device="cuda")
    ref_p = torch.nn.Parameter(w0.clone())
    opt = torch.optim.AdamW([ref_p], lr=1e-2, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    w, m, v = w0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for step in range(1, 4):
        grad = torch.randn(n, generator=g, device="cuda") * 3.0
        acc = torch.zeros((), dtype=torch.float64, device="cuda")
        ops.sumsq_(grad, acc)
        want = grad.double().pow(2).sum()
        assert abs(float(acc) - float(want)) <= 1e-6 * float(want)       # fp32 per-thread partials, double across CTAs
        coef = torch.clamp(1.0 / (acc.sqrt().float() + 1e-6), max=1.0)
        ref_p.grad = grad * coef
        opt.step()
        ops.adamw_step_(grad, w, m, v, step, 1e-2, (0.9, 0.95), 1e-8, 0.05, coef)
        torch.testing.assert_close(w, ref_p.detach(), rtol=2e-6, atol=2e-7)
        st = opt.state[ref_p]
        torch.testing.assert_close(m, st["exp_avg"], rtol=2e-6, atol=1e-9)
        torch.testing.assert_close(v, st["exp_avg_sq"], rtol=2e-6, atol=1e-12)

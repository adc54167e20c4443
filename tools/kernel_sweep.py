"""BASELINE.json configs[4]: kernel sweep on one B200.
  * self-attention fwd and bwd, B=1, H heads, hd=128, L in {8192, 16384, 32760, 65536, 75600} (bf16, randn seed 0)
  * the memory-bound kernels at [L, 5120] against the measured HBM copy bandwidth
  * the GEMM shapes of a 14B block (fwd, dgrad, wgrad)
Library kernels the reference would use on the same box (flash-attn 2 `flash_attn_func`, cuBLAS via torch.matmul)
are timed beside ours when importable — they are the bar to beat, not part of the product.
Timing: CUDA events on the launching stream, 3 warm-up + N timed launches back to back (sustained clocks); inputs
are far larger than the 126 MB L2.   Usage: python tools/kernel_sweep.py [--heads 40] [--quick] > profiles/rNN_kernel_sweep.json
"""
import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from prfl_b200 import ops  # noqa: E402


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--heads", type=int, default=40)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--sections", default="attention,membound,gemm")
    args = ap.parse_args()
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak, hbm_peak = pk.get("bf16_tflops_sustained", 1400.0), pk.get("hbm_gbs", 6650.0)
    res = {"peaks": {"bf16_tflops_sustained": tf_peak, "hbm_gbs": hbm_peak, "source": "MEASURED_PEAKS.json" if pk else "fallback"},
           "attention": [], "membound": [], "gemm": []}
    try:
        from flash_attn import flash_attn_func
    except Exception:
        flash_attn_func = None
    H = args.heads
    g = torch.Generator(device="cuda").manual_seed(0)
    Ls = [8192, 32760] if args.quick else [8192, 16384, 32760, 65536, 75600]
    sections = args.sections.split(",")
    if "attention" not in sections:
        Ls = []
    for L in Ls:
        q, k, v, do = (torch.randn(L, H, 128, generator=g, device="cuda").bfloat16() for _ in range(4))
        iters = max(2, int(3e9 / (L * L)))
        fl_f = 4.0 * L * L * 128 * H
        t_f = timeit(lambda: ops.attn_fwd(q, k, v), iters)
        o, lse = ops.attn_fwd(q, k, v, need_lse=True)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        t_b = timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse, dq=dq, dk=dk, dv=dv), iters)
        row = {"L": L, "heads": H, "fwd_ms": t_f, "fwd_tflops": fl_f / t_f / 1e9, "fwd_frac_of_peak": fl_f / t_f / 1e9 / tf_peak,
               "bwd_ms": t_b, "bwd_tflops_algorithmic": 2.5 * fl_f / t_b / 1e9, "bwd_frac_of_peak": 2.5 * fl_f / t_b / 1e9 / tf_peak}
        if flash_attn_func is not None:
            q4, k4, v4 = (t[None].clone().requires_grad_(True) for t in (q, k, v))
            t_ff = timeit(lambda: flash_attn_func(q4, k4, v4), iters)
            of = flash_attn_func(q4, k4, v4)
            t_fb = timeit(lambda: torch.autograd.grad(of, (q4, k4, v4), do[None], retain_graph=True), iters)
            row.update({"fa2_fwd_ms": t_ff, "fa2_fwd_tflops": fl_f / t_ff / 1e9, "fa2_bwd_ms": t_fb,
                        "fa2_bwd_tflops_algorithmic": 2.5 * fl_f / t_fb / 1e9, "speedup_fwd_vs_fa2": t_ff / t_f,
                        "speedup_bwd_vs_fa2": t_fb / t_b})
            del q4, k4, v4, of
        res["attention"].append(row)
        print(json.dumps(row), file=sys.stderr)
        del q, k, v, do, o, lse, dq, dk, dv
        torch.cuda.empty_cache()

    # cross-attention shapes (short KV: 512 T5 tokens / 257 CLIP tokens, model.py:206-226, 244-271)
    if "cross" in sections:
        res["cross_attention"] = []
        for Lq in ([32760] if args.quick else [32760, 75600]):
            for Lk in (512, 257):
                q = torch.randn(Lq, H, 128, generator=g, device="cuda").bfloat16()
                k, v = (torch.randn(Lk, H, 128, generator=g, device="cuda").bfloat16() for _ in range(2))
                do = torch.randn(Lq, H, 128, generator=g, device="cuda").bfloat16()
                fl = 4.0 * Lq * Lk * 128 * H
                t_f = timeit(lambda: ops.attn_fwd(q, k, v), 20)
                o, lse = ops.attn_fwd(q, k, v, need_lse=True)
                t_b = timeit(lambda: ops.attn_bwd(q, k, v, o, do, lse), 10)
                row = {"Lq": Lq, "Lk": Lk, "heads": H, "fwd_ms": t_f, "fwd_tflops": fl / t_f / 1e9, "bwd_ms": t_b,
                       "bwd_tflops_algorithmic": 2.5 * fl / t_b / 1e9,
                       "io_bytes_fwd": 2.0 * (2 * Lq + 2 * Lk) * H * 128, "fwd_GBps": 2.0 * (2 * Lq + 2 * Lk) * H * 128 / t_f / 1e6}
                if flash_attn_func is not None:
                    t_ff = timeit(lambda: flash_attn_func(q[None], k[None], v[None]), 20)
                    row.update({"fa2_fwd_ms": t_ff, "speedup_fwd_vs_fa2": t_ff / t_f})
                res["cross_attention"].append(row)
                print(json.dumps(row), file=sys.stderr)
                del q, k, v, do, o, lse

    # memory-bound kernels at the 480P token count
    M, C = 32760, 5120
    if "membound" not in sections:
        print(json.dumps(res, indent=1))
        return
    x = torch.randn(M, C, generator=g, device="cuda")
    sh, sc = torch.randn(C, device="cuda"), torch.randn(C, device="cuda") * 0.1
    t = timeit(lambda: ops.ln_mod(x, sh, sc), 20)
    res["membound"].append({"kernel": "ln_mod_fwd", "bytes": 6 * C * M, "ms": t, "GBps": 6 * C * M / t / 1e6, "frac": 6 * C * M / t / 1e6 / hbm_peak})
    qkv = torch.randn(M, 3 * C, generator=g, device="cuda").bfloat16()
    w = torch.ones(C, device="cuda")
    from prfl_b200.rope import rope_tables
    cos, sin = rope_tables((21, 30, 52), torch.device("cuda"))
    t = timeit(lambda: ops.rmsnorm_rope_(qkv[:, :C], w, cos, sin, 1e-6, M, 0), 20)
    res["membound"].append({"kernel": "rmsnorm_rope_fwd", "bytes": 4 * C * M, "ms": t, "GBps": 4 * C * M / t / 1e6, "frac": 4 * C * M / t / 1e6 / hbm_peak})
    dy = torch.randn(M, C, generator=g, device="cuda").bfloat16()
    _, mean, rstd = ops.ln_mod(x, sh, sc, save_stats=True)
    dx = torch.zeros(M, C, device="cuda")
    t = timeit(lambda: ops.ln_mod_bwd(x, dy, sc, None, mean, rstd, dx, False), 20)
    res["membound"].append({"kernel": "ln_mod_bwd (dx accumulate)", "bytes": 14 * C * M, "ms": t, "GBps": 14 * C * M / t / 1e6, "frac": 14 * C * M / t / 1e6 / hbm_peak})
    qkv_g = torch.randn(M, 3 * C, generator=g, device="cuda").bfloat16()
    rstd_q = torch.rand(M, device="cuda") + 0.5
    t = timeit(lambda: ops.rmsnorm_rope_bwd_(qkv[:, :C], w, cos, sin, qkv_g[:, :C], rstd_q, M, 0, need_dw=False), 20)
    res["membound"].append({"kernel": "rmsnorm_rope_bwd (dx in place)", "bytes": 6 * C * M, "ms": t, "GBps": 6 * C * M / t / 1e6, "frac": 6 * C * M / t / 1e6 / hbm_peak})
    del qkv_g
    wk = torch.randn(8, C, device="cuda") * 0.01
    t = timeit(lambda: ops.sq_pool(x, wk), 20)
    res["membound"].append({"kernel": "sq_pool_fwd (2 passes)", "bytes": 8 * C * M, "ms": t, "GBps": 8 * C * M / t / 1e6, "frac": 8 * C * M / t / 1e6 / hbm_peak})
    t = timeit(lambda: ops.cast_bf16(x), 20)
    res["membound"].append({"kernel": "cast_f32_bf16", "bytes": 6 * C * M, "ms": t, "GBps": 6 * C * M / t / 1e6, "frac": 6 * C * M / t / 1e6 / hbm_peak})
    del qkv, dx, dy
    for r in res["membound"]:
        print(json.dumps(r), file=sys.stderr)
    if "gemm" not in sections:
        print(json.dumps(res, indent=1))
        return

    # GEMMs of one 14B block at M = 32760
    def gemm_row(name, M_, N_, K_, a_t, b_t, epi):
        a = torch.randn((K_, M_) if a_t else (M_, K_), generator=g, device="cuda").bfloat16()
        b = torch.randn((K_, N_) if b_t else (N_, K_), generator=g, device="cuda").bfloat16() * 0.02
        out = torch.zeros(M_, N_, device="cuda", dtype=torch.float32 if epi in (ops.EPI_F32, ops.EPI_RESIDUAL) else torch.bfloat16)
        bias = torch.zeros(N_, device="cuda")
        t_ = timeit(lambda: ops.gemm(a, b, a_trans=a_t, b_trans=b_t, bias=bias, epi=epi, out=out), 10)
        fl = 2.0 * M_ * N_ * K_
        am, bm = (a.t() if a_t else a), (b if b_t else b.t())
        t_c = timeit(lambda: torch.matmul(am, bm), 10)
        r = {"gemm": name, "M": M_, "N": N_, "K": K_, "a_trans": a_t, "b_trans": b_t, "ms": t_, "tflops": fl / t_ / 1e9,
             "frac_of_peak": fl / t_ / 1e9 / tf_peak, "cublas_ms": t_c, "cublas_tflops": fl / t_c / 1e9}
        res["gemm"].append(r)
        print(json.dumps(r), file=sys.stderr)

    gemm_row("qkv fwd (bias)", M, 3 * C, C, False, False, ops.EPI_BF16)
    gemm_row("o fwd (gated residual)", M, C, C, False, False, ops.EPI_RESIDUAL)
    gemm_row("ffn.0 fwd (GELU)", M, 13824, C, False, False, ops.EPI_BF16_GELU)
    gemm_row("ffn.2 fwd (gated residual)", M, C, 13824, False, False, ops.EPI_RESIDUAL)
    if not args.quick:
        gemm_row("ffn.2 dgrad", M, 13824, C, False, True, ops.EPI_BF16)
        gemm_row("ffn.0 wgrad", 13824, C, M, True, True, ops.EPI_F32)
        gemm_row("qkv wgrad", 3 * C, C, M, True, True, ops.EPI_F32)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

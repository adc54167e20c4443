"""GEMM shapes of a 14B block at M tokens, timed with CUDA events (and, when run under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_bf16`, profiled for DRAM
traffic).  The tile raster is chosen per process with PRFL_GEMM_GROUP_M (> 0: row-tile groups walked m-fastest; < 0:
column-tile groups walked n-fastest): run once per setting.  `--cublas` times torch.matmul (cuBLAS) on the same shapes."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prfl_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=32760)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--cublas", action="store_true")
ap.add_argument("--once", action="store_true", help="launch every shape exactly once (for ncu)")
a = ap.parse_args()
M, C, F = a.M, 5120, 13824
g = torch.Generator(device="cuda").manual_seed(0)
shapes = {"qkv [M,15360,5120] bf16": (3 * C, C, "bf16"), "o+resid [M,5120,5120] f32 rmw": (C, C, "resid"),
          "ffn0+gelu [M,13824,5120]": (F, C, "gelu"), "ffn2+resid [M,5120,13824]": (C, F, "resid"),
          "dgrad qkv [M,5120,15360] nn": (C, 3 * C, "nn"), "wgrad ffn0 [13824,5120,M] tt f32": (F, C, "wgrad")}
out = {"M": M, "group_m": os.environ.get("PRFL_GEMM_GROUP_M", "8 (default)"), "shapes": {}}
x = torch.zeros(M, C, device="cuda")
gate, bias = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
for name, (N, K, kind) in shapes.items():
    A = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    W = (torch.randn(N, K, generator=g, device="cuda") * 0.02).bfloat16()
    if kind == "bf16":
        fn = lambda: ops.gemm(A, W, epi=ops.EPI_BF16)
    elif kind == "resid":
        fn = lambda: ops.gemm(A, W, bias=bias, epi=ops.EPI_RESIDUAL, out=x, gate=gate)
    elif kind == "gelu":
        fn = lambda: ops.gemm(A, W, epi=ops.EPI_BF16_GELU)
    elif kind == "nn":          # dgrad: dx[M, N] = dy[M, K] W[K, N]  (W stored [K, N], b_trans)
        Wt = (torch.randn(K, N, generator=g, device="cuda") * 0.02).bfloat16()
        fn = lambda: ops.gemm(A, Wt, b_trans=True, epi=ops.EPI_BF16)
    else:                       # wgrad: dW[N, K] = dy[M, N]^T x[M, K]
        dY = torch.randn(M, N, generator=g, device="cuda").bfloat16()
        fn = lambda: ops.gemm(dY, A, a_trans=True, b_trans=True, epi=ops.EPI_F32)
    if a.cublas:
        Wc = W.t() if kind != "wgrad" else None
        fn = (lambda: torch.matmul(A, Wc)) if kind != "wgrad" else (lambda: torch.matmul(dY.t(), A))
    if a.once:
        fn()
        torch.cuda.synchronize()
        continue
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    out["shapes"][name] = {"ms": round(ms, 4), "tflops": round(2.0 * M * N * K / ms / 1e9, 1)}
    del A, W
print(json.dumps(out))

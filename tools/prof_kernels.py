"""Launch the hot kernels a few times each in isolation (for `ncu --set full -k regex:... python tools/prof_kernels.py`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prfl_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator(device="cuda").manual_seed(0)
if which in ("all", "attn"):
    L, H = 16384, 40
    q, k, v, do = (torch.randn(L, H, 128, generator=g, device="cuda").bfloat16() for _ in range(4))
    for _ in range(3):
        o, lse = ops.attn_fwd(q, k, v, need_lse=True)
    for _ in range(2):
        ops.attn_bwd(q, k, v, o, do, lse)
if which in ("all", "gemm"):
    M, C = 32760, 5120
    a = torch.randn(M, C, generator=g, device="cuda").bfloat16()
    w = torch.randn(C, C, generator=g, device="cuda").bfloat16() * 0.02
    w3 = torch.randn(3 * C, C, generator=g, device="cuda").bfloat16() * 0.02
    x = torch.zeros(M, C, device="cuda")
    gate = torch.ones(C, device="cuda")
    bias = torch.zeros(C, device="cuda")
    for _ in range(3):
        ops.gemm(a, w, bias=bias, epi=ops.EPI_RESIDUAL, out=x, gate=gate)
    for _ in range(3):
        ops.gemm(a, w3, epi=ops.EPI_BF16)
torch.cuda.synchronize()
print("done")

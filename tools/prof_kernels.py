"""Launch each hot kernel ONCE in isolation at its full-size shape, for
`ncu --set full --clock-control none --import-source on -k regex:attn_fwd|attn_bwd|gemm_bf16|ln_mod|rmsnorm_rope python tools/prof_kernels.py`
(self-attention fwd / bwd at L = 32 760 x 40 heads; the gated-residual and QKV GEMMs and the four row kernels at [32 760, 5 120])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prfl_b200 import ops  # noqa: E402
from prfl_b200.rope import rope_tables  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator(device="cuda").manual_seed(0)
M, C, H = 32760, 5120, 40
if which in ("all", "attn"):
    q, k, v, do = (torch.randn(M, H, 128, generator=g, device="cuda").bfloat16() for _ in range(4))
    o, lse = ops.attn_fwd(q, k, v, need_lse=True)
    ops.attn_bwd(q, k, v, o, do, lse)
    del q, k, v, do, o, lse
if which in ("all", "gemm"):
    a = torch.randn(M, C, generator=g, device="cuda").bfloat16()
    w = torch.randn(C, C, generator=g, device="cuda").bfloat16() * 0.02
    w3 = torch.randn(3 * C, C, generator=g, device="cuda").bfloat16() * 0.02
    x = torch.zeros(M, C, device="cuda")
    ops.gemm(a, w, bias=torch.zeros(C, device="cuda"), epi=ops.EPI_RESIDUAL, out=x, gate=torch.ones(C, device="cuda"))
    ops.gemm(a, w3, epi=ops.EPI_BF16)
    del a, w, w3, x
if which in ("all", "rows"):
    x = torch.randn(M, C, generator=g, device="cuda")
    sh, sc = torch.randn(C, device="cuda"), torch.randn(C, device="cuda") * 0.1
    _, mean, rstd = ops.ln_mod(x, sh, sc, save_stats=True)
    dy = torch.randn(M, C, generator=g, device="cuda").bfloat16()
    ops.ln_mod_bwd(x, dy, sc, None, mean, rstd, torch.zeros(M, C, device="cuda"), False)
    qkv = torch.randn(M, 3 * C, generator=g, device="cuda").bfloat16()
    cos, sin = rope_tables((21, 30, 52), torch.device("cuda"))
    wn = torch.ones(C, device="cuda")
    _, rq = ops.rmsnorm_rope_(qkv[:, :C], wn, cos, sin, 1e-6, M, 0, save_rstd=True)
    ops.rmsnorm_rope_bwd_(qkv[:, C:2 * C], wn, cos, sin, qkv[:, 2 * C:], rq, M, 0, need_dw=False)
torch.cuda.synchronize()
print("done")

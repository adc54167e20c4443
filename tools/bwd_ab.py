"""A/B timing of the attention-backward kernels (per kernel, via torch.profiler) — development helper.
Usage: [PRFL_ATTN_BWD_SS=1] python tools/bwd_ab.py [L] [H]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prfl_b200 import ops  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32760
H = int(sys.argv[2]) if len(sys.argv) > 2 else 40
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v, do = (torch.randn(L, H, 128, generator=g, device="cuda").bfloat16() for _ in range(4))
o, lse = ops.attn_fwd(q, k, v, need_lse=True)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
for _ in range(3):
    ops.attn_bwd(q, k, v, o, do, lse, dq=dq, dk=dk, dv=dv)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 5
a.record()
for _ in range(iters):
    ops.attn_bwd(q, k, v, o, do, lse, dq=dq, dk=dk, dv=dv)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
fl = 10.0 * L * L * 128 * H
print(f"dq={os.environ.get('PRFL_ATTN_BWD_DQ', 'quad')} L={L} H={H}: bwd {ms:.2f} ms = {fl / ms / 1e9:.0f} TFLOP/s algorithmic")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        ops.attn_bwd(q, k, v, o, do, lse, dq=dq, dk=dk, dv=dv)
    torch.cuda.synchronize()
for e in prof.key_averages():
    if "attn" in e.key:
        t = e.device_time_total / e.count / 1e3
        nm = e.key[:70]
        # hardware flops: dK/dV kernel (attn_bwd_kernel<true, ...>) runs 4 tile-GEMMs per tile pair, dQ (<false, ...>) 3
        hw = (8.0 if "attn_bwd_kernel<true" in e.key else 6.0) * L * L * 128 * H if "bwd" in e.key else 0
        print(f"  {nm:70s} {t:8.3f} ms" + (f"  hardware {hw / t / 1e9:.0f} TFLOP/s" if hw else ""))

"""A/B timing of the attention-forward kernel variants (development helper).  Usage: PRFL_ATTN_EXP_FMA=0|25|37|50 python tools/fwd_ab.py [L] [H]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from prfl_b200 import _lib  # noqa: E402
if os.environ.get("PRFL_LIB"):                      # A/B against another build of the library
    _lib.LIB_PATH = os.environ["PRFL_LIB"]
from prfl_b200 import ops  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 32760
H = int(sys.argv[2]) if len(sys.argv) > 2 else 40
g = torch.Generator(device="cuda").manual_seed(0)
q, k, v = (torch.randn(L, H, 128, generator=g, device="cuda").bfloat16() for _ in range(3))
for _ in range(3):
    ops.attn_fwd(q, k, v)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
iters = 10
a.record()
for _ in range(iters):
    ops.attn_fwd(q, k, v)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / iters
print(f"variant={os.environ.get("PRFL_ATTN_FWD", "default(quad12)")} L={L} H={H}: fwd {ms:.3f} ms = {4.0 * L * L * 128 * H / ms / 1e9:.0f} TFLOP/s")

"""Import the UNMODIFIED reference modules from /root/reference (build container only).
TEST INFRASTRUCTURE ONLY — used by tests/golden/make_golden.py to generate fixtures.  Nothing
that runs on the GPU box imports this (the reference checkout does not exist there).

Shims (none touch reference files; recipe from SURVEY.md Appendix C):
  1. stub `diffusers.{configuration_utils,models.modeling_utils,models.normalization}` which
     model.py:8-9 and network.py:6 import but this image lacks;
  2. pre-register namespace packages so `diffusers_lite/wan/__init__.py` (easydict, ftfy, T5 ...)
     never executes;
  3. `WanModel.enable_teacache = False` exactly as the trainers do (train_pavrm.py:237);
  4. CPU only: `flash_attention` asserts CUDA (attention.py:54) -> replaced by SDPA, legitimate
     here because the path uses q_lens=None, non-causal, no dropout, default scale;
  5. `torch.cuda.synchronize` -> no-op so communication.py:80,113 run under gloo.
"""
from __future__ import annotations

import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF = os.environ.get("PRFL_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "diffusers_lite"))


def load():
    if "diffusers_lite.wan.modules.model" in sys.modules:
        m = sys.modules["diffusers_lite.wan.modules.model"]
        return m, sys.modules["diffusers_lite.utils.network"]
    d = types.ModuleType("diffusers")
    cu = types.ModuleType("diffusers.configuration_utils")
    mu = types.ModuleType("diffusers.models")
    mmu = types.ModuleType("diffusers.models.modeling_utils")
    nm = types.ModuleType("diffusers.models.normalization")
    cu.ConfigMixin = type("ConfigMixin", (), {})
    cu.register_to_config = lambda f: f
    mmu.ModelMixin = type("ModelMixin", (nn.Module,), {})

    class FP32LayerNorm(nn.LayerNorm):
        def forward(self, x):
            return F.layer_norm(x.float(), self.normalized_shape,
                                None if self.weight is None else self.weight.float(),
                                None if self.bias is None else self.bias.float(), self.eps).to(x.dtype)

    nm.FP32LayerNorm = FP32LayerNorm
    d.configuration_utils, d.models = cu, mu
    mu.modeling_utils, mu.normalization = mmu, nm
    sys.modules.update({"diffusers": d, "diffusers.configuration_utils": cu, "diffusers.models": mu,
                        "diffusers.models.modeling_utils": mmu, "diffusers.models.normalization": nm})
    for name, sub in [("diffusers_lite", "diffusers_lite"), ("diffusers_lite.wan", "diffusers_lite/wan"),
                      ("diffusers_lite.wan.modules", "diffusers_lite/wan/modules"),
                      ("diffusers_lite.utils", "diffusers_lite/utils")]:
        mod = types.ModuleType(name)
        mod.__path__ = [os.path.join(REF, sub)]
        sys.modules[name] = mod
    import diffusers_lite.wan.modules.model as M
    import diffusers_lite.utils.network as N
    M.WanModel.enable_teacache = False

    def sdpa_flash(q, k, v, q_lens=None, k_lens=None, dropout_p=0., softmax_scale=None, q_scale=None,
                   causal=False, window_size=(-1, -1), deterministic=False, dtype=torch.bfloat16,
                   version=None):
        assert q_lens is None and not causal and dropout_p == 0. and softmax_scale is None and q_scale is None
        outs = []
        for i in range(q.shape[0]):
            lk = k.shape[1] if k_lens is None else int(k_lens[i])
            o = F.scaled_dot_product_attention(q[i:i + 1].transpose(1, 2), k[i:i + 1, :lk].transpose(1, 2),
                                               v[i:i + 1, :lk].transpose(1, 2))
            outs.append(o.transpose(1, 2))
        return torch.cat(outs).type(q.dtype)

    if not torch.cuda.is_available():
        M.flash_attention = sdpa_flash
        torch.cuda.synchronize = lambda *a, **k: None
    return M, N

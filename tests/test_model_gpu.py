"""GPU: the drop-in modules (prfl_b200.model / network / pavrm), running the CUDA kernels through the C ABI,
against (a) the committed outputs of the real reference (tests/golden) and (b) the CPU oracle, block by block.
Tolerance is north_star's: cosine >= 0.999, max|a-b|/max|b| <= 2e-2 (bf16 compute vs the fp32 reference);
reward logits within 1e-2."""
import pytest
import torch

from conftest import cos_rel, golden
from oracle import synth
from oracle import wan_oracle as O

pytestmark = pytest.mark.gpu
COS, REL = 0.999, 2e-2


def _check(a, b, what):
    cos, rel = cos_rel(a.float().cpu(), b)
    assert cos >= COS and rel <= REL, (what, cos, rel)
    return cos, rel


def _model(cfg, sd):
    from prfl_b200.model import WanModel
    m = WanModel(**cfg.kwargs())
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.mark.parametrize("name", ["tiny_t2v", "tiny_i2v"])
def test_forward_vs_reference_golden(name):
    fx = golden(name)
    cfg = O.WanConfig(**fx["cfg"])
    sd = synth.make_wan_state_dict(cfg, fx["seed_w"])
    inp = synth.make_inputs(cfg, fx["latent"], fx["seed_in"])
    m = _model(cfg, sd)
    kw = dict(x=[u.cuda() for u in inp["x"]], t=inp["t"].cuda(), context=[c.cuda() for c in inp["context"]],
              seq_len=inp["seq_len"], clip_fea=None if inp["clip_fea"] is None else inp["clip_fea"].cuda(),
              y=None if inp["y"] is None else [u.cuda() for u in inp["y"]])
    with torch.no_grad():
        out = m(**kw)
        feats = m(**kw, output_features=True, selected_layers=fx["selected"])
    assert out[0].dtype == torch.float32 and out[0].shape == fx["out"][0].shape
    _check(out[0], fx["out"][0], "noise_pred")
    assert len(feats) == len(fx["features"])
    for f, r in zip(feats, fx["features"]):
        _check(f, r, "features")


def test_blocks_vs_oracle():
    """Per-block activations: every block output of the CUDA path vs the fp32 oracle on the same weights."""
    cfg = synth.tiny_cfg("t2v", heads=4, layers=4, ffn=768)
    sd = synth.make_wan_state_dict(cfg, 70)
    inp = synth.make_inputs(cfg, (4, 14, 18), 71)
    with torch.no_grad():
        _, ref_blocks = O.wan_forward(sd, cfg, inp["x"], inp["t"], inp["context"], inp["seq_len"], return_block_outputs=True)
    m = _model(cfg, sd)
    with torch.no_grad():
        feats = m(x=[u.cuda() for u in inp["x"]], t=inp["t"].cuda(), context=[c.cuda() for c in inp["context"]],
                  seq_len=inp["seq_len"], output_features=True, selected_layers=[1, 2, 3, 4])
    for i, (f, r) in enumerate(zip(feats, ref_blocks)):
        _check(f, r, f"block {i}")


@pytest.mark.parametrize("name,tol", [("tiny_reward", 1e-2), ("cfg0_reward", 1e-2)])
def test_reward_logit(name, tol):
    from prfl_b200.pavrm import PavrmScorer
    fx = golden(name)
    cfg = O.WanConfig(**fx["cfg"])
    sd = synth.make_wan_state_dict(cfg, fx["seed_w"])
    qa, mlp = synth.make_reward_state_dicts(cfg.dim, fx["seed_w"] + 1)
    inp = synth.make_inputs(cfg, fx["latent"], fx["seed_in"])
    scorer = PavrmScorer.from_state_dicts(cfg.kwargs(), sd, qa, mlp, num_blocks=fx["nblocks"])
    logit, feats = scorer.score([u.cuda() for u in inp["x"]], inp["t"].cuda(), [c.cuda() for c in inp["context"]],
                                inp["seq_len"], return_features=True)
    assert logit.shape == (1, 1, 1)
    if "features" in fx:
        _check(feats, fx["features"], "features")
    else:
        _check(feats[0, 0, ::97, ::13], fx["features_slice"], "features slice")
    assert abs(float(logit) - float(fx["logit"])) <= tol, (float(logit), float(fx["logit"]))


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_ragged_batch_vs_reference_golden():
    """B = 2 samples of different sizes (300 and 105 tokens) zero-padded to seq_len = 320: per-sample grids, RoPE tables,
    key masking to each sample's length and per-sample timesteps, against the real reference's output."""
    fx = golden("tiny_t2v_ragged")
    cfg = O.WanConfig(**fx["cfg"])
    sd = synth.make_wan_state_dict(cfg, fx["seed_w"])
    g = torch.Generator().manual_seed(fx["seed_in"])
    x = [torch.randn(16, *lat, generator=g) for lat in fx["latents"]]
    ctx = [torch.randn(n, cfg.text_dim, generator=g) * 0.08 for n in (40, 17)]
    t = torch.tensor([400.0, 725.0])
    m = _model(cfg, sd)
    with torch.no_grad():
        out = m(x=[u.cuda() for u in x], t=t.cuda(), context=[c.cuda() for c in ctx], seq_len=fx["seq_len"])
        feats = m(x=[u.cuda() for u in x], t=t.cuda(), context=[c.cuda() for c in ctx], seq_len=fx["seq_len"],
                  output_features=True, selected_layers=[2])
    for i, (o, r) in enumerate(zip(out, fx["out"])):
        assert o.shape == r.shape
        _check(o, r, f"sample {i}")
    # features include the padded rows (the reference computes them too); compare the valid rows of each sample
    for i, n in enumerate((300, 105)):
        _check(feats[0][i, :n], fx["features"][0][i, :n], f"features sample {i}")


def test_flash_attention_free_function_fwd_bwd():
    """`flash_attention` keeps the reference's signature (attention.py:24-38): fp32 in -> fp32 out, k_lens masking,
    differentiable; checked against the oracle's restatement."""
    from prfl_b200.attention import flash_attention
    g = torch.Generator().manual_seed(5)
    q = torch.randn(2, 150, 2, 128, generator=g)
    k = torch.randn(2, 200, 2, 128, generator=g)
    v = torch.randn(2, 200, 2, 128, generator=g)
    k_lens = torch.tensor([200, 131])
    qr, kr, vr = (t.clone().requires_grad_(True) for t in (q, k, v))
    prec = O._Prec(None)
    ref = torch.cat([O.flash_attention(qr[i:i + 1], kr[i:i + 1], vr[i:i + 1], prec, int(k_lens[i])) for i in range(2)])
    cot = torch.randn(ref.shape, generator=g)
    ref.backward(cot)
    qc, kc, vc = (t.cuda().requires_grad_(True) for t in (q, k, v))
    out = flash_attention(qc, kc, vc, k_lens=k_lens)
    assert out.dtype == torch.float32 and out.shape == ref.shape
    _check(out, ref.detach(), "flash_attention")
    out.backward(cot.cuda())
    for name, a, b in (("dq", qc.grad, qr.grad), ("dk", kc.grad, kr.grad), ("dv", vc.grad, vr.grad)):
        _check(a, b, name)
    assert float(kc.grad[1, 131:].abs().max()) == 0.0          # keys beyond k_lens get no gradient
    with pytest.raises(NotImplementedError):
        flash_attention(qc, kc, vc, causal=True)

"""GPU (1 device): the resident-bf16 training layout of prfl_b200.sharding against the replicated
fp32 layout it replaces — same kernels, so forwards must be bit-identical and the optimizer trajectories must agree to
rounding:
  * `make_resident` / `ShardedAdamW(resident)` re-point the block matrices to views of a flat bf16 buffer: forward equals
    the fp32-parameter model's forward exactly (both feed bf16(weight) to the GEMMs);
  * weight gradients written through the gradient sink equal the autograd-returned ones bit for bit, including fused
    QKV / context-KV operands and accumulation over two backward() calls;
  * two optimizer steps (clip 1.0) match dense torch.optim.AdamW on the fp32 model: masters to 1e-4 per tensor;
  * full_state_dict() returns fp32 masters under the reference's key names; frozen resident models (PRFL's reward
    transformer) back-propagate to their input without weight gradients.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _models(mt="t2v", seed=60):
    from oracle import synth
    from prfl_b200.model import WanModel
    cfg = synth.tiny_cfg(mt, heads=2, layers=2)
    sd = synth.make_wan_state_dict(cfg, seed)
    g = torch.Generator().manual_seed(seed + 1)
    sd["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=g) * 0.02
    inp = synth.make_inputs(cfg, (3, 8, 12), seed + 2)

    def make():
        m = WanModel(**cfg.kwargs())
        m.load_state_dict(sd, strict=True)
        return m.cuda().train()
    kw = dict(t=inp["t"].cuda(), context=[c.cuda() for c in inp["context"]], seq_len=inp["seq_len"])
    if mt == "i2v":
        kw.update(clip_fea=inp["clip_fea"].cuda(), y=[u.cuda() for u in inp["y"]])
    return cfg, sd, inp, make, kw


@pytest.mark.parametrize("mt", ["t2v", "i2v"])
def test_resident_forward_bit_identical_and_sink_grads_equal(mt):
    from prfl_b200.sharding import ShardedAdamW
    cfg, sd, inp, make, kw = _models(mt)
    a, b = make(), make()
    opt = ShardedAdamW(a, lr=1e-3).attach_hooks()
    assert opt.resident and a.blocks[0].self_attn.q.weight.dtype == torch.bfloat16
    # fused operands are views of the resident buffer, not copies
    wqkv, _ = a.blocks[0].self_attn._qkv_operands()
    assert wqkv.data_ptr() == a.blocks[0].self_attn.q.weight.data_ptr() and wqkv.shape == (3 * cfg.dim, cfg.dim)
    x = [u.cuda() for u in inp["x"]]
    with torch.no_grad():
        assert torch.equal(a(x=x, **kw)[0], b(x=x, **kw)[0])
    g = torch.Generator().manual_seed(5)
    for micro in range(2):                                          # two backward() calls accumulate
        cot = torch.randn(a(x=x, **kw)[0].shape, generator=g).cuda()
        (a(x=x, **kw)[0] * cot).sum().backward()
        (b(x=x, **kw)[0] * cot).sum().backward()
    shards = opt.reduce_gradients()
    names = dict(b.named_parameters())
    for ui, u in enumerate(opt.units):
        for n, (o, cnt, shp) in u.offsets.items():
            full = (u.sink.prefix + n) if u.kind == "resident" else n
            want = names[full].grad
            got = shards[ui][o:o + cnt].view(shp)
            if want is None:
                assert float(got.abs().max()) == 0.0, full
            else:
                assert torch.equal(got, want.float()), full
    assert all(p.grad is None for p in a.parameters())


def test_sharded_adamw_resident_matches_dense_adamw():
    from prfl_b200.sharding import ShardedAdamW
    cfg, sd, inp, make, kw = _models("t2v", 70)
    a, b = make(), make()
    opt_a = ShardedAdamW(a, lr=1e-3, weight_decay=0.01).attach_hooks()
    opt_b = torch.optim.AdamW(b.parameters(), lr=1e-3, weight_decay=0.01)
    for step in range(2):
        g = torch.Generator().manual_seed(300 + step)
        x = [torch.randn(inp["x"][0].shape, generator=g).cuda()]
        cot = torch.randn(16, *inp["x"][0].shape[1:], generator=g).cuda()
        (a(x=x, **kw)[0] * cot).sum().backward()
        na = opt_a.step(max_norm=1.0)
        (b(x=x, **kw)[0] * cot).sum().backward()
        for p in b.parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        nb = torch.nn.utils.clip_grad_norm_(b.parameters(), 1.0)
        opt_b.step()
        opt_b.zero_grad(set_to_none=True)
        assert abs(float(na) - float(nb)) <= 1e-5 * float(nb)
        full = opt_a.full_state_dict()
        assert set(full) == set(b.state_dict())
        for k, v in b.state_dict().items():
            assert full[k].dtype == torch.float32
            d = float((full[k].float() - v.float().cpu()).abs().max() / (v.float().abs().max().cpu() + 1e-12))
            assert d <= 2e-6, (step, k, d)
        # Re-synchronise the dense model to the sharded one's fp32 masters before the next step: Adam divides every
        # element's update by its own gradient magnitude, so the 1e-7 rounding difference between the two AdamW
        # implementations, through a bf16 ulp flip of a weight, turns the noise-dominated gradients of this model (the
        # cross-attention q / k path over ~470 identical padded text tokens cancels almost exactly) into O(lr) weight
        # differences at the following step.  With identical weights both sides compute bit-identical gradients again,
        # so step 2 (non-zero moments, t = 2) is checked as strictly as step 1.
        b.load_state_dict({k: v.cuda() for k, v in full.items()}, strict=True)
    # the resident bf16 copy is bf16(master)
    u = opt_a.units[0]
    assert torch.equal(u.wflat[:u.n].float(), u.master[:u.n].bfloat16().float())
    with torch.no_grad():
        oa, ob = a(x=x, **kw)[0], b(x=x, **kw)[0]
    assert torch.equal(oa, ob)                                      # same masters => same bf16 operands => same bits


def test_frozen_resident_model_backpropagates_to_input_only():
    from prfl_b200.sharding import make_resident
    cfg, sd, inp, make, kw = _models("t2v", 80)
    a, b = make(), make()
    for m in (a, b):
        for p in m.parameters():
            p.requires_grad_(False)
    make_resident(a)
    x1 = [inp["x"][0].cuda().requires_grad_(True)]
    x2 = [inp["x"][0].cuda().requires_grad_(True)]
    fa = a(x=x1, output_features=True, selected_layers=[2], **kw)[0]
    fb = b(x=x2, output_features=True, selected_layers=[2], **kw)[0]
    assert torch.equal(fa, fb)
    fa.square().sum().backward()
    fb.square().sum().backward()
    assert torch.equal(x1[0].grad, x2[0].grad)
    assert a.blocks[0].ffn[0].weight.dtype == torch.bfloat16 and a.blocks[0].ffn[0].weight.grad is None


def test_resident_block_without_optimizer_refuses_weight_grads():
    from prfl_b200.sharding import make_resident
    cfg, sd, inp, make, kw = _models("t2v", 90)
    a = make()
    make_resident(a, trainable=True)
    out = a(x=[u.cuda() for u in inp["x"]], **kw)[0]
    with pytest.raises(RuntimeError, match="gradient sink|ShardedAdamW"):
        out.sum().backward()


@pytest.mark.xfail(strict=False, reason="written after this round's GPU budget was spent: never executed on hardware (the oracle half "
                   "runs on CPU; the kernels it drives are covered by the B = 1 tests above). XPASS = it holds; a failure here "
                   "must not stop the `-x` run of the verified tests")
def test_ragged_batch_backward_sink_equals_autograd_and_oracle():
    """B = 2 samples of different sizes (300 and 105 tokens, zero-padded to 320): the per-sample loop of BlockFn.backward
    accumulates the second sample's weight gradients onto the first's (`beta` in the wgrad GEMMs of the sink path, `+` in
    the autograd path) and the selective checkpoint keeps one (attention output, LSE) pair per sample.  The resident /
    sink path must equal the autograd path bit for bit, and both must match the fp32 oracle's gradients."""
    from conftest import cos_rel, golden, within_bound_or_eager
    from oracle import synth
    from oracle import wan_oracle as O
    from prfl_b200.model import WanModel
    from prfl_b200.sharding import ShardedAdamW
    fx = golden("tiny_t2v_ragged")
    cfg = O.WanConfig(**fx["cfg"])
    sd = synth.make_wan_state_dict(cfg, fx["seed_w"])
    g = torch.Generator().manual_seed(fx["seed_in"])
    sd["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=g) * 0.02
    xs = [torch.randn(16, *lat, generator=g) for lat in fx["latents"]]
    ctx = [torch.randn(n, cfg.text_dim, generator=g) * 0.08 for n in (40, 17)]
    t = torch.tensor([400.0, 725.0])
    cots = [torch.randn(16, *lat, generator=g) for lat in fx["latents"]]
    keys = ["blocks.0.self_attn.q.weight", "blocks.1.ffn.0.weight", "blocks.1.cross_attn.v.weight", "blocks.0.modulation",
            "blocks.1.self_attn.o.bias", "patch_embedding.weight"]
    # fp32 oracle
    sdr = {k: v.clone().requires_grad_(k in keys) for k, v in sd.items()}
    ref = O.wan_forward(sdr, cfg, [u.clone() for u in xs], t, ctx, fx["seq_len"])
    sum((o * c).sum() for o, c in zip(ref, cots)).backward()

    def run(resident):
        m = WanModel(**cfg.kwargs())
        m.load_state_dict(sd, strict=True)
        m = m.cuda().train()
        opt = ShardedAdamW(m, lr=1e-3).attach_hooks() if resident else None
        out = m(x=[u.cuda() for u in xs], t=t.cuda(), context=[c.cuda() for c in ctx], seq_len=fx["seq_len"])
        sum((o * c.cuda()).sum() for o, c in zip(out, cots)).backward()
        if not resident:
            return {k: p.grad.float() for k, p in m.named_parameters() if p.grad is not None}
        shards, grads = opt.reduce_gradients(), {}
        for ui, u in enumerate(opt.units):
            for n, (o, cnt, shp) in u.offsets.items():
                grads[(u.sink.prefix + n) if u.kind == "resident" else n] = shards[ui][o:o + cnt].view(shp).clone()
        return grads

    ga, gs = run(False), run(True)
    for k, v in ga.items():
        assert torch.equal(gs[k], v), k
    report = {k: cos_rel(gs[k].cpu(), sdr[k].grad) for k in keys}
    bad = {k: v for k, v in report.items() if not within_bound_or_eager(v, None)}
    if bad:
        # outside north_star's fixed bound: hold those to the error the reference's OWN kernel stack (oracle in native mode:
        # bf16 cuBLAS F.linear + flash-attn 2) makes on this very case, as tests/test_backward_gpu.py does
        sde = {k: v.cuda().clone().requires_grad_(k in keys) for k, v in sd.items()}
        oe = O.wan_forward(sde, cfg, [u.cuda() for u in xs], t.cuda(), [c.cuda() for c in ctx], fx["seq_len"],
                           autocast_dtype=torch.bfloat16, native=True)
        sum((o * c.cuda()).sum() for o, c in zip(oe, cots)).backward()
        eager = {k: cos_rel(sde[k].grad.float().cpu(), sdr[k].grad) for k in keys}
        bad = {k: (v, eager[k]) for k, v in bad.items() if not within_bound_or_eager(v, eager[k])}
    assert not bad, bad


@pytest.mark.xfail(strict=False, reason="written after this round's GPU budget was spent: never executed on hardware. It ties the CPU "
                   "stand-ins of tests/ops_emulator.py (on which the CPU twins of the GPU tests run) to the real kernels; "
                   "XPASS = the emulation is faithful at model level")
def test_cpu_emulation_of_the_kernels_matches_the_cuda_path(monkeypatch):
    """The same tiny model, weights, inputs and cotangent through (a) the CUDA kernels and (b) the torch stand-ins that the
    CPU suite substitutes for them: outputs and gradients must agree to bf16 rounding (both sides round at the same points).
    Kept last in the last single-GPU file: the stand-ins are installed with monkeypatch and removed when the test ends."""
    import ops_emulator
    from conftest import cos_rel
    from prfl_b200 import model as pm
    cfg, sd, inp, make, kw = _models("i2v", 95)
    g = torch.Generator().manual_seed(96)
    cot = torch.randn(16, *inp["x"][0].shape[1:], generator=g)
    a = make()                                                       # CUDA
    xa = [u.cuda().requires_grad_(True) for u in inp["x"]]
    oa = a(x=xa, **kw)[0]
    (oa * cot.cuda()).sum().backward()
    torch.cuda.synchronize()
    ops_emulator.install(monkeypatch)                                # from here on prfl_b200.ops.* are the CPU stand-ins
    pm.bump_weight_epoch()
    b = pm.WanModel(**cfg.kwargs())
    b.load_state_dict(sd, strict=True)
    b.train()
    xb = [u.clone().requires_grad_(True) for u in inp["x"]]
    kw_cpu = {k: ([t.cpu() for t in v] if isinstance(v, list) else (v.cpu() if torch.is_tensor(v) else v)) for k, v in kw.items()}
    ob = b(x=xb, **kw_cpu)[0]
    (ob * cot).sum().backward()
    report = {"out": cos_rel(oa.detach().cpu(), ob.detach()), "grad_x": cos_rel(xa[0].grad.cpu(), xb[0].grad)}
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    for k in ("blocks.0.self_attn.q.weight", "blocks.1.ffn.0.weight", "blocks.1.cross_attn.k_img.weight", "blocks.0.modulation",
              "blocks.0.self_attn.norm_k.weight", "patch_embedding.weight", "head.head.weight"):
        report[k] = cos_rel(pa[k].grad.cpu(), pb[k].grad)
    bad = {k: v for k, v in report.items() if not (v[0] >= 0.9995 and v[1] <= 2e-2)}
    assert not bad, bad

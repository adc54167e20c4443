"""CPU: the package's HOST logic — model.py / engine.py (hand-written block backward, per-sample loops, checkpointing) /
network.py / scheduler.py / prfl.py / sampling.py — executed end to end with the C-ABI wrappers replaced by the torch
stand-ins of tests/ops_emulator.py (what tests/test_kernels_gpu.py holds each kernel to), against the SAME reference goldens
and oracle the GPU tests use.  What this pins without a GPU: the order and arguments of the kernel calls, in-place / strided
buffer handling, the backward chain of a block (every gradient the reference produces), ragged batches, the selective
activation checkpoint, the constant-prompt K/V cache, batched classifier-free guidance, the scheduler's bookkeeping and the
PRFL chain.  The numerics of the kernels themselves are the GPU tests' business."""
import pytest
import torch

import ops_emulator
from conftest import cos_rel, golden, within_bound_or_eager
from oracle import synth
from oracle import unipc_oracle as U
from oracle import wan_oracle as O

COS, REL = 0.999, 2e-2


@pytest.fixture()
def emu(monkeypatch):
    from prfl_b200 import model, rope
    ops = ops_emulator.install(monkeypatch)
    model.bump_weight_epoch()
    rope._dev_cache.clear()
    return ops


def _model(cfg, sd, train=False):
    from prfl_b200.model import WanModel
    m = WanModel(**cfg.kwargs())
    m.load_state_dict(sd, strict=True)
    return m.train() if train else m.eval()


def _kw(inp):
    return dict(t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], clip_fea=inp["clip_fea"], y=inp["y"])


@pytest.mark.parametrize("name", ["tiny_t2v", "tiny_i2v"])
def test_forward_and_gradients_vs_reference_golden(emu, name):
    fx = golden(name)
    cfg = O.WanConfig(**fx["cfg"])
    sd = synth.make_wan_state_dict(cfg, fx["seed_w"])
    inp = synth.make_inputs(cfg, fx["latent"], fx["seed_in"])
    m = _model(cfg, sd)
    with torch.no_grad():
        out = m(x=inp["x"], **_kw(inp))
        feats = m(x=inp["x"], **_kw(inp), output_features=True, selected_layers=fx["selected"])
    assert out[0].dtype == torch.float32 and out[0].shape == fx["out"][0].shape
    c, r = cos_rel(out[0], fx["out"][0])
    assert c >= COS and r <= REL, ("noise_pred", c, r)
    for f, ref in zip(feats, fx["features"]):
        c, r = cos_rel(f, ref)
        assert c >= COS and r <= REL, ("features", c, r)
    assert {"gemm", "attn_fwd", "ln_mod", "rmsnorm_rope_", "patchify", "unpatchify", "ln_mod_split"} <= set(ops_emulator.CALLS)
    # gradients: the hand-written backward of engine.py behind autograd vs what the real reference produced
    m.train()
    x = [u.clone().requires_grad_(True) for u in inp["x"]]
    out = m(x=x, **_kw(inp))
    g = torch.Generator().manual_seed(99)
    cot = [torch.randn(o.shape, generator=g) for o in out]
    sum((o * c_).sum() for o, c_ in zip(out, cot)).backward()
    report = {"grad_x": cos_rel(x[0].grad, fx["grad_x"][0])}
    params = dict(m.named_parameters())
    for k, ref in fx.items():
        if k.startswith("grad::"):
            assert params[k[6:]].grad is not None, k
            report[k[6:]] = cos_rel(params[k[6:]].grad, ref)
    # the bf16 stack's own error on this case (the oracle under bf16 autocast rounding) bounds the noise-level gradients
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xe = [u.clone().requires_grad_(True) for u in inp["x"]]
    oe = O.wan_forward(sdr, cfg, xe, inp["t"], inp["context"], inp["seq_len"], inp["clip_fea"], inp["y"], autocast_dtype=torch.bfloat16)
    sum((o * c_).sum() for o, c_ in zip(oe, cot)).backward()
    eager = {"grad_x": cos_rel(xe[0].grad, fx["grad_x"][0]), **{k: cos_rel(sdr[k].grad.float(), fx["grad::" + k]) for k in report if k != "grad_x"}}
    bad = {k: (v, eager[k]) for k, v in report.items() if not within_bound_or_eager(v, eager[k], slack=2.0)}
    assert not bad, bad
    assert {"attn_bwd", "ln_mod_bwd", "rmsnorm_rope_bwd_", "gate_bwd", "colsum", "unpatchify_bwd"} <= set(ops_emulator.CALLS)


def test_ragged_batch_forward_vs_golden_and_backward_vs_oracle(emu):
    """B = 2 samples of different sizes, zero-padded to seq_len: forward vs the real reference's output, gradients (the
    per-sample loop of BlockFn.backward, key-length masking, un-rotated padding rows) vs the fp32 oracle."""
    fx = golden("tiny_t2v_ragged")
    cfg = O.WanConfig(**fx["cfg"])
    sd = synth.make_wan_state_dict(cfg, fx["seed_w"])
    g = torch.Generator().manual_seed(fx["seed_in"])
    xs = [torch.randn(16, *lat, generator=g) for lat in fx["latents"]]
    ctx = [torch.randn(n, cfg.text_dim, generator=g) * 0.08 for n in (40, 17)]
    t = torch.tensor([400.0, 725.0])
    m = _model(cfg, sd)
    with torch.no_grad():
        out = m(x=xs, t=t, context=ctx, seq_len=fx["seq_len"])
        feats = m(x=xs, t=t, context=ctx, seq_len=fx["seq_len"], output_features=True, selected_layers=[2])
    for o, ref in zip(out, fx["out"]):
        c, r = cos_rel(o, ref)
        assert o.shape == ref.shape and c >= COS and r <= REL, (c, r)
    c, r = cos_rel(feats[0], fx["features"][0])
    assert c >= COS and r <= REL, ("features", c, r)
    # backward (non-zero head so that gradients reach the blocks)
    sd = dict(sd)
    sd["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=g) * 0.02
    cots = [torch.randn(16, *lat, generator=g) for lat in fx["latents"]]
    keys = ["blocks.0.self_attn.q.weight", "blocks.1.ffn.0.weight", "blocks.1.cross_attn.v.weight", "blocks.0.modulation",
            "blocks.1.self_attn.o.bias", "blocks.0.self_attn.norm_k.weight", "patch_embedding.weight"]
    sdr = {k: v.clone().requires_grad_(k in keys) for k, v in sd.items()}
    xr = [u.clone().requires_grad_(True) for u in xs]
    ref = O.wan_forward(sdr, cfg, xr, t, ctx, fx["seq_len"])
    sum((o * c_).sum() for o, c_ in zip(ref, cots)).backward()
    mt = _model(cfg, sd, train=True)
    xg = [u.clone().requires_grad_(True) for u in xs]
    og = mt(x=xg, t=t, context=ctx, seq_len=fx["seq_len"])
    sum((o * c_).sum() for o, c_ in zip(og, cots)).backward()
    params = dict(mt.named_parameters())
    report = {k: cos_rel(params[k].grad, sdr[k].grad) for k in keys}
    report.update({f"grad_x{i}": cos_rel(xg[i].grad, xr[i].grad) for i in range(2)})
    sde = {k: v.clone().requires_grad_(k in keys) for k, v in sd.items()}
    xe = [u.clone().requires_grad_(True) for u in xs]
    oe = O.wan_forward(sde, cfg, xe, t, ctx, fx["seq_len"], autocast_dtype=torch.bfloat16)
    sum((o * c_).sum() for o, c_ in zip(oe, cots)).backward()
    eager = {k: cos_rel(sde[k].grad.float(), sdr[k].grad) for k in keys}
    eager.update({f"grad_x{i}": cos_rel(xe[i].grad, xr[i].grad) for i in range(2)})
    bad = {k: (v, eager[k]) for k, v in report.items() if not within_bound_or_eager(v, eager[k], slack=2.0)}
    assert not bad, bad


def test_selective_checkpoint_equals_full_recompute_bit_for_bit(emu):
    from prfl_b200 import engine
    fx = golden("tiny_t2v")
    cfg = O.WanConfig(**fx["cfg"])
    sd = synth.make_wan_state_dict(cfg, fx["seed_w"])
    inp = synth.make_inputs(cfg, fx["latent"], fx["seed_in"])
    g = torch.Generator().manual_seed(99)
    grads, counts, prev = {}, {}, engine.SAVE_ATTENTION
    try:
        for mode in (True, False):
            engine.SAVE_ATTENTION = mode
            m = _model(cfg, sd, train=True)
            x = [u.clone().requires_grad_(True) for u in inp["x"]]
            del ops_emulator.CALLS[:]
            out = m(x=x, **_kw(inp))
            if mode:
                cot = [torch.randn(o.shape, generator=g) for o in out]
            sum((o * c).sum() for o, c in zip(out, cot)).backward()
            counts[mode] = ops_emulator.CALLS.count("attn_fwd")
            grads[mode] = {"x": x[0].grad.clone(), **{k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}}
    finally:
        engine.SAVE_ATTENTION = prev
    assert set(grads[True]) == set(grads[False]) and all(torch.equal(grads[True][k], grads[False][k]) for k in grads[True])
    assert counts[False] - counts[True] == cfg.num_layers           # the recompute skipped one self-attention per block


def test_reward_chain_vs_reference_golden(emu):
    from prfl_b200.pavrm import PavrmScorer
    fx = golden("tiny_reward")
    cfg = O.WanConfig(**fx["cfg"])
    sd = synth.make_wan_state_dict(cfg, fx["seed_w"])
    qa, mlp = synth.make_reward_state_dicts(cfg.dim, fx["seed_w"] + 1)
    inp = synth.make_inputs(cfg, fx["latent"], fx["seed_in"])
    scorer = PavrmScorer.from_state_dicts(cfg.kwargs(), sd, qa, mlp, num_blocks=fx["nblocks"], device="cpu")
    logit, feats = scorer.score(inp["x"], inp["t"], inp["context"], inp["seq_len"], return_features=True)
    c, r = cos_rel(feats, fx["features"])
    assert c >= COS and r <= REL and abs(float(logit) - float(fx["logit"])) <= 1e-2, (c, r, float(logit), float(fx["logit"]))
    scorer.train()
    f_ref = fx["features"].clone().requires_grad_(True)
    lg = scorer.mlp(scorer.query_attention(f_ref))
    prob = torch.sigmoid(lg)
    loss = torch.nn.functional.binary_cross_entropy(prob, torch.ones_like(prob))
    assert abs(float(lg.detach()) - float(fx["logit"])) <= 1e-4 and abs(float(loss.detach()) - float(fx["loss"])) <= 1e-4
    loss.backward()
    c, r = cos_rel(f_ref.grad, fx["grad_features"])
    assert c >= 0.99999 and r <= 1e-3, ("grad_features", c, r)
    x = [u.clone().requires_grad_(True) for u in inp["x"]]
    fe = scorer.features(x, inp["t"], inp["context"], inp["seq_len"])
    fe.backward(gradient=fx["grad_features"])
    c, r = cos_rel(x[0].grad, fx["grad_x"][0])
    assert c >= COS and r <= REL, ("grad_x", c, r)
    # frozen blocks (PRFL's reward model): dgrad only — no weight-gradient GEMM, no column reduction is issued
    for p in scorer.parameters():
        p.requires_grad_(False)
    del ops_emulator.CALLS[:]
    x2 = [u.clone().requires_grad_(True) for u in inp["x"]]
    scorer.features(x2, inp["t"], inp["context"], inp["seq_len"]).backward(gradient=fx["grad_features"])
    assert "colsum" not in ops_emulator.CALLS and torch.equal(x2[0].grad, x[0].grad)


def test_scheduler_class_reproduces_the_reference_chain(emu):
    """The real FlowUniPCMultistepScheduler object (bookkeeping of model_outputs / last_sample / orders + folded coefficients)
    stepping a chain, against the trajectory the unmodified reference scheduler produced (tests/golden/unipc.pt)."""
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    fx = golden("unipc")
    g = torch.Generator().manual_seed(fx["seed"])                   # same inputs as tests/golden/make_golden.py::case_unipc
    x_init = torch.randn(fx["shape"], generator=g)
    w = torch.randn(fx["shape"], generator=g) * 0.5
    for (steps, shift, st, order), ch in fx["chains"].items():
        s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False, solver_type=st, solver_order=order)
        s.set_timesteps(steps, device="cpu", shift=shift)
        x = x_init.clone()
        for i, t in enumerate(s.timesteps):
            x = s.step(U.toy_velocity(x, t, w), t, x, return_dict=False)[0]
            if torch.isfinite(ch["traj"][i]).all():
                torch.testing.assert_close(x, ch["traj"][i], rtol=2e-5, atol=2e-5)
    assert "unipc_step" in ops_emulator.CALLS


def test_prepared_context_batched_cfg_and_refl_chain(emu):
    """Constant-prompt K/V cache == recomputing; cond + uncond as one B = 2 forward == two forwards; the PRFL
    chain runs m no-grad steps + the differentiable step + the frozen reward model and back-propagates to the VGM only."""
    from prfl_b200.network import MLP, QueryAttention
    from prfl_b200.prfl import refl_chain
    from prfl_b200.sampling import sample_loop
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    cfg = synth.tiny_cfg("t2v", heads=2, layers=2)
    sd = synth.make_wan_state_dict(cfg, 5)
    g = torch.Generator().manual_seed(6)
    sd["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=g) * 0.02
    inp = synth.make_inputs(cfg, (2, 8, 8), 7)
    m = _model(cfg, sd)
    with torch.no_grad():
        plain = m(x=inp["x"], **_kw(inp))[0]
        pc = m.prepare_context(inp["context"])
        del ops_emulator.CALLS[:]
        first = m(x=inp["x"], t=inp["t"], context=pc, seq_len=inp["seq_len"])[0]
        n_first = ops_emulator.CALLS.count("gemm")
        del ops_emulator.CALLS[:]
        again = m(x=inp["x"], t=inp["t"], context=pc, seq_len=inp["seq_len"])[0]
        n_again = ops_emulator.CALLS.count("gemm")
    assert torch.equal(plain, first) and torch.equal(plain, again) and n_first - n_again == cfg.num_layers   # one K/V GEMM per block saved
    ctx_null = [torch.randn(9, cfg.text_dim, generator=g) * 0.08]
    noise = torch.randn(16, 2, 8, 8, generator=g)
    a = sample_loop(m, noise, inp["context"], ctx_null, inp["seq_len"], sampling_steps=3, batch_cfg=True)[0]
    b = sample_loop(m, noise, inp["context"], ctx_null, inp["seq_len"], sampling_steps=3, batch_cfg=False, cache_context=False)[0]
    # bit-identical on the GPU (tests/test_scheduler_gpu.py: per-sample kernels see the same operands); here the fp32 CPU BLAS
    # behind the emulation rounds a 2-row and a 1-row product differently, and three bf16 steps amplify that last bit
    c, r = cos_rel(a, b)
    assert torch.isfinite(a).all() and c >= 0.99999 and r <= 1e-2, (c, r)
    # PRFL chain: VGM trainable, reward model frozen
    vgm = _model(cfg, sd, train=True)
    lrm = _model(cfg, synth.make_wan_state_dict(cfg, 8))
    lrm.head = None
    qa, mlp = QueryAttention(cfg.dim, 1, 2, dropout=0.0, return_type="query").eval(), MLP(cfg.dim).eval()
    for mod in (lrm, qa, mlp):
        for p in mod.parameters():
            p.requires_grad_(False)
    sched = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    loss, reward = refl_chain(vgm, lrm, qa, mlp, sched, noise[None], torch.stack(inp["context"]), inp["seq_len"], 2, flow_shift=5.0,
                              feature_layer=[2], inference_steps=8)
    loss.backward()
    assert reward.shape == (1, 1, 1) and torch.isfinite(loss)
    got = [n for n, p in vgm.named_parameters() if p.grad is not None and float(p.grad.abs().max()) > 0]
    assert "blocks.0.self_attn.q.weight" in got and "head.head.weight" in got and "patch_embedding.weight" in got
    assert all(p.grad is None for p in lrm.parameters())


@pytest.mark.parametrize("case", ["two_backwards", "ragged_batch", "two_forwards_one_backward"])
def test_resident_layout_and_gradient_sink_equal_the_autograd_path(emu, monkeypatch, case):
    """sharding.ResidentUnit / _BlockSink over the emulated kernels (the CUDA-only guard lifted by the test hook): parameters
    become views of one flat bf16 buffer and the fused operands views of the same bytes; the forward equals the fp32-parameter
    model's bit for bit; weight gradients written through the sink (one fused [3C, C] / [2C, C] wgrad, `beta` accumulation over
    a second backward() or over the second sample of a ragged batch) equal the autograd-returned ones bit for bit.  The
    ragged case is the CPU twin of tests/test_sharding_gpu.py::test_ragged_batch_backward_sink_equals_autograd_and_oracle."""
    from prfl_b200 import sharding
    from prfl_b200.sharding import ShardedAdamW
    monkeypatch.setattr(sharding, "_ALLOW_CPU_UNITS", True)
    g = torch.Generator().manual_seed(5)
    if case == "ragged_batch":
        fx = golden("tiny_t2v_ragged")
        cfg = O.WanConfig(**fx["cfg"])
        xs = [torch.randn(16, *lat, generator=g) for lat in fx["latents"]]
        kw = dict(t=torch.tensor([400.0, 725.0]), context=[torch.randn(n, cfg.text_dim, generator=g) * 0.08 for n in (40, 17)], seq_len=fx["seq_len"])
    else:
        cfg = synth.tiny_cfg("i2v", heads=2, layers=2)
        inp = synth.make_inputs(cfg, (3, 8, 12), 62)
        xs, kw = inp["x"], _kw(inp)
    sd = synth.make_wan_state_dict(cfg, 60)
    sd["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=g) * 0.02
    a, b = _model(cfg, sd, train=True), _model(cfg, sd, train=True)
    opt = ShardedAdamW(a, lr=1e-3).attach_hooks()
    assert opt.resident and a.blocks[0].self_attn.q.weight.dtype == torch.bfloat16
    wqkv, _ = a.blocks[0].self_attn._qkv_operands()
    assert wqkv.data_ptr() == a.blocks[0].self_attn.q.weight.data_ptr() and wqkv.shape == (3 * cfg.dim, cfg.dim)
    with torch.no_grad():
        assert all(torch.equal(u, v) for u, v in zip(a(x=xs, **kw), b(x=xs, **kw)))
    if case == "two_forwards_one_backward":
        # the Bradley-Terry branch of PAVRM training (train_pavrm.py:816-845): a "win" and a "lose" forward through the same blocks,
        # ONE backward — every block's backward node runs twice inside it and the second run accumulates onto the first
        xs2 = [torch.randn(u.shape, generator=g) for u in xs]
        c1, c2 = ([torch.randn(16, *u.shape[1:], generator=g) for u in xs] for _ in range(2))
        for mdl in (a, b):
            (sum((o * c).sum() for o, c in zip(mdl(x=xs, **kw), c1)) - sum((o * c).sum() for o, c in zip(mdl(x=xs2, **kw), c2))).backward()
    for micro in range({"two_backwards": 2, "ragged_batch": 1}.get(case, 0)):
        cots = [torch.randn(16, *u.shape[1:], generator=g) for u in xs]
        sum((o * c).sum() for o, c in zip(a(x=xs, **kw), cots)).backward()
        sum((o * c).sum() for o, c in zip(b(x=xs, **kw), cots)).backward()
    shards = opt.reduce_gradients()
    names = dict(b.named_parameters())
    checked = 0
    for ui, u in enumerate(opt.units):
        for n, (o, cnt, shp) in u.offsets.items():
            full = (u.sink.prefix + n) if u.kind == "resident" else n
            want, got = names[full].grad, shards[ui][o:o + cnt].view(shp)
            if want is None:
                assert float(got.abs().max()) == 0.0, full
            else:
                assert torch.equal(got, want.float()), full
                checked += 1
    assert checked > 40 and all(p.grad is None for p in a.parameters())


def test_refl_chain_vs_the_oracle_chain(emu):
    """prfl.refl_chain (train_prfl.py:631-798: m no-grad denoising steps, the differentiable step, the scheduler step, the frozen
    reward model, 0.1 * relu(2 - r)) against the same chain built from the two oracles on identical weights and noise: the latent
    the differentiable step starts from, the reward, the loss, and the VGM gradients (CPU twin of
    tests/test_scheduler_gpu.py::test_refl_chain_vs_oracle)."""
    from prfl_b200.network import MLP, QueryAttention
    from prfl_b200.prfl import refl_chain
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    cfg = synth.tiny_cfg("t2v", heads=2, layers=2)
    sd_v, sd_l = synth.make_wan_state_dict(cfg, 80), synth.make_wan_state_dict(cfg, 81)
    g = torch.Generator().manual_seed(84)
    sd_v["head.head.weight"] = torch.randn(sd_v["head.head.weight"].shape, generator=g) * 0.02
    qa_sd, mlp_sd = synth.make_reward_state_dicts(cfg.dim, 82)
    inp = synth.make_inputs(cfg, (3, 8, 12), 83)
    noise, steps, mid, shift = inp["x"][0], 8, 2, 5.0
    # the oracle chain (fp32)
    sd_vg = {k: v.clone().requires_grad_(True) for k, v in sd_v.items()}
    osch = U.UniPCOracle()
    osch.set_timesteps(steps, shift=shift)
    lat = noise[None].clone()
    with torch.no_grad():
        for i in range(mid):
            t = osch.timesteps[i]
            lat = osch.step(O.wan_forward(sd_v, cfg, [lat[0]], t[None], inp["context"], inp["seq_len"])[0][None], t, lat)
    t = osch.timesteps[mid]
    lat_o = osch.step(O.wan_forward(sd_vg, cfg, [lat[0]], t[None], inp["context"], inp["seq_len"])[0][None], t, lat)
    logit_o, _ = O.pavrm_reward(sd_l, cfg, qa_sd, mlp_sd, [lat_o[0]], osch.timesteps[mid + 1][None], inp["context"], inp["seq_len"],
                                selected_layers=(2,), num_blocks=2)
    reward_o = torch.sigmoid(logit_o.float())
    loss_o = U.prfl_loss(reward_o)
    loss_o.backward()
    # the product chain over the emulated kernels
    vgm = _model(cfg, sd_v, train=True)
    lrm = _model(cfg, sd_l)
    lrm.head = None
    qa = QueryAttention(cfg.dim, 1, 8, dropout=0.0, return_type="query").eval()
    qa.load_state_dict(qa_sd, strict=True)
    mlp = MLP(cfg.dim).eval()
    mlp.load_state_dict(mlp_sd, strict=True)
    for mod in (lrm, qa, mlp):
        for p in mod.parameters():
            p.requires_grad_(False)
    sch = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    loss, reward = refl_chain(vgm, lrm, qa, mlp, sch, noise[None], torch.stack(inp["context"]), inp["seq_len"], mid,
                              inference_steps=steps, flow_shift=shift, feature_layer=[2])
    loss.backward()
    c, r = cos_rel(sch.last_sample, osch.last_sample)
    assert c >= COS and r <= REL, ("latent", c, r)
    assert abs(float(reward.detach()) - float(reward_o.detach())) <= 1e-2 and abs(float(loss.detach()) - float(loss_o.detach())) <= 1e-3
    if float(loss_o.detach()) > 0:                                  # the hinge is active: gradients flow, compare the large ones
        params = dict(vgm.named_parameters())
        big = max(float(v.grad.abs().max()) for v in sd_vg.values() if v.grad is not None)
        for k in ("head.head.weight", "blocks.1.ffn.2.weight", "blocks.0.self_attn.o.weight"):
            if float(sd_vg[k].grad.abs().max()) >= 1e-2 * big:
                c, r = cos_rel(params[k].grad, sd_vg[k].grad)
                assert c >= 0.99 and r <= 0.15, (k, c, r)          # through the reward MLP's ReLU masks: the GPU test's fixed slack


def test_i2v_sampling_loop_batched_cfg_and_context_cache(emu):
    """Image-to-video sampling (image2video.py:357-388): CLIP tokens + conditioning latents ride along; cond + uncond as one B = 2
    forward with the prompt pair prepared once == two forwards per step without any cache (same values up to CPU BLAS rounding)."""
    from prfl_b200.sampling import sample_loop
    cfg = synth.tiny_cfg("i2v", heads=2, layers=2)
    sd = synth.make_wan_state_dict(cfg, 15)
    g = torch.Generator().manual_seed(16)
    sd["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=g) * 0.02
    inp = synth.make_inputs(cfg, (2, 8, 8), 17)
    m = _model(cfg, sd)
    ctx_null = [torch.randn(11, cfg.text_dim, generator=g) * 0.08]
    noise = torch.randn(16, 2, 8, 8, generator=g)
    traj = []
    kw = dict(sampling_steps=4, shift=3.0, guide_scale=4.0, clip_fea=inp["clip_fea"], y=inp["y"])
    a = sample_loop(m, noise, inp["context"], ctx_null, inp["seq_len"], trajectory=traj, **kw)[0]
    b = sample_loop(m, noise, inp["context"], ctx_null, inp["seq_len"], batch_cfg=False, cache_context=False, **kw)[0]
    c_, r_ = cos_rel(a, b)
    assert a.shape == noise.shape and len(traj) == 4 and torch.isfinite(a).all() and c_ >= 0.99999 and r_ <= 1e-2, (c_, r_)
    # against the oracle loop: oracle DiT (fp32) + the scheduler oracle, guidance as text2video.py:295-296
    osch = U.UniPCOracle()
    osch.set_timesteps(4, shift=3.0)
    lat = noise[None].clone()
    with torch.no_grad():
        for t in osch.timesteps:
            vc = O.wan_forward(sd, cfg, [lat[0]], t[None], inp["context"], inp["seq_len"], inp["clip_fea"], inp["y"])[0]
            vu = O.wan_forward(sd, cfg, [lat[0]], t[None], ctx_null, inp["seq_len"], inp["clip_fea"], inp["y"])[0]
            lat = osch.step((vu + 4.0 * (vc - vu))[None], t, lat)
    c_, r_ = cos_rel(a, lat[0])
    assert c_ >= COS and r_ <= REL, ("vs oracle loop", c_, r_)


@pytest.mark.skipif(not __import__("oracle.ref_shim", fromlist=["x"]).available(), reason="needs the reference checkout (/root/reference)")
@pytest.mark.parametrize("shape,with_e", [((1, 150, 256), False), ((2, 1, 150, 256), False), ((2, 1, 150, 256), True), ((3, 256), False)])
def test_query_attention_equals_the_live_reference_module(emu, shape, with_e):
    """The algebraic collapse of the single-query pooling (two streaming passes instead of the K/V in-projection GEMM + hd-wide
    attention) against the UNMODIFIED reference `QueryAttention` (nn.MultiheadAttention inside) on the same weights, for every
    input rank the reference handles (network.py:58-69), the optional query offset `e`, and `return_type='query'`: values and
    the gradients w.r.t. features, the learnable query and the in-projection weights."""
    from oracle import ref_shim
    from prfl_b200.network import QueryAttention
    _, N = ref_shim.load()
    qa_sd, _ = synth.make_reward_state_dicts(256, 91)
    ref = N.QueryAttention(256, num_queries=1, num_heads=8, dropout=0.0, return_type="query").eval()
    ours = QueryAttention(256, num_queries=1, num_heads=8, dropout=0.0, return_type="query").eval()
    ref.load_state_dict(qa_sd, strict=True)
    ours.load_state_dict(qa_sd, strict=True)
    g = torch.Generator().manual_seed(92)
    x = torch.randn(*shape, generator=g)
    e = torch.randn(1, 256, generator=g) * 0.1 if with_e else None
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    out_r, out_o = ref(xr, e=e), ours(xo, e=e)
    assert out_o.shape == out_r.shape
    torch.testing.assert_close(out_o, out_r, rtol=1e-4, atol=1e-5)
    cot = torch.randn(out_r.shape, generator=g)
    (out_r * cot).sum().backward()
    (out_o * cot).sum().backward()
    for name, a, b in (("features", xo.grad, xr.grad), ("queries", ours.queries.grad, ref.queries.grad),
                       ("in_proj_weight", ours.multihead_attn.in_proj_weight.grad, ref.multihead_attn.in_proj_weight.grad),
                       ("out_proj.weight", ours.multihead_attn.out_proj.weight.grad, ref.multihead_attn.out_proj.weight.grad)):
        c, r = cos_rel(a, b)
        assert c >= 0.99999 and r <= 1e-3, (name, c, r)


@pytest.mark.skipif(not __import__("oracle.ref_shim", fromlist=["x"]).available(), reason="needs the reference checkout (/root/reference)")
def test_flf2v_model_type_vs_the_live_reference(emu):
    """`model_type='flf2v'` (first-last-frame-to-video, task flf2v-14b-720p of NAME_MAPPING, train_prfl.py:86-93): two CLIP images
    per sample go through `MLPProj` with its positional embedding (model.py:392-410) and 514 image tokens precede the text in
    the cross-attention context.  Product (emulated kernels) vs the unmodified reference model on the same weights."""
    from oracle import ref_shim
    from prfl_b200.model import WanModel
    M, _ = ref_shim.load()
    kw_model = dict(model_type="flf2v", in_dim=36, dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=64)
    ref = M.WanModel(**kw_model).eval()
    g = torch.Generator().manual_seed(21)
    with torch.no_grad():
        for p in ref.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.05 if p.dim() > 1 else 0.02))
        for n_, p in ref.named_parameters():
            if n_.endswith("norm_q.weight") or n_.endswith("norm_k.weight") or n_.endswith("norm_k_img.weight") or "norm3.weight" in n_ \
                    or n_.endswith("proj.0.weight") or n_.endswith("proj.4.weight"):
                p.add_(1.0)
    ours = WanModel(**kw_model).eval()
    assert set(ours.state_dict()) == set(ref.state_dict())
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = [torch.randn(16, 2, 8, 8, generator=g)]
    y = [torch.randn(20, 2, 8, 8, generator=g)]
    clip = torch.randn(2, 257, 1280, generator=g)                    # first + last frame
    ctx = [torch.randn(30, 64, generator=g) * 0.5]
    t = torch.tensor([500.0])
    with torch.no_grad():
        want = ref(x, t=t, context=ctx, seq_len=32, clip_fea=clip, y=y)[0]
        got = ours(x, t=t, context=ctx, seq_len=32, clip_fea=clip, y=y)[0]
    c, r = cos_rel(got, want)
    assert got.shape == want.shape and c >= COS and r <= REL, (c, r)


def test_edge_inputs_empty_prompt_full_prompt_integer_timestep(emu):
    """Edge inputs of WanModel.forward against the oracle: a prompt of zero tokens (all 512 context rows are padding,
    model.py:597-603), a prompt that fills text_len exactly, an int64 timestep (PRFL passes scheduler timesteps, train_prfl.py:671)
    and a float one give the same result."""
    cfg = synth.tiny_cfg("t2v", heads=2, layers=2)
    sd = synth.make_wan_state_dict(cfg, 33)
    g = torch.Generator().manual_seed(34)
    sd["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=g) * 0.02
    x = [torch.randn(16, 2, 8, 12, generator=g)]
    m = _model(cfg, sd)
    for n_ctx in (0, 512):
        ctx = [torch.randn(n_ctx, cfg.text_dim, generator=g) * 0.08]
        with torch.no_grad():
            got = m(x=x, t=torch.tensor([731]), context=ctx, seq_len=48)[0]
            got_f = m(x=x, t=torch.tensor([731.0]), context=ctx, seq_len=48)[0]
            want = O.wan_forward(sd, cfg, x, torch.tensor([731]), ctx, 48)[0]
        assert torch.equal(got, got_f)
        c, r = cos_rel(got, want)
        assert c >= COS and r <= REL, (n_ctx, c, r)
    with pytest.raises(AssertionError):                               # a sequence longer than seq_len is the caller's error (model.py:586)
        m(x=x, t=torch.tensor([731]), context=ctx, seq_len=40)


def test_resident_training_state_checkpoint_and_bit_exact_resume(emu, monkeypatch, tmp_path):
    """Resident bf16 weights + fp32 masters through the reference's checkpoint layout and the optimizer shard file: after two
    training steps, `save_checkpoint(state_dict=opt.full_state_dict())` + `save_optimizer`; a fresh process-equivalent
    (`WanModel.from_pretrained` + `ShardedAdamW` + `load_optimizer`) continues bit for bit like the uninterrupted run."""
    from prfl_b200 import sharding
    from prfl_b200.checkpoint import load_optimizer, save_checkpoint, save_optimizer
    from prfl_b200.model import WanModel
    from prfl_b200.sharding import ShardedAdamW
    monkeypatch.setattr(sharding, "_ALLOW_CPU_UNITS", True)
    cfg = synth.tiny_cfg("t2v", heads=2, layers=2)
    sd = synth.make_wan_state_dict(cfg, 44)
    g = torch.Generator().manual_seed(45)
    sd["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=g) * 0.02
    inp = synth.make_inputs(cfg, (2, 8, 8), 46)

    def steps(m, opt, n, seed):
        gi = torch.Generator().manual_seed(seed)
        for _ in range(n):
            x = [torch.randn(inp["x"][0].shape, generator=gi)]
            cot = torch.randn(inp["x"][0].shape, generator=gi)
            (m(x=x, **_kw(inp))[0] * cot).sum().backward()
            opt.step(max_norm=1.0)

    a = _model(cfg, sd, train=True)
    oa = ShardedAdamW(a, lr=1e-3, weight_decay=0.01).attach_hooks()
    steps(a, oa, 2, 1)
    d = save_checkpoint(a, 0, str(tmp_path), 2, state_dict=oa.full_state_dict())
    save_optimizer(oa, str(tmp_path), 2)
    b = WanModel.from_pretrained(d).train()
    assert b.blocks[0].ffn[0].weight.dtype == torch.float32          # the checkpoint holds the fp32 masters, as the reference's does
    ob = ShardedAdamW(b, lr=1e-3, weight_decay=0.01).attach_hooks()
    load_optimizer(ob, d)
    ua, ub = oa.units[0], ob.units[0]
    assert torch.equal(ua.master, ub.master) and torch.equal(ua.wflat, ub.wflat) and ua.t == ub.t == 2
    steps(a, oa, 1, 2)
    steps(b, ob, 1, 2)
    fa, fb = oa.full_state_dict(), ob.full_state_dict()
    assert all(torch.equal(fa[k], fb[k]) for k in fa)
    with torch.no_grad():
        assert torch.equal(a(x=inp["x"], **_kw(inp))[0], b(x=inp["x"], **_kw(inp))[0])

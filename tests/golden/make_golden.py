"""Generate tests/golden/*.pt by running the UNMODIFIED reference (via oracle/ref_shim.py).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures hold only reference OUTPUTS; weights and inputs are re-created by oracle/synth.py
from the seeds recorded in each fixture.  All runs are CPU fp32 (the only mode the reference can
execute here), torch.manual_seed-free: every tensor comes from a seeded torch.Generator.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_model(M, cfg, sd):
    m = M.WanModel(**cfg.kwargs())
    missing, unexpected = m.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return m.eval()


def case_model(M, name, cfg, latent, seed_w, seed_in, selected, with_grad=True):
    sd = synth.make_wan_state_dict(cfg, seed_w)
    inp = synth.make_inputs(cfg, latent, seed_in)
    m = ref_model(M, cfg, sd)
    x = [u.clone().requires_grad_(with_grad) for u in inp["x"]]
    kw = dict(x=x, t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], clip_fea=inp["clip_fea"], y=inp["y"])
    out = m(**kw)
    fx = dict(cfg=cfg.kwargs(), latent=latent, seed_w=seed_w, seed_in=seed_in, selected=selected,
              out=[o.detach().clone() for o in out])
    if with_grad:
        # deterministic "loss": fixed pseudo-random cotangent
        g = torch.Generator().manual_seed(99)
        cot = [torch.randn(o.shape, generator=g) for o in out]
        loss = sum((o * c).sum() for o, c in zip(out, cot))
        loss.backward()
        fx["grad_x"] = [u.grad.clone() for u in x]
        for k in ("blocks.0.self_attn.q.weight", "blocks.0.self_attn.norm_k.weight", "blocks.0.modulation",
                  "blocks.1.ffn.0.weight", "blocks.1.cross_attn.v.weight", "blocks.0.norm3.weight",
                  "patch_embedding.weight", "head.head.weight", "blocks.1.cross_attn.o.bias"):
            p = dict(m.named_parameters())[k]
            fx["grad::" + k] = p.grad.clone()
        m.zero_grad()
    with torch.no_grad():
        feats = m(**{**kw, "x": inp["x"]}, output_features=True, selected_layers=selected)
    fx["features"] = [f.clone() for f in feats]
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "out", [tuple(o.shape) for o in out], "feat", [tuple(f.shape) for f in feats])


def case_ragged(M, name, cfg, latents, seq_len, seed_w, seed_in):
    """Two samples of different sizes in one call, zero-padded to seq_len (model.py:578-587; k_lens masking :188-193)."""
    sd = synth.make_wan_state_dict(cfg, seed_w)
    g = torch.Generator().manual_seed(seed_in)
    x = [torch.randn(16, *lat, generator=g) for lat in latents]
    ctx = [torch.randn(n, cfg.text_dim, generator=g) * 0.08 for n in (40, 17)]
    t = torch.tensor([400.0, 725.0])
    m = ref_model(M, cfg, sd)
    with torch.no_grad():
        out = m(x=x, t=t, context=ctx, seq_len=seq_len)
        feats = m(x=x, t=t, context=ctx, seq_len=seq_len, output_features=True, selected_layers=[2])
    fx = dict(cfg=cfg.kwargs(), latents=latents, seq_len=seq_len, seed_w=seed_w, seed_in=seed_in,
              out=[o.clone() for o in out], features=[f.clone() for f in feats])
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "out", [tuple(o.shape) for o in out], "feat", [tuple(f.shape) for f in feats])


def case_reward(M, N, name, cfg, latent, nblocks, seed_w, seed_in, slice_only=False):
    sd = synth.make_wan_state_dict(cfg, seed_w)
    inp = synth.make_inputs(cfg, latent, seed_in)
    qa_sd, mlp_sd = synth.make_reward_state_dicts(cfg.dim, seed_w + 1)
    m = ref_model(M, cfg, sd)
    m.blocks = torch.nn.ModuleList([m.blocks[i] for i in range(nblocks)])      # train_pavrm.py:215-231
    m.head = None                                                               # train_pavrm.py:233-235
    qa = N.QueryAttention(cfg.dim, num_queries=1, num_heads=8, dropout=0.0, return_type="query").eval()
    qa.load_state_dict(qa_sd, strict=True)
    mlp = N.MLP(cfg.dim).eval()
    mlp.load_state_dict(mlp_sd, strict=True)
    x = [u.clone().requires_grad_(True) for u in inp["x"]]
    feats = m(x=x, t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], clip_fea=inp["clip_fea"],
              y=inp["y"], output_features=True, selected_layers=[nblocks])
    stacked = torch.stack(feats)                                                # list2batch
    stacked.retain_grad()
    pooled = qa(stacked)
    logit = mlp(pooled)
    prob = N.forward_mlp(mlp, pooled)
    loss = torch.nn.functional.binary_cross_entropy(prob, torch.ones_like(prob))
    loss.backward()
    fx = dict(cfg=cfg.kwargs(), latent=latent, nblocks=nblocks, seed_w=seed_w, seed_in=seed_in,
              logit=logit.detach().clone(), prob=prob.detach().clone(), pooled=pooled.detach().clone(),
              loss=loss.detach().clone())
    if slice_only:
        fx["features_slice"] = stacked.detach()[0, 0, ::97, ::13].clone()
        fx["grad_x_slice"] = x[0].grad[:, :, ::3, ::5].clone()
    else:
        fx["features"] = stacked.detach().clone()
        fx["grad_features"] = stacked.grad.clone()
        fx["grad_x"] = [u.grad.clone() for u in x]
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "logit", logit.flatten().tolist(), "prob", prob.flatten().tolist())


def case_ops(M, name):
    g = torch.Generator().manual_seed(7)
    fx = {}
    # rope_apply (model.py:60-103) on a grid whose length is not a multiple of anything nice
    grid = (3, 5, 7)
    L = 3 * 5 * 7
    x = torch.randn(1, L + 4, 2, 128, generator=g)          # 4 pass-through pad tokens
    freqs = torch.cat([M.rope_params(1024, 128 - 4 * (128 // 6)), M.rope_params(1024, 2 * (128 // 6)),
                       M.rope_params(1024, 2 * (128 // 6))], dim=1)
    fx["rope_in"], fx["rope_grid"] = x, grid
    fx["rope_out"] = M.rope_apply(x, torch.tensor([grid]), freqs)
    # WanRMSNorm (model.py:106-122) fp32 and bf16 inputs
    xr = torch.randn(9, 256, generator=g)
    w = 1 + 0.1 * torch.randn(256, generator=g)
    n = M.WanRMSNorm(256, eps=1e-6)
    n.weight.data.copy_(w)
    fx["rms_in"], fx["rms_w"] = xr, w
    fx["rms_out"] = n(xr).detach()
    fx["rms_out_bf16in"] = n(xr.bfloat16()).detach()
    # WanLayerNorm (model.py:125-135)
    ln = M.WanLayerNorm(256, 1e-6, elementwise_affine=True)
    ln.weight.data.copy_(w)
    ln.bias.data.copy_(0.1 * w)
    fx["ln_out"] = ln(xr).detach()
    fx["ln_plain_out"] = M.WanLayerNorm(256, 1e-6)(xr).detach()
    # sinusoidal_embedding_1d (model.py:22-32)
    fx["sin_t"] = torch.tensor([0.0, 400.0, 999.0])
    fx["sin_out"] = M.sinusoidal_embedding_1d(256, fx["sin_t"])
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "ok")


def _a2a_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ref_shim.load()
    from diffusers_lite.utils import parallel_states as PS
    from diffusers_lite.utils import communication as C
    PS.initialize_sequence_parallel_state(world)
    g = torch.Generator().manual_seed(1234)
    full = torch.randn(1, 12 * world, 2 * world, 8, generator=g)              # [b, L, H, d]
    mine = full.chunk(world, dim=1)[rank].clone().requires_grad_(True)
    out = C.all_to_all_4D(mine, scatter_dim=2, gather_dim=1)                  # [b, L, H/P, d]
    back = C.all_to_all_4D(out, scatter_dim=1, gather_dim=2)
    (out * (rank + 1)).sum().backward()
    gathered = C.all_gather(mine.detach(), dim=1)
    q.put((rank, out.detach(), back.detach(), mine.grad.clone(), gathered))
    dist.barrier()
    dist.destroy_process_group()


def case_a2a(name, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_a2a_worker, args=(r, world, 29631, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    [p.join() for p in procs]
    fx = dict(world=world, out=[r[1] for r in res], back=[r[2] for r in res], grad=[r[3] for r in res],
              gathered=[r[4] for r in res])
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "ok")


def case_unipc(name):
    """the reference FlowUniPCMultistepScheduler (fm_solvers_unipc.py) driven the way train_step_refl does
    (train_prfl.py:631-735): set_timesteps(40, shift), m no-grad steps, one step with autograd through the model output,
    then the PRFL loss glue.  The 'model' is oracle.unipc_oracle.toy_velocity (seeded weights)."""
    from oracle.unipc_oracle import toy_velocity
    S = ref_shim.load_scheduler()
    g = torch.Generator().manual_seed(70)
    shape = (1, 8, 2, 4, 6)
    x_init = torch.randn(shape, generator=g)
    w = torch.randn(shape, generator=g) * 0.5
    fx = dict(shape=shape, seed=70, chains={})
    for steps, shift, solver_type, order in [(40, 3.0, "bh2", 2), (40, 5.0, "bh2", 2), (12, 3.0, "bh1", 3), (8, 1.0, "bh2", 1)]:
        sch = S.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False,
                                            solver_type=solver_type, solver_order=order)
        sch.set_timesteps(steps, device="cpu", shift=shift)
        x = x_init.clone()
        traj, x0s = [], []
        with torch.no_grad():
            for t in sch.timesteps:
                v = toy_velocity(x, t, w)
                x = sch.step(v, t, x, return_dict=False)[0]
                traj.append(x.clone())
                x0s.append(sch.model_outputs[-1].clone())
        fx["chains"][(steps, shift, solver_type, order)] = dict(timesteps=sch.timesteps.clone(), sigmas=sch.sigmas.clone(),
                                                                traj=traj, x0=x0s)
    # PRFL-shaped: m no-grad steps, then a differentiable step; gradient w.r.t. the toy model weight and the latent
    grads = {}
    for m in (0, 1, 7, 38):
        sch = S.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
        sch.set_timesteps(40, device="cpu", shift=3.0)
        x = x_init.clone()
        with torch.no_grad():
            for i in range(m):
                t = sch.timesteps[i]
                x = sch.step(toy_velocity(x, t, w), t, x, return_dict=False)[0]
        wg = w.clone().requires_grad_(True)
        xg = x.clone().requires_grad_(True)
        t_mid = sch.timesteps[m]
        prev = sch.step(toy_velocity(xg, t_mid, wg), t_mid, xg, return_dict=False)[0]
        reward = torch.tanh(prev.mean(dim=(1, 2, 3, 4)) * 5.0)        # stand-in for the reward model (|r| < 1: loss active)
        loss = 0.1 * torch.relu(-reward.squeeze() + 2).mean()       # train_prfl.py:796-798
        loss.backward()
        grads[m] = dict(prev=prev.detach().clone(), loss=loss.detach().clone(), grad_w=wg.grad.clone(), grad_x=xg.grad.clone())
    fx["prfl"] = grads
    # add_noise (fm_solvers_unipc.py:758-797): by timestep lookup, with a begin index, and after a step
    sch = S.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    sch.set_timesteps(40, device="cpu", shift=3.0)
    clean = torch.randn((3,) + shape[1:], generator=g)
    eps = torch.randn((3,) + shape[1:], generator=g)
    ts = sch.timesteps[[0, 17, 39]]
    an = {"by_timestep": sch.add_noise(clean, eps, ts).clone()}
    sch.set_begin_index(5)
    an["begin_index"] = sch.add_noise(clean, eps, ts).clone()
    sch.step(toy_velocity(x_init, sch.timesteps[5], w), sch.timesteps[5], x_init)
    an["after_step"] = sch.add_noise(clean, eps, ts).clone()
    fx["add_noise"] = an
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "ok", {k: float(v["loss"]) for k, v in grads.items()})


def case_unipc_fresh(name):
    """add_noise on a FRESHLY CONSTRUCTED scheduler (no set_timesteps): its schedule is the float tensor
    `sigmas * num_train_timesteps` (fm_solvers_unipc.py:96-103) whose entries share integer parts once shift > 1
    squeezes them together near sigma ~ 1, so the lookup must compare exactly (fm_solvers_unipc.py:628-641)."""
    S = ref_shim.load_scheduler()
    g = torch.Generator().manual_seed(71)
    fx = {"cases": {}}
    for shift in (1.0, 5.0):
        sch = S.FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=shift, use_dynamic_shifting=False)
        idx = [0, 1, 2, 3, 10, 500, 998, 999]
        ts = sch.timesteps[idx].clone()
        clean = torch.randn(len(idx), 4, 1, 2, 3, generator=g)
        eps = torch.randn(len(idx), 4, 1, 2, 3, generator=g)
        fx["cases"][shift] = dict(idx=idx, timesteps=ts, schedule=sch.timesteps.clone(), clean=clean, eps=eps,
                                  out=sch.add_noise(clean, eps, ts).clone())
    torch.save(fx, os.path.join(HERE, name + ".pt"))
    print(name, "ok", {k: int((v["schedule"].long().unique().numel())) for k, v in fx["cases"].items()}, "distinct integer parts of 1000")


def case_checkpoint(name):
    """run the UNMODIFIED reference save_checkpoint (diffusers_lite/utils/model_utils.py:70-126) on the
    tiny T2V model and record what it wrote (directory name, file list, tensor keys, config.json).  Shims: `peft` (unused
    import) stubbed; FSDP.state_dict_type -> null context, because a single un-wrapped process already holds the full state."""
    import contextlib
    import importlib.util
    import json
    import tempfile
    import types
    from safetensors import safe_open
    peft = types.ModuleType("peft")
    peft.get_peft_model_state_dict = lambda *a, **k: {}
    sys.modules.setdefault("peft", peft)
    pkg = types.ModuleType("_refutils")
    pkg.__path__ = [os.path.join(ref_shim.REF, "diffusers_lite/utils")]
    sys.modules["_refutils"] = pkg
    tu = types.ModuleType("_refutils.torch_utils")
    tu.set_logging = lambda *a, **k: None
    sys.modules["_refutils.torch_utils"] = tu
    spec = importlib.util.spec_from_file_location("_refutils.model_utils", os.path.join(ref_shim.REF, "diffusers_lite/utils/model_utils.py"))
    mu = importlib.util.module_from_spec(spec)
    sys.modules["_refutils.model_utils"] = mu
    spec.loader.exec_module(mu)
    mu.FSDP.state_dict_type = staticmethod(lambda *a, **k: contextlib.nullcontext())
    M, _ = ref_shim.load()
    cfg = synth.tiny_cfg("t2v")
    m = ref_model(M, cfg, synth.make_wan_state_dict(cfg, 7))
    m.config = cfg.kwargs() | {"dtype": "bf16"}
    with tempfile.TemporaryDirectory() as d:
        mu.save_checkpoint(m, 0, d, 7)
        sub = os.listdir(d)
        assert len(sub) == 1
        files = sorted(os.listdir(os.path.join(d, sub[0])))
        with safe_open(os.path.join(d, sub[0], "diffusion_pytorch_model.safetensors"), "pt") as f:
            keys = sorted(f.keys())
        config = json.load(open(os.path.join(d, sub[0], "config.json")))
        back = mu.load_state_dict(os.path.join(d, sub[0]))
        assert all(torch.equal(back[k], v) for k, v in m.state_dict().items())
    json.dump({"dir": sub[0], "files": files, "keys": keys, "config": config}, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
    print(name, sub[0], files, len(keys), "keys")


if __name__ == "__main__":
    assert ref_shim.available(), "needs the reference checkout"
    torch.set_num_threads(8)
    M, N = ref_shim.load()
    if len(sys.argv) > 1 and sys.argv[1] == "unipc":      # regenerate only the scheduler fixture
        case_unipc("unipc")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "unipc_fresh":
        case_unipc_fresh("unipc_fresh")
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "checkpoint":
        case_checkpoint("checkpoint_ref")
        sys.exit(0)
    case_unipc("unipc")
    case_unipc_fresh("unipc_fresh")
    case_checkpoint("checkpoint_ref")
    case_ops(M, "ops")
    case_model(M, "tiny_t2v", synth.tiny_cfg("t2v"), (5, 12, 20), 10, 11, [1, 2])
    case_model(M, "tiny_i2v", synth.tiny_cfg("i2v"), (3, 10, 14), 20, 21, [2])
    case_reward(M, N, "tiny_reward", synth.tiny_cfg("t2v", heads=2, layers=3), (5, 12, 20), 2, 30, 31)
    # BASELINE.json configs[0]: 1.3B architecture, 8 blocks, 17f x 240 x 416 -> latent 5 x 30 x 52
    case_reward(M, N, "cfg0_reward", synth.cfg_1_3b(layers=8), (5, 30, 52), 8, 40, 41, slice_only=True)
    case_a2a("a2a_gloo2")
    case_ragged(M, "tiny_t2v_ragged", synth.tiny_cfg("t2v"), [(5, 12, 20), (3, 10, 14)], 320, 60, 61)

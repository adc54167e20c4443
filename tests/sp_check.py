"""torchrun --nproc-per-node P tests/sp_check.py : Ulysses sequence-parallel forward + backward on P GPUs against
the oracle (tiny models) and the sharded optimizer against a dense one.  `run_checks()` is also what `bench.py` calls
during warm-up at N > 1 (the driver's GPU test box has one GPU), printing the numbers into its JSON line as `parity`.

Checks, on every rank:
  sp_fwd        noise prediction + gathered features vs the oracle (cos >= 0.999, max-rel <= 2e-2); reward logit
                (sp-local pooling == gathered pooling to 1e-5, vs oracle within 1e-2)
  sp_bwd        through the training-path exchanges (peer stores; NCCL all-to-all under PRFL_ULYSSES=nccl): the SUM over SP ranks of the partial weight / input gradients equals
                the oracle's SP=1 gradient, same tolerances
  sharded_adamw `ShardedAdamW` (resident bf16 weights, reduce-scattered fp32 gradient shards, 1/W fp32 masters,
                bf16 all-gather) vs dense torch.optim.AdamW on the replicated fp32 model, 2 steps with clip_grad_norm_(1.0):
                (i) the gradient shards, gathered back, equal the all-reduced(AVG) dense gradients to 1e-5 of each tensor's
                largest gradient (NCCL's reduce-scatter and all-reduce sum in different orders for W > 2); (ii) given the
                SAME reduced gradients the fp32 masters agree to max-rel <= 2e-6 per tensor after each step (the dense
                model is re-synchronised to the sharded masters in between) and the two models' outputs stay equal.  Feeding
                each side its own reduction instead lets Adam amplify the order-of-summation noise of this model's
                noise-level gradients (measured at W = 8: masters differ by 3e-5 with both sides correct)
  usp_fwd       (even world sizes) the same no-grad forward with the ranks arranged as Ulysses (world / 2) x Ring 2
                (`parallel.initialize_usp_state`): K / V blocks travel round the ring, partial results merged by
                `prfl_attn_merge`; vs the oracle, same tolerances
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))     # test infrastructure: the only place besides bench.py / smoke() that imports oracle/
from conftest import cos_rel  # noqa: E402
from oracle import synth  # noqa: E402
from oracle import wan_oracle as O  # noqa: E402

COS_MIN, REL_MAX, LOGIT_TOL, ADAMW_REL, GRAD_REL = 0.999, 2e-2, 1e-2, 2e-6, 1e-5


def run_checks(world: int, rank: int, verbose: bool = True) -> dict:
    """Requires an initialised NCCL process group + `parallel.initialize_sequence_parallel_state(world)`."""
    from prfl_b200 import parallel as _parallel
    from prfl_b200.model import WanModel
    from prfl_b200.pavrm import PavrmScorer
    from prfl_b200.sharding import ShardedAdamW
    dev = torch.device("cuda", torch.cuda.current_device())
    res = {"tolerance": {"cos_min": COS_MIN, "max_rel": REL_MAX, "logit_abs": LOGIT_TOL, "adamw_max_rel": ADAMW_REL, "grad_max_rel": GRAD_REL}}

    def say(*a):
        if verbose:
            print(f"[rank {rank}]", *a, flush=True)

    # heads must divide by P and tokens by P: a max(4, P)-head model on a latent with 192 tokens
    cfg = synth.tiny_cfg("t2v", heads=max(4, world), layers=2, ffn=768)
    sd = synth.make_wan_state_dict(cfg, 50)
    inp = synth.make_inputs(cfg, (4, 12, 16), 51)           # grid 4 x 6 x 8 = 192 tokens
    assert inp["seq_len"] % world == 0
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xr = [u.clone().requires_grad_(True) for u in inp["x"]]
    ref = O.wan_forward(sdr, cfg, xr, inp["t"], inp["context"], inp["seq_len"])
    g = torch.Generator().manual_seed(99)
    cot = [torch.randn(o.shape, generator=g) for o in ref]
    sum((o * c).sum() for o, c in zip(ref, cot)).backward()

    m = WanModel(**cfg.kwargs())
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).train()
    x = [u.to(dev).requires_grad_(True) for u in inp["x"]]
    ctx = [c.to(dev) for c in inp["context"]]
    out = m(x=x, t=inp["t"].to(dev), context=ctx, seq_len=inp["seq_len"])
    c_f, r_f = cos_rel(out[0].detach().cpu(), ref[0].detach())
    say(f"SP={world} forward vs oracle: cos={c_f:.6f} rel={r_f:.4f}")
    sum((o * cc.to(dev)).sum() for o, cc in zip(out, cot)).backward()
    # every rank back-propagates only its token chunk -> sum over ranks == full gradient
    worst_c, worst_r = 1.0, 0.0
    gx = x[0].grad.clone()
    dist.all_reduce(gx)
    c, r = cos_rel(gx.cpu(), xr[0].grad)
    worst_c, worst_r = min(worst_c, c), max(worst_r, r)
    say(f"sum-over-ranks grad_x: cos={c:.6f} rel={r:.4f}")
    params = dict(m.named_parameters())
    for k in ("blocks.0.self_attn.q.weight", "blocks.1.ffn.0.weight", "blocks.0.modulation", "blocks.1.cross_attn.v.weight",
              "blocks.0.self_attn.norm_k.weight", "blocks.1.self_attn.o.bias", "patch_embedding.weight"):
        gp = params[k].grad.clone()
        dist.all_reduce(gp)
        c, r = cos_rel(gp.cpu(), sdr[k].grad)
        worst_c, worst_r = min(worst_c, c), max(worst_r, r)
        if rank == 0:
            say(f"  sum-over-ranks grad {k}: cos={c:.6f} rel={r:.4f}")
    res["sp_bwd"] = {"cos": worst_c, "max_rel": worst_r, "ok": bool(worst_c >= COS_MIN and worst_r <= REL_MAX),
                     "what": "worst over grad_x + 7 weight grads, sum over SP ranks vs oracle SP=1 (training exchanges: "
                             + ("NCCL all-to-all" if _parallel._p2p_disabled else "peer stores into symmetric memory") + ")"}
    m.zero_grad(set_to_none=True)

    with torch.no_grad():
        feats = m(x=[u.to(dev) for u in inp["x"]], t=inp["t"].to(dev), context=ctx, seq_len=inp["seq_len"],
                  output_features=True, selected_layers=[2])
        rf = O.wan_forward(sd, cfg, inp["x"], inp["t"], inp["context"], inp["seq_len"], output_features=True, selected_layers=[2])
    c_g, r_g = cos_rel(feats[0].cpu(), rf[0])
    say(f"gathered features: cos={c_g:.6f} rel={r_g:.4f} shape={tuple(feats[0].shape)}")
    # reward scoring: pooled per rank + merged (QueryAttention sp_local) vs the oracle chain and vs the gathered path
    qa_sd, mlp_sd = synth.make_reward_state_dicts(cfg.dim, 52)
    scorer = PavrmScorer.from_state_dicts(cfg.kwargs(), sd, qa_sd, mlp_sd, num_blocks=2, device=dev)
    args = ([u.to(dev) for u in inp["x"]], inp["t"].to(dev), ctx, inp["seq_len"])
    logit_sp = scorer.score(*args)
    logit_g, _ = scorer.score(*args, return_features=True)
    with torch.no_grad():
        logit_o, _ = O.pavrm_reward(sd, cfg, qa_sd, mlp_sd, inp["x"], inp["t"], inp["context"], inp["seq_len"],
                                    selected_layers=(2,), num_blocks=2)
    say(f"reward logit: sp-local {float(logit_sp):.6f} gathered {float(logit_g):.6f} oracle {float(logit_o):.6f}")
    dl = abs(float(logit_sp) - float(logit_o))
    ok_f = (c_f >= COS_MIN and r_f <= REL_MAX and c_g >= COS_MIN and r_g <= REL_MAX and feats[0].shape == rf[0].shape
            and abs(float(logit_sp) - float(logit_g)) <= 1e-5 and dl <= LOGIT_TOL)
    res["sp_fwd"] = {"cos": min(c_f, c_g), "max_rel": max(r_f, r_g), "dlogit": dl, "ok": bool(ok_f),
                     "what": "noise prediction + gathered features + reward logit vs oracle (peer-store exchange in no-grad, NCCL in train)"}

    # ---- sharded optimizer vs dense ---------------------------------------------------------------------------
    def fresh():
        mm = WanModel(**cfg.kwargs())
        sd2 = dict(sd)
        gg = torch.Generator().manual_seed(7)
        sd2["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=gg) * 0.02   # non-zero head => grads reach the blocks
        mm.load_state_dict(sd2, strict=True)
        return mm.to(dev).train()

    a, b = fresh(), fresh()
    opt_a = ShardedAdamW(a, lr=1e-3, weight_decay=0.01).attach_hooks()                  # resident bf16 + sharded state (auto)
    assert opt_a.resident
    opt_b = torch.optim.AdamW(b.parameters(), lr=1e-3, weight_decay=0.01)
    names_b = dict(b.named_parameters())
    worst, worst_g = 0.0, 0.0
    for step in range(2):
        gi = torch.Generator().manual_seed(200 + step)
        xin = [torch.randn(inp["x"][0].shape, generator=gi).to(dev)]
        cot2 = torch.randn(ref[0].shape, generator=gi).to(dev)
        (a(x=xin, t=inp["t"].to(dev), context=ctx, seq_len=inp["seq_len"])[0] * cot2).sum().backward()
        (b(x=xin, t=inp["t"].to(dev), context=ctx, seq_len=inp["seq_len"])[0] * cot2).sum().backward()
        # (i) the reduce-scattered gradient shards, gathered back, against plain all-reduced(AVG) gradients of the dense model:
        #     checks the staging layout, the sink and the collectives (summation order differs between NCCL's reduce-scatter
        #     and all-reduce for W > 2, hence a tolerance relative to each tensor's largest gradient)
        shards = opt_a.reduce_gradients()
        for ui, u in enumerate(opt_a.units):
            full = torch.empty(u.shard * world, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(full, shards[ui].contiguous())
            for n, (o, cnt, shp) in u.offsets.items():
                p_b = names_b[(u.sink.prefix + n) if u.kind == "resident" else n]
                if p_b.grad is None:
                    p_b.grad = torch.zeros_like(p_b)             # FSDP flat-parameter semantics (unused parameters still decay)
                dist.all_reduce(p_b.grad, op=dist.ReduceOp.AVG)
                got = full[o:o + cnt].view(shp)
                worst_g = max(worst_g, float((got - p_b.grad.float()).abs().max() / (p_b.grad.float().abs().max() + 1e-30)))
                # (ii) hand the dense optimizer the SAME reduced gradient, so that the optimizer comparison below is not at the
                #      mercy of Adam dividing rounding-level differences of noise-level gradients by their own magnitude
                p_b.grad.copy_(got)
        opt_a.step(max_norm=1.0)
        torch.nn.utils.clip_grad_norm_(b.parameters(), 1.0)
        opt_b.step()
        opt_b.zero_grad(set_to_none=True)
        full = opt_a.full_state_dict(to_cpu=True)
        for k, v in b.state_dict().items():
            d = float((full[k].float() - v.detach().float().cpu()).abs().max() / (v.detach().float().abs().max().cpu() + 1e-12))
            worst = max(worst, d)
        # re-synchronise the dense model to the sharded masters (see tests/test_sharding_gpu.py: with identical weights both
        # sides compute the same gradients again and step 2 is as strict as step 1)
        b.load_state_dict({k: v.to(dev) for k, v in full.items()}, strict=True)
    with torch.no_grad():
        oa = a(x=xin, t=inp["t"].to(dev), context=ctx, seq_len=inp["seq_len"])[0]
        ob = b(x=xin, t=inp["t"].to(dev), context=ctx, seq_len=inp["seq_len"])[0]
    c_o, r_o = cos_rel(oa.cpu(), ob.cpu())
    say(f"ShardedAdamW (resident bf16, W={world}) vs dense AdamW, 2 steps: gradient shards vs all-reduce max-rel {worst_g:.2e}; "
        f"masters max-rel {worst:.2e}; outputs cos={c_o:.6f} rel={r_o:.2e}")
    res["sharded_adamw"] = {"grad_max_rel": worst_g, "master_max_rel": worst, "out_cos": c_o, "out_max_rel": r_o,
                            "ok": bool(worst_g <= GRAD_REL and worst <= ADAMW_REL and c_o >= 0.99999 and r_o <= 1e-3),
                            "what": "reduce-scattered fp32 gradient shards vs all-reduced(AVG) dense gradients; fp32 masters vs dense torch "
                                    "AdamW given the same reduced gradients, 2 steps, clip 1.0"}
    # ---- Ulysses x Ring (USP): ring degree 2 over the same ranks, no-grad forward vs the oracle ------
    keys = ["sp_fwd", "sp_bwd", "sharded_adamw"]
    if world % 2 == 0:
        from prfl_b200 import parallel
        U, R = world // 2, 2
        parallel.initialize_usp_state(U, R)
        try:
            with torch.no_grad():
                out_u = m(x=[u.to(dev) for u in inp["x"]], t=inp["t"].to(dev), context=ctx, seq_len=inp["seq_len"])
            c_u, r_u = cos_rel(out_u[0].cpu(), ref[0].detach())
        finally:
            parallel.initialize_sequence_parallel_state(world)          # back to plain Ulysses for the caller
        say(f"USP (ulysses {U} x ring {R}) forward vs oracle: cos={c_u:.6f} rel={r_u:.4f}")
        res["usp_fwd"] = {"ulysses": U, "ring": R, "cos": c_u, "max_rel": r_u, "ok": bool(c_u >= COS_MIN and r_u <= REL_MAX),
                          "what": "no-grad forward under Ulysses x Ring (K/V blocks round the ring, LSE merge kernel) vs oracle"}
        keys.append("usp_fwd")
    flag = torch.tensor([1.0 if all(res[k]["ok"] for k in keys) else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res["ok"] = bool(float(flag) == 1.0)
    return res


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    from prfl_b200 import parallel
    parallel.initialize_sequence_parallel_state(world)
    res = run_checks(world, rank)
    dist.barrier()
    dist.destroy_process_group()
    if not res["ok"]:
        print("SP CHECK FAILED", res)
        sys.exit(1)
    if rank == 0:
        import json
        print(json.dumps(res))
        print("SP CHECK OK")


if __name__ == "__main__":
    main()

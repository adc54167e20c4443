"""torchrun --nproc-per-node P tests/sp_check.py : Ulysses sequence-parallel forward + backward on P GPUs against
the reference goldens (tiny_t2v / tiny_i2v / tiny_reward-sized models) — the same fixtures the 1-GPU tests use.
Checks, on every rank: noise prediction, features; and that the SUM over SP ranks of the partial weight gradients
equals the reference's SP=1 gradient."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))     # test infrastructure: the only place besides bench.py / smoke() that imports oracle/
from conftest import cos_rel, golden  # noqa: E402
from oracle import synth  # noqa: E402
from oracle import wan_oracle as O  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl")
    from prfl_b200 import parallel
    from prfl_b200.model import WanModel
    parallel.initialize_sequence_parallel_state(world)
    ok = True
    # heads must divide by P and tokens by P: a 4-head model on a latent with 4*k tokens
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
    m = m.cuda().train()
    x = [u.cuda().requires_grad_(True) for u in inp["x"]]
    out = m(x=x, t=inp["t"].cuda(), context=[c.cuda() for c in inp["context"]], seq_len=inp["seq_len"])
    c, r = cos_rel(out[0].detach().cpu(), ref[0].detach())
    print(f"[rank {rank}] SP={world} forward vs oracle: cos={c:.6f} rel={r:.4f}")
    ok &= c >= 0.999 and r <= 2e-2
    sum((o * cc.cuda()).sum() for o, cc in zip(out, cot)).backward()
    # latents' grads: every rank back-propagates only its token chunk -> sum over ranks == full gradient
    gx = x[0].grad.clone()
    dist.all_reduce(gx)
    c, r = cos_rel(gx.cpu(), xr[0].grad)
    print(f"[rank {rank}] sum-over-ranks grad_x: cos={c:.6f} rel={r:.4f}")
    ok &= c >= 0.999 and r <= 2e-2
    for k in ("blocks.0.self_attn.q.weight", "blocks.1.ffn.0.weight", "blocks.0.modulation", "blocks.1.cross_attn.v.weight",
              "blocks.0.self_attn.norm_k.weight", "patch_embedding.weight"):
        gp = dict(m.named_parameters())[k].grad.clone()
        dist.all_reduce(gp)
        c, r = cos_rel(gp.cpu(), sdr[k].grad)
        if rank == 0:
            print(f"  sum-over-ranks grad {k}: cos={c:.6f} rel={r:.4f}")
        ok &= c >= 0.999 and r <= 2e-2
    with torch.no_grad():
        feats = m(x=[u.cuda() for u in inp["x"]], t=inp["t"].cuda(), context=[c.cuda() for c in inp["context"]],
                  seq_len=inp["seq_len"], output_features=True, selected_layers=[2])
        rf = O.wan_forward(sd, cfg, inp["x"], inp["t"], inp["context"], inp["seq_len"], output_features=True, selected_layers=[2])
    c, r = cos_rel(feats[0].cpu(), rf[0])
    print(f"[rank {rank}] gathered features: cos={c:.6f} rel={r:.4f} shape={tuple(feats[0].shape)}")
    ok &= c >= 0.999 and r <= 2e-2 and feats[0].shape == rf[0].shape
    # reward scoring: pooled per rank + merged (QueryAttention sp_local) vs the oracle chain and vs the gathered path
    from prfl_b200.pavrm import PavrmScorer
    qa_sd, mlp_sd = synth.make_reward_state_dicts(cfg.dim, 52)
    scorer = PavrmScorer.from_state_dicts(cfg.kwargs(), sd, qa_sd, mlp_sd, num_blocks=2)
    args = ([u.cuda() for u in inp["x"]], inp["t"].cuda(), [c.cuda() for c in inp["context"]], inp["seq_len"])
    logit_sp = scorer.score(*args)
    logit_g, _ = scorer.score(*args, return_features=True)
    with torch.no_grad():
        logit_o, _ = O.pavrm_reward(sd, cfg, qa_sd, mlp_sd, inp["x"], inp["t"], inp["context"], inp["seq_len"],
                                    selected_layers=(2,), num_blocks=2)
    print(f"[rank {rank}] reward logit: sp-local {float(logit_sp):.6f} gathered {float(logit_g):.6f} oracle {float(logit_o):.6f}")
    ok &= abs(float(logit_sp) - float(logit_g)) <= 1e-5 and abs(float(logit_sp) - float(logit_o)) <= 1e-2
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if float(t) != 1.0:
        print("SP CHECK FAILED")
        sys.exit(1)
    if rank == 0:
        print("SP CHECK OK")


if __name__ == "__main__":
    main()

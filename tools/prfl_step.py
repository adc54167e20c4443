"""A PRFL `train_step_refl`-shaped step (scripts/prfl/train_prfl.py:585-898 of the reference) on the B200 path, at a
depth that fits one GPU with fp32 master weights + fp32 gradients:

  1. `m` no-grad DiT forwards of the trainable video model (VGM, `--blocks` Wan-14B blocks)      [train_prfl.py:665-699]
  2. one forward WITH grad (per-block recompute in backward)                                      [:723-725]
  3. the differentiable FlowUniPC scheduler step (prfl_b200.scheduler, one fused kernel)           [:734]
  4. frozen reward model: 8-block Wan-14B features -> QueryAttention -> MLP -> loss 0.1*relu(2 - r) [:764-798]
  5. backward through 4 -> 3 -> 2 (dgrad only through the reward model; its weight grads are never used, Appendix B.10)

Prints one JSON object with per-phase milliseconds and algorithmic TFLOP/s (SURVEY.md Appendix A per-block numbers).
Not the bench.py headline (that is BASELINE configs[1]); numbers go to profiles/.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=8)
    ap.add_argument("--nograd-forwards", dest="m", type=int, default=2)
    ap.add_argument("--latent", default="21,60,104")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--i2v", action="store_true", help="image-to-video architecture (in_dim 36, CLIP tokens, y conditioning)")
    ap.add_argument("--profile", action="store_true", help="after timing, run one more step under torch.profiler and add a per-kernel table")
    ap.add_argument("--opt", action="store_true", help="also run the sharded-gradient AdamW step (reduce-scatter, clip, update, all-gather)")
    args = ap.parse_args()
    import torch.distributed as dist
    from prfl_b200 import _lib, parallel
    from prfl_b200.model import WanModel
    from prfl_b200.network import MLP, QueryAttention
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        parallel.initialize_sequence_parallel_state(world)
    fr, hh, ww = (int(v) for v in args.latent.split(","))
    L = fr * (hh // 2) * (ww // 2)
    torch.manual_seed(0)
    with torch.device(dev):
        mt, ind = ("i2v", 36) if args.i2v else ("t2v", 16)
        vgm = WanModel(model_type=mt, in_dim=ind, dim=5120, ffn_dim=13824, num_heads=40, num_layers=args.blocks)
        vgm.head.head.weight.data.normal_(0, 0.02)                  # the reference zero-inits it (model.py:729)
        lrm = WanModel(model_type=mt, in_dim=ind, dim=5120, ffn_dim=13824, num_heads=40, num_layers=8)
        lrm.head = None
        qa = QueryAttention(5120, 1, 8, dropout=0.0, return_type="query")
        mlp = MLP(5120)
    for mod in (lrm, qa, mlp):
        for p in mod.parameters():
            p.requires_grad_(False)                                 # frozen reward model: dgrad only
    vgm.train()
    opt = None
    if args.opt:
        from prfl_b200.sharding import ShardedAdamW
        # AdamW over transformer params only (train_prfl.py:482-491); gradients are reduce-scattered block by block inside backward
        opt = ShardedAdamW(vgm, lr=1e-6, weight_decay=0.0).attach_hooks()
    latent = torch.randn(16, fr, hh, ww, device=dev)
    ctx = [torch.randn(512, 4096, device=dev) * 0.08]
    extra = {}
    if args.i2v:
        mask = torch.zeros(4, fr, hh, ww, device=dev)
        mask[:, 0] = 1.0                                            # train_prfl.py:537-542
        extra = dict(clip_fea=torch.randn(1, 257, 1280, device=dev), y=[torch.cat([mask, torch.randn(16, fr, hh, ww, device=dev)])])
    from prfl_b200.prfl import refl_chain
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    sched = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)   # train_prfl.py:411-413
    cond = dict(image_embeds=extra.get("clip_fea"), latents_condition=torch.stack(extra["y"]) if args.i2v else None)

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def one_step():
        marks = {}
        loss, _ = refl_chain(vgm, lrm, qa, mlp, sched, latent[None], torch.stack(ctx), L, args.m, flow_shift=5.0,
                             feature_layer=[8], marks=marks, **cond)
        loss.backward()
        marks["bwd_done"] = ev()
        if opt is not None:
            opt.step(max_norm=1.0)                                  # clip_grad_norm_(1.0) + optimizer.step (train_prfl.py:825-830)
            marks["opt_done"] = ev()
        return marks, float(loss.detach())

    one_step()                                                      # warm-up (operand caches, allocator)
    vgm.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    rows = []
    _lib.launch_count_reset()
    for _ in range(args.steps):
        marks, loss = one_step()
        torch.cuda.synchronize()
        k = list(marks)
        rows.append({k[i + 1]: marks[k[i]].elapsed_time(marks[k[i + 1]]) for i in range(len(k) - 1)})
        vgm.zero_grad(set_to_none=True)
    avg = {k: sum(r[k] for r in rows) / len(rows) for k in rows[0]}
    # algorithmic FLOPs per block at this L (SURVEY.md Appendix A formulae)
    lin = 597.7e6 * L + 4 * 5120 * 5120 * 512
    att = 4.0 * L * L * 128 * 40 + 4.0 * L * (512 + (257 if args.i2v else 0)) * 128 * 40
    fwd_blk = (lin + att) / world
    bwd_blk = (2 * lin + 2.5 * att) / world
    nb = args.blocks
    out = {
        "workload": f"PRFL refl-shaped step ({'i2v' if args.i2v else 't2v'}), VGM {nb} blocks (14B dims) + frozen 8-block reward model, L={L}, m={args.m}, SP={world}",
        "ms": avg, "loss": loss, "gpu_launches_per_step": _lib.launch_count() / args.steps,
        "nograd_fwd_tflops": args.m * nb * fwd_blk / (avg["nograd_done"] * 1e-3) / 1e12 if args.m else None,
        "grad_fwd_tflops": nb * fwd_blk / (avg["grad_fwd_done"] * 1e-3) / 1e12,
        "lrm_fwd_tflops": 8 * fwd_blk / (avg["lrm_fwd_done"] * 1e-3) / 1e12,
        # backward = recompute fwd + bwd of the VGM blocks, and recompute + dgrad (here: full bwd kernels) of 8 LRM blocks
        "bwd_tflops_algorithmic": ((nb + 8) * (fwd_blk + bwd_blk)) / (avg["bwd_done"] * 1e-3) / 1e12,
        "step_ms": sum(avg.values()),
        "step_tokens_per_s": L * (args.m + 2) / (sum(avg.values()) * 1e-3),   # DiT forwards per step: m no-grad + 1 grad + 1 reward
        "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
    }
    if args.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            one_step()
            torch.cuda.synchronize()
        vgm.zero_grad(set_to_none=True)
        evs = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in evs)
        out["kernels"] = [{"name": e.key[:90], "launches": e.count, "total_ms": e.device_time_total / 1e3,
                           "share": e.device_time_total / tot} for e in evs[:25]]
        out["kernels_total_ms"] = tot / 1e3
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out, indent=1))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""A PRFL `train_step_refl`-shaped step (scripts/prfl/train_prfl.py:585-898 of the reference) on the B200 path:

  1. `m` no-grad DiT forwards of the trainable video model (VGM, `blocks` Wan-14B blocks)        [train_prfl.py:665-699]
  2. one forward WITH grad (per-block recompute in backward)                                      [:723-725]
  3. the differentiable FlowUniPC scheduler step (prfl_b200.scheduler, one fused kernel)           [:734]
  4. frozen reward model: 8-block Wan-14B features -> QueryAttention -> MLP -> loss 0.1*relu(2 - r) [:764-798]
  5. backward through 4 -> 3 -> 2 (dgrad only through the reward model; its weight grads are never used, Appendix B.10)
  6. clip_grad_norm_(1.0) + sharded AdamW step                                                     [:822-830]

`measure()` is what `bench.py` folds into its JSON line as `prfl_step` (the headline metric of BASELINE.json: PRFL train
s/step, 14B, 720P x 81 frames, sequence-parallel over the ranks of the job); `python tools/prfl_step.py` prints the same
object standalone (optionally with a torch.profiler per-kernel table).  Algorithmic FLOPs follow SURVEY.md Appendix A.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LATENT_720P = (21, 90, 160)      # 81 frames x 720 x 1280 -> tokens = 21 * 45 * 80 = 75 600
LATENT_480P = (21, 60, 104)


def algorithmic_flops(L: int, i2v: bool, world: int):
    """Per-rank algorithmic FLOPs of one 14B block forward / backward at L tokens."""
    lin = 597.7e6 * L + 4 * 5120 * 5120 * (512 + (257 if i2v else 0))
    att = 4.0 * L * L * 128 * 40 + 4.0 * L * (512 + (257 if i2v else 0)) * 128 * 40
    return (lin + att) / world, (2 * lin + 2.5 * att) / world


def fit_blocks(world: int, L: int, want: int = 40, budget_gb: float = 135.0) -> int:
    """Largest VGM depth (<= want) whose training state fits `budget_gb` per GPU: bf16 resident weights + 1/W fp32
    master / moment shards + the per-block saved fp32 inputs (activation checkpointing) + one block's recompute stash."""
    per_block_params = 351.4e6 + 52.4e6
    M = L / world
    fixed = 8 * per_block_params * 2 + 2 * per_block_params * 4 + 50 * M * 5120 * 2 + 8e9       # reward model, grad buffers, one block's stash + backward temporaries, slack
    # bf16 resident + (master, m, v, grad shard) fp32 / W + saved fp32 block input + saved bf16 attention output (selective ckpt)
    per_block = per_block_params * (2 + 16.0 / world) + M * 5120 * 4 + M * 5120 * 2
    n = int((budget_gb * 1e9 - fixed) // per_block)
    return max(1, min(want, n))


def extrapolate_m(runs: dict) -> dict:
    """The reference draws m = mid_timestep uniformly from [0, 38) (train_prfl.py:640, E[m] = 19; BASELINE configs[2] asks for
    m in {0, 19, 38}).  The m no-grad forwards are identical, independent DiT passes, so s/step is affine in m: from two measured
    values the others follow.  Pure host arithmetic on measured numbers; {} unless two different m were timed."""
    ms = sorted((int(k[1:]), v["s_per_step"]) for k, v in runs.items() if k[:1] == "m" and k[1:].isdigit())
    if len(ms) < 2 or ms[0][0] == ms[-1][0]:
        return {}
    (m0, s0), (m1, s1) = ms[0], ms[-1]
    per = (s1 - s0) / (m1 - m0)
    return {"extrapolated_s_per_step": {
        "per_nograd_forward_s": per, **{f"m{m}": s0 + (m - m0) * per for m in (0, 19, 38)},
        "how": f"affine in m through the measured m{m0} and m{m1}: every no-grad forward is the same DiT pass; m19 = E[m] of the "
               "reference's randint(0, 38) (train_prfl.py:640)"}}


def measure(blocks: int, m_list=(2,), latent=LATENT_720P, steps: int = 2, i2v: bool = True, opt: bool = True,
            profile: bool = False, legacy_fp32: bool = False, warmup: int = 1):
    """Build the VGM (`blocks` 14B blocks) + frozen reward model on the current device / process group, run the step for
    every m in m_list and return a dict (rank-independent: times are max over ranks).  Frees everything on return."""
    import torch.distributed as dist
    from prfl_b200 import _lib, engine, sharding
    from prfl_b200.model import WanModel
    from prfl_b200.network import MLP, QueryAttention
    from prfl_b200.prfl import refl_chain
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = torch.device("cuda", torch.cuda.current_device())
    fr, hh, ww = latent
    L = fr * (hh // 2) * (ww // 2)
    torch.manual_seed(0)
    mt, ind = ("i2v", 36) if i2v else ("t2v", 16)

    torch.cuda.reset_peak_memory_stats()
    vgm = sharding.build_wan(mt, ind, blocks, dev, resident_bf16=not legacy_fp32, head=True)
    vgm.head.head.weight.data.normal_(0, 0.02)                      # the reference zero-inits it (model.py:729)
    lrm = sharding.build_wan(mt, ind, 8, dev, resident_bf16=not legacy_fp32, head=False, frozen=True)
    with torch.device(dev):
        qa = QueryAttention(5120, 1, 8, dropout=0.0, return_type="query")
        mlp = MLP(5120)
    for mod in (lrm, qa, mlp):
        for p in mod.parameters():
            p.requires_grad_(False)                                 # frozen reward model: dgrad only
    vgm.train()
    optim = None
    if opt:
        # AdamW over transformer params only (train_prfl.py:482-491); gradients are reduce-scattered block by block inside backward
        optim = sharding.ShardedAdamW(vgm, lr=1e-6, weight_decay=0.0, resident_bf16=not legacy_fp32).attach_hooks()
    g = torch.Generator(device=dev).manual_seed(1)
    lat = torch.randn(16, fr, hh, ww, device=dev, generator=g)
    ctx = [torch.randn(512, 4096, device=dev, generator=g) * 0.08]
    extra = {}
    if i2v:
        mask = torch.zeros(4, fr, hh, ww, device=dev)
        mask[:, 0] = 1.0                                            # train_prfl.py:537-542
        extra = dict(clip_fea=torch.randn(1, 257, 1280, device=dev, generator=g),
                     y=[torch.cat([mask, torch.randn(16, fr, hh, ww, device=dev, generator=g)])])
    sched = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)   # train_prfl.py:411-413
    cond = dict(image_embeds=extra.get("clip_fea"), latents_condition=torch.stack(extra["y"]) if i2v else None)

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def one_step(m):
        marks = {}
        loss, _ = refl_chain(vgm, lrm, qa, mlp, sched, lat[None], torch.stack(ctx), L, m, flow_shift=5.0,
                             feature_layer=[8], marks=marks, **cond)
        loss.backward()
        marks["bwd_done"] = ev()
        if optim is not None:
            optim.step(max_norm=1.0)                                # clip_grad_norm_(1.0) + optimizer.step (train_prfl.py:825-830)
            marks["opt_done"] = ev()
        else:
            vgm.zero_grad(set_to_none=True)
        return marks, loss.detach()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fwd_blk, bwd_blk = algorithmic_flops(L, i2v, world)
    pk = 1370.9
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", pk)
    except Exception:
        pass
    out = {"workload": f"PRFL {'I2V' if i2v else 'T2V'} {'720P' if latent == LATENT_720P else 'x'.join(map(str, latent))}x81f train step "
                       f"(train_step_refl), Wan2.1-14B dims, L={L}", "blocks": blocks, "reward_blocks": 8, "sp": world, "L": L,
           "optimizer": "ShardedAdamW (fp32 master/moment shards 1/W, bf16 resident weights)" if opt and not legacy_fp32 else
                        ("ShardedAdamW (legacy fp32 replicated params)" if opt else "none (fwd+bwd only)"),
           "checkpointing": "selective: per block the fp32 input + the bf16 self-attention output / LSE are kept, everything else is recomputed "
                            "(FLOPs below count what is executed)" if engine.SAVE_ATTENTION else
                            "full: only the fp32 block input is kept (the reference's per-block activation checkpointing)",
           "runs": {}}
    if world > 1:
        from prfl_b200 import parallel
        out["ulysses_exchange"] = "NCCL all_to_all_single (PRFL_ULYSSES=nccl)" if parallel._p2p_disabled else "peer stores into symmetric memory"
        out["gradient_reduce_scatter"] = ("overlapped with the next block's backward on a side stream" if optim is not None and optim.overlap
                                          else "on the compute stream (PRFL_RS=serial)")
    if optim is not None:
        sync()
        optim.warmup_collectives()                                  # big buffers + first-use NCCL setup at a quiescent point
        sync()
    for _ in range(warmup):
        one_step(max(m_list))                                       # warm-up (operand caches, allocator, NCCL)
    sync()
    for m in m_list:
        rows, losses = [], []
        _lib.launch_count_reset()
        for _ in range(steps):
            sync()
            marks, loss = one_step(m)
            sync()
            k = list(marks)
            row = {k[i + 1]: marks[k[i]].elapsed_time(marks[k[i + 1]]) for i in range(len(k) - 1)}
            if world > 1:                                           # device-timed, max over ranks per phase
                t = torch.tensor(list(row.values()), device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                row = dict(zip(row, t.tolist()))
            rows.append(row)
            losses.append(float(loss))
        avg = {k: sum(r[k] for r in rows) / len(rows) for k in rows[0]}
        step_ms = sum(avg.values())
        # algorithmic work per rank: m + 1 VGM forwards, 1 reward forward, backward = recompute + bwd of the VGM blocks and
        # recompute + dgrad-only (2/3 of the linear backward) of the 8 reward blocks
        lin_f = (597.7e6 * L + 4 * 5120 * 5120 * (512 + (257 if i2v else 0))) / world
        dgrad_only_blk = bwd_blk - lin_f                            # no weight-gradient GEMMs for the frozen reward blocks
        # selective checkpointing keeps the self-attention output: the recompute does not re-run that kernel, so it is not counted
        recompute_blk = fwd_blk - (4.0 * L * L * 128 * 40 / world if engine.SAVE_ATTENTION else 0.0)
        flops = ((m + 1) * blocks + 8) * fwd_blk + blocks * (recompute_blk + bwd_blk) + 8 * (recompute_blk + dgrad_only_blk)
        out["runs"][f"m{m}"] = {
            "s_per_step": step_ms / 1e3, "ms": avg, "loss": losses[-1],
            "dit_tokens_per_s": L * (m + 2) / (step_ms * 1e-3),    # DiT forwards per step: m no-grad + 1 grad + 1 reward
            "gpu_launches_per_step": _lib.launch_count() / steps,
            "nograd_fwd_tflops_per_gpu": m * blocks * fwd_blk / (avg["nograd_done"] * 1e-3) / 1e12 if m else None,
            "grad_fwd_tflops_per_gpu": blocks * fwd_blk / (avg["grad_fwd_done"] * 1e-3) / 1e12,
            "bwd_tflops_per_gpu": (blocks * (recompute_blk + bwd_blk) + 8 * (recompute_blk + dgrad_only_blk)) / (avg["bwd_done"] * 1e-3) / 1e12,
            "executed_pflop_per_gpu": flops / 1e15,
            "step_tflops_per_gpu": flops / (step_ms * 1e-3) / 1e12,
            "frac_of_sustained_bf16_peak": flops / (step_ms * 1e-3) / 1e12 / pk,
        }
    try:
        out.update(extrapolate_m(out["runs"]))
    except Exception:                                               # bookkeeping must never cost the measurement
        pass
    peak = torch.tensor([torch.cuda.max_memory_allocated() / 1e9], device=dev)
    if world > 1:
        dist.all_reduce(peak, op=dist.ReduceOp.MAX)
    out["peak_mem_gb"] = float(peak)
    if world > 1 and parallel._p2p_disabled:                        # symmetric memory could not be set up: the run fell back
        out["ulysses_exchange"] = "NCCL all_to_all_single"
    if profile:
        from torch.profiler import ProfilerActivity, profile as tprofile
        with tprofile(activities=[ProfilerActivity.CUDA]) as prof:
            one_step(max(m_list))
            torch.cuda.synchronize()
        evs = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
        tot = sum(e.device_time_total for e in evs)
        out["kernels"] = [{"name": e.key[:90], "launches": e.count, "total_ms": e.device_time_total / 1e3,
                           "share": e.device_time_total / tot} for e in evs[:30]]
        out["kernels_total_ms"] = tot / 1e3
    del vgm, lrm, qa, mlp, optim, lat, ctx, extra, cond
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=0, help="VGM depth (0 = the largest that fits this world size)")
    ap.add_argument("--nograd", dest="m", default="2", help="comma-separated numbers of no-grad forwards (mid_timestep) to time")
    ap.add_argument("--latent", default="21,90,160")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--t2v", action="store_true", help="text-to-video architecture (default: I2V, in_dim 36 + CLIP tokens + y)")
    ap.add_argument("--profile", action="store_true", help="one more step under torch.profiler -> per-kernel table")
    ap.add_argument("--no-opt", action="store_true")
    ap.add_argument("--legacy-fp32", action="store_true", help="round-1 layout: replicated fp32 parameters + bf16 operand caches")
    ap.add_argument("--out", default="", help="rank 0 also writes the result object (one JSON document) to this file, atomically")
    args = ap.parse_args()
    parent = os.environ.get("PRFL_CHILD_OF")
    if parent:                                                      # started by bench.py: never outlive the bench process (an orphan would keep the GPU)
        import ctypes
        import signal
        try:
            ctypes.CDLL("libc.so.6", use_errno=True).prctl(1, int(signal.SIGKILL), 0, 0, 0)     # PR_SET_PDEATHSIG
        except OSError:
            pass
        if os.getppid() != int(parent):
            sys.exit(4)
    import torch.distributed as dist
    from prfl_b200 import parallel
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        parallel.initialize_sequence_parallel_state(world)
    latent = tuple(int(v) for v in args.latent.split(","))
    L = latent[0] * (latent[1] // 2) * (latent[2] // 2)
    blocks = args.blocks or fit_blocks(world, L)
    out = measure(blocks, tuple(int(v) for v in args.m.split(",")), latent, args.steps, not args.t2v, not args.no_opt,
                  args.profile, args.legacy_fp32)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out, indent=1))
        if args.out:
            with open(args.out + ".tmp", "w") as f:
                json.dump(out, f)
            os.replace(args.out + ".tmp", args.out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Sampling s/step of the CFG denoising loop (text2video.py:283-304) on the B200 path: Wan2.1-14B
architecture, 40 blocks, 480P x 81 frames (L = 32 760) or 720P (L = 75 600), conditional + unconditional forward per step
(batched as B = 2 or sequential), prompt K/V cached, guidance fused into the scheduler kernel.  Random-init bf16-resident
weights (28 GB).  Usage: [torchrun --nproc-per-node N] python tools/sample_step.py [--latent 21,60,104] [--steps 4] [--i2v]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--latent", default="21,60,104")
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--blocks", type=int, default=40)
    ap.add_argument("--i2v", action="store_true")
    args = ap.parse_args()
    import torch.distributed as dist
    from prfl_b200 import parallel, sharding
    from prfl_b200.sampling import sample_loop
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
    mt, ind = ("i2v", 36) if args.i2v else ("t2v", 16)
    model = sharding.build_wan(mt, ind, args.blocks, dev, frozen=True).eval()
    model.head.head.weight.data.normal_(0, 0.02)
    g = torch.Generator(device=dev).manual_seed(1)
    noise = torch.randn(16, fr, hh, ww, device=dev, generator=g)
    ctx = [torch.randn(512, 4096, device=dev, generator=g) * 0.08]
    ctx_null = [torch.randn(126, 4096, device=dev, generator=g) * 0.08]
    extra = {}
    if args.i2v:
        mask = torch.zeros(4, fr, hh, ww, device=dev)
        mask[:, 0] = 1.0
        extra = dict(clip_fea=torch.randn(1, 257, 1280, device=dev, generator=g), y=[torch.cat([mask, torch.randn(16, fr, hh, ww, device=dev, generator=g)])])
    out = {"workload": f"CFG sampling loop, Wan2.1-14B dims, {args.blocks} blocks, {mt}, L={L}, SP={world}", "runs": {}}
    for name, batch in (("batched_b2", True), ("sequential", False)):
        sample_loop(model, noise, ctx, ctx_null, L, sampling_steps=2, batch_cfg=batch, **extra)      # warm-up (also fills the K/V cache path)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        sample_loop(model, noise, ctx, ctx_null, L, sampling_steps=args.steps, batch_cfg=batch, **extra)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        flops = 2 * args.blocks * (597.7e6 * L + 4.0 * L * L * 128 * 40 + 4.0 * L * 512 * 128 * 40) / world
        out["runs"][name] = {"s_per_step": ms / args.steps / 1e3, "dit_tokens_per_s": 2 * L * args.steps / (ms * 1e-3),
                             "tflops_per_gpu": flops * args.steps / (ms * 1e-3) / 1e12}
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out, indent=1))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

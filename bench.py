#!/usr/bin/env python
"""bench.py — the hot path of HY-Video-PRFL on B200: PAVRM latent reward scoring with the Wan2.1-14B
architecture (BASELINE.json configs[1]; the largest single-GPU configuration of the north-star path).

A "step" = one scoring pass: patchify -> 8 Wan-DiT blocks (14B dims) over L = 32 760 video tokens
(480P x 81 frames) -> single-query reward attention -> MLP -> reward logit.
  value : DiT tokens/s (L x forwards / time) with the inputs already resident in HBM
  e2e   : the same through the public call with HOST (pinned) inputs: H2D copy of latents / text states / t
          and D2H read of the logit inside the timed region
  roofline : the dominant kernel (self-attention forward, tcgen05): algorithmic 4*L^2*128*40 flops per launch
             / its live CUDA-event duration, against the measured bf16 peak in MEASURED_PEAKS.json
  cpu_baseline : the oracle port (torch CPU fp32) timed on this box's host cores on a bounded sample
N > 1 (torchrun): the same sample, tokens sharded over N ranks with Ulysses sequence parallelism ("strong").
`--impl reference` times the reference algorithm's CPU port (oracle/) instead — reported baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

LATENT_480P = (21, 60, 104)      # 81 frames x 480 x 832 -> VAE latent; tokens = 21 * 30 * 52 = 32 760
NUM_BLOCKS = 8                   # lrm.trainable_blocks [0..7], feature_layer [8]
HEADS, HD = 40, 128
WORKLOAD = "PAVRM T2V 480Px81f reward scoring, Wan2.1-14B arch (dim 5120, ffn 13824, 40 heads), 8 blocks + reward head, L=32760, batch 1"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sust=d.get("bf16_tflops_sustained", 1400.0), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (ts, r) in self.rows if t0 <= ts <= t1 + 0.2 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------
def cpu_sample(threads=None, blocks=NUM_BLOCKS, latent=(3, 30, 52), repeats=1):
    """The oracle (CPU restatement of the reference path) on a bounded sample of the same workload:
    14B architecture, `blocks` blocks + reward head, a short clip (latent 3x30x52 -> 1170 tokens)."""
    from oracle import synth
    from oracle import wan_oracle as O
    if threads:
        torch.set_num_threads(threads)
    cfg = synth.cfg_14b("t2v", layers=blocks)
    g = torch.Generator().manual_seed(0)
    sd, by_shape = {}, {}
    for k, v in _shapes_14b(cfg).items():       # cheap init: values do not matter for timing, so tensors of one
        if len(v) > 1:                            # shape share storage (reads still stream the full matrix per use)
            if v not in by_shape:
                by_shape[v] = torch.empty(v).uniform_(-0.02, 0.02, generator=g)
            sd[k] = by_shape[v]
        else:
            sd[k] = torch.zeros(v)
        if k.endswith("norm_q.weight") or k.endswith("norm_k.weight") or k.endswith("norm3.weight"):
            sd[k] = torch.ones(v)
    qa, mlp = synth.make_reward_state_dicts(cfg.dim, 1)
    inp = synth.make_inputs(cfg, latent, 2, text_tokens=512)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            logit, _ = O.pavrm_reward(sd, cfg, qa, mlp, inp["x"], inp["t"], inp["context"], inp["seq_len"],
                                      selected_layers=(blocks,), num_blocks=blocks)
            times.append(time.perf_counter() - t0)
    return inp["seq_len"], times, float(logit)


def _shapes_14b(cfg):
    from oracle import synth
    tiny = synth.WanConfig(**{**cfg.kwargs(), "num_layers": 1})
    # shapes from a 1-layer dict, replicated per block, without materialising random weights twice
    d, f = cfg.dim, cfg.ffn_dim
    shp = {"patch_embedding.weight": (d, cfg.in_dim, 1, 2, 2), "patch_embedding.bias": (d,),
           "text_embedding.0.weight": (d, cfg.text_dim), "text_embedding.0.bias": (d,), "text_embedding.2.weight": (d, d),
           "text_embedding.2.bias": (d,), "time_embedding.0.weight": (d, cfg.freq_dim), "time_embedding.0.bias": (d,),
           "time_embedding.2.weight": (d, d), "time_embedding.2.bias": (d,), "time_projection.1.weight": (6 * d, d),
           "time_projection.1.bias": (6 * d,)}
    for l in range(cfg.num_layers):
        p = f"blocks.{l}."
        for a in ("self_attn", "cross_attn"):
            for nm in ("q", "k", "v", "o"):
                shp[p + f"{a}.{nm}.weight"], shp[p + f"{a}.{nm}.bias"] = (d, d), (d,)
            shp[p + f"{a}.norm_q.weight"] = shp[p + f"{a}.norm_k.weight"] = (d,)
        shp[p + "norm3.weight"] = shp[p + "norm3.bias"] = (d,)
        shp[p + "ffn.0.weight"], shp[p + "ffn.0.bias"] = (f, d), (f,)
        shp[p + "ffn.2.weight"], shp[p + "ffn.2.bias"] = (d, f), (d,)
        shp[p + "modulation"] = (1, 6, d)
    del tiny
    return shp


def run_reference(args):
    """`--impl reference`: the reference algorithm's CPU implementation (oracle port — the reference is
    Python/PyTorch and is not present on the GPU box) with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    L, _, _ = cpu_sample(threads, repeats=max(1, min(args.warmup, 1)))
    L, times, _ = cpu_sample(threads, repeats=max(1, args.steps))
    ms = 1e3 * sum(times) / len(times)
    val = L / (ms / 1e3)
    sample = f"14B arch, {NUM_BLOCKS} blocks + reward head, fp32, latent 16x3x30x52 -> {L} tokens (attention cost grows with L^2: a short clip flatters the CPU)"
    print(json.dumps({"impl": "reference", "metric": "dit_tokens_per_s", "value": val, "unit": "tokens/s", "n_gpus": args.gpus,
                      "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
                      "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from prfl_b200 import _lib, ops, parallel
    from prfl_b200.model import WanModel
    from prfl_b200.network import MLP, QueryAttention
    from prfl_b200.pavrm import PavrmScorer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a B200 (no CPU fallback); use --impl reference for the CPU port"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        parallel.initialize_sequence_parallel_state(world)
    assert world == args.gpus or world == 1

    torch.manual_seed(0)
    with torch.device(dev):
        m = WanModel(model_type="t2v", dim=5120, ffn_dim=13824, num_heads=HEADS, num_layers=NUM_BLOCKS)
        for blk in m.blocks:                                      # random init incl. non-trivial norms
            blk.norm3.weight.data.normal_(1.0, 0.1)
        m.head = None
        qa = QueryAttention(5120, num_queries=1, num_heads=8, dropout=0.0, return_type="query")
        mlp = MLP(5120)
    scorer = PavrmScorer(m, qa, mlp, NUM_BLOCKS).to(dev).eval()

    fr, hh, ww = LATENT_480P
    L = fr * (hh // 2) * (ww // 2)
    g = torch.Generator().manual_seed(1)
    x_host = torch.randn(16, fr, hh, ww, generator=g).pin_memory()
    ctx_host = (torch.randn(512, 4096, generator=g) * 0.08).pin_memory()
    t_host = torch.tensor([400.0]).pin_memory()
    x_dev, ctx_dev, t_dev = x_host.to(dev), ctx_host.to(dev), t_host.to(dev)
    h2d = x_host.numel() * 4 + ctx_host.numel() * 4 + 4
    logit_host = torch.empty(1, 1, 1).pin_memory()

    def step_resident():
        return scorer.score([x_dev], t_dev, [ctx_dev], L)

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        cd = ctx_host.to(dev, non_blocking=True)
        td = t_host.to(dev, non_blocking=True)
        logit_host.copy_(scorer.score([xd], td, [cd], L), non_blocking=True)
        return logit_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, timer=None):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.TIMER = timer
        t0 = time.time()
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        t1 = time.time()
        ops.TIMER = None
        ms = a.elapsed_time(b)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt)
        return ms, t0, t1

    for _ in range(max(args.warmup, 3)):
        step_resident()
    step_e2e()
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.launch_count_reset()
    timer = ops.KernelTimer(["attn_fwd_self"])
    ms, t0, t1 = timed(step_resident, args.steps, timer)
    launches = _lib.launch_count()
    clocks = sampler.stop(t0, t1) if sampler else None
    ms_e2e, _, _ = timed(step_e2e, args.steps)
    logit = float(logit_host)

    if rank == 0:
        pk = peaks()
        attn_ms = timer.elapsed_ms("attn_fwd_self")
        attn_avg = sum(attn_ms) / max(1, len(attn_ms))
        heads_local = HEADS // world
        attn_flops = 4.0 * L * L * HD * heads_local                      # algorithmic, per launch
        achieved = attn_flops / (attn_avg * 1e-3) / 1e12 if attn_avg > 0 else 0.0
        per_step = ms / args.steps
        line = {
            "metric": "dit_tokens_per_s", "value": L / (per_step * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "parallelism": f"ulysses_sp{world}", "weights": "random init",
                       "l2": "per-step working set (5.6 GB bf16 weights + >2 GB activations) >> 126 MB L2; no explicit flush",
                       "reward_logit": logit},
            "e2e": {"value": L / (ms_e2e / args.steps * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "attn_fwd_kernel (self-attention fwd, tcgen05)", "bound": "tensor", "achieved": achieved,
                         "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"],
                         # DRAM bytes per launch from the committed `ncu --set full` capture of this kernel at this shape
                         # (profiles/r01_ncu_full_v2.csv: 1.049 GB read + 0.462 GB written; algorithmic Q,K,V,O = 1.342 GB;
                         # the first capture of the round, r01_ncu_full_attn_gemm.csv, read 1.031 + 0.322 GB)
                         "traffic": 1.511e9 if world == 1 else None,
                         "peak_source": pk["src"] + " sustained bf16", "launch_ms": attn_avg, "launches_timed": len(attn_ms),
                         "share_of_step": attn_avg * len(attn_ms) / max(ms, 1e-9)},
            "step_tflops": (8 * 41.96e12 + 3.4e12 * 0) / world / (per_step * 1e-3) / 1e12,
        }
        if world == 1 and not args.no_cpu:
            try:
                cores = min(os.cpu_count() or 1, 64)
                Ls, times, _ = cpu_sample(cores)
                line["cpu_baseline"] = {"value": Ls / times[0], "unit": "tokens/s", "cores": cores, "kind": "port",
                                        "sample": f"oracle (torch CPU fp32), 14B arch, {NUM_BLOCKS} blocks + reward head, latent 16x3x30x52 -> {Ls} tokens, 1 pass = {times[0]:.1f} s"}
            except Exception as e:  # the GPU line must not be lost to a host-side problem
                line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

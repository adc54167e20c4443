#!/usr/bin/env python
"""bench.py — the hot path of HY-Video-PRFL on B200.

Top-level line (same contract as round 1, so rounds compare): PAVRM latent reward scoring with the Wan2.1-14B
architecture (BASELINE.json configs[1], the configuration labelled "1 B200").  A "step" = one scoring pass: patchify ->
8 Wan-DiT blocks (14B dims) over L = 32 760 video tokens (480P x 81 frames) -> single-query reward attention -> MLP ->
reward logit.
  value    : DiT tokens/s (L x forwards / time) with the inputs already resident in HBM
  e2e      : the same through the public call with HOST (pinned) inputs: H2D copy of latents / text states / t and D2H
             read of the logit inside the timed region
  roofline : the dominant kernel (self-attention forward, tcgen05): algorithmic 4*L^2*128*heads flops per launch / its
             live CUDA-event duration, against the measured sustained bf16 peak in MEASURED_PEAKS.json; `traffic` is read
             from the committed `ncu --set full` capture under profiles/
  parity   : driver-visible numerics (a) N = 1: the SAME weights on the CPU oracle (fp32) and on the GPU, 14B dims
             (5120 / 40 heads / ffn 13 824), 8 blocks + reward head, on the bounded sample the CPU leg times:
             features cos / max-rel, |d logit|; (b) N > 1: Ulysses forward, sum-over-ranks gradients through the NCCL
             all-to-all path and ShardedAdamW vs dense AdamW (tests/sp_check.py) + the same-weights check of (a) under SP
  prfl_step: the HEADLINE metric of BASELINE.json — "PRFL train s/step & DiT tokens/s, 14B 720P x 81f, 1/2/4/8 B200":
             train_step_refl-shaped step (tools/prfl_step.py) at L = 75 600, I2V architecture, Ulysses SP over the N ranks,
             sharded fp32 master / AdamW state, for m in {0, 2} no-grad denoising forwards; 40 blocks where the training
             state fits (N = 8), else the largest depth that fits with headroom, stated.  At N > 1 the leg runs in CHILD
             processes with their own process group (`prfl_step_in_children`): a cross-rank hang cannot be cancelled from
             inside a process, so the parents put a deadline on the children, kill exactly those PIDs, agree on the outcome
             over their own (idle, healthy) group; after a hang the second and last attempt runs in the bench processes in a
             conservative mode (NCCL all-to-all exchange, no side-stream overlap) under the leg watchdog, after a crash in the
             default mode; `prfl_step.attempts` records what happened
  cpu_baseline : the oracle port (torch CPU fp32) timed on this box's host cores on the bounded sample (+ the GPU's
             throughput on that same sample, `gpu_same_sample`, for a like-for-like ratio)
  gpu_baseline : the same oracle port run on THIS GPU with eager PyTorch kernels — bf16 `F.linear` (cuBLAS), flash-attn 2
             (`flash_attn_func`), ATen norms / fp64 RoPE — i.e. the reference's own kernel stack on the
             full 32 760-token workload; `speedup` = ours / eager
N > 1 (torchrun): one sample, tokens sharded over N ranks with Ulysses sequence parallelism ("strong").
`--impl reference` times the reference algorithm's CPU port (oracle/) instead — reported baseline.
"""
import argparse
import glob
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
SAMPLE_LATENT = (3, 32, 52)      # bounded sample for the CPU legs / same-weights parity: 3 * 16 * 26 = 1 248 tokens (divisible by 8)
NUM_BLOCKS = 8                   # lrm.trainable_blocks [0..7], feature_layer [8]
HEADS, HD = 40, 128
WORKLOAD = "PAVRM T2V 480Px81f reward scoring, Wan2.1-14B arch (dim 5120, ffn 13824, 40 heads), 8 blocks + reward head, L=32760, batch 1"


def bench_config(world):
    """Identical in both arms (`--impl ours` / `--impl reference`): the workload both are measured on."""
    return {"workload": WORKLOAD, "parallelism": f"ulysses_sp{world}", "weights": "random init (seed 0)",
            "l2": "per-step working set (5.6 GB bf16 weights + >2 GB activations) >> 126 MB L2; no explicit flush",
            "bounded_sample": "CPU legs run the same architecture and depth on latent 16x3x32x52 -> 1248 tokens"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        if os.path.exists(p):
            d = json.load(open(p))
            return dict(hbm=float(d.get("hbm_gbs", 6650.0)), tf_burst=float(d.get("bf16_tflops", 1590.0)),
                        tf_sust=float(d.get("bf16_tflops_sustained", 1400.0)), src="measured")
    except Exception:                                            # unreadable file: the profiling recipe's stated fallback
        pass
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def traffic_from_profiles(kernel_prefix="attn_fwd_kernel"):
    """DRAM bytes per launch (read + write) of the dominant kernel from the newest committed `ncu --set full` summary
    (profiles/r*_ncu_full*.csv, written by tools/summarize_ncu.py): (bytes, file) or (None, None)."""
    unit = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}
    def age(path):                     # newest round first (file name r<NN>_...), then newest file: a fresh checkout levels the mtimes
        digits = "".join(ch for ch in os.path.basename(path)[1:3] if ch.isdigit())
        return (int(digits) if digits else -1, os.path.getmtime(path))
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full*.csv")), key=age, reverse=True):
        try:
            import csv
            rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("#"))]
            hdr, units = rows[0], rows[1]
            ir, iw = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            for r in rows[2:]:
                if r and r[0].startswith(kernel_prefix):
                    return (float(r[ir]) * unit[units[ir].lower()] + float(r[iw]) * unit[units[iw].lower()],
                            os.path.relpath(path, ROOT))
        except Exception:
            continue
    return None, None


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
        try:
            return self._stop(t0, t1)
        except Exception as e:                                   # an unparsable field ("[N/A]") must not cost the measurement
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"clock sampling failed: {type(e).__name__}: {e}"]}

    def _stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (ts, r) in self.rows if t0 <= ts <= t1 + 0.2 and len(r) >= 7] or [r for (_, r) in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        def num(v):
            try:
                return float(v)
            except ValueError:                                   # "[N/A]" on boards that do not report the field
                return None
        sm = sorted(x for x in (num(r[0]) for r in rows) if x is not None)
        power = [x for x in (num(r[2]) for r in rows) if x is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": num(rows[0][1]), "power_w_max": max(power) if power else None,
                "samples": len(rows), "reasons": reasons}


# ---------------------------------------------------------------------------------------------------
# CPU legs (the oracle: the only code below that touches oracle/)
# ---------------------------------------------------------------------------------------------------
def sample_inputs():
    from oracle import synth
    cfg = synth.cfg_14b("t2v", layers=NUM_BLOCKS)
    return cfg, synth.make_inputs(cfg, SAMPLE_LATENT, 2, text_tokens=512)


def cpu_oracle_pass(sd, qa, mlp, threads=None, repeats=1):
    """One PAVRM scoring pass of the oracle (CPU fp32) on the bounded sample with the given weights.
    Returns (tokens, [seconds], logit, features)."""
    from oracle import wan_oracle as O
    if threads:
        torch.set_num_threads(threads)
    cfg, inp = sample_inputs()
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            logit, feats = O.pavrm_reward(sd, cfg, qa, mlp, inp["x"], inp["t"], inp["context"], inp["seq_len"],
                                          selected_layers=(NUM_BLOCKS,), num_blocks=NUM_BLOCKS)
            times.append(time.perf_counter() - t0)
    return inp["seq_len"], times, float(logit), feats


def synthetic_cpu_weights():
    """14B-dims state dict for the TIMING-only reference arm: values do not matter, so tensors of one shape share
    storage (reads still stream the full matrix per use) — building 2.8 G random fp32 parameters would take longer than
    the measurement."""
    from oracle import synth
    cfg = synth.cfg_14b("t2v", layers=NUM_BLOCKS)
    g = torch.Generator().manual_seed(0)
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
    sd, by_shape = {}, {}
    for k, v in shp.items():
        if len(v) > 1:
            if v not in by_shape:
                by_shape[v] = torch.empty(v).uniform_(-0.02, 0.02, generator=g)
            sd[k] = by_shape[v]
        else:
            sd[k] = torch.zeros(v)
        if k.endswith("norm_q.weight") or k.endswith("norm_k.weight") or k.endswith("norm3.weight"):
            sd[k] = torch.ones(v)
    qa, mlp = synth.make_reward_state_dicts(cfg.dim, 1)
    return sd, qa, mlp


SAMPLE_DESC = ("oracle port (torch CPU fp32) of the SAME workload architecture and depth — 14B dims, 8 blocks + reward head — "
               "on a bounded sample: latent 16x3x32x52 -> 1248 tokens (attention cost grows with L^2: a short clip flatters the CPU)")


def run_reference(args):
    """`--impl reference`: the reference algorithm's CPU implementation (oracle port — the reference is Python/PyTorch and
    is not present on the GPU box) with all host threads; every step = one pass over the bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = min(cores, 64)
    sd, qa, mlp = synthetic_cpu_weights()
    cpu_oracle_pass(sd, qa, mlp, threads, repeats=max(1, min(args.warmup, 1)))
    steps = max(1, min(args.steps, 10))                       # bounded: ~3-4 s per pass on 16 cores
    L, times, _, _ = cpu_oracle_pass(sd, qa, mlp, threads, repeats=steps)
    ms = 1e3 * sum(times) / len(times)
    val = L / (ms / 1e3)
    print(json.dumps({"impl": "reference", "metric": "dit_tokens_per_s", "value": val, "unit": "tokens/s", "n_gpus": args.gpus,
                      "steps": len(times), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(max(world, args.gpus)),
                      "cpu_baseline": {"value": val, "unit": "tokens/s", "cores": threads, "kind": "port", "sample": SAMPLE_DESC},
                      "e2e": {"value": val, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ---------------------------------------------------------------------------------------------------
# GPU baseline: the oracle port on CUDA with eager PyTorch kernels (cuBLAS bf16 + flash-attn 2), BASELINE.md §2.2
# ---------------------------------------------------------------------------------------------------
def gpu_eager_baseline(sd_dev, qa_dev, mlp_dev, x_dev, t_dev, ctx_dev, L, ours_ms, steps=3):
    from oracle import synth
    from oracle import wan_oracle as O
    try:
        import flash_attn  # noqa: F401
    except Exception as e:
        return {"unavailable": f"flash_attn import failed: {type(e).__name__}: {e}"}
    cfg = synth.cfg_14b("t2v", layers=NUM_BLOCKS)

    def one():
        with torch.no_grad():
            logit, _ = O.pavrm_reward(sd_dev, cfg, qa_dev, mlp_dev, [x_dev], t_dev, [ctx_dev], L, selected_layers=(NUM_BLOCKS,),
                                      num_blocks=NUM_BLOCKS, autocast_dtype=torch.bfloat16, native=True)
        return logit
    try:
        logit = one()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            one()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / steps
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
    return {"what": "oracle port on this GPU with eager PyTorch kernels: bf16 F.linear (cuBLAS, weights pre-cast once), "
                    "flash_attn_func (FA2 2.8), ATen LayerNorm / RMSNorm, float64 RoPE as the reference does — same workload, same weights",
            "ms_per_step": ms, "value": L / (ms * 1e-3), "unit": "tokens/s", "reward_logit": float(logit),
            "speedup_ours_over_eager": ms / ours_ms}


# ---------------------------------------------------------------------------------------------------
# N > 1: the PRFL training-step leg runs in CHILD processes (one per rank, their own process group)
# ---------------------------------------------------------------------------------------------------
# Why: a cross-rank hang inside a collective cannot be cancelled from inside the process — it wedges the NCCL communicator
# of the job, costs the NCCL watchdog's ten minutes and ends in SIGABRT (profiles/r02_bench_n4_deadlock.log).  With the leg in
# children, the parents (whose own process group stays healthy and idle) put a deadline on it, kill exactly the PIDs they
# started, AGREE on the outcome over their own group, and can run the leg again in a more conservative mode.
CHILD_MODES = [
    ("default: peer-store Ulysses exchange, gradient reduce-scatter overlapped on a side stream", {}),
    ("conservative: NCCL all-to-all exchange, reduce-scatter on the compute stream, NCCL without NVLS multicast",
     {"PRFL_ULYSSES": "nccl", "PRFL_RS": "serial", "NCCL_NVLS_ENABLE": "0"}),
]


def free_port():
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def child_env(rank, local, world, port, extra):
    """Environment of one child rank: the launcher's variables with a rendezvous of the children's own (rank 0 of the
    children hosts a fresh TCP store on `port`; torchrun's TORCHELASTIC_* variables would point them at the agent's store and
    the parents' key space, so they are dropped)."""
    env = {k: v for k, v in os.environ.items() if not k.startswith("TORCHELASTIC_")}
    env.update(RANK=str(rank), LOCAL_RANK=str(local), WORLD_SIZE=str(world), LOCAL_WORLD_SIZE=str(world),
               MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PRFL_CHILD_OF=str(os.getpid()))
    env.update(extra)
    return env


def wait_child(proc, deadline_s, fail_flag, tick=None):
    """0 = exited with code 0; 1 = exited non-zero (or a peer's child did: `fail_flag` exists); 2 = still running at the
    deadline.  In cases 1 / 2 the child — exactly the PID started here — is killed."""
    t_end = time.time() + deadline_s
    while True:
        rc = proc.poll()
        if rc is not None:
            if rc != 0:
                try:
                    open(fail_flag, "w").close()                  # tell the other parents: their children wait for a dead peer
                except OSError:
                    pass
            return 0 if rc == 0 else 1
        crashed_elsewhere = os.path.exists(fail_flag)
        if crashed_elsewhere or time.time() > t_end:
            proc.kill()
            proc.wait()
            return 1 if crashed_elsewhere else 2
        if tick:
            tick()
        time.sleep(0.5)


def prfl_step_in_children(make_cmd, world, rank, local, dev, deadline_s, begin_leg=None, tick=None, modes=None):
    """Run the leg as `world` child processes (this parent starts the one of its rank).  Collective over the parents'
    default process group.  Returns (result or None, attempts, fallback_in_process):
      * every child exits 0 -> rank 0 gets the result object its child wrote, the others get {};
      * some child is still running at the deadline (a hang) -> all are killed and the next mode of `modes` is tried;
      * some child exits non-zero (deterministic: out of memory, environment) -> no retry in another mode; the caller may run
        the leg in-process (`fallback_in_process`)."""
    import tempfile
    import torch.distributed as dist
    attempts = []
    for name, extra in (modes or CHILD_MODES):
        if begin_leg:
            begin_leg(f"prfl_step in child processes ({name.split(':')[0]})")
        port_t = torch.zeros(1, dtype=torch.int64, device=dev)
        if rank == 0:
            port_t[0] = free_port()
        dist.broadcast(port_t, 0)
        port = int(port_t)
        base = os.path.join(tempfile.gettempdir(), f"prfl_step_{port}")
        out, fail_flag, errf = base + ".json", base + ".failed", f"{base}.rank{rank}.err"
        t0 = time.time()
        with open(errf, "w") as ef:
            proc = subprocess.Popen(make_cmd(out), env=child_env(rank, local, world, port, extra), stdout=subprocess.DEVNULL, stderr=ef,
                                    cwd=ROOT)
            st = wait_child(proc, deadline_s, fail_flag, tick)
        flags = torch.tensor([int(st == 1), int(st == 2)], dtype=torch.int32, device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MAX)
        crashed, hung = bool(int(flags[0])), bool(int(flags[1]))
        att = {"mode": name, "seconds": round(time.time() - t0, 1),
               "outcome": "ok" if not (crashed or hung) else ("a child exited non-zero" if crashed else f"no result within {deadline_s} s: children killed")}
        res = None
        if not crashed and not hung:
            res = {}
            if rank == 0:
                try:
                    res = json.load(open(out))
                except Exception as e:
                    att["outcome"], crashed = f"children exited 0 but the result is unreadable: {type(e).__name__}: {e}", True
            ok_t = torch.tensor([0 if (rank == 0 and crashed) else 1], dtype=torch.int32, device=dev)
            dist.broadcast(ok_t, 0)
            if int(ok_t) == 0:
                res, crashed = None, True
        if res is None and rank == 0:
            try:
                att["stderr_tail_rank0"] = open(errf).read()[-400:]
            except OSError:
                pass
        attempts.append(att)
        for f in (errf, fail_flag if rank == 0 else None, out if rank == 0 else None):
            try:
                if f:
                    os.remove(f)
            except OSError:
                pass
        if res is not None:
            return res, attempts, False
        if crashed:
            return None, attempts, True
    return None, attempts, False


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

    # ---- the measurement proper comes FIRST; everything after it (parity legs, eager baseline, PRFL training step) is
    #      guarded by a wall-clock watchdog that prints the line with what is finished and exits 0, so that a hang in an
    #      optional leg (a cross-rank deadlock costs the NCCL watchdog's 10 minutes and a SIGABRT otherwise) can never lose it ----
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
    per_step = ms / args.steps

    pk = peaks()
    attn_ms = timer.elapsed_ms("attn_fwd_self")
    attn_avg = sum(attn_ms) / max(1, len(attn_ms))
    heads_local = HEADS // world
    attn_flops = 4.0 * L * L * HD * heads_local                      # algorithmic, per launch
    achieved = attn_flops / (attn_avg * 1e-3) / 1e12 if attn_avg > 0 else 0.0
    try:
        traffic, traffic_src = traffic_from_profiles() if world == 1 else (None, None)
    except Exception:                                            # evidence lookup must never cost the measurement
        traffic, traffic_src = None, None
    parity = {"reward_logit_full_workload": logit}
    line = {
        "metric": "dit_tokens_per_s", "value": L / (per_step * 1e-3), "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": bench_config(world),
        "e2e": {"value": L / (ms_e2e / args.steps * 1e-3), "unit": "tokens/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": "attn_fwd_kernel (self-attention fwd, tcgen05)", "bound": "tensor", "achieved": achieved,
                     "peak": pk["tf_sust"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sust"],
                     # DRAM bytes per launch of this kernel at this shape, read from the committed `ncu --set full` summary
                     "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": 4.0 * L * HEADS * HD * 2,
                     "peak_source": pk["src"] + " sustained bf16", "launch_ms": attn_avg, "launches_timed": len(attn_ms),
                     "share_of_step": attn_avg * len(attn_ms) / max(ms, 1e-9)},
        "step_tflops": (8 * 41.96e12) / world / (per_step * 1e-3) / 1e12,
        "parity": parity,
        "prfl_step": None,
    }
    emit_lock = threading.Lock()
    emitted = [False]
    leg = ["(none)", time.time(), None]          # name, start, own time limit in seconds (None: --leg-timeout)

    def emit(note=None):
        with emit_lock:
            if emitted[0]:
                return
            emitted[0] = True
            if note:
                line["watchdog"] = note
            if rank == 0:
                try:
                    txt = json.dumps(line)
                except Exception as e:          # a half-built optional leg must not cost the measurement
                    slim = {k: v for k, v in line.items() if k not in ("parity", "prfl_step", "gpu_baseline", "cpu_baseline")}
                    slim["watchdog"] = f"{note or ''} (optional legs dropped: {type(e).__name__}: {e})"
                    txt = json.dumps(slim)
                print(txt, flush=True)

    def watchdog():
        while not emitted[0]:
            time.sleep(1.0)
            limit = leg[2] or args.leg_timeout
            if time.time() - leg[1] > limit:
                emit(f"leg '{leg[0]}' did not finish within {limit} s; line printed without it and the process ended")
                time.sleep(2.0 if rank == 0 else 6.0)
                os._exit(0)                               # a hung collective cannot be cancelled: leave without the NCCL teardown
    threading.Thread(target=watchdog, daemon=True).start()

    def begin_leg(name, limit=None):
        leg[0], leg[1], leg[2] = name, time.time(), limit

    # ---- N > 1: sequence-parallel parity on this very process group (NCCL, symmetric-memory exchange, training path, sharded
    #      optimizer, Ulysses x Ring) — the driver's GPU test box has one GPU, so this is where those numbers become visible ----
    if world > 1 and not args.no_parity:
        begin_leg("parity.sp_checks")
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        try:
            import sp_check
            parity.update(sp_check.run_checks(world, rank, verbose=False))
        except Exception as e:
            parity["sp_checks_error"] = f"{type(e).__name__}: {str(e)[:300]}"

    # ---- same-weights parity on the bounded sample (and the CPU baseline it doubles as) ------------------------------
    if not args.no_cpu and not args.no_parity:
        begin_leg("parity.same_weights_14b / cpu_baseline")
        try:
            cfg_s, inp_s = sample_inputs()
            xs = [u.to(dev) for u in inp_s["x"]]
            cs = [c.to(dev) for c in inp_s["context"]]
            ts = inp_s["t"].to(dev)
            Ls = inp_s["seq_len"]
            for _ in range(2):
                lg, fg = scorer.score(xs, ts, cs, Ls, return_features=True)
            ms_s, _, _ = timed(lambda: scorer.score(xs, ts, cs, Ls), 5)
            fg_cpu = fg.float().cpu()
            if rank == 0:
                sd = {k: v.detach().float().cpu() for k, v in m.state_dict().items()}
                qa_sd = {k: v.detach().float().cpu() for k, v in qa.state_dict().items()}
                mlp_sd = {k: v.detach().float().cpu() for k, v in mlp.state_dict().items()}
                cores = min(os.cpu_count() or 1, 64)
                _, times, logit_o, feats_o = cpu_oracle_pass(sd, qa_sd, mlp_sd, cores)
                del sd
                a, b = fg_cpu.flatten().double(), feats_o.flatten().double()
                cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
                rel = float((a - b).abs().max() / b.abs().max())
                dl = abs(float(lg) - logit_o)
                parity["same_weights_14b"] = {
                    "what": f"same weights on both sides: oracle (CPU fp32) vs GPU, 14B dims, {NUM_BLOCKS} blocks + reward head, {Ls} tokens"
                            + (f", GPU side Ulysses SP={world}" if world > 1 else ""),
                    "features_cos": cos, "features_max_rel": rel, "logit_gpu": float(lg), "logit_oracle": logit_o, "dlogit": dl,
                    "tolerance": {"cos_min": 0.999, "max_rel": 2e-2, "logit_abs": 1e-2},
                    "ok": bool(cos >= 0.999 and rel <= 2e-2 and dl <= 1e-2)}
                if world == 1:
                    line["cpu_baseline"] = {"value": Ls / times[0], "unit": "tokens/s", "cores": cores, "kind": "port",
                                            "sample": SAMPLE_DESC + f"; 1 pass = {times[0]:.1f} s",
                                            "gpu_same_sample": {"value": Ls / (ms_s / 5 * 1e-3), "unit": "tokens/s", "ms_per_pass": ms_s / 5,
                                                                "note": "this arm on the CPU leg's exact sample and weights: the like-for-like ratio is gpu_same_sample / cpu_baseline"}}
            del xs, cs, fg
        except Exception as e:  # the GPU line must not be lost to a host-side problem
            parity["same_weights_14b"] = {"ok": False, "error": f"{type(e).__name__}: {str(e)[:300]}"}
        barrier()

    # ---- eager-PyTorch GPU baseline (N = 1) ---------------------------------------------------------------------------------
    if world == 1 and not args.no_gpu_baseline:
        begin_leg("gpu_baseline")
        try:
            keep32 = ("time_embedding", "time_projection", "norm", "modulation", "bias")
            sd_dev = {k: (v.detach() if any(s in k for s in keep32) else v.detach().to(torch.bfloat16)) for k, v in m.state_dict().items()}
            qa_dev = {k: v.detach() for k, v in qa.state_dict().items()}
            mlp_dev = {k: v.detach() for k, v in mlp.state_dict().items()}
            gpu_base = gpu_eager_baseline(sd_dev, qa_dev, mlp_dev, x_dev, t_dev, ctx_dev, L, per_step)
            if "reward_logit" in gpu_base:
                gpu_base["dlogit_vs_ours"] = abs(gpu_base["reward_logit"] - logit)
            del sd_dev
        except Exception as e:
            gpu_base = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
        line["gpu_baseline"] = gpu_base

    # ---- free the scoring model, then the headline: PRFL 720P training step ---------------------------------------------------
    del scorer, m, qa, mlp
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if not args.no_prfl:
        begin_leg("prfl_step")
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        attempts = None
        try:
            import prfl_step
            Lp = 21 * 45 * 80
            blocks = args.prfl_blocks or prfl_step.fit_blocks(world, Lp)
            prfl, in_process, hung_children = None, world == 1, False
            if world > 1 and not args.prfl_in_process:
                # the parents stay idle meanwhile; what they still hold (CUDA context, NCCL buffers, the small symmetric exchange
                # buffers of the scoring step) is a few GB next to the children's <= 135 GB
                cmd = lambda out: [sys.executable, os.path.join(ROOT, "tools", "prfl_step.py"), "--blocks", str(blocks), "--nograd", "0,2",
                                   "--steps", str(args.prfl_steps), "--out", out]
                prfl, attempts, in_process = prfl_step_in_children(cmd, world, rank, local, dev, args.prfl_timeout, begin_leg,
                                                                   tick=lambda: leg.__setitem__(1, time.time()), modes=CHILD_MODES[:1])
                if prfl is None and not in_process:
                    # the children hung and were killed (every rank agrees: the flags were all-reduced).  Whether the cause is the
                    # training step (the one N = 4 hang of round 2) or the child set-up next to the parents' communicators, the second
                    # and last attempt runs HERE, in the conservative mode — NCCL all-to-all exchange, collectives in stream order —
                    # under the leg watchdog: it is the last leg, so a second hang costs --prfl-timeout and nothing that is already measured
                    hung_children, in_process = True, True
                    os.environ["PRFL_RS"] = "serial"
                    parallel._p2p_disabled = True
                    time.sleep(5.0)                               # the killed children's device memory comes back asynchronously
                    torch.cuda.empty_cache()
                    barrier()
            if prfl is None and (in_process or args.prfl_in_process):
                # after hung children this attempt gets the children's deadline, not the longer general one: a second hang must
                # not double the time the bench takes
                begin_leg("prfl_step (in process)", args.prfl_timeout if hung_children else None)
                t_in = time.time()
                prfl = prfl_step.measure(blocks, (0, 2), prfl_step.LATENT_720P, steps=args.prfl_steps, i2v=True, opt=True)
                if attempts is not None:
                    attempts.append({"mode": ("conservative: NCCL all-to-all exchange, reduce-scatter on the compute stream" if hung_children
                                              else "default") + ", in the bench processes", "seconds": round(time.time() - t_in, 1), "outcome": "ok"})
            if prfl is None:
                prfl = {"error": "the training-step leg did not finish in any mode (see attempts); the measurement above is unaffected"}
            if attempts:
                prfl["attempts"] = attempts
            prfl["metric"] = "PRFL train s/step (BASELINE.json headline), I2V 720Px81f, 14B dims"
            m19 = (prfl.get("extrapolated_s_per_step") or {}).get("m19")
            if blocks == 40 and m19:
                # the one number the reference publishes for this metric (assets/efficiency.png, "PRFL, full 81 frames,
                # 14B, 720P", 43.69 s per step without the SFT loss; hardware and GPU count NOT stated) next to the full-depth step here
                prfl["vs_published"] = {"reference_s_per_step_without_sft": 43.69, "ours_s_per_step_at_E_m_19": m19, "ratio": 43.69 / m19,
                                        "caveat": "the reference's hardware and GPU count are unstated (README recommends >= 80 GB GPUs; "
                                                  "shipped config sp_size 4): orientation only, not a like-for-like baseline"}
            if blocks < 40:
                prfl["note"] = (f"{blocks} of 40 VGM blocks: the largest depth whose bf16 weights + 1/{world} fp32 master/AdamW shards + "
                                "checkpointed activations fit 180 GB at this N with headroom; per-block cost is depth-independent")
        except Exception as e:
            prfl = {"error": f"{type(e).__name__}: {str(e)[:400]}"}
            if attempts:
                prfl["attempts"] = attempts
        line["prfl_step"] = prfl

    begin_leg("teardown")
    emit()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def build_parser():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / same-weights oracle leg")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity legs (profiling runs)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-PyTorch (cuBLAS + flash-attn 2) leg")
    ap.add_argument("--no-prfl", action="store_true", help="skip the PRFL 720P training-step leg")
    ap.add_argument("--prfl-blocks", type=int, default=0, help="VGM depth of the training-step leg (0 = the largest that fits)")
    ap.add_argument("--prfl-steps", type=int, default=2)
    ap.add_argument("--prfl-timeout", type=int, default=210, help="N > 1: seconds one attempt of the training-step leg (child processes) may take")
    ap.add_argument("--prfl-in-process", action="store_true", help="N > 1: run the training-step leg inside the bench processes (no children)")
    ap.add_argument("--leg-timeout", type=int, default=300, help="seconds any post-measurement leg may take before the line is "
                    "printed without it and the process exits (0 hangs are expected; this bounds the cost of one)")
    return ap


if __name__ == "__main__":
    a = build_parser().parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

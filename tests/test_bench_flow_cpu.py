"""CPU: control flow of bench.py's GPU arm with the device layer faked — the measurement is assembled first, every optional
leg (parity, eager baseline, PRFL step) may fail or HANG without costing the JSON line: a failing leg becomes an `error` entry,
a hanging leg trips the wall-clock watchdog, which prints the line with a `watchdog` note and ends the process with exit code 0
(a hung collective cannot be cancelled; the alternative is the NCCL watchdog's SIGABRT ten minutes later and no line at all)."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = textwrap.dedent('''
    import sys, time, types, threading
    import torch, torch.nn as nn
    ROOT = %r
    sys.path.insert(0, ROOT)
    HANG = %r
    EXTRA = %r
    # ---- fake device layer -------------------------------------------------------------------------
    class FakeEvent:
        def __init__(self, enable_timing=False): self.t = None
        def record(self, *a): self.t = time.time()
        def elapsed_time(self, other): return max((other.t - self.t) * 1e3, 1e-3)
    import torch.distributed as dist
    _real_init = dist.init_process_group
    dist.init_process_group = lambda backend, device_id=None, **k: _real_init("gloo")      # N > 1 harness: gloo stands in for NCCL
    torch.cuda.is_available = lambda: True
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.Event = FakeEvent
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    _real_device = torch.device
    class _Dev:
        def __new__(cls, *a, **k):
            return _real_device("cpu")
    torch.device = _Dev
    # ---- fake product package ------------------------------------------------------------------------
    pkg = types.ModuleType("prfl_b200"); pkg.__path__ = []
    lib = types.ModuleType("prfl_b200._lib"); lib.launch_count_reset = lambda: None; lib.launch_count = lambda: 42
    ops = types.ModuleType("prfl_b200.ops"); ops.TIMER = None
    class KernelTimer:
        def __init__(self, names): self.names = names
        def elapsed_ms(self, name): return [1.0, 1.0]
    ops.KernelTimer = KernelTimer
    par = types.ModuleType("prfl_b200.parallel"); par.initialize_sequence_parallel_state = lambda n: None
    class Blk(nn.Module):
        def __init__(self): super().__init__(); self.norm3 = nn.LayerNorm(4)
    class WanModel(nn.Module):
        def __init__(self, **kw): super().__init__(); self.blocks = nn.ModuleList([Blk() for _ in range(2)]); self.head = nn.Linear(2, 2)
    mod = types.ModuleType("prfl_b200.model"); mod.WanModel = WanModel
    net = types.ModuleType("prfl_b200.network")
    net.MLP = lambda d: nn.Linear(2, 2)
    net.QueryAttention = lambda *a, **k: nn.Linear(2, 2)
    calls = [0]
    class PavrmScorer(nn.Module):
        def __init__(self, m, qa, mlp, n): super().__init__(); self.m, self.qa, self.mlp = m, qa, mlp
        def score(self, x, t, c, L, return_features=False):
            calls[0] += 1
            if return_features:
                if HANG: time.sleep(3600)
                return torch.zeros(1, 1, 1), torch.zeros(1, 1, L, 8)
            return torch.zeros(1, 1, 1)
    pav = types.ModuleType("prfl_b200.pavrm"); pav.PavrmScorer = PavrmScorer
    for name, m_ in [("prfl_b200", pkg), ("prfl_b200._lib", lib), ("prfl_b200.ops", ops), ("prfl_b200.parallel", par),
                     ("prfl_b200.model", mod), ("prfl_b200.network", net), ("prfl_b200.pavrm", pav)]:
        sys.modules[name] = m_
    pkg._lib, pkg.ops, pkg.parallel = lib, ops, par
    if "--prfl-blocks" in EXTRA:          # success path of the training-step leg: a canned result in place of tools/prfl_step.py
        ps = types.ModuleType("prfl_step"); ps.LATENT_720P = (21, 90, 160); ps.fit_blocks = lambda w, L: 8
        def _measure(blocks, m_list, latent, steps=2, i2v=True, opt=True):
            runs = {"m0": {"s_per_step": 3.5}, "m2": {"s_per_step": 5.0}}
            return {"blocks": blocks, "runs": runs, "extrapolated_s_per_step": {"per_nograd_forward_s": 0.75, "m0": 3.5, "m19": 17.75, "m38": 32.0}}
        ps.measure = _measure
        sys.modules["prfl_step"] = ps
    import bench
    import os
    if os.environ.get("HARNESS_HUNG_CHILDREN"):      # the children of the first attempt hung and were killed (agreed by all ranks)
        bench.prfl_step_in_children = lambda *a, **k: (None, [{"mode": "default", "seconds": 1.0, "outcome": "no result within 1 s: children killed"}], False)
        par._p2p_disabled = False
        _m = sys.modules["prfl_step"].measure
        def _measure_checked(*a, **k):               # the in-process attempt must run in the conservative mode
            assert os.environ.get("PRFL_RS") == "serial" and par._p2p_disabled is True
            return _m(*a, **k)
        sys.modules["prfl_step"].measure = _measure_checked
    bench.ClockSampler = lambda i: types.SimpleNamespace(stop=lambda a, b: {"sm_mhz": 1.0, "sm_max_mhz": 2.0, "reasons": []})
    sys.argv = ["bench.py", "--steps", "2", "--warmup", "1", "--leg-timeout", %r] + EXTRA
    bench.run_ours(bench.build_parser().parse_args())
    print("CLEAN EXIT")
''')


def _run(hang, leg_timeout, extra=(), env=None, wait=True):
    code = HARNESS % (ROOT, hang, list(extra), str(leg_timeout))
    env = dict(os.environ, OMP_NUM_THREADS="4", **(env or {}))
    if not wait:
        return subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, env=env)
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)


def _line(res):
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, (res.stdout[-2000:], res.stderr[-3000:])       # exactly ONE JSON line
    return json.loads(lines[0])


def test_failing_legs_become_error_entries_and_the_line_is_printed_once():
    res = _run(False, 300)
    assert res.returncode == 0, res.stderr[-3000:]
    line = _line(res)
    assert "CLEAN EXIT" in res.stdout and "watchdog" not in line
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "parity", "prfl_step"):
        assert k in line, k
    assert line["gpu_launches"] == 42 and line["roofline"]["launches_timed"] == 2
    # with a faked device the optional legs cannot succeed: each must have turned into an error entry, not an exception
    assert line["parity"]["same_weights_14b"]["ok"] is False and "error" in line["parity"]["same_weights_14b"]
    assert "unavailable" in line["gpu_baseline"] or "ms_per_step" in line["gpu_baseline"]
    assert "error" in line["prfl_step"]


def test_successful_training_step_leg_is_folded_into_the_line():
    for blocks, expect_published in (("40", True), ("8", False)):
        res = _run(False, 300, extra=["--prfl-blocks", blocks])
        assert res.returncode == 0, res.stderr[-3000:]
        p = _line(res)["prfl_step"]
        assert p["blocks"] == int(blocks) and p["runs"]["m2"]["s_per_step"] == 5.0 and p["metric"].startswith("PRFL train s/step")
        assert ("vs_published" in p) is expect_published and ("note" in p) is (not expect_published)
        if expect_published:
            assert abs(p["vs_published"]["ratio"] - 43.69 / 17.75) < 1e-9 and "unstated" in p["vs_published"]["caveat"]


def test_hanging_leg_trips_the_watchdog_line_still_printed_exit_code_zero():
    res = _run(True, 3)
    assert res.returncode == 0, res.stderr[-3000:]
    line = _line(res)
    assert "CLEAN EXIT" not in res.stdout
    assert "did not finish within 3 s" in line["watchdog"] and "same_weights_14b" in line["watchdog"]
    assert line["value"] > 0 and line["prfl_step"] is None


# ---------------------------------------------------------------------------------------------------
# N > 1: the training-step leg in child processes (bench.prfl_step_in_children), 2 parents over gloo with fake children
# ---------------------------------------------------------------------------------------------------
CHILD = textwrap.dedent('''
    import json, os, sys, time
    out, scenario = sys.argv[1], sys.argv[2]
    rank, serial = int(os.environ["RANK"]), os.environ.get("PRFL_RS") == "serial"
    assert "TORCHELASTIC_USE_AGENT_STORE" not in os.environ and os.environ["MASTER_ADDR"] == "127.0.0.1"
    if scenario == "hang_then_ok" and not serial:
        time.sleep(3600)                                   # every rank of the first attempt waits for ever (a cross-rank deadlock)
    if scenario == "crash" and rank == 1:
        sys.exit(3)                                        # one rank dies; its peer would wait for it for ever
    if scenario == "crash":
        time.sleep(3600)
    if rank == 0:
        json.dump({"mode": os.environ.get("PRFL_ULYSSES", "p2p"), "port": os.environ["MASTER_PORT"], "world": os.environ["WORLD_SIZE"]},
                  open(out, "w"))
''')

PARENT = textwrap.dedent('''
    import json, os, sys, time
    import torch, torch.distributed as dist
    sys.path.insert(0, %r)
    import bench
    rank = int(os.environ["RANK"])
    dist.init_process_group("gloo")
    legs = []
    res, attempts, in_process = bench.prfl_step_in_children(
        lambda out: [sys.executable, "-c", %r, out, %r], 2, rank, rank, torch.device("cpu"), deadline_s=4, begin_leg=legs.append)
    print("RESULT " + json.dumps({"rank": rank, "res": res, "attempts": attempts, "in_process": in_process, "legs": legs}), flush=True)
    dist.barrier()                                         # the parents' own group is still healthy
    dist.destroy_process_group()
''')


def _parents(scenario):
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1",
                   TORCHELASTIC_USE_AGENT_STORE="False", TORCHELASTIC_RUN_ID="x")      # what torchrun leaves in the environment
        procs.append(subprocess.Popen([sys.executable, "-c", PARENT % (ROOT, CHILD, scenario)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True, cwd=ROOT))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=300)
        assert p.returncode == 0, e[-3000:]
        outs.append(json.loads([l for l in o.splitlines() if l.startswith("RESULT ")][0][7:]))
    return sorted(outs, key=lambda d: d["rank"])


def test_children_hang_is_killed_at_the_deadline_and_retried_in_the_conservative_mode():
    t0 = __import__("time").time()
    r0, r1 = _parents("hang_then_ok")
    assert __import__("time").time() - t0 < 120
    for r in (r0, r1):
        assert [a["outcome"][:9] for a in r["attempts"]] == ["no result", "ok"], r["attempts"]
        assert r["attempts"][1]["mode"].startswith("conservative") and r["in_process"] is False and len(r["legs"]) == 2
    assert r0["res"]["mode"] == "nccl" and r0["res"]["world"] == "2" and r1["res"] == {}
    assert r0["attempts"] == r1["attempts"] or [a["outcome"] for a in r0["attempts"]] == [a["outcome"] for a in r1["attempts"]]


def test_children_crash_is_not_retried_and_peers_are_released_early():
    t0 = __import__("time").time()
    r0, r1 = _parents("crash")
    for r in (r0, r1):
        assert len(r["attempts"]) == 1 and r["attempts"][0]["outcome"] == "a child exited non-zero", r["attempts"]
        assert r["res"] is None and r["in_process"] is True
        assert r["attempts"][0]["seconds"] < 4                      # rank 0's child was killed when rank 1's died, not at the deadline


def test_two_rank_flow_children_crash_without_a_gpu_then_in_process_fallback_errors_and_rank0_prints_the_line():
    """The N > 1 wiring of run_ours end to end (gloo standing in for NCCL, device layer faked): measurement, then the
    training-step leg as REAL children (`tools/prfl_step.py`), which cannot start without a GPU -> a crash, agreed on by both
    parents, no conservative retry, in-process fallback, whose failure becomes an error entry; exactly one line, from rank 0,
    both ranks leave through the normal teardown (barrier + destroy_process_group on the parents' own group)."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [_run(False, 300, extra=["--gpus", "2", "--no-parity", "--no-cpu", "--prfl-timeout", "120"], wait=False,
                  env=dict(RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port)))
             for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0 and "CLEAN EXIT" in o, (o[-1500:], e[-3000:])
    lines0 = [l for l in outs[0][0].splitlines() if l.startswith("{")]
    assert len(lines0) == 1 and not [l for l in outs[1][0].splitlines() if l.startswith("{")]
    line = json.loads(lines0[0])
    assert line["n_gpus"] == 2 and "watchdog" not in line and line["config"]["parallelism"] == "ulysses_sp2"
    att = line["prfl_step"]["attempts"]
    assert len(att) == 1 and att[0]["outcome"] == "a child exited non-zero" and att[0]["mode"].startswith("default")
    assert "error" in line["prfl_step"]                          # the in-process fallback cannot run on the faked device either


def test_two_rank_flow_children_hang_then_conservative_attempt_in_the_bench_processes():
    """After a hang of the children (killed at the deadline, agreed by both parents) the second and last attempt of the
    training-step leg runs in the bench processes themselves in the conservative mode (PRFL_RS=serial, NCCL exchange) under the
    leg watchdog; its result is folded into the line with both attempts recorded."""
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    procs = [_run(False, 300, extra=["--gpus", "2", "--no-parity", "--no-cpu", "--prfl-blocks", "40"], wait=False,
                  env=dict(RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), HARNESS_HUNG_CHILDREN="1"))
             for r in range(2)]
    outs = [p.communicate(timeout=600) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0 and "CLEAN EXIT" in o, (o[-1500:], e[-3000:])
    line = json.loads([l for l in outs[0][0].splitlines() if l.startswith("{")][0])
    p = line["prfl_step"]
    assert "error" not in p and p["blocks"] == 40 and "vs_published" in p
    att = p["attempts"]
    assert len(att) == 2 and "children killed" in att[0]["outcome"] and att[1]["outcome"] == "ok"
    assert att[1]["mode"].startswith("conservative") and att[1]["mode"].endswith("in the bench processes")

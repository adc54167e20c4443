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
    # ---- fake device layer -------------------------------------------------------------------------
    class FakeEvent:
        def __init__(self, enable_timing=False): self.t = None
        def record(self, *a): self.t = time.time()
        def elapsed_time(self, other): return max((other.t - self.t) * 1e3, 1e-3)
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
    import bench
    bench.ClockSampler = lambda i: types.SimpleNamespace(stop=lambda a, b: {"sm_mhz": 1.0, "sm_max_mhz": 2.0, "reasons": []})
    sys.argv = ["bench.py", "--steps", "2", "--warmup", "1", "--leg-timeout", %r]
    import argparse
    ap = argparse.ArgumentParser()
    for flag, typ, dflt in [("--gpus", int, 1), ("--steps", int, 5), ("--warmup", int, 3), ("--prfl-blocks", int, 0), ("--prfl-steps", int, 2),
                            ("--leg-timeout", int, 300)]:
        ap.add_argument(flag, type=typ, default=dflt)
    ap.add_argument("--impl", default="ours")
    for flag in ("--no-cpu", "--no-parity", "--no-gpu-baseline", "--no-prfl"):
        ap.add_argument(flag, action="store_true")
    bench.run_ours(ap.parse_args())
    print("CLEAN EXIT")
''')


def _run(hang, leg_timeout):
    code = HARNESS % (ROOT, hang, str(leg_timeout))
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT,
                          env=dict(os.environ, OMP_NUM_THREADS="4"))


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


def test_hanging_leg_trips_the_watchdog_line_still_printed_exit_code_zero():
    res = _run(True, 3)
    assert res.returncode == 0, res.stderr[-3000:]
    line = _line(res)
    assert "CLEAN EXIT" not in res.stdout
    assert "did not finish within 3 s" in line["watchdog"] and "same_weights_14b" in line["watchdog"]
    assert line["value"] > 0 and line["prfl_step"] is None

"""CPU: the parts of bench.py that need no GPU — the reference arm (`--impl reference`: the oracle port on the host cores)
prints one JSON line that carries the contract's keys and the SAME `config` dict as the GPU arm; `roofline.traffic` is read
from the committed ncu summary; the PRFL-step depth rule."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_config_match():
    env = dict(os.environ, OMP_NUM_THREADS="8")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads([l for l in res.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "dit_tokens_per_s" and line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.bench_config(1)                      # the GPU arm prints bench_config(world) too


def test_traffic_comes_from_the_committed_ncu_summary():
    sys.path.insert(0, ROOT)
    import bench
    traffic, src = bench.traffic_from_profiles()
    assert src is not None and src.startswith("profiles/") and os.path.exists(os.path.join(ROOT, src))
    rounds = sorted(int(os.path.basename(f)[1:3]) for f in __import__("glob").glob(os.path.join(ROOT, "profiles", "r*_ncu_full*.csv")))
    assert int(os.path.basename(src)[1:3]) == rounds[-1]                  # the latest round's capture, whatever the files' mtimes
    algorithmic = 4.0 * 32760 * 40 * 128 * 2                              # Q, K, V, O once each
    assert algorithmic <= traffic <= 1.5 * algorithmic


def test_prfl_step_depth_rule():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import prfl_step
    L = 21 * 45 * 80
    d = [prfl_step.fit_blocks(w, L) for w in (1, 2, 4, 8)]
    assert d[3] == 40 and 1 <= d[0] < d[1] < d[2] <= 40
    f8, b8 = prfl_step.algorithmic_flops(L, True, 8)
    assert abs(f8 * 8 / 163.08e12 - 1) < 0.02                            # 163.08 TFLOP per 720P block forward (+CLIP tokens)


def test_prfl_step_extrapolation_in_m():
    """s/step is affine in the number of no-grad forwards: m19 (= E[m] of the reference's randint(0, 38)) and m38 follow from
    two measured m; nothing is reported from a single m."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import prfl_step
    n8 = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_n8.json")))["prfl_step"]["runs"]
    ex = prfl_step.extrapolate_m(n8)["extrapolated_s_per_step"]
    s0, s2 = n8["m0"]["s_per_step"], n8["m2"]["s_per_step"]
    assert abs(ex["per_nograd_forward_s"] - (s2 - s0) / 2) < 1e-12 and abs(ex["m0"] - s0) < 1e-12
    assert abs(ex["m19"] - (s0 + 19 * (s2 - s0) / 2)) < 1e-9 and 16.0 < ex["m19"] < 17.5 and ex["m38"] > ex["m19"]
    assert prfl_step.extrapolate_m({"m2": {"s_per_step": 5.0}}) == {}

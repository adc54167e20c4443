"""CPU: SURVEY §8 row a16 (scheduler step + PRFL loss glue).
  * the oracle restatement (oracle/unipc_oracle.py) against the committed outputs of the UNMODIFIED reference scheduler
    (tests/golden/unipc.pt, made by tests/golden/make_golden.py) — fp32 vs fp32, bit-exact expected, 1e-6 allowed;
  * the host logic of the product scheduler (schedule construction, step bookkeeping, the folded coefficients that
    prfl_unipc_step consumes) against the same goldens, by evaluating the linear combinations with torch CPU ops HERE in
    the test (the product itself has no CPU path: its `step` raises without the CUDA library)."""
import pytest
import torch

from conftest import golden
from oracle.unipc_oracle import UniPCOracle, prfl_loss, toy_velocity


def _inputs(fx):
    g = torch.Generator().manual_seed(fx["seed"])
    x = torch.randn(fx["shape"], generator=g)
    w = torch.randn(fx["shape"], generator=g) * 0.5
    return x, w


def test_oracle_chains_match_reference():
    fx = golden("unipc")
    x_init, w = _inputs(fx)
    for (steps, shift, st, order), ch in fx["chains"].items():
        o = UniPCOracle(solver_order=order, solver_type=st)
        o.set_timesteps(steps, shift=shift)
        assert torch.equal(o.timesteps, ch["timesteps"]) and torch.equal(o.sigmas, ch["sigmas"])
        x = x_init.clone()
        for i, t in enumerate(o.timesteps):
            x = o.step(toy_velocity(x, t, w), t, x)
            # equal_nan: with solver_type "bh1" the reference's LAST step is NaN (B_h = -h = -inf times pred_res = 0,
            # fm_solvers_unipc.py:437,471-478); the restatement reproduces it, the product returns the limit x0 instead
            torch.testing.assert_close(x, ch["traj"][i], rtol=1e-6, atol=1e-6, equal_nan=True)
            torch.testing.assert_close(o.model_outputs[-1], ch["x0"][i], rtol=1e-6, atol=1e-6)


def test_oracle_prfl_gradients_match_reference():
    fx = golden("unipc")
    x_init, w = _inputs(fx)
    for m, gd in fx["prfl"].items():
        o = UniPCOracle()
        o.set_timesteps(40, shift=3.0)
        x = x_init.clone()
        with torch.no_grad():
            for i in range(m):
                t = o.timesteps[i]
                x = o.step(toy_velocity(x, t, w), t, x)
        wg, xg = w.clone().requires_grad_(True), x.clone().requires_grad_(True)
        t = o.timesteps[m]
        prev = o.step(toy_velocity(xg, t, wg), t, xg)
        loss = prfl_loss(torch.tanh(prev.mean(dim=(1, 2, 3, 4)) * 5.0))
        loss.backward()
        torch.testing.assert_close(prev, gd["prev"], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(loss, gd["loss"], rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(wg.grad, gd["grad_w"], rtol=1e-5, atol=1e-9)
        torch.testing.assert_close(xg.grad, gd["grad_x"], rtol=1e-5, atol=1e-9)


def _host_chain(s, steps, order, x, w, upto=None):
    """Drive the product scheduler's bookkeeping + folded coefficients, evaluating the lincombs with torch CPU ops."""
    hist, last, lower, this_order = [], None, 0, None
    traj = []
    for i, t in enumerate(s.timesteps[:upto]):
        v = toy_velocity(x, t, w)
        use_c = i > 0 and last is not None
        po = min(min(order, steps - i), lower + 1)
        c = s.step_coefficients(i, use_c, this_order, po)
        hs = hist + [torch.zeros_like(x)] * 3
        x0 = x - c["sigma"] * v
        xc = x
        if use_c:
            cc = c["corr"]
            xc = cc[0] * last + cc[1] * x0 + cc[2] * hs[0] + cc[3] * hs[1] + cc[4] * hs[2]
        p = c["pred"]
        xp = p[0] * xc + p[1] * x0 + p[2] * hs[0] + p[3] * hs[1] + p[4] * hs[2]
        hist = ([x0] + hist)[:order]
        last, this_order, lower, x = xc, po, min(lower + 1, order), xp
        traj.append((xp, c))
    return traj


def test_product_host_logic_matches_reference():
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    fx = golden("unipc")
    x_init, w = _inputs(fx)
    for (steps, shift, st, order), ch in fx["chains"].items():
        s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False, solver_type=st,
                                        solver_order=order)
        s.set_timesteps(steps, device="cpu", shift=shift)
        assert torch.equal(s.timesteps, ch["timesteps"]) and s.timesteps.dtype == torch.int64
        assert torch.equal(s.sigmas, ch["sigmas"]) and s.sigmas.device.type == "cpu"
        assert s.num_inference_steps == steps and s.step_index is None and s.config.solver_order == order
        for i, (xp, _) in enumerate(_host_chain(s, steps, order, x_init.clone(), w)):
            if torch.isfinite(ch["traj"][i]).all():
                torch.testing.assert_close(xp, ch["traj"][i], rtol=2e-6, atol=2e-6)
            else:       # reference quirk (bh1, final step: -inf * 0): the product lands on the x0 prediction, the intended limit
                assert i == steps - 1 and st == "bh1"
                torch.testing.assert_close(xp, ch["x0"][i], rtol=2e-6, atol=2e-6)
        assert s.index_for_timestep(s.timesteps[3]) == 3


def test_product_step_derivatives_match_autograd():
    """d prev / d model_output and d prev / d sample of the folded step (what _StepFn.backward scales by)."""
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    s.set_timesteps(40, device="cpu", shift=3.0)
    for i, use_c, oc, po in [(0, False, None, 1), (1, True, 1, 2), (7, True, 2, 2), (38, True, 2, 2), (39, True, 2, 1)]:
        c = s.step_coefficients(i, use_c, oc, po)
        v = torch.randn(5, dtype=torch.float64, requires_grad=True)
        x = torch.randn(5, dtype=torch.float64, requires_grad=True)
        last, h0, h1 = (torch.randn(5, dtype=torch.float64) for _ in range(3))
        x0 = x - c["sigma"] * v
        xc = x
        if use_c:
            cc = c["corr"]
            xc = cc[0] * last + cc[1] * x0 + cc[2] * h0 + cc[3] * h1
        p = c["pred"]
        prev = p[0] * xc + p[1] * x0 + p[2] * h0 + p[3] * h1
        gv, gx = torch.autograd.grad(prev.sum(), (v, x))
        assert abs(float(gv[0]) - c["d_model_output"]) < 1e-9 and abs(float(gx[0]) - c["d_sample"]) < 1e-9
    # last step of the chain lands exactly on the x0 prediction (sigma -> 0)
    c = s.step_coefficients(39, True, 2, 1)
    assert c["pred"][0] == 0.0 and abs(c["pred"][1] - 1.0) < 1e-6


def test_unsupported_configurations_raise():
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    for kw in (dict(predict_x0=False), dict(thresholding=True), dict(prediction_type="epsilon"), dict(solver_order=4),
               dict(solver_type="nope")):
        with pytest.raises(NotImplementedError):
            FlowUniPCMultistepScheduler(**kw)
    with pytest.raises(ValueError):
        FlowUniPCMultistepScheduler().step(torch.zeros(1), 0, torch.zeros(1))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_step_has_no_cpu_fallback():
    from prfl_b200 import _lib
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    s = FlowUniPCMultistepScheduler()
    s.set_timesteps(4, device="cpu", shift=3.0)
    with pytest.raises(_lib.PrflError):
        s.step(torch.zeros(1, 4), s.timesteps[0], torch.zeros(1, 4))


def test_add_noise_matches_reference():
    """fm_solvers_unipc.py:758-797 (cold path, plain torch ops: runs on CPU): index by timestep / begin index / step index."""
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    fx = golden("unipc")
    g = torch.Generator().manual_seed(fx["seed"])
    shape = fx["shape"]
    torch.randn(shape, generator=g)
    torch.randn(shape, generator=g)                          # x_init, w: advance the generator as the fixture did
    clean = torch.randn((3,) + tuple(shape[1:]), generator=g)
    eps = torch.randn((3,) + tuple(shape[1:]), generator=g)
    s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=1, use_dynamic_shifting=False)
    s.set_timesteps(40, device="cpu", shift=3.0)
    ts = s.timesteps[[0, 17, 39]]
    torch.testing.assert_close(s.add_noise(clean, eps, ts), fx["add_noise"]["by_timestep"], rtol=0, atol=0)
    s.set_begin_index(5)
    torch.testing.assert_close(s.add_noise(clean, eps, ts), fx["add_noise"]["begin_index"], rtol=0, atol=0)
    s._step_index = 6                                        # the state after one step from begin index 5
    torch.testing.assert_close(s.add_noise(clean, eps, ts), fx["add_noise"]["after_step"], rtol=0, atol=0)
    assert len(s) == 1000 and s.scale_model_input(clean) is clean


def test_add_noise_on_fresh_scheduler_float_schedule():
    """A scheduler that never saw set_timesteps() has the float schedule `sigmas * 1000`; with shift = 5 many entries share
    an integer part (golden from the unmodified reference: tests/golden/make_golden.py case_unipc_fresh)."""
    from prfl_b200.scheduler import FlowUniPCMultistepScheduler
    fx = golden("unipc_fresh")
    for shift, c in fx["cases"].items():
        s = FlowUniPCMultistepScheduler(num_train_timesteps=1000, shift=shift, use_dynamic_shifting=False)
        assert torch.equal(s.timesteps, c["schedule"])
        for i, t in zip(c["idx"], c["timesteps"]):
            assert s.index_for_timestep(t, s.timesteps) == i
        torch.testing.assert_close(s.add_noise(c["clean"], c["eps"], c["timesteps"]), c["out"], rtol=0, atol=0)

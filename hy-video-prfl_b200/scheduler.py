"""`FlowUniPCMultistepScheduler` with the reference's API (diffusers_lite/wan/utils/fm_solvers_unipc.py), the step fused
into one CUDA kernel (the differentiable scheduler step between the video model and the reward
model in PRFL, train_prfl.py:690,734; also the sampling loop, text2video.py:298-303).

What is kept: constructor arguments, `set_timesteps(num_inference_steps, device, sigmas, mu, shift)`, `step(model_output,
timestep, sample, return_dict, generator)`, `convert_model_output`, `timesteps` (int64), `sigmas` (fp32, CPU),
`model_outputs`, `last_sample`, `lower_order_nums`, `step_index`, `begin_index`, `this_order`, `config.*`.
What changes: the reference evaluates a step as ~25 ATen elementwise launches; every quantity of a step is linear in
(sample, model_output, last_sample, previous x0 predictions) with scalar coefficients that depend only on the sigma
schedule, so the scalars are folded on the host (`step_coefficients`, fp32 sigma arithmetic as in the reference, folding
in float64) and `prfl_unipc_step` reads each tensor once.  Autograd: the step is linear, so its backward with respect
to (model_output, sample) is two scalings (`prfl_scale2_f32`).

Restricted to what the reference instantiates (predict_x0, flow_prediction, bh1/bh2, no thresholding, no solver_p,
final_sigmas_type "zero", solver_order <= 3); anything else raises NotImplementedError rather than silently differing.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from . import ops

__all__ = ["FlowUniPCMultistepScheduler", "SchedulerOutput", "prfl_loss"]


class SchedulerOutput:
    def __init__(self, prev_sample):
        self.prev_sample = prev_sample


def prfl_loss(reward_scores: torch.Tensor, target_reward: float = 2.0, weight: float = 0.1) -> torch.Tensor:
    """train_prfl.py:796-798: `0.1 * F.relu(-reward_scores.squeeze() + target_reward).mean()` (a [B]-sized scalar op)."""
    return weight * torch.relu(-reward_scores.squeeze() + target_reward).mean()


class _StepFn(torch.autograd.Function):
    """prev = UniPC step; d prev / d model_output = cv, d prev / d sample = cx (scalars)."""

    @staticmethod
    def forward(ctx, model_output, sample, sch, coef):
        x0, corrected, prev = ops.unipc_step(sample.detach(), model_output.detach(), sch.last_sample, sch._hist(), coef["sigma"],
                                             coef["corr"], coef["pred"])
        ctx.cv, ctx.cx = coef["d_model_output"], coef["d_sample"]
        ctx.need = (model_output.requires_grad, sample.requires_grad)
        ctx.mark_non_differentiable(x0)
        if corrected is None:
            return prev, x0, sample.detach()
        ctx.mark_non_differentiable(corrected)
        return prev, x0, corrected

    @staticmethod
    def backward(ctx, g, _g0, _g1):
        gv, gx = ops.scale2(g.float(), ctx.cv, ctx.cx if ctx.need[1] else None)
        return (gv if ctx.need[0] else None), gx, None, None


class FlowUniPCMultistepScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, prediction_type: str = "flow_prediction",
                 shift: Optional[float] = 1.0, use_dynamic_shifting=False, thresholding: bool = False,
                 dynamic_thresholding_ratio: float = 0.995, sample_max_value: float = 1.0, predict_x0: bool = True,
                 solver_type: str = "bh2", lower_order_final: bool = True, disable_corrector: List[int] = [],
                 solver_p=None, timestep_spacing: str = "linspace", steps_offset: int = 0,
                 final_sigmas_type: Optional[str] = "zero"):
        # fm_solvers_unipc.py:77-132
        if solver_type not in ("bh1", "bh2"):
            if solver_type in ("midpoint", "heun", "logrho"):
                solver_type = "bh2"
            else:
                raise NotImplementedError(f"{solver_type} is not implemented for {self.__class__}")
        if not predict_x0 or prediction_type != "flow_prediction" or thresholding or solver_p is not None or \
                final_sigmas_type != "zero" or solver_order > 3:
            raise NotImplementedError("prfl_b200 FlowUniPCMultistepScheduler covers the configuration the reference "
                                      "instantiates: predict_x0, flow_prediction, no thresholding / solver_p, order <= 3")
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, solver_order=solver_order,
                                      prediction_type=prediction_type, shift=shift, use_dynamic_shifting=use_dynamic_shifting,
                                      thresholding=thresholding, dynamic_thresholding_ratio=dynamic_thresholding_ratio,
                                      sample_max_value=sample_max_value, predict_x0=predict_x0, solver_type=solver_type,
                                      lower_order_final=lower_order_final, disable_corrector=disable_corrector,
                                      solver_p=solver_p, timestep_spacing=timestep_spacing, steps_offset=steps_offset,
                                      final_sigmas_type=final_sigmas_type)
        self.predict_x0 = predict_x0
        self.num_inference_steps = None
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sigmas = torch.from_numpy(1.0 - alphas).to(dtype=torch.float32)
        if not use_dynamic_shifting:
            sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.sigmas = sigmas.to("cpu")
        self.timesteps = sigmas * num_train_timesteps
        self._timesteps_host = self.timesteps.tolist()          # float schedule until set_timesteps() (fm_solvers_unipc.py:103)
        self.model_outputs = [None] * solver_order
        self.timestep_list = [None] * solver_order
        self.lower_order_nums = 0
        self.disable_corrector = disable_corrector
        self.solver_p = None
        self.last_sample = None
        self._step_index = None
        self._begin_index = None
        self.this_order = None
        self.sigma_min = self.sigmas[-1].item()
        self.sigma_max = self.sigmas[0].item()

    # ---- bookkeeping identical to the reference ----------------------------------------------------
    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def __len__(self):
        return self.config.num_train_timesteps

    def time_shift(self, mu: float, sigma: float, t):
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device=None, sigmas=None, mu=None, shift=None):
        # fm_solvers_unipc.py:160-227
        if self.config.use_dynamic_shifting and mu is None:
            raise ValueError(" you have to pass a value for `mu` when `use_dynamic_shifting` is set to be `True`")
        if sigmas is None:
            sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if self.config.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            if shift is None:
                shift = self.config.shift
            sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.config.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas).to("cpu")
        ts_host = torch.from_numpy(timesteps).to(dtype=torch.int64)
        self.timesteps = ts_host.to(device=device)
        self.num_inference_steps = len(timesteps)
        self.model_outputs = [None] * self.config.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self._step_index = None
        self._begin_index = None
        self._timesteps_host = ts_host.tolist()                 # exact values of the schedule's own dtype (int64 here), no device sync

    def index_for_timestep(self, timestep, schedule_timesteps=None):
        # fm_solvers_unipc.py:628-641 (exact `==` on the schedule's values, second hit if the value repeats), on the host
        # copy of the schedule (no device sync per step).  No truncation: a float schedule (the constructor's
        # `sigmas * num_train_timesteps`, dynamic shifting) has entries that share an integer part.
        ts = self._timesteps_host if schedule_timesteps is None else schedule_timesteps.tolist()
        if torch.is_tensor(timestep):
            timestep = timestep.item()
        if schedule_timesteps is not None and torch.is_tensor(schedule_timesteps) and schedule_timesteps.dtype.is_floating_point:
            timestep = torch.tensor(timestep, dtype=schedule_timesteps.dtype).item()       # compare in the schedule's dtype, as torch's == does
        hits = [i for i, u in enumerate(ts) if u == timestep]
        return hits[1 if len(hits) > 1 else 0]

    def _init_step_index(self, timestep):
        self._step_index = self.index_for_timestep(timestep) if self.begin_index is None else self._begin_index

    def scale_model_input(self, sample, *args, **kwargs):
        return sample

    def add_noise(self, original_samples, noise, timesteps):
        # fm_solvers_unipc.py:758-797: x_t = (1 - sigma) x_0 + sigma eps  ([B]-indexed scalars; cold path, torch ops)
        sigmas = self.sigmas.to(device=original_samples.device, dtype=original_samples.dtype)
        sched = self.timesteps.to(original_samples.device)
        timesteps = timesteps.to(original_samples.device)
        if self.begin_index is None:
            idx = [self.index_for_timestep(t, sched) for t in timesteps]
        elif self.step_index is not None:
            idx = [self.step_index] * timesteps.shape[0]
        else:
            idx = [self.begin_index] * timesteps.shape[0]
        sigma = sigmas[idx].flatten()
        while len(sigma.shape) < len(original_samples.shape):
            sigma = sigma.unsqueeze(-1)
        return (1 - sigma) * original_samples + sigma * noise

    def _hist(self):
        """previous x0 predictions, newest first"""
        return [m for m in reversed(self.model_outputs) if m is not None]

    def convert_model_output(self, model_output, *args, sample=None, **kwargs):
        """fm_solvers_unipc.py:279-321 (x0 = sample - sigma_t * v).  `step` fuses this; the stand-alone call exists for API
        parity and evaluates the same kernel with an identity predictor."""
        if sample is None:
            sample = args[1] if len(args) > 1 else None
        if sample is None:
            raise ValueError("missing `sample` as a required keyward argument")
        x0, _, _ = ops.unipc_step(sample.float().contiguous(), model_output.float().contiguous(), None, [],
                                  float(self.sigmas[self.step_index]), None, [1.0, 0.0, 0.0, 0.0, 0.0])
        return x0

    # ---- the step's scalars ---------------------------------------------------------------------------
    def _lam(self, s):
        return torch.log(1 - s) - torch.log(s)

    def _rb(self, rks, hh, order):
        # fm_solvers_unipc.py:431-453 / 566-588
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = hh if self.config.solver_type == "bh1" else torch.expm1(hh)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return torch.stack(R), torch.tensor(b), float(h_phi_1), float(B_h)

    def step_coefficients(self, step_index: int, use_corrector: bool, corr_order: Optional[int], pred_order: int) -> dict:
        """Fold one step (fm_solvers_unipc.py:318-321, 486-626, 350-484) into the coefficient vectors of prfl_unipc_step.
        Layout: corr = [last_sample, x0, hist0, hist1, hist2], pred = [sample', x0, hist0, hist1, hist2] with hist_k the
        k-th newest previous x0 prediction (before this step's own x0 is pushed).  Pure host arithmetic."""
        sig = self.sigmas
        i = step_index
        out = {"sigma": float(sig[i]), "corr": None}
        a1 = 0.0   # d corrected / d x0
        if use_corrector:
            oc = corr_order
            s_t, s_s0 = sig[i], sig[i - 1]
            lam_s0 = self._lam(s_s0)
            h = self._lam(s_t) - lam_s0
            rks = [(self._lam(sig[i - (k + 1)]) - lam_s0) / h for k in range(1, oc)] + [1.0]
            R, b, h_phi_1, B_h = self._rb(torch.tensor(rks), -h, oc)
            rhos = [0.5] if oc == 1 else [float(r) for r in torch.linalg.solve(R, b)]
            al = float(1 - s_t)
            c = [float(s_t / s_s0), -al * B_h * rhos[-1], -al * h_phi_1 + al * B_h * rhos[-1], 0.0, 0.0]
            for k in range(oc - 1):                    # D1_k = (m_{k+1} - m_0) / rk_k
                w = al * B_h * rhos[k] / float(rks[k])
                c[2] += w
                c[3 + k] -= w
            out["corr"] = c
            a1 = c[1]
        op = pred_order
        s_t, s_s0 = sig[i + 1], sig[i]
        lam_s0 = self._lam(s_s0)
        h = self._lam(s_t) - lam_s0
        rks = [(self._lam(sig[i - k]) - lam_s0) / h for k in range(1, op)] + [1.0]
        R, b, h_phi_1, B_h = self._rb(torch.tensor(rks), -h, op)
        al = float(1 - s_t)
        p = [float(s_t / s_s0), -al * h_phi_1, 0.0, 0.0, 0.0]
        if op > 1:
            rhos = [0.5] if op == 2 else [float(r) for r in torch.linalg.solve(R[:-1, :-1], b[:-1])]
            for k in range(op - 1):                    # D1_k = (hist_k - x0) / rk_k   (hist_k = old model_outputs[-(k+1)])
                w = al * B_h * rhos[k] / float(rks[k])
                p[1] += w
                p[2 + k] -= w
        out["pred"] = p
        # x0 = sample - sigma v; prev = p0 * (corrected | sample) + p1 * x0 + ..., corrected = ... + a1 * x0
        dx0 = p[1] + (p[0] * a1 if use_corrector else 0.0)
        out["d_model_output"] = -out["sigma"] * dx0
        out["d_sample"] = dx0 + (0.0 if use_corrector else p[0])
        return out

    # ---- the step ----------------------------------------------------------------------------------------
    def step(self, model_output: torch.Tensor, timestep: Union[int, torch.Tensor], sample: torch.Tensor, return_dict: bool = True,
             generator=None, model_output_uncond: Optional[torch.Tensor] = None, guide_scale: float = 1.0
             ) -> Union[SchedulerOutput, Tuple]:
        """fm_solvers_unipc.py:655-739.  Extension (no-grad only): with `model_output_uncond` the classifier-free-guidance
        combination uncond + guide_scale * (model_output - uncond) of the sampling loop (text2video.py:295-296) is
        evaluated inside the same kernel instead of as three more elementwise launches."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if self.step_index is None:
            self._init_step_index(timestep)
        i = self.step_index
        use_corrector = i > 0 and (i - 1) not in self.disable_corrector and self.last_sample is not None
        corr_order = self.this_order
        if self.config.lower_order_final:
            this_order = min(self.config.solver_order, len(self._timesteps_host) - i)
        else:
            this_order = self.config.solver_order
        pred_order = min(this_order, self.lower_order_nums + 1)
        assert pred_order > 0
        coef = self.step_coefficients(i, use_corrector, corr_order, pred_order)
        mo = model_output if model_output.dtype == torch.float32 else model_output.float()
        sm = sample if sample.dtype == torch.float32 else sample.float()
        mo, sm = mo.contiguous(), sm.contiguous()
        if torch.is_grad_enabled() and (mo.requires_grad or sm.requires_grad):
            if model_output_uncond is not None:
                raise NotImplementedError("classifier-free guidance inside step() is a no-grad (sampling) feature")
            prev, x0, new_sample = _StepFn.apply(mo, sm, self, coef)
        else:
            mu = None if model_output_uncond is None else model_output_uncond.float().contiguous()
            x0, corrected, prev = ops.unipc_step(sm, mo, self.last_sample, self._hist(), coef["sigma"], coef["corr"], coef["pred"],
                                                 mu, guide_scale)
            new_sample = corrected if corrected is not None else sm
        for k in range(self.config.solver_order - 1):
            self.model_outputs[k] = self.model_outputs[k + 1]
            self.timestep_list[k] = self.timestep_list[k + 1]
        self.model_outputs[-1] = x0.detach()
        self.timestep_list[-1] = timestep
        self.this_order = pred_order
        self.last_sample = new_sample.detach()
        if self.lower_order_nums < self.config.solver_order:
            self.lower_order_nums += 1
        self._step_index += 1
        prev = prev.to(sample.dtype)
        if not return_dict:
            return (prev,)
        return SchedulerOutput(prev_sample=prev)

"""CPU oracle for SURVEY.md §8 row a16: the differentiable scheduler step between the video model and the reward
model in PRFL — `FlowUniPCMultistepScheduler` (reference diffusers_lite/wan/utils/fm_solvers_unipc.py) — and the PRFL
loss glue (scripts/prfl/train_prfl.py:796-798).

TEST INFRASTRUCTURE ONLY: imported by tests/, tests/golden/make_golden.py and nothing else.  Pinned: the
golden fixture tests/golden/unipc.pt holds outputs of the UNMODIFIED reference scheduler run in the build container
(oracle/ref_shim.load_scheduler); tests/test_oracle_golden.py holds this restatement to 1e-6 of them.

Plain torch-CPU fp32 restatement, state kept in a small object, every function citing the lines it follows.  Only the
configuration the reference actually instantiates is restated (predict_x0, flow_prediction, bh1 / bh2, no thresholding,
no solver_p, final_sigmas_type "zero"; text2video.py:257-263, train_prfl.py:411-413).
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch


class UniPCOracle:
    def __init__(self, num_train_timesteps=1000, solver_order=2, shift=1.0, solver_type="bh2", lower_order_final=True,
                 disable_corrector=()):
        # fm_solvers_unipc.py:77-132
        self.num_train_timesteps = num_train_timesteps
        self.solver_order = solver_order
        self.shift = shift
        self.solver_type = solver_type
        self.lower_order_final = lower_order_final
        self.disable_corrector = list(disable_corrector)
        alphas = np.linspace(1, 1 / num_train_timesteps, num_train_timesteps)[::-1].copy()
        sigmas = torch.from_numpy(1.0 - alphas).to(torch.float32)
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        self.sigmas = sigmas
        self.sigma_min, self.sigma_max = sigmas[-1].item(), sigmas[0].item()
        self.timesteps = sigmas * num_train_timesteps
        self.reset()

    def reset(self):
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = None
        self.this_order = None

    def set_timesteps(self, num_inference_steps, shift=None):
        # fm_solvers_unipc.py:160-227 (sigmas=None, no dynamic shifting, final sigma 0; timesteps truncated to int64)
        sigmas = np.linspace(self.sigma_max, self.sigma_min, num_inference_steps + 1).copy()[:-1]
        if shift is None:
            shift = self.shift
        sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
        timesteps = sigmas * self.num_train_timesteps
        sigmas = np.concatenate([sigmas, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sigmas)
        self.timesteps = torch.from_numpy(timesteps).to(torch.int64)
        self.num_inference_steps = len(timesteps)
        self.reset()

    # ---- helpers -------------------------------------------------------------------------------
    def _lambda(self, sigma):
        # fm_solvers_unipc.py:272-273, 409-413: alpha = 1 - sigma, lambda = log(alpha) - log(sigma)
        return torch.log(1 - sigma) - torch.log(sigma)

    def _rb(self, rks, hh, order):
        # fm_solvers_unipc.py:431-453 / 566-588: R[i] = rks^i, b[i] = h phi_{i+1}(h) i! / B(h)
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = hh if self.solver_type == "bh1" else torch.expm1(hh)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(torch.pow(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return torch.stack(R), torch.tensor(b), h_phi_1, B_h

    def convert_model_output(self, model_output, sample):
        # fm_solvers_unipc.py:318-321: x0 = sample - sigma_t * v
        return sample - self.sigmas[self.step_index] * model_output

    def _uni_p(self, sample, order):
        # fm_solvers_unipc.py:350-484 (predict_x0 branch)
        m0, x = self.model_outputs[-1], sample
        sigma_t, sigma_s0 = self.sigmas[self.step_index + 1], self.sigmas[self.step_index]
        alpha_t = 1 - sigma_t
        lambda_s0 = self._lambda(sigma_s0)
        h = self._lambda(sigma_t) - lambda_s0
        rks, D1s = [], []
        for i in range(1, order):
            mi = self.model_outputs[-(i + 1)]
            rk = (self._lambda(self.sigmas[self.step_index - i]) - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks)
        hh = -h
        R, b, h_phi_1, B_h = self._rb(rks, hh, order)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        if D1s:
            D1s = torch.stack(D1s, dim=1)
            rhos_p = torch.tensor([0.5], dtype=x.dtype) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1]).to(x.dtype)
            pred_res = torch.einsum("k,bkc...->bc...", rhos_p, D1s)
        else:
            pred_res = 0
        return (x_t_ - alpha_t * B_h * pred_res).to(x.dtype)

    def _uni_c(self, this_model_output, last_sample, this_sample, order):
        # fm_solvers_unipc.py:486-626 (predict_x0 branch)
        m0, x, model_t = self.model_outputs[-1], last_sample, this_model_output
        sigma_t, sigma_s0 = self.sigmas[self.step_index], self.sigmas[self.step_index - 1]
        alpha_t = 1 - sigma_t
        lambda_s0 = self._lambda(sigma_s0)
        h = self._lambda(sigma_t) - lambda_s0
        rks, D1s = [], []
        for i in range(1, order):
            mi = self.model_outputs[-(i + 1)]
            rk = (self._lambda(self.sigmas[self.step_index - (i + 1)]) - lambda_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(1.0)
        rks = torch.tensor(rks)
        hh = -h
        R, b, h_phi_1, B_h = self._rb(rks, hh, order)
        rhos_c = torch.tensor([0.5], dtype=x.dtype) if order == 1 else torch.linalg.solve(R, b).to(x.dtype)
        x_t_ = sigma_t / sigma_s0 * x - alpha_t * h_phi_1 * m0
        corr_res = torch.einsum("k,bkc...->bc...", rhos_c[:-1], torch.stack(D1s, dim=1)) if D1s else 0
        return (x_t_ - alpha_t * B_h * (corr_res + rhos_c[-1] * (model_t - m0))).to(x.dtype)

    def step(self, model_output, timestep, sample):
        # fm_solvers_unipc.py:655-739
        if self.step_index is None:
            idx = (self.timesteps == timestep).nonzero()          # :628-641
            self.step_index = idx[1 if len(idx) > 1 else 0].item()
        use_corrector = self.step_index > 0 and (self.step_index - 1) not in self.disable_corrector and self.last_sample is not None
        m_conv = self.convert_model_output(model_output, sample)
        if use_corrector:
            sample = self._uni_c(m_conv, self.last_sample, sample, self.this_order)
        for i in range(self.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = m_conv
        this_order = min(self.solver_order, len(self.timesteps) - self.step_index) if self.lower_order_final else self.solver_order
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev = self._uni_p(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev


def prfl_loss(reward_scores: torch.Tensor, target_reward: float = 2.0, weight: float = 0.1) -> torch.Tensor:
    """train_prfl.py:796-798: loss = 0.1 * relu(target - r).mean()"""
    return weight * torch.relu(-reward_scores.squeeze() + target_reward).mean()


def toy_velocity(x: torch.Tensor, t, w: torch.Tensor) -> torch.Tensor:
    """A cheap differentiable stand-in for the DiT in scheduler tests: v = tanh(x * w) * (0.3 + t / 2000)."""
    return torch.tanh(x * w) * (0.3 + float(t) / 2000.0)

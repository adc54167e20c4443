"""PRFL chain glue: the hot part of the reference's `train_step_refl`
(scripts/prfl/train_prfl.py:631-798) — m no-grad denoising steps, one denoising step with autograd, the differentiable
scheduler step, the frozen reward model on the next latent, and the PRFL loss — expressed over the drop-in modules
(`WanModel`, `QueryAttention`, `MLP`, `FlowUniPCMultistepScheduler`).  The trainer's control plane (data, logging,
empty_cache / gc / barrier calls, optimizer scheduling) is out of scope and stays in the reference's script; what is
dropped here on purpose are exactly those stalls (train_prfl.py:645-646, 694-699, 738-741: ~10 synchronising calls per
step) — the chain below never synchronises the host with the device.

    loss, reward = refl_chain(transformer, lrm_transformer, query_attention, mlp, noise_scheduler, latent, text_states,
                              max_sequence_length, mid_timestep, flow_shift=5.0, feature_layer=[8])
    loss.backward()
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from .network import forward_mlp
from .scheduler import prfl_loss

__all__ = ["batch2list", "list2batch", "refl_chain"]


def batch2list(batch):
    """diffusers_lite/utils/diffusion_utils.py:378-379"""
    return [item for item in batch]


def list2batch(items):
    """diffusers_lite/utils/diffusion_utils.py:381-382"""
    return torch.stack(items)


def refl_chain(transformer, lrm_transformer, query_attention, mlp, noise_scheduler, latent: torch.Tensor,
               text_states: torch.Tensor, max_sequence_length: int, mid_timestep: int, *, image_embeds=None,
               latents_condition: Optional[torch.Tensor] = None, inference_steps: int = 40, flow_shift: float = 5.0,
               feature_layer: Sequence[int] = (8,), target_reward: float = 2.0, marks: Optional[dict] = None
               ) -> Tuple[torch.Tensor, torch.Tensor]:
    """latent [B, 16, F, H, W] fp32 noise; text_states [B, <=512, 4096]; returns (loss, reward_scores).
    `marks`, if given, receives CUDA events at the phase boundaries (for tools/prfl_step.py); recording an event does not
    synchronise."""
    dev = latent.device

    def mark(name):
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks[name] = e

    noise_scheduler.set_timesteps(num_inference_steps=inference_steps, device=dev, shift=flow_shift)      # :632
    timesteps = noise_scheduler.timesteps
    host_t = [int(t) for t in timesteps.tolist()]                   # one D2H copy per chain instead of one per step
    cond = dict(context=batch2list(text_states), seq_len=max_sequence_length, clip_fea=image_embeds,
                y=batch2list(latents_condition) if latents_condition is not None else None)
    mark("start")
    # 1. no-grad denoising to the mid timestep (:665-699); the prompt's embeddings and cross-attention K/V are constants of
    #    these m forwards (weights only change at optimizer.step), so they are computed once (WanModel.prepare_context)
    with torch.no_grad():
        if mid_timestep > 0 and hasattr(transformer, "prepare_context"):
            cond_ng = dict(cond, context=transformer.prepare_context(cond["context"], image_embeds), clip_fea=None)
        else:
            cond_ng = cond
        for i in range(mid_timestep):
            t = torch.tensor([host_t[i]], device=dev)
            noise_pred = list2batch(transformer(x=batch2list(latent), t=t, cond_flag=True, **cond_ng))
            latent = noise_scheduler.step(noise_pred, host_t[i], latent, return_dict=False)[0]
    mark("nograd_done")
    # 2. the step whose gradient is kept (:703-725)
    t_mid = torch.tensor([host_t[mid_timestep]], device=dev)
    noise_pred = list2batch(transformer(x=batch2list(latent), t=t_mid, cond_flag=True, **cond))
    mark("grad_fwd_done")
    # 3. differentiable scheduler step (:734-735)
    latent = noise_scheduler.step(noise_pred, host_t[mid_timestep], latent, return_dict=False)[0]
    # 4. frozen reward model on the next latent at the next timestep (:745-798)
    t_next = torch.tensor([host_t[mid_timestep + 1]], device=dev)
    feats = lrm_transformer(x=batch2list(latent), t=t_next, output_features=True, selected_layers=list(feature_layer), **cond)
    feats = list2batch(feats)
    reward = forward_mlp(mlp, query_attention(feats))
    loss = prfl_loss(reward, target_reward)
    mark("lrm_fwd_done")
    return loss, reward

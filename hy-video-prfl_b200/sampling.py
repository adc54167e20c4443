"""The denoising loop of `WanT2V.generate` / `WanI2V.generate` (diffusers_lite/wan/text2video.py:253-304,
image2video.py:309-388) over the drop-in modules: `sampling_steps` x (conditional forward, unconditional forward,
classifier-free guidance, FlowUniPC step).  Everything before (T5 / CLIP / VAE encode, noise) and after (VAE decode) the
loop is outside the hot path and stays in the reference's pipeline classes, which can call this instead of their loop.

B200-side differences, all exact (same values as evaluating the reference's loop with these modules):
  * the two prompts are embedded once (`WanModel.prepare_context`) and every block's cross-attention K / V of the text /
    CLIP context is computed on the first step only — the reference recomputes them in each of the 2 x steps forwards;
  * guidance and the scheduler update are one kernel (`prfl_unipc_step` with the unconditional output as an extra operand)
    instead of 3 + ~25 elementwise launches;
  * no host-device synchronisation inside the loop (timesteps are read back once);
  * `batch_cfg=True` (default): the conditional and the unconditional branch run as ONE forward with batch 2
    (x = [latent, latent], context = [prompt, negative prompt]) — one pass over the block list, the time-embedding MLP and
    the per-block modulation computed once per step instead of twice; per-sample kernels are unchanged, so the result is
    bit-identical to the two sequential forwards of the reference (text2video.py:290-293).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .scheduler import FlowUniPCMultistepScheduler

__all__ = ["sample_loop"]


@torch.no_grad()
def sample_loop(model, noise: torch.Tensor, context: Sequence[torch.Tensor], context_null: Sequence[torch.Tensor], seq_len: int, *,
                sampling_steps: int = 50, shift: float = 5.0, guide_scale: float = 5.0, clip_fea: Optional[torch.Tensor] = None,
                y: Optional[List[torch.Tensor]] = None, num_train_timesteps: int = 1000, sample_solver: str = "unipc",
                cache_context: bool = True, trajectory: Optional[list] = None, batch_cfg: bool = True) -> List[torch.Tensor]:
    """noise: [16, F, H, W] fp32 latent; context / context_null: lists with one [<=512, 4096] tensor (T5 states of the
    prompt / the negative prompt); clip_fea [1, 257, 1280] and y = [[20, F, H, W]] for image-to-video.  Returns
    `[x0]` like the reference (`x0 = latents`, text2video.py:306).  `trajectory`, if a list, receives every step's latent."""
    if sample_solver != "unipc":
        raise NotImplementedError("Unsupported solver.")            # the reference's other option (dpm++) is not on this path
    dev = noise.device
    scheduler = FlowUniPCMultistepScheduler(num_train_timesteps=num_train_timesteps, shift=1, use_dynamic_shifting=False)
    scheduler.set_timesteps(sampling_steps, device=dev, shift=shift)
    host_t = [int(t) for t in scheduler.timesteps.tolist()]
    if batch_cfg:
        assert len(context) == 1 and len(context_null) == 1, "batch_cfg pairs one prompt with one negative prompt"
        clip2 = None if clip_fea is None else torch.cat([clip_fea, clip_fea], dim=0)
        y2 = None if y is None else [y[0], y[0]]
        if cache_context:
            arg_b = dict(context=model.prepare_context([context[0], context_null[0]], clip2), seq_len=seq_len, y=y2, cond_flag=True)
        else:
            arg_b = dict(context=[context[0], context_null[0]], clip_fea=clip2, seq_len=seq_len, y=y2, cond_flag=True)
    elif cache_context:
        ctx_c, ctx_n = model.prepare_context(context, clip_fea), model.prepare_context(context_null, clip_fea)
        arg_c = dict(context=ctx_c, seq_len=seq_len, y=y, cond_flag=True)
        arg_null = dict(context=ctx_n, seq_len=seq_len, y=y, cond_flag=False)
    else:
        arg_c = dict(context=context, clip_fea=clip_fea, seq_len=seq_len, y=y, cond_flag=True)
        arg_null = dict(context=context_null, clip_fea=clip_fea, seq_len=seq_len, y=y, cond_flag=False)
    latent = noise
    for t in host_t:
        if batch_cfg:
            cond, uncond = model([latent, latent], t=torch.tensor([t, t], device=dev), **arg_b)
        else:
            timestep = torch.tensor([t], device=dev)
            cond = model([latent], t=timestep, **arg_c)[0]
            uncond = model([latent], t=timestep, **arg_null)[0]
        # noise_pred = uncond + guide_scale * (cond - uncond); latent = scheduler.step(noise_pred, t, latent)  -- one kernel
        latent = scheduler.step(cond.unsqueeze(0), t, latent.unsqueeze(0), return_dict=False, model_output_uncond=uncond.unsqueeze(0),
                                guide_scale=guide_scale)[0].squeeze(0)
        if trajectory is not None:
            trajectory.append(latent)
    return [latent]

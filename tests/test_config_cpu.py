"""CPU: the `configs/*.yaml` surface mapped onto the drop-in objects by `prfl_b200.config`.
An inline YAML with the shipped files' hot-path keys runs everywhere; in the build container every one of the reference's own
`configs/train_*.yaml` / `infer_*.yaml` is loaded too (skipped where /root/reference does not exist)."""
import glob
import os

import pytest
import torch

from oracle import ref_shim

YAML = """
task: "i2v-14b-720p"
model:
  base_path: weights/Wan2.1-I2V-14B-720P
  lrm_mlp_path: null
  patch_size: [1, 2, 2]
  fsdp:
    fsdp_sharding_startegy: full
  gradient_checkpointing: true
  selective_checkpointing: 1.0
extra_model:
  scheduler:
    flow_shift: 5.0
    num_train_timesteps: 1000
dataset:
  sp_size: 4
optimizer:
  learning_rate: 5e-6
  adam_beta1: 0.9
  adam_beta2: 0.999
  weight_decay: 0.01
train:
  seed: 110221
  precision: bf16
  gradient_accumulation_steps: 5.
lrm:
  query_attention:
    num_queries: 1
    num_heads: 8
    dropout: 0.
    return_type: query
  feature_layer: [8]
  pool: q_attn
  mlp_dim: 256
  trainable_blocks: [0, 1]
"""


def _check_common(cfg):
    from prfl_b200 import config as C
    assert isinstance(cfg.optimizer.learning_rate, float) and 0 < cfg.optimizer.learning_rate < 1e-3     # `5e-6` is a number, not a string
    assert C.accumulation_steps(cfg) >= 1 and isinstance(C.accumulation_steps(cfg), int)
    kw = C.refl_kwargs(cfg)
    assert kw["flow_shift"] == float(cfg.extra_model.scheduler.flow_shift) and all(isinstance(v, int) for v in kw["feature_layer"])
    arch = C.arch_from_task(cfg.task)
    assert arch["dim"] // arch["num_heads"] == 128 and arch["model_type"] in ("t2v", "i2v", "flf2v")
    s = C.scheduler_from_config(cfg)
    assert s.config.num_train_timesteps == cfg.extra_model.scheduler.num_train_timesteps and s.config.solver_order == 2
    s.set_timesteps(num_inference_steps=40, device="cpu", shift=kw["flow_shift"])                        # train_prfl.py:632
    assert len(s.timesteps) == 40
    assert int(cfg.dataset.sp_size) >= 1 and hasattr(cfg.lrm, "trainable_blocks")


def test_inline_yaml_drives_the_drop_in_objects(tmp_path):
    from prfl_b200 import config as C
    from prfl_b200.model import WanModel
    from prfl_b200.network import MLP, QueryAttention
    from prfl_b200.sharding import ShardedAdamW
    p = tmp_path / "train_prfl.yaml"
    p.write_text(YAML)
    cfg = C.load_config(str(p))
    _check_common(cfg)
    assert cfg.optimizer.learning_rate == 5e-6 and C.accumulation_steps(cfg) == 5 and cfg.model.lrm_mlp_path is None
    assert getattr(cfg.model, "resume_transformer_path", None) is None and not hasattr(cfg.model, "nope")
    qa, mlp = C.reward_head_from_config(cfg)
    assert isinstance(qa, QueryAttention) and isinstance(mlp, MLP) and not qa.training and not mlp.training
    assert (qa.feature_dim, qa.num_queries, qa.num_heads, qa.return_type, qa.multihead_attn.dropout) == (256, 1, 8, "query", 0.0)
    # the reward transformer as train_prfl.py:219-258 cuts it
    m = WanModel(model_type="i2v", in_dim=36, dim=256, ffn_dim=512, num_heads=2, num_layers=3, text_dim=64)
    kept = [m.blocks[0], m.blocks[1]]
    C.truncate_reward_transformer(m, cfg)
    assert len(m.blocks) == 2 and m.blocks[0] is kept[0] and m.blocks[1] is kept[1] and m.head is None
    assert not any(p_.requires_grad for p_ in m.parameters())                                            # PRFL: dgrad-only reward model
    m2 = WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=3, text_dim=64)
    C.truncate_reward_transformer(m2, cfg, trainable=True)                                               # PAVRM training
    assert all(p_.requires_grad for p_ in m2.blocks.parameters()) and not any(p_.requires_grad for p_ in m2.patch_embedding.parameters())
    # the optimizer that replaces the FSDP-wrapped AdamW (CPU toy: flat fp32 units)
    toy = torch.nn.Module()
    toy.blocks = torch.nn.ModuleList([torch.nn.Linear(4, 4)])
    opt = C.optimizer_from_config(cfg, toy)
    assert isinstance(opt, ShardedAdamW) and (opt.lr, tuple(opt.betas), opt.wd, opt.eps) == (5e-6, (0.9, 0.999), 0.01, 1e-8)
    # feature_layer default when the YAML has none (train_prfl.py:233-235)
    del cfg.lrm["feature_layer"]
    assert C.refl_kwargs(cfg)["feature_layer"] == [6, 7]
    with pytest.raises(KeyError):
        C.arch_from_task("t2v-7b")


@pytest.mark.skipif(not ref_shim.available(), reason="needs the reference checkout (/root/reference)")
def test_every_shipped_reference_yaml_loads():
    from prfl_b200 import config as C
    files = sorted(glob.glob(os.path.join(ref_shim.REF, "configs", "train_*.yaml")) + glob.glob(os.path.join(ref_shim.REF, "configs", "infer_*.yaml")))
    assert len(files) >= 9
    for f in files:
        cfg = C.load_config(f)
        if "optimizer" in cfg:
            _check_common(cfg)
        qa, mlp = C.reward_head_from_config(cfg)
        assert qa.feature_dim == cfg.lrm.mlp_dim == 5120 and qa.return_type == "query" and mlp.fc1.in_features == 5120
        assert list(cfg.lrm.feature_layer) == [8] and int(cfg.dataset.sp_size) == 4

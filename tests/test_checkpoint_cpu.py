"""CPU: checkpoint wire formats.  The on-disk layout is the reference's (model_utils.py:70-141): names,
greedy 5 GiB sharding over sorted keys, index JSON, config.json.  Pinned against the committed golden listing the
UNMODIFIED reference `save_checkpoint` produced for the same tiny model (tests/golden/checkpoint_ref.json, generator in
tests/golden/make_golden.py: FSDP.state_dict_type replaced by a null context — a single process holds the full state),
and against the restated sharding rule for the > max_bytes branch (the reference hard-codes 5 GiB)."""
import json
import os

import torch

from conftest import GOLDEN
from oracle import synth


def _tiny():
    from prfl_b200.model import WanModel
    cfg = synth.tiny_cfg("t2v")
    m = WanModel(**cfg.kwargs())
    m.load_state_dict(synth.make_wan_state_dict(cfg, 7), strict=True)
    m.config = cfg.kwargs() | {"dtype": "bf16"}
    return m, cfg


def test_single_file_layout_matches_reference(tmp_path):
    from safetensors import safe_open
    from prfl_b200.checkpoint import load_state_dict, save_checkpoint
    ref = json.load(open(os.path.join(GOLDEN, "checkpoint_ref.json")))
    m, cfg = _tiny()
    d = save_checkpoint(m, 0, str(tmp_path), 7)
    assert os.path.basename(d) == ref["dir"] and sorted(os.listdir(d)) == ref["files"]
    with safe_open(os.path.join(d, "diffusion_pytorch_model.safetensors"), "pt") as f:
        keys = sorted(f.keys())
    assert keys == ref["keys"]
    assert json.load(open(os.path.join(d, "config.json"))) == ref["config"]           # "dtype" dropped, as the reference does
    back = load_state_dict(d)
    for k, v in m.state_dict().items():
        assert torch.equal(back[k], v), k
    assert save_checkpoint(m, 1, str(tmp_path), 8) is None and not os.path.exists(tmp_path / "checkpoint-8")   # rank > 0 writes nothing
    assert os.path.basename(save_checkpoint(m, 0, str(tmp_path), 9, ema=True)) == "checkpoint-9-ema"


def test_sharded_layout_and_roundtrip(tmp_path):
    from prfl_b200.checkpoint import load_state_dict, save_checkpoint
    m, _ = _tiny()
    sd = m.state_dict()
    max_bytes = 600_000
    d = save_checkpoint(m, 0, str(tmp_path), 3, max_bytes=max_bytes)
    files = sorted(n for n in os.listdir(d) if n.endswith(".safetensors"))
    # restatement of model_utils.py:92-101: sorted keys, greedy fill, a tensor larger than the limit gets its own shard
    expect, cur, size = [], [], 0
    for k in sorted(sd):
        b = sd[k].numel() * sd[k].element_size()
        if size + b > max_bytes and cur:
            expect.append(cur)
            cur, size = [], 0
        cur.append(k)
        size += b
    expect.append(cur)
    n = len(expect)
    assert n > 2 and files == [f"diffusion_pytorch_model-{i:05}-of-{n:05}.safetensors" for i in range(1, n + 1)]
    idx = json.load(open(os.path.join(d, "diffusion_pytorch_model.safetensors.index.json")))
    assert idx["metadata"]["total_size"] == sum(v.numel() * v.element_size() for v in sd.values())
    for i, ks in enumerate(expect, start=1):
        for k in ks:
            assert idx["weight_map"][k] == f"diffusion_pytorch_model-{i:05}-of-{n:05}.safetensors"
    back = load_state_dict(d)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)


def test_optimizer_shard_roundtrip(tmp_path):
    import torch.nn as nn
    from prfl_b200.checkpoint import load_optimizer, load_state_dict, save_checkpoint, save_optimizer
    from prfl_b200.sharding import ShardedAdamW

    def make():
        torch.manual_seed(3)
        m = nn.Module()
        m.blocks = nn.ModuleList([nn.Linear(6, 6) for _ in range(2)])
        m.head = nn.Linear(6, 2)
        m.config = {"dim": 6}
        return m

    def run(m, opt, steps, seed):
        g = torch.Generator().manual_seed(seed)
        for _ in range(steps):
            x = torch.randn(4, 6, generator=g)
            for b in m.blocks:
                x = x + b(x)
            m.head(x).pow(2).mean().backward()
            opt.step(max_norm=1.0)

    a = make()
    oa = ShardedAdamW(a, lr=1e-2, weight_decay=0.01)
    run(a, oa, 3, 5)
    d = save_checkpoint(a, 0, str(tmp_path), 3)
    p = save_optimizer(oa, str(tmp_path), 3)
    assert os.path.basename(p) == "optimizer-rank00000-of-00001.safetensors" and os.path.dirname(p) == d
    b = make()
    b.load_state_dict(load_state_dict(d), strict=True)              # the optimizer shard is not mistaken for weights
    ob = ShardedAdamW(b, lr=1e-2, weight_decay=0.01)
    load_optimizer(ob, d)
    run(a, oa, 2, 6)
    run(b, ob, 2, 6)
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.equal(pa, pb)                                   # resumed run == uninterrupted run, bit for bit


def test_from_pretrained_reads_what_save_checkpoint_writes(tmp_path):
    """`WanModel.from_pretrained(dir)` is how every reference trainer obtains its models (train_prfl.py:182-217); the directory
    it reads is the one `save_checkpoint` writes (single file and sharded), plus published Wan2.1-style config files that omit
    the constructor arguments the reference lists in `ignore_for_config` and carry `_class_name` / `_diffusers_version`."""
    import pytest
    from prfl_b200.checkpoint import save_checkpoint
    from prfl_b200.model import WanModel
    m, cfg = _tiny()
    sd = m.state_dict()
    for kw in ({}, {"max_bytes": 600_000}):                                  # one file / index + shards
        d = save_checkpoint(m, 0, str(tmp_path / ("s" if kw else "o")), 11, **kw)
        # an optimizer shard in the same directory is not a weight file
        torch.save({}, os.path.join(d, "scratch.bin"))
        b = WanModel.from_pretrained(d)
        assert not b.training and dict(b.config) == {k: v for k, v in m.config.items() if k != "dtype"}
        got = b.state_dict()
        assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    # a published-style config.json: ignored keys absent, bookkeeping keys present, an annotation the trainers add
    ignored = ("patch_size", "cross_attn_norm", "qk_norm", "text_dim", "window_size")
    full = json.load(open(os.path.join(d, "config.json")))
    slim = {k: v for k, v in full.items() if k not in ignored} | {"_class_name": "WanModel", "_diffusers_version": "0.30.0", "lora_rank": 4}
    # text_dim is not the default in the tiny model: keep it (a real 14B config relies on the default 4096)
    slim["text_dim"] = full["text_dim"]
    json.dump(slim, open(os.path.join(d, "config.json"), "w"))
    c = WanModel.from_pretrained(str(tmp_path / "s"), subfolder="checkpoint-11", torch_dtype=torch.bfloat16)
    assert c.config.lora_rank == 4 and c.config["patch_size"] == (1, 2, 2) and "_class_name" not in c.config
    assert c.blocks[0].ffn[0].weight.dtype == torch.bfloat16
    assert torch.equal(c.blocks[0].ffn[0].weight, sd["blocks.0.ffn.0.weight"].bfloat16())
    c.config.lora_alpha = 8                                                  # train_prfl.py:355-357 assigns attributes
    assert dict(c.config)["lora_alpha"] == 8
    # strict by default; strict=False reports like load_state_dict (train_prfl.py:209 uses the non-strict form by hand)
    from safetensors.torch import load_file, save_file
    one = str(tmp_path / "o" / "checkpoint-11")
    st = load_file(os.path.join(one, "diffusion_pytorch_model.safetensors"))
    st.pop("blocks.1.modulation")
    st["not.a.key"] = torch.zeros(1)
    save_file(st, os.path.join(one, "diffusion_pytorch_model.safetensors"))
    with pytest.raises(RuntimeError, match="1 missing.*1 unexpected"):
        WanModel.from_pretrained(one)
    _, missing, unexpected = WanModel.from_pretrained(one, strict=False)
    assert missing == ["blocks.1.modulation"] and unexpected == ["not.a.key"]
    with pytest.raises(FileNotFoundError):
        WanModel.from_pretrained(str(tmp_path / "nowhere"))


def test_save_pretrained_roundtrip_and_reference_loader(tmp_path):
    """save_pretrained -> from_pretrained, and the reference's own `load_state_dict(model_dir)` (model_utils.py:128-141, restated
    in checkpoint.load_state_dict) reads the same directory."""
    from prfl_b200.checkpoint import load_state_dict
    from prfl_b200.model import WanModel
    m, _ = _tiny()
    m.save_pretrained(tmp_path / "p", max_shard_size=600_000)
    cfg = json.load(open(tmp_path / "p" / "config.json"))
    assert cfg["_class_name"] == "WanModel" and "dtype" not in cfg and cfg["dim"] == m.config["dim"]
    b = WanModel.from_pretrained(tmp_path / "p")
    sd, back = m.state_dict(), load_state_dict(str(tmp_path / "p"))
    assert all(torch.equal(b.state_dict()[k], sd[k]) and torch.equal(back[k], sd[k]) for k in sd)


def test_update_ema_model_rule_and_cache_invalidation():
    """model_utils.py:172-175 restated: trainable parameters only, p_ema <- d * p_ema + (1 - d) * p; the operand caches of the
    averaged model are invalidated (the rule writes through `.data`); bf16 averaged parameters are refused."""
    import pytest
    from prfl_b200 import model as Mdl
    from prfl_b200.checkpoint import update_ema_model
    a, b = torch.nn.Linear(3, 3), torch.nn.Linear(3, 3)
    a.bias.requires_grad_(False)
    wa, wb, bb = a.weight.detach().clone(), b.weight.detach().clone(), b.bias.detach().clone()
    epoch = Mdl._WEIGHT_EPOCH[0]
    update_ema_model(a, b, 0.9)
    assert torch.allclose(b.weight, 0.9 * wb + 0.1 * wa) and torch.equal(b.bias, bb) and Mdl._WEIGHT_EPOCH[0] == epoch + 1
    with pytest.raises(NotImplementedError):
        update_ema_model(a, torch.nn.Linear(3, 3).bfloat16(), 0.9)

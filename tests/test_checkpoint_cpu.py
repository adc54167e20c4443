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

"""`prfl_b200.plugin.install()` on the UNMODIFIED reference WanModel (build container only: needs
/root/reference; skipped elsewhere).  No device here, so what is checked is the contract of the patch itself: every
reference block gets a shadow that shares the reference's own Parameters (same names, same tensors, no copies), the
ModulePlugin convention (`old_forward` kept, enable flag) holds, the model's state dict is unchanged, and uninstall
restores the reference forward.  The numerics of the installed path are covered on the GPU (tests/test_plugin_gpu.py)."""
import pytest
import torch

from oracle import ref_shim, synth

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="needs the reference checkout (/root/reference)")


@pytest.mark.parametrize("mt", ["t2v", "i2v"])
def test_install_on_real_reference_model(mt):
    from prfl_b200.plugin import install, uninstall
    M, _ = ref_shim.load()
    cfg = synth.tiny_cfg(mt)
    ref = M.WanModel(**cfg.kwargs())
    ref.load_state_dict(synth.make_wan_state_dict(cfg, 3), strict=True)
    keys = list(ref.state_dict())
    cls_forward = type(ref.blocks[0]).forward
    install(ref)
    assert list(ref.state_dict()) == keys                                  # no duplicate / extra keys
    for blk in ref.blocks:
        fast = blk._prfl_b200_fast
        mine, theirs = dict(fast.named_parameters()), dict(blk.named_parameters())
        assert set(mine) == set(theirs)
        assert all(mine[k] is theirs[k] for k in mine)                       # the SAME Parameter objects
        assert blk.old_forward.__func__ is cls_forward and blk.forward is not blk.old_forward
    install(ref)                                                            # idempotent
    # the flag routes back to the reference forward (CPU can run that one)
    ref.prfl_b200_enable(False)
    inp = synth.make_inputs(cfg, (2, 4, 4), 4)
    kw = dict(clip_fea=inp["clip_fea"], y=inp["y"]) if mt != "t2v" else {}
    with torch.no_grad():
        a = ref(inp["x"], t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], **kw)[0]
    uninstall(ref)
    assert not hasattr(ref.blocks[0], "old_forward") and ref.blocks[0].forward.__func__ is cls_forward
    with torch.no_grad():
        b = ref(inp["x"], t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], **kw)[0]
    assert torch.equal(a, b)
    # enabled without a GPU: the B200 path refuses instead of falling back
    install(ref)
    if not torch.cuda.is_available():
        from prfl_b200._lib import PrflError
        with pytest.raises(PrflError):
            with torch.no_grad():
                ref(inp["x"], t=inp["t"], context=inp["context"], seq_len=inp["seq_len"], **kw)


def test_patched_model_adopts_the_reference_sequence_parallel_state():
    """A reference trainer initialises the REFERENCE's `parallel_states` (train_prfl.py:119), not this package's; the patched
    blocks read this package's.  install() / the patched block 0 mirror the reference's state (flag, group, sizes, ranks), follow
    it when it changes, and leave a natively initialised state (class-swap route) alone."""
    import importlib
    from prfl_b200 import parallel
    ref_shim.load()
    ps = importlib.import_module("diffusers_lite.utils.parallel_states")
    saved = (ps._SEQUENCE_PARALLEL_STATE, ps.nccl_info.group, ps.nccl_info.sp_size, ps.nccl_info.rank_within_group, ps.nccl_info.group_id,
             ps.nccl_info.global_rank)
    try:
        assert parallel.adopt_reference_state() is False and not parallel.get_sequence_parallel_state()   # both off: nothing to do
        group = object()                                                    # stands in for the ProcessGroup the reference created
        ps.set_sequence_parallel_state(True)
        ps.nccl_info.group, ps.nccl_info.sp_size, ps.nccl_info.rank_within_group = group, 4, 3
        ps.nccl_info.group_id, ps.nccl_info.global_rank = 1, 7
        assert parallel.adopt_reference_state() is True
        n = parallel.nccl_info
        assert parallel.get_sequence_parallel_state() and n.group is group and (n.sp_size, n.rank_within_group, n.group_id, n.global_rank) == (4, 3, 1, 7)
        assert (n.ulysses_degree, n.ring_degree, n.ulysses_rank) == (4, 1, 3) and n.ulysses_group is group
        assert parallel.adopt_reference_state() is False                    # in sync: a no-op
        ps.set_sequence_parallel_state(False)                               # the mirror follows the reference back
        assert parallel.adopt_reference_state() is True and not parallel.get_sequence_parallel_state() and parallel.nccl_info.sp_size == 1
        # a state this package was given directly is not undone by an idle reference module
        parallel.set_sequence_parallel_state(True)
        assert parallel.adopt_reference_state() is False and parallel.get_sequence_parallel_state()
    finally:
        parallel.set_sequence_parallel_state(False)
        parallel._adopted[0] = False
        (ps._SEQUENCE_PARALLEL_STATE, ps.nccl_info.group, ps.nccl_info.sp_size, ps.nccl_info.rank_within_group, ps.nccl_info.group_id,
         ps.nccl_info.global_rank) = saved


def test_install_refuses_wrapped_blocks():
    """Under the reference's FSDP / checkpoint wrap (train_prfl.py:346-362) a block's parameters are replaced on every forward;
    install() must refuse rather than share stale tensors."""
    from torch.distributed.algorithms._checkpoint.checkpoint_wrapper import checkpoint_wrapper
    from prfl_b200.plugin import install
    M, _ = ref_shim.load()
    cfg = synth.tiny_cfg("t2v")
    ref = M.WanModel(**cfg.kwargs())
    ref.blocks[1] = checkpoint_wrapper(ref.blocks[1])
    with pytest.raises(RuntimeError, match="wrapped"):
        install(ref)
    assert not hasattr(ref.blocks[0], "_prfl_b200_fast")              # nothing was patched


@pytest.mark.parametrize("mt", ["t2v", "i2v"])
def test_installed_path_matches_the_reference_forward_and_trains_its_parameters(mt, monkeypatch):
    """Route B end to end on the REAL reference model (CPU, kernels emulated by tests/ops_emulator.py): the reference's own
    forward — its patch embedding, time / text / CLIP embeddings, autocast blocks, head and unpatchify — drives the patched
    blocks with exactly the arguments it passes its own; outputs must match the unpatched reference within north_star's
    bound, and a backward through the patched model must leave gradients on the reference's own Parameters."""
    import ops_emulator
    from conftest import cos_rel
    from prfl_b200 import model as pm
    from prfl_b200.plugin import install, uninstall
    M, _ = ref_shim.load()
    ops_emulator.install(monkeypatch)
    pm.bump_weight_epoch()
    cfg = synth.tiny_cfg(mt)
    ref = M.WanModel(**cfg.kwargs())
    sd = synth.make_wan_state_dict(cfg, 3)
    g = torch.Generator().manual_seed(4)
    sd["head.head.weight"] = torch.randn(sd["head.head.weight"].shape, generator=g) * 0.02
    ref.load_state_dict(sd, strict=True)
    inp = synth.make_inputs(cfg, (3, 8, 12), 5)
    kw = dict(t=inp["t"], context=inp["context"], seq_len=inp["seq_len"])
    if mt != "t2v":
        kw.update(clip_fea=inp["clip_fea"], y=inp["y"])
    with torch.no_grad():
        want = ref(inp["x"], **kw)[0]
        want_f = ref(inp["x"], **kw, output_features=True, selected_layers=[2])[0]
    install(ref)
    del ops_emulator.CALLS[:]
    with torch.no_grad():
        got = ref(inp["x"], **kw)[0]
        got_f = ref(inp["x"], **kw, output_features=True, selected_layers=[2])[0]
    assert ops_emulator.CALLS.count("attn_fwd") >= 2 * 2 * len(ref.blocks)           # the patched blocks did the work
    for a, b, what in ((got, want, "noise_pred"), (got_f, want_f, "features")):
        c, r = cos_rel(a, b)
        assert a.shape == b.shape and c >= 0.999 and r <= 2e-2, (what, c, r)
    ref.train()
    x = [u.clone().requires_grad_(True) for u in inp["x"]]
    ref(x, **kw)[0].square().sum().backward()
    named = dict(ref.named_parameters())
    for k in ("blocks.0.self_attn.q.weight", "blocks.1.ffn.2.bias", "blocks.0.modulation", "patch_embedding.weight", "time_projection.1.weight"):
        assert named[k].grad is not None and float(named[k].grad.abs().max()) > 0, k
    assert x[0].grad is not None and torch.isfinite(x[0].grad).all()
    uninstall(ref)

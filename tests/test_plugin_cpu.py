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

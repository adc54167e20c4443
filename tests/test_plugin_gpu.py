"""GPU: numerics of `prfl_b200.plugin.install()`.  /root/reference does not exist on the GPU box, so the
patched object is a reference-SHAPED stand-in: a module tree with the reference's block attributes and parameter names
whose own forward is the oracle's restatement of WanAttentionBlock.forward (fp32 torch ops on the GPU).  After install()
the same call runs the B200 kernels on the same Parameter objects; the enable flag switches back."""
import pytest
import torch
import torch.nn as nn

from conftest import cos_rel
from oracle import synth
from oracle import wan_oracle as O

pytestmark = pytest.mark.gpu


class RefLikeBlock(nn.Module):
    def __init__(self, cfg, sd, prefix):
        super().__init__()
        self.dim, self.ffn_dim, self.num_heads = cfg.dim, cfg.ffn_dim, cfg.num_heads
        self.window_size, self.qk_norm, self.cross_attn_norm, self.eps = (-1, -1), True, True, cfg.eps
        self.cfg = cfg
        for k, v in sd.items():
            if k.startswith(prefix):
                name = k[len(prefix):]
                mod = self
                parts = name.split(".")
                for part in parts[:-1]:
                    if not hasattr(mod, part):
                        mod.add_module(part, nn.Module())
                    mod = getattr(mod, part)
                mod.register_parameter(parts[-1], nn.Parameter(v.clone().cuda()))

    def forward(self, x, e, seq_lens, grid_sizes, freqs, context, context_lens):
        sd = {k: v for k, v in self.named_parameters()}
        grids = [tuple(int(a) for a in g) for g in grid_sizes.tolist()]
        return O.attention_block(sd, "", x.float(), e, grids, [int(s) for s in seq_lens], context.float(), self.cfg, O._Prec(None))


@pytest.mark.parametrize("mt", ["t2v", "i2v"])
def test_install_runs_b200_kernels_on_shared_parameters(mt):
    from prfl_b200 import _lib
    from prfl_b200.plugin import install, uninstall
    cfg = synth.tiny_cfg(mt, heads=2, layers=2)
    sd = synth.make_wan_state_dict(cfg, 33)
    ref = nn.Module()
    ref.model_type = mt
    ref.blocks = nn.ModuleList([RefLikeBlock(cfg, sd, f"blocks.{i}.") for i in range(2)])
    g = torch.Generator().manual_seed(1)
    L, n_ctx = 3 * 4 * 6, 512 + (257 if mt != "t2v" else 0)
    x = torch.randn(1, L, cfg.dim, generator=g).cuda()
    e = (torch.randn(1, 6, cfg.dim, generator=g) * 0.1).cuda()
    ctx = (torch.randn(1, n_ctx, cfg.dim, generator=g) * 0.5).cuda()
    args = (e, torch.tensor([L]), torch.tensor([[3, 4, 6]]), None, ctx, None)
    with torch.no_grad():
        want = ref.blocks[1](ref.blocks[0](x, *args), *args)
    install(ref)
    _lib.launch_count_reset()
    with torch.no_grad():
        got = ref.blocks[1](ref.blocks[0](x.clone(), *args), *args)
    assert _lib.launch_count() > 20                                     # the B200 kernels ran
    cos, rel = cos_rel(got.cpu(), want.cpu())
    assert cos >= 0.999 and rel <= 2e-2, (cos, rel)
    assert ref.blocks[0]._prfl_b200_fast.ffn[0].weight is ref.blocks[0].get_submodule("ffn.0").weight
    ref.prfl_b200_enable(False)
    with torch.no_grad():
        back = ref.blocks[1](ref.blocks[0](x, *args), *args)
    assert torch.equal(back, want)
    uninstall(ref)
    assert not hasattr(ref.blocks[0], "old_forward")

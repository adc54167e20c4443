"""CPU: the C-ABI shared library loads and exports every symbol include/prfl_b200.h declares; the ctypes
binding covers them all; the product refuses to run without a GPU (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "prfl_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(prfl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from prfl_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/prfl_b200.h but not exported"
    assert set(names) == set(_lib._SIGS), set(names) ^ set(_lib._SIGS)
    assert _lib.lib().prfl_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from prfl_b200 import _lib, ops
    from prfl_b200.model import WanModel
    with pytest.raises(_lib.PrflError):
        ops.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    m = WanModel(dim=256, ffn_dim=512, num_heads=2, num_layers=1, text_dim=64)
    with pytest.raises(_lib.PrflError):
        m(x=[torch.randn(16, 1, 4, 4)], t=torch.tensor([1.0]), context=[torch.randn(3, 64)], seq_len=4)
    # the raw entry points report PRFL_E_ARCH instead of computing anything
    rc = _lib.lib().prfl_cast_f32_bf16(None, None, 0, None)
    assert rc == -3 and b"no CUDA device" in _lib.lib().prfl_last_error_string()


def test_state_dict_keys_match_reference_layout():
    """Key names are those of the reference modules — the golden generator loaded the same
    synthetic state dicts into the real reference with strict=True."""
    from oracle import synth
    from prfl_b200.model import WanModel
    from prfl_b200.network import MLP, QueryAttention
    for mt in ("t2v", "i2v"):
        cfg = synth.tiny_cfg(mt)
        res = WanModel(**cfg.kwargs()).load_state_dict(synth.make_wan_state_dict(cfg, 0), strict=True)
        assert not res.missing_keys and not res.unexpected_keys
    qa, mlp = synth.make_reward_state_dicts(256)
    QueryAttention(256, 1, 8, dropout=0.0, return_type="query").load_state_dict(qa, strict=True)
    MLP(256).load_state_dict(mlp, strict=True)


def test_rope_apply_free_function_matches_reference_golden():
    """`rope_apply(x, grid_sizes, freqs)` keeps the reference signature (model.py:61); torch ops, so it runs on CPU."""
    from conftest import golden
    from prfl_b200.model import rope_apply
    fx = golden("ops")
    out = rope_apply(fx["rope_in"], torch.tensor([fx["rope_grid"]]), None)
    assert out.dtype == torch.float32
    torch.testing.assert_close(out, fx["rope_out"], rtol=1e-5, atol=1e-5)

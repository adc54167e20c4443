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
    assert _lib.lib().prfl_abi_version() == 2


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


def test_attention_split_policy():
    """Host logic of the forward attention's key-split tail (csrc/attention_fwd.cu, "Wave quantisation"), no device needed."""
    from prfl_b200 import _lib

    def plan(Lq, Lk, H, sms=148):
        out = (ctypes.c_int * 4)()
        _lib.lib().prfl_attn_fwd_split_plan(Lq, Lk, H, sms, out)
        return tuple(out)

    assert plan(32760, 32760, 5) == (640, 592, 48, 3)          # one rank's share at 8 GPUs: 4 waves + 48 units, split 3 ways
    assert plan(32760, 32760, 20) == (2560, 2516, 44, 3)       # 2 GPUs
    assert plan(32760, 32760, 40) == (5120, 5120, 0, 1)        # remainder 88 > half a wave: plain launch
    assert plan(75600, 75600, 5) == (1480, 1480, 0, 1)         # 720P at 8 GPUs: exactly 10 waves
    assert plan(4096, 512, 10)[2] == 0                          # cross-attention-sized key axis: never split
    assert plan(300, 300, 2) == (4, 4, 0, 1)                    # fewer units than SMs
    for Lq in (2100, 4096, 8200, 32760):
        for Lk in (4096, 4100, 8200, 32760):
            for H in (1, 3, 5, 10, 17, 40):
                units, main, tail, pieces = plan(Lq, Lk, H)
                n_kv = -(-Lk // 128)
                assert units == main + tail and main % 148 == 0 or tail == 0
                if tail:
                    per = -(-n_kv // pieces)
                    assert 2 <= pieces <= 4 and 2 * tail <= 148 and (pieces - 1) * per < n_kv      # no empty piece


def test_python_surface_the_reference_callers_import():
    """Every name the reference trainers / pipelines import from the modules this package replaces (train_prfl.py:28-31,
    41-58, 97; train_pavrm.py:37-52; inference_prfl.py:71-82) exists here with the reference's call signature."""
    import inspect
    from prfl_b200 import checkpoint, model, network, parallel, scheduler
    for mod, names in ((model, ("WanModel", "WanAttentionBlock", "WanSelfAttention", "WanT2VCrossAttention", "WanI2VCrossAttention",
                                "WanRMSNorm", "WanLayerNorm", "Head", "MLPProj", "sinusoidal_embedding_1d", "rope_params", "rope_apply")),
                       (network, ("MLP", "QueryAttention", "forward_mlp", "forward_siamese")),
                       (parallel, ("nccl_info", "COMM_INFO", "initialize_sequence_parallel_state", "initialize_sequence_parallel_group",
                                   "set_sequence_parallel_state", "get_sequence_parallel_state", "destroy_sequence_parallel_group",
                                   "broadcast", "all_gather", "all_to_all_4D", "SeqAllToAll4D", "initialize_usp_state")),
                       (scheduler, ("FlowUniPCMultistepScheduler",)),
                       (checkpoint, ("save_checkpoint", "load_state_dict"))):
        for n in names:
            assert hasattr(mod, n), (mod.__name__, n)
    sig = lambda f: list(inspect.signature(f).parameters)
    assert sig(model.WanModel.forward)[:10] == ["self", "x", "t", "context", "seq_len", "clip_fea", "y", "cond_flag", "output_features",
                                                "selected_layers"]
    assert sig(model.WanModel.__init__)[1:] == ["model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim",
                                                "out_dim", "num_heads", "num_layers", "window_size", "qk_norm", "cross_attn_norm", "eps"]
    assert sig(model.WanAttentionBlock.forward)[:8] == ["self", "x", "e", "seq_lens", "grid_sizes", "freqs", "context", "context_lens"]
    assert sig(network.QueryAttention.__init__)[1:] == ["feature_dim", "num_queries", "num_heads", "dropout", "layer_norm", "return_type",
                                                       "product_text", "text_dim"]
    assert sig(network.QueryAttention.forward)[:4] == ["self", "x", "e", "text"]
    assert sig(parallel.all_to_all_4D) == ["input_", "scatter_dim", "gather_dim"] and sig(parallel.all_gather) == ["input_", "dim"]
    assert sig(checkpoint.save_checkpoint)[:5] == ["transformer", "rank", "output_dir", "step", "ema"]
    for cm in ("from_pretrained", "from_config", "save_pretrained"):
        assert callable(getattr(model.WanModel, cm))


def test_header_is_plain_c_and_links_from_a_c_host(tmp_path):
    """The boundary is a C ABI: `include/prfl_b200.h` compiles as strict C99 (`-pedantic -Werror`: no C++-isms, no torch or CUDA
    runtime types in a signature) and a C program links the shared library and calls it.  Without a GPU the call must come back
    with an error code and a message — never compute anything."""
    import shutil
    import subprocess
    from prfl_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "host.c"
    src.write_text('#include <stdio.h>\n#include "prfl_b200.h"\n'
                   "int main(void) {\n"
                   "  int rc = prfl_cast_f32_bf16(NULL, NULL, 0, NULL);\n"
                   '  printf("%d %d %s\\n", prfl_abi_version(), rc, prfl_last_error_string());\n'
                   "  return 0;\n}\n")
    exe = tmp_path / "host"
    libdir = os.path.dirname(_lib.LIB_PATH)
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                          "-L", libdir, "-lprfl_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout.split(" ", 2)
    assert out[0] == "2"
    if not torch.cuda.is_available():
        assert out[1] == "-3" and "no CPU fallback" in out[2]


def test_off_path_options_are_refused_not_silently_ignored():
    """Options the kernels do not implement raise at construction / call time (no silent difference from the reference)."""
    from prfl_b200.attention import flash_attention
    from prfl_b200.model import WanModel, WanSelfAttention
    from prfl_b200.network import QueryAttention
    with pytest.raises(NotImplementedError):
        WanSelfAttention(256, 2, window_size=(128, 128))
    with pytest.raises(AssertionError):
        WanSelfAttention(192, 2)                                        # head_dim 96
    with pytest.raises(AssertionError):
        WanModel(patch_size=(1, 4, 4), dim=256, ffn_dim=512, num_heads=2, num_layers=1)
    with pytest.raises(AssertionError):
        QueryAttention(256, num_queries=2)
    q = torch.zeros(1, 4, 2, 128)
    for kw in (dict(causal=True), dict(dropout_p=0.1), dict(window_size=(4, 4)), dict(q_lens=torch.tensor([4]))):
        with pytest.raises(NotImplementedError):
            flash_attention(q, q, q, **kw)


def test_operand_cache_follows_parameter_updates():
    """The bf16 / fused operand copies (what autocast re-casts on every Linear call in the reference) are cached per module and must
    refresh whenever the parameter changes: in-place updates through the Parameter (torch.optim, load_state_dict) bump its version
    counter, `.to()` / `.data = ...` change its storage, and `bump_weight_epoch()` covers writers that bypass both (updates
    through `.data`, e.g. the reference's EMA rule model_utils.py:172-175, and ShardedAdamW's flat-buffer updates)."""
    from prfl_b200.model import _OperandCache, bump_weight_epoch
    p = torch.nn.Parameter(torch.ones(4, 4))
    cache, builds = _OperandCache(), []

    def get():
        return cache.get("w", [p], lambda: builds.append(1) or p.detach().clone())
    a = get()
    assert get() is a and len(builds) == 1                                  # hit
    with torch.no_grad():
        p.add_(1.0)                                                         # what torch.optim / load_state_dict do
    assert torch.equal(get(), torch.full((4, 4), 2.0)) and len(builds) == 2
    p.data = torch.zeros(4, 4)                                              # .to() / re-pointing: new storage
    assert torch.equal(get(), torch.zeros(4, 4)) and len(builds) == 3
    p.data.add_(5.0)                                                        # bypasses the version counter ...
    assert len(builds) == 3 and torch.equal(get(), torch.zeros(4, 4))       # ... so the copy is stale until told:
    bump_weight_epoch()
    assert torch.equal(get(), torch.full((4, 4), 5.0)) and len(builds) == 4


def test_forwards_run_with_the_callers_autocast_switched_off():
    """The trainers call the model under torch.autocast(bf16) (train_prfl.py:669-748); the drop-in's precision is explicit, so
    its entry points switch CUDA autocast off for their own body (and restore the caller's state afterwards)."""
    from prfl_b200.model import WanModel, no_autocast
    from prfl_b200.network import QueryAttention
    assert hasattr(WanModel.forward, "__wrapped__") and hasattr(QueryAttention.forward, "__wrapped__")
    seen = []

    @no_autocast
    def body(v):
        seen.append(torch.is_autocast_enabled("cuda"))
        return v + 1
    assert body(1) == 2 and seen == [False]
    if torch.cuda.is_available():                                        # only a CUDA box can switch the outer context on
        with torch.autocast("cuda", dtype=torch.bfloat16):
            body(1)
            assert torch.is_autocast_enabled("cuda")
        assert seen == [False, False]

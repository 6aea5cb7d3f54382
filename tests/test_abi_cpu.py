"""CPU checks of the drop-in boundary: the C-ABI library builds / loads without a GPU and
exports every symbol include/ergm_b200.h declares; the Python model mirrors the reference's
module tree; the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ergm_b200 import build
    build.build()
    from ergm_b200 import _lib
    return _lib.lib()


def test_header_symbols_exported(lib):
    from ergm_b200 import _lib
    protos = _lib.header_prototypes()
    assert len(protos) >= 27
    for name in protos:
        assert hasattr(lib, name), name
    assert lib.ergm_abi_version() == 2


def test_header_is_plain_c():
    """The boundary must be consumable by a C compiler: no C++ / torch types in the signatures."""
    import subprocess
    import tempfile
    src = '#include "ergm_b200.h"\nint main(void){ergm_gemm_args a; (void)a; return ERGM_OK;}\n'
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "t.c")
        open(p, "w").write(src)
        r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", p,
                            "-o", os.path.join(d, "t.o")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_gemm_args_struct_matches_header(tmp_path):
    """The ctypes mirror of ergm_gemm_args has the size and field offsets a C compiler gives the header's struct."""
    from ergm_b200 import _lib
    src = tmp_path / "sz.c"
    fields = [f[0] for f in _lib.GemmArgs._fields_]
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ergm_b200.h"\nint main(void) {\n'
                   '  printf("%zu", sizeof(ergm_gemm_args));\n'
                   + "".join('  printf(" %%zu", offsetof(ergm_gemm_args, %s));\n' % f for f in fields)
                   + "  return 0;\n}\n")
    exe = tmp_path / "sz"
    r = subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    nums = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    assert ctypes.sizeof(_lib.GemmArgs) == nums[0]
    assert [getattr(_lib.GemmArgs, f).offset for f in fields] == nums[1:]


def test_argument_errors_without_gpu(lib):
    from ergm_b200 import _lib
    a = _lib.GemmArgs()
    assert lib.ergm_gemm_bf16(ctypes.byref(a), None) == -1  # null operands -> ERGM_ERR_ARG, nothing launched
    assert lib.ergm_ln_fwd(None, None, None, None, None, None, None, 4, 128, 1e-5, None, None, None) == -1
    assert lib.ergm_attn_fwd(1, 8, 0, 1, 8, 0, 1, 8, 0, 1, 8, None, None, None, 1, 1, 8, 8, 32, 1, 0, 0.0, 0, 0, None, 0, None) == -2
    assert lib.ergm_pack_plan(None, 4, 16, None, None, None, None, None, 64, None) == -1
    assert lib.ergm_zero_rows_dyn(None, 16, None, 4, None) == -1
    # decode-side entry points: argument errors are detected before anything touches a device
    assert lib.ergm_dec_pack_weight(None, 768, 768, 768, 0, None, None, None, None, None, None) == -1
    assert lib.ergm_dec_gemm(None, None, 768, 1e-5, None, 768, 768, None, None, 768, 0, 0, 64, None) == -1
    assert lib.ergm_dec_gemm(8, None, 768, 1e-5, 16, 768, 768, None, 16, 768, 0, 0, 65, None) == -2   # M > 64
    assert lib.ergm_dec_gemm(8, None, 768, 1e-5, 16, 768, 768, None, 16, 768, 2, 1, 64, None) == -1   # gelu with "+="
    assert lib.ergm_mm_pool_fwd(None, 0, 0, None, 4, 113, 768, None, 0, None, 0, None) == -1
    assert lib.ergm_mm_pool_fwd(16, 768 * 113, 768, None, 4, 113, 770, 16, 768, None, 0, None) == -1  # D % 4
    assert lib.ergm_sample(16, 50304, 4, 50260, 5, 0.8, 1.0, 0, None, 0, None, 0, None, None, None, -1, None) == -1  # top-k AND top-p
    assert lib.ergm_sample(16, 50304, 4, 50260, 0, 0.0, 1.0, 0, None, 0, None, 0, None, None, None, -1, None) == -1  # top_p <= 0
    assert lib.ergm_decode_layers(None, 12, 768, 3072, 12, 64, None, None, None, None, None, None, None, 12, 0, 1e-5, None, None) == -1


def test_model_tree_matches_reference_keys():
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    from oracle import ergm_oracle as O
    m = GPT2LMHeadModel(GPT2Config(vocab_size=1024, n_positions=64, n_embd=128, n_layer=2, n_head=2))
    want = dict(O.param_shapes(O.OracleConfig(1024, 64, 128, 2, 2)))
    want["lm_head.weight"] = want["transformer.wte.weight"]
    sd = m.state_dict()
    assert set(sd) == set(want)
    assert all(tuple(sd[k].shape) == want[k] for k in want)
    assert m.lm_head.weight is m.transformer.wte.weight
    assert m.config.n_ctx == 64 and m.config.add_cross_attention is True
    # init distributions of model.py:359-375
    assert abs(m.transformer.h[0].attn.c_proj.weight.std().item() - 0.02 / 2.0) < 2e-3
    assert abs(m.transformer.h[0].attn.c_attn.weight.std().item() - 0.02) < 2e-3


def test_no_cpu_fallback():
    from transformers import GPT2Config
    from ergm_b200.model import GPT2LMHeadModel
    from ergm_b200._lib import ErgmError
    m = GPT2LMHeadModel(GPT2Config(vocab_size=1024, n_positions=64, n_embd=128, n_layer=1, n_head=2))
    with pytest.raises(ErgmError):
        m(input_ids=torch.zeros(1, 4, dtype=torch.long))


def test_product_never_imports_oracle():
    for dp, _, files in os.walk(os.path.join(ROOT, "ergm_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_reference_compat_shim_exports():
    """`from model import *` (main.py:22) must find the reference's public names."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("model", os.path.join(ROOT, "compat", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name in ("GPT2LMHeadModel", "GPT2Model", "CausalLMOutputWithEmotionClassification", "torch", "nn", "F"):
        assert hasattr(mod, name)


def test_lmhead_ce_workspace_query_without_gpu():
    """The workspace queries of the composed LM-head + CE entry points are pure host arithmetic."""
    from ergm_b200 import ops
    fwd = ops.lmhead_ce_layout(8192, 768, 50260, False)
    both = ops.lmhead_ce_layout(8192, 768, 50260, True)
    assert fwd[:8] == both[:8] and all(o % 256 == 0 for o in both)
    assert all(b > a for a, b in zip(both[:7], both[1:8])) and both[-1] > fwd[-1]
    ldl = (50260 + 63) // 64 * 64
    assert both[5] - both[4] == 8192 * ldl * 2            # bf16 logits of the scored rows (capacity = all rows)
    assert both[-1] - both[8] == 8192 * 768 * 4


def test_kv_page_allocator_host_logic():
    """Free-list page allocator of the generation K/V pool (no kernels involved: runs on the CPU device)."""
    import types
    import torch
    from ergm_b200 import generation
    eng = types.SimpleNamespace(nh=2, L=3, device=torch.device("cpu"))
    alloc = generation.page_allocator(eng)
    assert generation.page_allocator(eng) is alloc
    t1, ids1 = alloc.alloc_table(3, 4)
    assert t1.shape == (3, 4) and sorted(ids1) == list(range(12)) and alloc.n_pages == 12
    assert t1[:, 0].tolist() == [0, 1, 2] and t1[0].tolist() == [0, 3, 6, 9]       # page j of every sequence first
    assert len(alloc.pools) == 3 and alloc.pools[0].shape == (12, 2, 2, generation.PAGE, 64)
    t2, ids2 = alloc.alloc_table(2, 2)                                             # grows: 4 more pages, ids1 stay reserved
    assert alloc.n_pages == 16 and sorted(ids2) == [12, 13, 14, 15] and not set(ids1) & set(ids2)
    alloc.release(ids1)
    t3, ids3 = alloc.alloc_table(4, 3)                                             # reuses the released pages, no growth
    assert alloc.n_pages == 16 and sorted(ids3) == sorted(ids1)
    assert len(alloc.free) == 0
    alloc.release(ids2)
    alloc.release(ids3)
    assert sorted(alloc.free) == list(range(16))

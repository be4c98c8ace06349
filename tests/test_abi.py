"""CPU checks of the drop-in boundary: liblecb.so loads without a GPU, exports every symbol declared
in include/lecb.h with the arity bound in lecb200/_lib.py, and fails loudly (never silently falls back)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import lecb200
    return lecb200


def _header_functions():
    src = open(os.path.join(ROOT, "include", "lecb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*|unsigned long long)\s+(lecb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_header_and_library_agree(built):
    from lecb200 import _lib
    decl = _header_functions()
    assert len(decl) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name, nargs in decl.items():
        assert hasattr(lib, name), f"{name} declared in lecb.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes binding"
        assert len(_lib.SIGNATURES[name][1]) == nargs, f"{name}: header has {nargs} args, binding {len(_lib.SIGNATURES[name][1])}"
    assert set(_lib.SIGNATURES) == set(decl), set(_lib.SIGNATURES) ^ set(decl)
    assert lib.lecb_abi_version() == 1


def test_header_cites_reference_call_sites():
    src = open(os.path.join(ROOT, "include", "lecb.h")).read()
    for token in ("T:409-410", "M:44", "T:456-470", "U:126-173", "U:85-93", "M:89-127", "T:444-448"):
        assert token in src, token


def test_argument_errors_are_reported_without_a_gpu(built):
    from lecb200 import _lib
    lib = _lib.lib
    # null pointers / bad shapes are rejected before any CUDA call
    st = lib.lecb_gemm_bf16(0, 0, 0, 0, 0, 0, 128, 64, 64, 0, 0)
    assert st == -1 and b"null" in lib.lecb_last_error()
    st = lib.lecb_gemm_bf16_dual(0, 64, 0, 64, 0, 0, 0, 128, 64, 0, 0)
    assert st == -1 and b"null" in lib.lecb_last_error()
    st = lib.lecb_gemm_bf16_dual(16, 64, 16, 96, 16, 0, 16, 128, 64, 0, 0)              # K2 not a multiple of 64
    assert st == -1 and b"multiples of 64" in lib.lecb_last_error()
    st = lib.lecb_gemm_bf16_dual(16, 64, 16, 64, 16, 0, 16, 128, 64, _lib.EPI_QUICKGELU, 0)   # only ReLU is offered
    assert st == -1 and b"LECB_EPI_RELU" in lib.lecb_last_error()
    st = lib.lecb_head_aggregate(1, 240, 0, 0, 1, 0, 0, 1, 1, 500, 3, 4.0, 50.0, 0)
    assert st == -1 and b"K" in lib.lecb_last_error()
    # the fused average pool needs even H and W (checked before any CUDA call); the planning query never launches
    st = lib.lecb_conv3x3_bf16(1, 1, 0, 1, 2, 7, 8, 64, 64, _lib.EPI_RELU | _lib.EPI_AVGPOOL2, 0)
    assert st == -1 and b"even" in lib.lecb_last_error()
    assert lib.lecb_conv3x3_pool_fusable(256, 223, 224, 32, 64) == 0                   # odd H
    assert lib.lecb_conv3x3_pool_fusable(256, 224, 224, 48, 64) == 0                   # Cin not a multiple of 32
    assert lib.lecb_conv3x3_pool_fusable(256, 224, 224, 32, 64) in (0, 1)              # 1 on a GPU box, 0 without a device


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback(built):
    from lecb200 import LecbError, ops
    a = torch.zeros((128, 64), dtype=torch.bfloat16)
    with pytest.raises(LecbError):
        ops.gemm(a, a)
    with pytest.raises(LecbError):
        ops.l2norm_rows(torch.zeros((4, 64))) if False else ops.gemm(a, a)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "language-enhanced-clip-for-multi-label-image-recognition_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn

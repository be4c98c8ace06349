"""ctypes binding of csrc/liblecb.so (the C ABI declared in include/lecb.h).

There is no fallback of any kind: if the library is missing the import fails, and if a compute entry
point is called without a CUDA device it returns LECB_ERR_CUDA, surfaced here as RuntimeError."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LECB_LIB_PATH") or os.path.join(_HERE, "csrc", "liblecb.so")   # override: A/B runs of two builds

c_void_p, c_int, c_i64, c_uint, c_float = C.c_void_p, C.c_int, C.c_int64, C.c_uint, C.c_float

# name -> (restype, argtypes); mirrors include/lecb.h one to one (tests/test_abi.py checks the header)
SIGNATURES = {
    "lecb_abi_version": (c_int, []),
    "lecb_last_error": (C.c_char_p, []),
    "lecb_launch_count": (C.c_ulonglong, []),
    "lecb_gemm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_int,
                               c_uint, c_void_p]),
    "lecb_gemm_bf16_dual": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_uint,
                                    c_void_p]),
    "lecb_conv3x3_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_uint, c_void_p]),
    "lecb_set_pair_gemm": (c_int, [c_int]),
    "lecb_set_attn_poly": (c_int, [c_int]),
    "lecb_conv3x3_pool_fusable": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "lecb_stem_conv1": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "lecb_stem_conv1_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "lecb_avgpool2x2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "lecb_token_mean": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "lecb_l2norm_rows": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_int, c_void_p]),
    "lecb_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64,
                                   c_int, c_float, c_void_p]),
    "lecb_attnpool_query0": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "lecb_causal_attn_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "lecb_attn_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lecb_patchify": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lecb_vit_embed_ln": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_float, c_void_p]),
    "lecb_copy_cols": (c_int, [c_void_p, c_i64, c_int, c_void_p, c_i64, c_i64, c_int, c_void_p]),
    "lecb_head_aggregate": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                    c_int, c_int, c_float, c_float, c_void_p]),
    "lecb_global_logits": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "lecb_asl_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_float, c_float, c_float,
                                 c_float, c_float, c_float, c_int, c_void_p]),
    "lecb_ranking_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float,
                                     c_void_p]),
    "lecb_ranking_cooc_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float,
                                          c_void_p]),
    "lecb_kl_softmax_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int, c_float, c_void_p]),
    "lecb_resample_bce_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i64, c_int,
                                          c_float, c_float, c_float, c_float, c_int, c_float, c_float, c_float, c_void_p]),
    "lecb_ema_update": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p]),
    "lecb_pack_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "lecb_unpack_scale_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_void_p]),
    "lecb_sgd_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_float, c_void_p]),
    "lecb_split_f16_hilo": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "lecb_gemm_topk10": (c_int, [c_void_p, c_void_p, c_i64, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "lecb_topk10_merge": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "lecb_window_plan_size": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "lecb_window_plan": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, C.c_longlong]),
    "lecb_crop_resize_u8": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p]),
    "lecb_split_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "lecb_topk10": (c_int, [c_void_p, c_i64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "lecb_gather_mean10": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "lecb_quick_gelu_fwd": (c_int, [c_void_p, c_void_p, c_i64, c_void_p]),
    "lecb_quick_gelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "lecb_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_i64, c_int, c_void_p]),
    "lecb_causal_attn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "lecb_attn_causal_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "lecb_residual_relu_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_void_p]),
    "lecb_relu_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_i64, c_void_p]),
    "lecb_l2norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_i64, c_int, c_void_p]),
    "lecb_head_aggregate_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_int, c_float, c_float, c_void_p]),
    "lecb_block_fuse": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float,
                                c_void_p]),
    "lecb_cooc_adjust": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "lecb_resize_ksize": (c_int, [c_int, c_int, c_int]),
    "lecb_resize_plan": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_int]),
    "lecb_tn_gemm_small": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_int,
                                   c_void_p]),
}

EPI_RELU, EPI_QUICKGELU, EPI_OUT_F32, EPI_RES_F32, GEMM_F16_OPERANDS, EPI_AVGPOOL2, EPI_MUL_QGELU_GRAD = 1, 2, 4, 8, 16, 32, 64


class LecbError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). lecb200 has no CPU or PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib.lecb_last_error()
        raise LecbError(f"{what or 'lecb'} failed with status {status}: {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(lib.lecb_launch_count())

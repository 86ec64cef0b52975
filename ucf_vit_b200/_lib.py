"""ctypes binding of the C-ABI library (include/ucf_vit_b200.h).

The product path has NO fallback: if the shared library is missing, or a launcher returns a
non-zero code, a RuntimeError is raised.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libucfvit_b200.so")
_lib = None

UCF_DTYPE_F32, UCF_DTYPE_BF16, UCF_DTYPE_U8, UCF_DTYPE_F64, UCF_DTYPE_I64 = 0, 1, 2, 3, 4
UCF_LAYOUT_K_MAJOR, UCF_LAYOUT_MN_MAJOR = 0, 1
EPI_BIAS, EPI_BIAS_RESIDUAL, EPI_BIAS_GELU_AUX, EPI_DGELU, EPI_F32_ADD = 0, 1, 2, 3, 4
PATCH_MSE_MAX_BLOCKS = 4096

_LL = c_longlong


class BlockParams(ctypes.Structure):
    """ucf_block_params (include/ucf_vit_b200.h)."""
    _fields_ = ([(n, c_int) for n in ("B", "N", "D", "H", "hidden")] + [("eps1", c_float), ("eps2", c_float),
                ("ln_dtype", c_int), ("bias_dtype", c_int)] +
                [(n, c_void_p) for n in ("n1_w", "n1_b", "n2_w", "n2_b", "qkv_b", "proj_b", "fc1_b", "fc2_b",
                                         "qkv_w", "proj_w", "fc1_w", "fc2_w",
                                         "qkv_w_master", "proj_w_master", "fc1_w_master", "fc2_w_master")])


class BlockActs(ctypes.Structure):
    """ucf_block_acts."""
    _fields_ = [(n, c_void_p) for n in ("x", "h1", "qkv", "o", "x1", "h2", "z", "u", "y", "mean1", "rstd1", "mean2",
                                        "rstd2", "lse")]


class BlockGrads(ctypes.Structure):
    """ucf_block_grads."""
    _fields_ = [(n, c_void_p) for n in ("dy", "dx", "g_n1_w", "g_n1_b", "g_qkv_w", "g_qkv_b", "g_proj_w", "g_proj_b",
                                        "g_n2_w", "g_n2_b", "g_fc1_w", "g_fc1_b", "g_fc2_w", "g_fc2_b",
                                        "ws_a", "ws_b", "ws_c", "dq_acc", "delta")]


_SIGNATURES = {
    "ucf_block_fwd": (c_int, [ctypes.POINTER(BlockParams), ctypes.POINTER(BlockActs), c_void_p]),
    "ucf_block_bwd": (c_int, [ctypes.POINTER(BlockParams), ctypes.POINTER(BlockActs), ctypes.POINTER(BlockGrads), c_void_p]),
    "ucf_wgrad_splits": (c_int, [c_int, c_int, _LL]),
    "ucf_abi_version": (c_int, []),
    "ucf_last_error": (c_char_p, []),
    "ucf_launch_count": (c_ulonglong, []),
    "ucf_gemm_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                              _LL, _LL, _LL, _LL, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ucf_gemm_dgrad_delta_supported": (c_int, [c_int, c_int, c_int, c_int]),
    "ucf_gemm_dgrad_delta": (c_int, [c_void_p] * 5 + [c_int] * 3 + [_LL] * 4 + [c_int, c_int, c_void_p]),
    "ucf_attention_bwd_with_delta": (c_int, [c_void_p] * 11 + [c_int] * 5 + [_LL] * 21 + [c_float, c_void_p]),
    "ucf_layernorm_stats": (c_int, [c_void_p, c_void_p, c_void_p, _LL, c_int, c_float, c_int, c_void_p]),
    "ucf_ln_gemm_supported": (c_int, [c_int, c_int, c_int]),
    "ucf_ln_gemm": (c_int, [c_void_p] * 7 + [c_int] * 3 + [_LL] * 3 + [c_void_p]),
    "ucf_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, _LL, c_int,
                                  c_float, c_int, c_int, c_void_p]),
    "ucf_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, _LL, c_int, c_int, c_void_p]),
    "ucf_attention_fwd": (c_int, [c_void_p] * 5 + [c_int] * 5 + [_LL] * 12 + [c_float, c_void_p]),
    "ucf_attention_bwd": (c_int, [c_void_p] * 11 + [c_int] * 5 + [_LL] * 21 + [c_float, c_void_p]),
    "ucf_var_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, _LL, c_int, c_int, c_int, c_int, c_int,
                                      c_float, c_void_p]),
    "ucf_var_attention_bwd": (c_int, [c_void_p] * 7 + [_LL, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "ucf_sap_build_tree_host": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, ctypes.c_double, c_int, c_void_p,
                                        c_void_p]),
    "ucf_sap_build_tree_batch_host": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_double, c_int,
                                              c_void_p, c_void_p, c_void_p, c_int]),
    "ucf_sap_gather": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p]),
    "ucf_sap_scatter": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p,
                                c_void_p]),
    "ucf_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, _LL, c_void_p]),
    "ucf_cast_f32_to_bf16_multi": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ucf_cast_bf16_to_f32": (c_int, [c_void_p, c_void_p, _LL, c_int, c_void_p]),
    "ucf_colsum_bf16": (c_int, [c_void_p, c_void_p, _LL, c_int, _LL, c_int, c_void_p]),
    "ucf_assemble_tokens": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, _LL, c_int,
                                    c_int, c_void_p]),
    "ucf_add_bcast": (c_int, [c_void_p, c_void_p, c_void_p, _LL, c_int, c_int, c_int, _LL, _LL, _LL, c_int, c_void_p]),
    "ucf_mask_plan": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ucf_gather_tokens": (c_int, [c_void_p] * 5 + [c_int] * 4 + [_LL, c_int, c_void_p]),
    "ucf_scatter_tokens": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p]),
    "ucf_patch_mse_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p] + [c_int] * 8 + [c_void_p] * 3),
    "ucf_patch_mse_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p] + [c_int] * 8 +
                          [c_void_p] * 2),
    "ucf_dice_bce_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, _LL, c_float, c_float, c_int, c_void_p,
                                 c_void_p, c_void_p]),
    "ucf_dice_bce_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, _LL, c_float, c_int,
                                 c_void_p, c_void_p]),
    "ucf_conv3d_wgrad_supported": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "ucf_conv3d_wgrad_ctas": (c_int, [c_int, c_int, c_int, c_int]),
    "ucf_conv3d_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "ucf_pointwise_conv_supported": (c_int, [c_int, c_int]),
    "ucf_pointwise_conv_ctas": (c_int, [_LL]),
    "ucf_pointwise_conv": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, _LL, c_int, c_int, c_void_p]),
    "ucf_pointwise_conv_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, _LL, c_int, c_int, c_void_p, c_void_p]),
    "ucf_gaussian_blur_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ucf_canny_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                     c_void_p, c_void_p]),
    "ucf_inorm_chunks": (c_int, [c_int, _LL, c_int]),
    "ucf_inorm_stats": (c_int, [c_void_p, c_int, _LL, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "ucf_inorm_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, _LL, c_int, c_float, c_void_p]),
    "ucf_inorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, _LL, c_int,
                      c_float, c_void_p, c_void_p, c_void_p]),
    "ucf_dice_ce_blocks_per_sample": (c_int, [c_int, _LL]),
    "ucf_dice_ce_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, _LL, c_int, c_float, c_float, c_float, c_float,
                                c_void_p, c_void_p, c_int, c_void_p]),
    "ucf_dice_ce_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, _LL, c_int, c_void_p, c_int,
                                c_void_p]),
    "ucf_adamw_multi": (c_int, [c_int] + [c_void_p] * 5 + [c_double] * 5 + [_LL, c_int, c_void_p]),
    "ucf_adamw_multi_dev": (c_int, [c_int] + [c_void_p] * 6 + [c_double] * 4 + [c_void_p, c_int, c_void_p]),
    "ucf_patchify": (c_int, [c_void_p, c_void_p] + [c_int] * 10 + [_LL, c_int, c_void_p]),
}


def declared_symbols():
    """Every entry point include/ucf_vit_b200.h declares (checked by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"ucf_vit_b200: CUDA library {LIB_PATH} is missing. Build it with "
                "`python -m ucf_vit_b200._build` (or __graft_entry__.build()). There is no CPU fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            if not hasattr(l, name):
                continue   # tests/test_abi.py reports missing symbols explicitly
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().ucf_last_error()
        raise RuntimeError(f"ucf_vit_b200.{what} failed (code {rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(lib().ucf_launch_count())

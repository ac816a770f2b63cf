"""ctypes binding of libvgslm.so — the C-ABI kernel library declared in ``include/vgslm.h``.

The product path has NO CPU fallback: if the shared library is missing, or a tensor handed to an op
is not a CUDA tensor, this module raises.  Build the library with ``python __graft_entry__.py`` (or
``make -C vae_gslm_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# VGSLM_LIB: path of an alternative build of the same library (A/B runs of experiment switches, tools/build_variant.sh)
LIB_PATH = os.environ.get("VGSLM_LIB") or os.path.join(_HERE, "libvgslm.so")

VG_F32, VG_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU, ACT_SILU, ACT_MULT = 0, 1, 2, 3, 4
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f32 = C.c_float
_sz = C.c_size_t


class GemmArgs(C.Structure):
    _fields_ = [
        ("M", _i64), ("N", _i64), ("K", _i64),
        ("A", _p), ("lda", _i64), ("trans_a", _i32),
        ("B", _p), ("ldb", _i64), ("trans_b", _i32),
        ("C", _p), ("ldc", _i64),
        ("ab_dtype", _i32), ("c_dtype", _i32),
        ("bias", _p),
        ("act", _i32), ("dact", _i32),
        ("preact", _p), ("ld_preact", _i64),
        ("dact_src", _p), ("ld_dact", _i64),
        ("residual", _p), ("ld_res", _i64),
        ("row_mask", _p),
        ("mask_before_residual", _i32),
        ("preact_is_grad", _i32),
        ("beta", _f32),
    ]


class SkinnyLinearArgs(C.Structure):
    _fields_ = [
        ("x", _p), ("ldx", _i64),
        ("w", _p), ("ldw", _i64),
        ("y", _p), ("ldy", _i64),
        ("bias", _p),
        ("residual", _p), ("ld_res", _i64),
        ("row_mask", _p),
        ("B", _i64), ("N", _i64), ("K", _i64),
        ("y_dtype", _i32), ("act", _i32), ("mask_before_residual", _i32), ("ss_inv_k", _f32), ("ss_eps", _f32),
        ("row_ss_in", _p), ("row_ss_out", _p), ("zero_ss", _p),
    ]


class DecodeLinearArgs(C.Structure):
    _fields_ = [
        ("B", _i64), ("N", _i64), ("K", _i64),
        ("x", _p), ("ldx", _i64),
        ("W", _p), ("ldw", _i64),
        ("norm_scale", _p), ("x_ss", _p), ("norm_eps", _f32),
        ("bias", _p),
        ("act", _i32),
        ("residual", _p), ("ld_res", _i64),
        ("y", _p), ("ldy", _i64),
        ("y_f32", _p), ("ldy_f32", _i64),
        ("y_ss", _p),
        ("zero_ss", _p),
        ("allow_overlap", _i32),
    ]


class LatentFrontArgs(C.Structure):
    _fields_ = [
        ("B", _i64), ("T", _i64),
        ("latent_dim", _i32), ("emb_dim", _i32), ("vocab", _i32),
        ("h_enc", _p), ("eps", _p), ("ids", _p), ("mask", _p), ("init_state", _p),
        ("w_mean", _p), ("b_mean", _p), ("w_logstd", _p), ("b_logstd", _p),
        ("tok_emb", _p), ("w_fuse", _p), ("b_fuse", _p),
        ("temperature", _f32),
        ("mean", _p), ("logstd", _p), ("z", _p), ("log_q", _p),
        ("u", _p), ("u_shift", _p),
        ("act_dtype", _i32),
    ]


class LatentFrontBwdArgs(C.Structure):
    _fields_ = [
        ("f", LatentFrontArgs),
        ("d_z", _p), ("d_log_q", _p), ("d_mean_out", _p), ("d_logstd_out", _p),
        ("d_u", _p), ("d_u_shift", _p),
        ("d_h_enc", _p),
        ("d_w_mean", _p), ("d_b_mean", _p), ("d_w_logstd", _p), ("d_b_logstd", _p),
        ("d_tok_emb", _p), ("d_w_fuse", _p), ("d_b_fuse", _p),
    ]


class LatentBackArgs(C.Structure):
    _fields_ = [
        ("M", _i64),
        ("latent_dim", _i32), ("hidden", _i32), ("n_layers", _i32),
        ("head", _p), ("head_ld", _i64),
        ("z", _p), ("log_q", _p), ("mask", _p),
        ("w1", _p), ("b1", _p), ("ln_w", _p), ("ln_b", _p), ("w2", _p), ("b2", _p),
        ("ln_eps", _f32), ("scale_lo", _f32), ("scale_hi", _f32),
        ("log_p", _p), ("y", _p), ("kl_frame", _p), ("kl_sum", _p),
    ]


class LatentBackBwdArgs(C.Structure):
    _fields_ = [
        ("f", LatentBackArgs),
        ("d_log_p", _p), ("d_head", _p), ("d_z", _p),
        ("d_w1", _p), ("d_b1", _p), ("d_ln_w", _p), ("d_ln_b", _p), ("d_w2", _p), ("d_b2", _p),
    ]


# name -> (restype, argtypes); mirrors include/vgslm.h one to one (tests/test_abi.py checks the set)
_SIGNATURES = {
    "vg_version": (C.c_int, []),
    "vg_last_error_string": (C.c_char_p, []),
    "vg_device_is_sm100": (C.c_int, []),
    "vg_rmsnorm_fwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _f32, C.c_int, C.c_int, _p]),
    "vg_rmsnorm_bwd_workspace": (_sz, [_i64, _i64]),
    "vg_rmsnorm_bwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _p, _f32, _p, _sz, _i64, _i64, C.c_int, C.c_int, _p]),
    "vg_gemm_workspace": (_sz, [C.POINTER(GemmArgs), C.c_int]),
    "vg_gemm": (C.c_int, [C.POINTER(GemmArgs), C.c_int, _p, _sz, _p]),
    "vg_set_gemm_sm_budget": (C.c_int, [C.c_int]),
    "vg_set_pdl_mode": (C.c_int, [C.c_int]),
    "vg_skinny_linear": (C.c_int, [_p, _p]),
    "vg_im2col_fwd": (C.c_int, [_p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, C.c_int, C.c_int, _p]),
    "vg_im2col_bwd": (C.c_int, [_p, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, _i64, C.c_int, _p]),
    "vg_colsum_workspace": (_sz, [_i64, _i64]),
    "vg_colsum": (C.c_int, [_p, _i64, _p, _i64, _i64, C.c_int, _f32, _p, _sz, _p]),
    "vg_mask_rows": (C.c_int, [_p, _p, _p, _i64, _i64, C.c_int, _p]),
    "vg_act_bwd": (C.c_int, [_p, _p, _p, _i64, C.c_int, C.c_int, _p]),
    "vg_attn_fwd": (C.c_int, [_p, _p, _p, _i64, _i64, _p, _i64, _p, _p, _p,
                              _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _f32, C.c_int, _p]),
    "vg_set_attn_backend": (C.c_int, [C.c_int]),
    "vg_attn_bwd_workspace": (_sz, [_i64, _i64, _i64, _i64, _i64]),
    "vg_attn_bwd": (C.c_int, [_p, _i64, _p, _p, _p, _i64, _i64, _p, _i64, _p, _p, _p, _p, _i64, _i64, _p, _p,
                              _i64, _i64, _i64, _i64, _i64, _i64, _f32, C.c_int, _p, _sz, _p]),
    "vg_attn_decode_workspace": (_sz, [_i64, _i64, _i64, _i64]),
    "vg_attn_decode": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _i64, _i64, _p, _i64, _f32, C.c_int,
                                 _p, _sz, _p, _p]),
    "vg_add_i32": (C.c_int, [_p, _i32, _p]),
    "vg_decode_linear_workspace": (_sz, [_i64, _i64]),
    "vg_decode_linear": (C.c_int, [C.POINTER(DecodeLinearArgs), _p, _sz, _p]),
    "vg_debug_decode_linear_trace": (C.c_int, [_p]),
    "vg_debug_attn_trace": (C.c_int, [_p]),
    "vg_kv_append": (C.c_int, [_p, _p, _i64, _p, _p, _i64, _i64, _i64, _i64, _i64, _i64, C.c_int, _p]),
    "vg_latent_front_fwd": (C.c_int, [C.POINTER(LatentFrontArgs), _p]),
    "vg_latent_front_bwd_workspace": (_sz, [_i64, _i64, _i32, _i32, _i32]),
    "vg_latent_front_bwd": (C.c_int, [C.POINTER(LatentFrontBwdArgs), _p, _sz, _p]),
    "vg_latent_back_workspace": (_sz, [_i64, _i32, _i32, _i32]),
    "vg_latent_back_fwd": (C.c_int, [C.POINTER(LatentBackArgs), _p, _sz, _p]),
    "vg_latent_back_bwd": (C.c_int, [C.POINTER(LatentBackBwdArgs), _p, _sz, _p]),
    "vg_latent_prior_sample": (C.c_int, [C.POINTER(LatentBackArgs), _p, _f32, _p, _p]),
    "vg_dwconv_ln_fwd": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _p, _p, _i64, _i64, _i64, _i32, _i32, _f32,
                                   C.c_int, _p]),
    "vg_dwconv_ln_bwd_workspace": (_sz, [_i64, _i64, _i64, _i32]),
    "vg_dwconv_ln_bwd": (C.c_int, [_p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _sz,
                                   _i64, _i64, _i64, _i32, _i32, C.c_int, _p]),
    "vg_softmax_ce_workspace": (_sz, [_i64]),
    "vg_softmax_ce_fwd": (C.c_int, [_p, _i64, _p, _p, _p, _p, _p, _i64, _i64, C.c_int, _p, _sz, _p]),
    "vg_softmax_ce_bwd": (C.c_int, [_p, _i64, _p, _p, _p, _p, _p, _i64, _i64, _i64, C.c_int, _p]),
    "vg_qsample": (C.c_int, [_p, _p, _p, _p, _p, _p, _f32, _p, _p, _i64, _i64, _i64, _p]),
    "vg_masked_l1_workspace": (_sz, [_i64, _i64, _i64]),
    "vg_masked_l1_fwd": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _i64, _p, _sz, _p]),
    "vg_masked_l1_bwd": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, _i64, _p]),
    "vg_sample_token": (C.c_int, [_p, _i64, _p, _f32, _p, _i64, _i64, C.c_int, _p]),
    "vg_adamw_step": (C.c_int, [_p, _p, _p, _p, _p, _i64, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _f32, _p, _p]),
    "vg_zero_segments": (C.c_int, [_p, _p, _p, C.c_int, _p]),
    "vg_cast_f32_to_bf16": (C.c_int, [_p, _p, _i64, _p]),
    "vg_decode_step_task_bytes": (_sz, []),
    "vg_decode_step_smem_bytes": (_sz, [_i32]),
    "vg_decode_step": (C.c_int, [_p, _p]),
}

_lib: Optional[C.CDLL] = None
launch_count = 0   # kernels of libvgslm launched by this process (bench.py reports it)
_KERNELS_PER_CALL = {"vg_rmsnorm_bwd": 2, "vg_colsum": 2, "vg_attn_bwd": 3, "vg_dwconv_ln_bwd": 5, "vg_latent_front_bwd": 3,
                     "vg_latent_back_fwd": 2, "vg_latent_back_bwd": 2, "vg_softmax_ce_fwd": 2, "vg_masked_l1_fwd": 2}


def load() -> C.CDLL:
    """dlopen libvgslm.so (once) and attach the signatures.  Raises if the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA kernel library has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` from the repo root. "
            "There is no CPU fallback for the product path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES.keys())


def last_error() -> str:
    return load().vg_last_error_string().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"libvgslm {what} failed (rc={rc}): {last_error()}")


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise on a non-zero return code."""
    global launch_count
    rc = getattr(load(), name)(*args)
    if rc != 0:
        raise RuntimeError(f"libvgslm {name} failed (rc={rc}): {last_error()}")
    launch_count += _KERNELS_PER_CALL.get(name, 1)


def set_attention_backend(backend: str) -> None:
    """'auto' (tcgen05 kernels for packed bf16, CUDA-core kernels otherwise), 'simt' or 'tcgen05'."""
    call("vg_set_attn_backend", {"auto": 0, "simt": 1, "tcgen05": 2}[backend])


def dtype_id(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return VG_F32
    if dt == torch.bfloat16:
        return VG_BF16
    raise TypeError(f"libvgslm supports float32 and bfloat16 activations, got {dt}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device pointer of a CUDA tensor (None → NULL).  Refuses CPU tensors: no CPU path exists."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("libvgslm ops need CUDA tensors: the product path has no CPU fallback")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


_ws_cache = {}


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Scratch buffer from the torch caching allocator (caller-owned memory, per the ABI contract).

    A fresh tensor per call keeps stream-ordering and CUDA-graph capture semantics trivially right."""
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)

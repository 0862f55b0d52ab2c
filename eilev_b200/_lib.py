"""ctypes binding of ``libvideoblip_b200.so`` (C ABI: ``include/videoblip_b200.h``).

The shared library is built in-tree by ``build.sh`` / ``__graft_entry__.build()``.  There is
no CPU or PyTorch fallback: if the library is missing, or a call fails, we raise.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_PKG = Path(__file__).resolve().parent
_ROOT = _PKG.parent
# VB_LIB_PATH: an instrumented build of the same library (scripts/micro/gemm_trace.sh); measurement only
LIB_PATH = Path(os.environ["VB_LIB_PATH"]) if os.environ.get("VB_LIB_PATH") else _PKG / "libvideoblip_b200.so"

VB_BF16, VB_F32, VB_F16 = 0, 1, 2
EPI_NONE, EPI_GELU, EPI_RELU = 0, 1, 2
EPI_GELU_BWD, EPI_RELU_BWD = 3, 4  # C = acc * act'(residual): activation backward fused into the dgrad GEMM
GEMM_AUTO, GEMM_TCGEN05, GEMM_GENERIC = 0, 1, 2

i64, i32, f32, vp = C.c_int64, C.c_int32, C.c_float, C.c_void_p


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", vp), ("b", vp), ("c", vp), ("bias", vp), ("residual", vp),
        ("m", i64), ("n", i64), ("k", i64),
        ("lda", i64), ("ldb", i64), ("ldc", i64), ("ldr", i64),
        ("alpha", f32), ("beta", f32),
        ("alpha_cols", i64), ("row_group", i64),
        ("epilogue", i32), ("out_dtype", i32), ("backend", i32), ("reserved", i32),
        ("dropout_p", f32), ("reserved2", i32), ("dropout_seed", vp), ("dropout_salt", C.c_uint64),
        ("ln_stats", vp), ("ln_colsum", vp), ("ln_eps", f32), ("reserved3", i32), ("stats_out", vp), ("stats_zero", vp),
    ]


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", vp), ("k", vp), ("v", vp), ("o", vp), ("lse", vp), ("key_mask", vp),
        ("batch", i64), ("heads", i64), ("sq", i64), ("skv", i64), ("d", i64),
        ("q_bs", i64), ("q_rs", i64), ("k_bs", i64), ("k_rs", i64),
        ("v_bs", i64), ("v_rs", i64), ("o_bs", i64), ("o_rs", i64),
        ("scale", f32), ("causal", i32),
        ("dropout_p", f32), ("reserved", i32), ("dropout_seed", vp), ("dropout_salt", C.c_uint64),
        ("rel_bias", vp), ("rel_bias_stride", i64),
    ]


class AttnBwdArgs(C.Structure):
    _fields_ = [
        ("fwd", AttnArgs), ("d_o", vp), ("dq", vp), ("dk", vp), ("dv", vp),
        ("dq_bs", i64), ("dq_rs", i64), ("dk_bs", i64), ("dk_rs", i64),
        ("dv_bs", i64), ("dv_rs", i64),
        ("delta", vp), ("dq_acc", vp), ("dq_scale", f32), ("reserved", i32),
    ]


# name -> argtypes; every entry is declared in include/videoblip_b200.h
SIGNATURES: dict[str, list] = {
    "vb_abi_version": [],
    "vb_last_error": [],
    "vb_device_arch": [],
    "vb_gemm": [C.POINTER(GemmArgs), vp],
    "vb_gemm_uses_tcgen05": [C.POINTER(GemmArgs)],
    "vb_layernorm": [vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, f32, vp],
    "vb_layernorm_bwd": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, f32, vp],
    "vb_row_stats": [vp, vp, i64, i64, i64, vp],
    "vb_layernorm_bwd_dropout": [vp, vp, vp, vp, vp, vp, vp, vp, f32, vp, C.c_uint64, i64, i64, vp],
    "vb_crop_resize_normalize_u8": [vp, i64, i64, i64, i64, i64, i64, i64, i64, i32, vp, i32, i64, i64, C.c_double,
                                    C.POINTER(C.c_float), C.POINTER(C.c_float), vp],
    "vb_attention_fwd": [C.POINTER(AttnArgs), vp],
    "vb_attention_probs": [C.POINTER(AttnArgs), vp, i32, vp],
    "vb_attention_uses_tcgen05": [C.POINTER(AttnArgs)],
    "vb_attention_bwd": [C.POINTER(AttnBwdArgs), vp],
    "vb_attention_bwd_uses_tcgen05": [C.POINTER(AttnBwdArgs)],
    "vb_patch_gather": [vp, i32, vp, i64, i64, i64, i64, i64, i64, i64, vp],
    "vb_patch_gather_u8": [vp, vp, i64, i64, i64, i64, i64, i64, i64, C.c_double, C.POINTER(f32), C.POINTER(f32), vp],
    "vb_resize_bicubic_ksize": [i64, i64],
    "vb_resize_bicubic_coeffs": [i64, i64, C.POINTER(i32), C.POINTER(i32), i64],
    "vb_resize_u8_pass": [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, i64, i64, i64, i32, vp],
    "vb_cls_rows": [vp, vp, vp, i64, i64, i64, vp],
    "vb_embed_splice": [vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, vp],
    "vb_splice_bwd": [vp, vp, vp, i64, i64, i64, vp],
    "vb_cross_entropy": [vp, i32, vp, vp, vp, vp, i64, i64, i64, i64, i32, vp],
    "vb_cross_entropy_bwd": [vp, i32, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i32, vp],
    "vb_rmsnorm": [vp, vp, vp, vp, i64, i64, i64, i64, f32, vp],
    "vb_rmsnorm_bwd": [vp, vp, vp, vp, vp, vp, i64, i64, vp],
    "vb_gated_gelu": [vp, vp, i64, i64, vp],
    "vb_gated_gelu_bwd": [vp, vp, vp, i64, i64, vp],
    "vb_embedding": [vp, vp, vp, i64, i64, i64, vp],
    "vb_attention_merge": [vp, vp, i64, vp, vp, i64, vp, i64, i64, i64, vp],
    "vb_token_logprob": [vp, i32, vp, vp, vp, i64, i64, i64, vp],
    "vb_transpose": [vp, vp, i64, i64, i64, i64, vp],
    "vb_convert": [vp, i32, vp, i32, i64, vp],
    "vb_act_bwd": [vp, vp, vp, i32, i64, vp],
    "vb_colsum": [vp, vp, i64, i64, i64, i32, vp],
    "vb_dropout": [vp, vp, i64, i64, i64, i64, f32, vp, C.c_uint64, vp],
    "vb_add": [vp, vp, vp, i64, vp],
    "vb_adamw": [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i64, vp, vp],
    "vb_sumsq": [vp, i64, vp, vp],
    "vb_gemv": [vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, f32, i64, i32, i32, vp, vp, f32, vp],
    "vb_decode_embed": [vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, vp],
    "vb_debug_decode_trace": [vp],
    "vb_decode_step": [vp, vp, i32, i32, vp, vp],
    "vb_paged_decode_attention": [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, f32,
                                  vp, i64, i64, vp],
    "vb_decode_cross_attention": [vp, i64, vp, vp, i64, vp, vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, f32, vp],
    "vb_paged_kv_write": [vp, vp, i64, vp, vp, vp, i64, i64, i64, i64, i64, vp],
}

_lib = None


class VbError(RuntimeError):
    pass


def build(verbose: bool = False) -> Path:
    """Compile the CUDA extension for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["bash", str(_ROOT / "build.sh")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise VbError(f"nvcc build failed (exit {r.returncode})")
    return LIB_PATH


def lib() -> C.CDLL:
    """Load (once) and return the shared library with typed entry points."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise VbError(
            f"{LIB_PATH} is missing: run ./build.sh (or __graft_entry__.build()). "
            "The VideoBLIP path has no CPU/PyTorch fallback."
        )
    handle = C.CDLL(os.fspath(LIB_PATH))
    for name, argtypes in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == "vb_last_error" else C.c_int
    if handle.vb_abi_version() != 6:
        raise VbError("ABI version mismatch between eilev_b200/_lib.py and the built library")
    _lib = handle
    return handle


# kernels launched per successful C-ABI call (lower bounds), for bench.py's gpu_launches
_KERNELS_PER_CALL = {"vb_cross_entropy": 2, "vb_embed_splice": 2, "vb_attention_bwd": 2}
_launches = 0


def launch_count() -> int:
    """Number of native kernel launches issued through this binding so far."""
    return _launches


def check(status: int, what: str) -> None:
    global _launches
    _launches += _KERNELS_PER_CALL.get(what, 1)
    if status != 0:
        msg = lib().vb_last_error()
        raise VbError(f"{what} failed: {msg.decode() if msg else 'unknown error'}")

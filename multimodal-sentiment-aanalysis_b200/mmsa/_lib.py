"""ctypes binding of libmmsa.so (the C ABI declared in include/mmsa.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the caller
gets an exception.  PyTorch is only used for device memory and streams."""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_double, c_float, c_int, c_int64, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MMSA_LIB_PATH") or os.path.join(_HERE, "libmmsa.so")      # override: A/B builds of the library (probes)
HEADER_PATH = os.path.normpath(os.path.join(_HERE, "..", "..", "include", "mmsa.h"))

F32, BF16 = 0, 1
ACT_NONE, ACT_SIGMOID, ACT_GELU, ACT_RELU = 0, 1, 2, 3
BN_THEN_GELU, RELU_THEN_BN, BN_ONLY = 0, 1, 2
LOSS_INFONCE, LOSS_SUPCON, LOSS_NTXENT = 0, 1, 2

P, I, L, F, U, Dbl = c_void_p, c_int, c_int64, c_float, c_uint64, c_double

# name -> (restype, argtypes); must cover every symbol include/mmsa.h declares
_PROTOS = {
    "mmsa_version": (ctypes.c_char_p, []),
    "mmsa_last_error": (ctypes.c_char_p, []),
    "mmsa_check_device": (I, []),
    "mmsa_launch_count": (L, []),
    "mmsa_prof_enable": (None, [I]),
    "mmsa_prof_collect": (I, [P, I, P, P, P, I]),
    "mmsa_cast": (I, [P, I, P, I, L, P]),
    "mmsa_cast_multi": (I, [I, P, P, P, I, I, P]),
    "mmsa_split3": (I, [P, L, L, L, P, I, P, I, P]),
    "mmsa_linear_fwd": (I, [I, L, L, L, L, P, L, P, L, P, L, P, P, L, I, P, L, I, P]),
    "mmsa_debug_gemm": (I, [I, I, L, L, L, P, L, P, L, P, L, I, I, P]),
    "mmsa_linear_dgrad": (I, [I, L, L, L, P, L, P, L, P, L, P, L, I, P]),
    "mmsa_linear_wgrad_workspace": (L, [I, L, L, L]),
    "mmsa_linear_wgrad": (I, [I, L, L, L, P, L, P, L, P, L, P, P, P]),
    "mmsa_linear_wgrad2": (I, [I, L, L, L, L, P, L, P, L, P, L, P, L, P, P, P]),
    "mmsa_debug_force_simt_attention": (None, [I]),
    "mmsa_debug_attention_engine": (None, [I]),
    "mmsa_debug_gemm_engine": (None, [I]),
    "mmsa_attn_fwd": (I, [I, L, L, L, L, L, P, L, P, L, P, L, P, L, P, P]),
    "mmsa_attn_bwd": (I, [I, L, L, L, L, L, P, L, P, L, P, L, P, L, P, L, P, P, P, L, P, L, P, L, P]),
    "mmsa_attn_dropout_fwd": (I, [I, L, L, L, L, L, P, L, P, L, P, L, P, L, P, F, P, U, U, P, P]),
    "mmsa_attn_dropout_bwd": (I, [I, L, L, L, L, L, P, L, P, L, P, L, P, L, P, L, P, P, P, L, P, L, P, L, F, P, U, U, P, P]),
    "mmsa_gate_ln_fwd": (I, [I, L, L, P, P, P, P, P, F, P, P, P, P, P]),
    "mmsa_gate_ln_bwd_blocks": (L, [L]),
    "mmsa_gate_ln_bwd": (I, [I, L, L, P, L, P, P, P, P, P, P, P, L, P, P, P, L, P, P, P, P, P]),
    "mmsa_gate_ln_pool_fwd": (I, [I, L, L, L, P, P, P, P, P, F, P, P, P, P, P, P, P]),
    "mmsa_gate_ln_pool_bwd": (I, [I, L, L, L, P, P, P, P, P, P, P, P, P, P, P, L, P, P, P, P, P]),
    "mmsa_add_ln_fwd": (I, [I, L, L, P, P, P, P, F, P, P, P, P]),
    "mmsa_add_ln_bwd": (I, [I, L, L, P, P, P, P, P, P, P, P, P, P, P]),
    "mmsa_add_rows": (I, [I, L, L, L, P, P, P, P]),
    "mmsa_pool_fwd": (I, [I, L, L, L, P, I, P, P, P]),
    "mmsa_pool_bwd": (I, [I, L, L, L, P, I, P, P, P]),
    "mmsa_modal_concat_fwd": (I, [I, L, L, I, P, P, P, P, P]),
    "mmsa_modal_concat_bwd": (I, [I, L, L, I, P, P, P, P, P, P]),
    "mmsa_modal_head_fwd": (I, [I, L, L, I, L, P, P, P, P, P, P, P, P, P]),
    "mmsa_modal_head_bwd": (I, [I, L, L, I, L, P, P, P, P, P, P, P, P, P]),
    "mmsa_act_fwd": (I, [I, L, P, I, P, P]),
    "mmsa_act_bwd": (I, [I, L, P, P, I, P, P]),
    "mmsa_linear_bn_act_supported": (I, [L, L, L, L, L]),
    "mmsa_linear_bn_act_fwd": (I, [L, L, L, P, L, P, L, P, P, P, P, P, P, F, F, I, I, F, P, I, U, U, P, P, I, P, P, P, P, P]),
    "mmsa_bn_act_fwd": (I, [I, L, L, I, P, P, P, P, P, P, F, F, I, F, P, I, U, U, P, P, P, P, P, P]),
    "mmsa_bn_act_bwd": (I, [I, L, L, I, P, P, P, P, P, P, I, F, P, P, P, P, P, P]),
    "mmsa_dropout": (I, [I, L, P, F, P, I, U, U, P, P, P]),
    "mmsa_rng_advance": (I, [P, U, P]),
    "mmsa_ce_fwd": (I, [L, L, P, P, P, L, P, P, P, P]),
    "mmsa_ce_bwd": (I, [I, L, L, P, P, P, P, P]),
    "mmsa_l2norm_fwd": (I, [I, L, L, P, P, P, P]),
    "mmsa_l2norm_bwd": (I, [I, L, L, P, P, P, P, P, P]),
    "mmsa_contrastive_fwd": (I, [I, L, L, L, P, P, P, P, F, L, P, P, P, P, P, P]),
    "mmsa_contrastive_bwd": (I, [I, L, L, L, P, P, P, P, F, L, P, P, P, P, P, I, P, P, P, P]),
    "mmsa_sumsq": (I, [P, L, P, L, P, P]),
    "mmsa_clip_adamw": (I, [P, P, P, P, L, P, F, Dbl, Dbl, Dbl, Dbl, Dbl, L, P]),
}

_lib = None


def header_symbols():
    """Every function name include/mmsa.h declares."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mmsa_[a-z0-9_]+)\s*\(", text)))


def load():
    """dlopen libmmsa.so and bind every prototype.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"mmsa: CUDA extension {LIB_PATH} is missing. Build it with "
            f"`python __graft_entry__.py` (nvcc, sm_100a). There is no CPU/PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MmsaError(RuntimeError):
    pass


def call(name: str, *args):
    """Invoke an int-returning entry point; raise MmsaError with mmsa_last_error() on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise MmsaError(f"{name} failed (code {rc}): {lib.mmsa_last_error().decode()}")


def launch_count() -> int:
    return int(load().mmsa_launch_count())


def prof_enable(on: bool) -> None:
    """Start (clearing earlier records) or stop per-launch device timing (include/mmsa.h)."""
    load().mmsa_prof_enable(1 if on else 0)


def prof_collect(max_entries: int = 128):
    """-> {kernel name: {"count", "ms", "work"}} summed since prof_enable(True); synchronises."""
    stride = 48
    names = ctypes.create_string_buffer(stride * max_entries)
    counts = (ctypes.c_int64 * max_entries)()
    ms = (ctypes.c_double * max_entries)()
    work = (ctypes.c_double * max_entries)()
    n = load().mmsa_prof_collect(ctypes.cast(names, c_void_p), stride, ctypes.cast(counts, c_void_p),
                                 ctypes.cast(ms, c_void_p), ctypes.cast(work, c_void_p), max_entries)
    out = {}
    for i in range(n):
        nm = names.raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode()
        out[nm] = {"count": int(counts[i]), "ms": float(ms[i]), "work": float(work[i])}
    return out

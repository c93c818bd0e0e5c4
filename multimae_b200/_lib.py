"""ctypes binding of libmultimae_b200.so (the C ABI declared in include/multimae_b200.h).

The library is the product: there is no Python/CPU fallback.  `lib()` raises if the shared object is missing
or was built for a different ABI version.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmultimae_b200.so")

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_i64 = ctypes.c_int64
c_float = ctypes.c_float


class GemmEpilogue(ctypes.Structure):
    _fields_ = [
        ("alpha", c_float),
        ("act", c_int),
        ("accumulate", c_int),
        ("reserved", c_int),
        ("bias", c_void_p),
        ("residual", c_void_p),
        ("dgelu_z", c_void_p),
        ("preact_bf16", c_void_p),
        ("out_f32", c_void_p),
        ("out_bf16", c_void_p),
        ("ld_residual", c_i64),
        ("ld_dgelu_z", c_i64),
        ("ld_preact", c_i64),
        ("ld_out_f32", c_i64),
        ("ld_out_bf16", c_i64),
    ]


class MmaeError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol of include/multimae_b200.h appears here
SIGNATURES = {
    "mmae_abi_version": (c_int, []),
    "mmae_last_error": (ctypes.c_char_p, []),
    "mmae_launch_count": (c_i64, []),
    "mmae_gemm_bf16": (c_int, [c_void_p, c_i64, c_int, c_void_p, c_i64, c_int, c_int, c_int, c_int, c_int,
                               ctypes.POINTER(GemmEpilogue), c_void_p]),
    "mmae_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_i64, c_void_p]),
    "mmae_cast_colsum_f32": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_int, c_int, c_void_p]),
    "mmae_colsum_bf16": (c_int, [c_void_p, c_i64, c_void_p, c_int, c_int, c_void_p]),
    "mmae_transpose_bf16": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_int, c_int, c_void_p]),
    "mmae_layernorm_forward": (c_int, [c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64,
                                       c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "mmae_layernorm_backward": (c_int, [c_void_p, c_int, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mmae_attention_forward": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p,
                                       c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "mmae_attention_backward": (c_int, [c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64, c_void_p,
                                        c_i64, c_void_p, c_void_p, c_void_p, c_i64, c_void_p, c_i64, c_void_p, c_i64,
                                        c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
}

ABI_VERSION = 1


def lib():
    """Load (once) and return the ctypes handle.  Fails loudly when the CUDA library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MmaeError(
            "multimae_b200: %s is missing - build it with `python -m multimae_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    ver = handle.mmae_abi_version()
    if ver != ABI_VERSION:
        raise MmaeError("multimae_b200: ABI version mismatch (library %d, binding %d)" % (ver, ABI_VERSION))
    _lib = handle
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().mmae_last_error()
        raise MmaeError("%s failed (rc=%d): %s" % (what or "mmae call", rc, (msg or b"").decode()))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()


def current_stream():
    import torch
    return torch.cuda.current_stream().cuda_stream

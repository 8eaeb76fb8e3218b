"""ctypes binding of include/merlot_b200.h (the C-ABI drop-in boundary).

The library is the product path: if it cannot be loaded this module raises -- there is no CPU or PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libmerlot_b200.so"

MERLOT_OK = 0
MERLOT_EINVAL = -1
MERLOT_ESHAPE = -2
MERLOT_ECUDA = -3
MERLOT_ENOTIMPL = -4

GEMM_OUT_F32 = 1
GEMM_ATOMIC = 2
GEMM_GELU = 4
GEMM_MUL_DGELU = 8
GEMM_DROPOUT = 16


class MerlotError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"merlot_b200 error {code}: {msg}")
        self.code = code


class MerlotShapeError(MerlotError, ValueError):
    """Mirrors the ValueError / assert the reference raises at graph-build time (utils/model_utils.py:29-56)."""


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("a", C.c_void_p), ("lda", C.c_int), ("a_mn_major", C.c_int),
        ("b", C.c_void_p), ("ldb", C.c_int), ("b_mn_major", C.c_int),
        ("out", C.c_void_p), ("ld_out", C.c_int),
        ("out2", C.c_void_p), ("ld_out2", C.c_int),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p), ("ld_resid", C.c_int),
        ("aux", C.c_void_p), ("ld_aux", C.c_int),
        ("alpha", C.c_float),
        ("flags", C.c_uint32),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64), ("dropout_site", C.c_uint32),
        ("splits", C.c_int),
        ("block_n", C.c_int),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load libmerlot_b200.so (built in-tree by merlot_b200.build). Raises if missing -- no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} not found: run `python -m merlot_b200.build` (or __graft_entry__.build()). "
            "merlot_b200 has no CPU/PyTorch fallback path.")
    l = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
    l.merlot_last_error.restype = C.c_char_p
    l.merlot_abi_version.restype = C.c_int
    l.merlot_launch_count.restype = C.c_longlong
    l.merlot_reset_launch_count.restype = None
    _lib = l
    return l


def check(rc: int) -> None:
    if rc == MERLOT_OK:
        return
    msg = lib().merlot_last_error().decode("utf-8", "replace")
    if rc == MERLOT_ESHAPE:
        raise MerlotShapeError(rc, msg)
    if rc == MERLOT_ENOTIMPL:
        raise NotImplementedError(msg)
    raise MerlotError(rc, msg)


def exported_symbols_from_header() -> list[str]:
    """Every function name declared in include/merlot_b200.h (used by the CPU test that checks the .so exports them)."""
    import re
    hdr = (_HERE.parent / "include" / "merlot_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(merlot_[a-z0-9_]+)\s*\(", hdr)))


class AttnDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("S", C.c_int), ("heads", C.c_int), ("head_dim", C.c_int),
        ("qkv", C.c_void_p), ("ld_qkv", C.c_int),
        ("valid", C.c_void_p),
        ("scale", C.c_float),
        ("ctx", C.c_void_p), ("ld_ctx", C.c_int),
        ("lse", C.c_void_p),
        ("d_ctx", C.c_void_p),
        ("dsum", C.c_void_p),
        ("dq_accum", C.c_void_p), ("ld_dq", C.c_int),
        ("dqkv", C.c_void_p), ("ld_dqkv", C.c_int),
        ("colsum", C.c_void_p),
    ]

"""ctypes binding of the C ABI declared in ``include/fragnet_b200.h``.

The product path has NO CPU fallback: if the library cannot be loaded the import of any module that
needs a kernel fails loudly here.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_vp, _i64, _i32, _sz, _u64, _f32 = C.c_void_p, C.c_int64, C.c_int, C.c_size_t, C.c_uint64, C.c_float

# name -> (restype, argtypes).  Mirrors include/fragnet_b200.h one to one (tests check the export list).
SIGNATURES = {
    "fnb_version": (C.c_int, []),
    "fnb_error_string": (C.c_char_p, [C.c_int]),
    "fnb_launch_count": (_u64, []),
    "fnb_scratch_bytes": (_sz, []),
    "fnb_csr_workspace_bytes": (_sz, [_i64, _i64]),
    "fnb_csr_build": (C.c_int, [_vp, _vp, _i64, _i64, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp]),
    "fnb_gather_rows": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "fnb_segment_offsets": (C.c_int, [_vp, _i64, _i64, _vp, _vp]),
    "fnb_narrow_index": (C.c_int, [_vp, _i64, _vp, _vp]),
    "fnb_proj_fwd": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _vp]),
    "fnb_proj_bwd": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "fnb_node_scalars": (C.c_int, [_vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp]),
    "fnb_edge_coef_fwd": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp]),
    "fnb_edge_coef_bwd": (C.c_int, [_vp, _vp, _i32, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "fnb_gat_fwd": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i64,
                              _vp, _i32, _vp, _vp]),
    "fnb_attn_by_source": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp]),
    "fnb_gat_bwd_dst": (C.c_int, [_vp, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fnb_gat_bwd_src": (C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp,
                                  _vp, _vp, _vp]),
    "fnb_edge_table_bwd": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "fnb_segment_sum": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp]),
    "fnb_segment_gather": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp]),
    "fnb_dropout_relu_fwd": (C.c_int, [_vp, _vp, _i64, _f32, _i32, _i32, _u64, _u64, _vp]),
    "fnb_dropout_relu_bwd": (C.c_int, [_vp, _vp, _vp, _i64, _f32, _i32, _vp]),
}

EDGE_NONE, EDGE_AFFINE1, EDGE_AFFINE6, EDGE_TABLE = 0, 1, 2, 3
PRECISION_FP32, PRECISION_TF32 = 0, 1

_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def load():
    """Load (building first if the in-tree library is missing or stale and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    if _build.nvcc_path() is not None:
        path = _build.build()
    if not os.path.isfile(path):
        raise RuntimeError(
            f"fragnet_b200: kernel library {path} is missing and nvcc is not available to build it. "
            "There is no CPU fallback; run `python -c 'import __graft_entry__ as g; g.build()'` first.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == ABI / header mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    if lib.fnb_version() != 1:
        raise RuntimeError(f"fragnet_b200: ABI version mismatch ({lib.fnb_version()} != 1)")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().fnb_error_string(rc).decode()
        raise RuntimeError(f"fragnet_b200.{what} failed: {msg} (code {rc})")

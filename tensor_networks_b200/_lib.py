"""ctypes binding of libttb200.so (the C ABI declared in include/ttb200.h).

The shared library is built in-tree by `__graft_entry__.build()` /
`make -C tensor_networks_b200/csrc`.  There is deliberately NO fallback: if the
library is missing or a call fails, the product path raises.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import (
    POINTER,
    Structure,
    c_char_p,
    c_double,
    c_int,
    c_int32,
    c_int64,
    c_size_t,
    c_uint64,
    c_void_p,
)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libttb200.so")


class TTBError(RuntimeError):
    """A ttb200 C-ABI call returned a non-zero status."""

    def __init__(self, status: int, message: str):
        super().__init__(f"ttb200 status {status}: {message}")
        self.status = status


class ttb_tt(Structure):
    _fields_ = [
        ("d", c_int32),
        ("n", POINTER(c_int64)),
        ("r", POINTER(c_int64)),
        ("core", POINTER(c_void_p)),
    ]


class ttb_tt_batch(Structure):
    _fields_ = [
        ("d", c_int32),
        ("batch", c_int64),
        ("n", POINTER(c_int64)),
        ("r", POINTER(c_int64)),
        ("core", POINTER(c_void_p)),
    ]


_lib = None


def _declare(lib):
    P = POINTER
    lib.ttb_version.restype = c_char_p
    lib.ttb_version.argtypes = []
    lib.ttb_last_error.restype = c_char_p
    lib.ttb_last_error.argtypes = []
    lib.ttb_launch_count.restype = c_uint64
    lib.ttb_launch_count.argtypes = []

    lib.ttb_gemm_workspace_bytes.restype = c_size_t
    lib.ttb_gemm_workspace_bytes.argtypes = [c_int64, c_int64, c_int64]
    lib.ttb_gemm_f64.restype = c_int
    lib.ttb_gemm_f64.argtypes = [
        c_int64, c_int64, c_int64, c_double, c_void_p, c_int64, c_int64, c_void_p, c_int64,
        c_int64, c_double, c_void_p, c_int64, c_void_p, c_size_t, c_void_p,
    ]
    lib.ttb_gemm_f64_ex.restype = c_int
    lib.ttb_gemm_f64_ex.argtypes = [
        c_int64, c_int64, c_int64, c_double, c_void_p, c_int64, c_int64, c_void_p, c_int64,
        c_int64, c_double, c_void_p, c_int64, c_int, c_int, c_void_p, c_size_t, c_void_p,
    ]

    lib.ttb_gemm_profile_enable.restype = c_int
    lib.ttb_gemm_profile_enable.argtypes = [c_int]
    lib.ttb_gemm_profile_read.restype = c_int
    lib.ttb_gemm_profile_read.argtypes = [P(c_double), P(c_double), P(c_uint64)]

    lib.ttb_inner_workspace_bytes.restype = c_size_t
    lib.ttb_inner_workspace_bytes.argtypes = [P(ttb_tt), P(ttb_tt)]
    lib.ttb_inner_f64.restype = c_int
    lib.ttb_inner_f64.argtypes = [P(ttb_tt), P(ttb_tt), c_void_p, c_void_p, c_size_t, c_void_p]

    lib.ttb_inner_streamed_workspace_bytes.restype = c_size_t
    lib.ttb_inner_streamed_workspace_bytes.argtypes = [P(ttb_tt), P(ttb_tt)]
    lib.ttb_inner_streamed_f64.restype = c_int
    lib.ttb_inner_streamed_f64.argtypes = [
        P(ttb_tt), P(ttb_tt), P(c_void_p), P(c_void_p), c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
    ]

    lib.ttb_round_workspace_bytes.restype = c_size_t
    lib.ttb_round_workspace_bytes.argtypes = [P(ttb_tt)]
    lib.ttb_round_f64.restype = c_int
    lib.ttb_round_f64.argtypes = [
        P(ttb_tt), c_double, c_int32, P(c_int64), P(c_double), P(c_int32), c_void_p, c_size_t, c_void_p,
    ]
    lib.ttb_right_orth_workspace_bytes.restype = c_size_t
    lib.ttb_right_orth_workspace_bytes.argtypes = [P(ttb_tt), c_int32]
    lib.ttb_right_orth_f64.restype = c_int
    lib.ttb_right_orth_f64.argtypes = [P(ttb_tt), c_int32, P(c_int64), c_void_p, c_size_t, c_void_p]
    lib.ttb_orth_rows_workspace_bytes.restype = c_size_t
    lib.ttb_orth_rows_workspace_bytes.argtypes = [c_int64, c_int64]
    lib.ttb_orth_rows_f64.restype = c_int
    lib.ttb_orth_rows_f64.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.ttb_delta_svd_workspace_bytes.restype = c_size_t
    lib.ttb_delta_svd_workspace_bytes.argtypes = [c_int64, c_int64]
    lib.ttb_delta_svd_f64.restype = c_int
    lib.ttb_delta_svd_f64.argtypes = [
        c_void_p, c_int64, c_int64, c_double, c_int32, c_int32, c_void_p, c_void_p, c_void_p,
        P(c_double), c_void_p, c_size_t, c_void_p,
    ]

    lib.ttb_gram_eig_batched_workspace_bytes.restype = c_size_t
    lib.ttb_gram_eig_batched_workspace_bytes.argtypes = [c_int32, c_int32]
    lib.ttb_gram_eig_batched_f64.restype = c_int
    lib.ttb_gram_eig_batched_f64.argtypes = [
        c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]

    lib.ttb_ttsvd_workspace_bytes.restype = c_size_t
    lib.ttb_ttsvd_workspace_bytes.argtypes = [c_int32, P(c_int64)]
    lib.ttb_ttsvd_f64.restype = c_int
    lib.ttb_ttsvd_f64.argtypes = [
        c_void_p, c_int32, P(c_int64), c_double, c_int32, c_void_p, c_size_t, P(c_int64), P(c_double),
        c_void_p, c_size_t, c_void_p,
    ]

    lib.ttb_inner_batched_workspace_bytes.restype = c_size_t
    lib.ttb_inner_batched_workspace_bytes.argtypes = [P(ttb_tt_batch), P(ttb_tt_batch)]
    lib.ttb_inner_batched_f64.restype = c_int
    lib.ttb_inner_batched_f64.argtypes = [P(ttb_tt_batch), P(ttb_tt_batch), c_void_p, c_void_p, c_size_t, c_void_p]

    lib.ttb_inner_batched_scatter_workspace_bytes.restype = c_size_t
    lib.ttb_inner_batched_scatter_workspace_bytes.argtypes = [P(ttb_tt_batch), P(ttb_tt_batch)]
    lib.ttb_inner_batched_scatter_f64.restype = c_int
    lib.ttb_inner_batched_scatter_f64.argtypes = [
        P(ttb_tt_batch), P(ttb_tt_batch), P(c_void_p), c_int32, c_int64, c_void_p, c_size_t, c_void_p,
    ]

    lib.ttb_round_batched_workspace_bytes.restype = c_size_t
    lib.ttb_round_batched_workspace_bytes.argtypes = [P(ttb_tt_batch)]
    lib.ttb_round_batched_f64.restype = c_int
    lib.ttb_round_batched_f64.argtypes = [
        P(ttb_tt_batch), c_double, c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
    ]

    lib.ttb_tt_to_dense_workspace_bytes.restype = c_size_t
    lib.ttb_tt_to_dense_workspace_bytes.argtypes = [P(ttb_tt)]
    lib.ttb_tt_to_dense_f64.restype = c_int
    lib.ttb_tt_to_dense_f64.argtypes = [P(ttb_tt), c_void_p, c_void_p, c_size_t, c_void_p]

    lib.ttb_h2d_staged.restype = c_int
    lib.ttb_h2d_staged.argtypes = [P(c_void_p), P(c_void_p), P(c_size_t), c_int32, c_void_p]
    lib.ttb_strided_copy_f64.restype = c_int
    lib.ttb_strided_copy_f64.argtypes = [c_void_p, c_void_p, c_int32, P(c_int64), P(c_int64), P(c_int64), c_void_p]
    lib.ttb_strided_op_f64.restype = c_int
    lib.ttb_strided_op_f64.argtypes = [
        c_void_p, c_void_p, c_int32, P(c_int64), P(c_int64), P(c_int64), c_int32, c_double, c_void_p,
    ]
    lib.ttb_fill_f64.restype = c_int
    lib.ttb_fill_f64.argtypes = [c_void_p, c_int64, c_double, c_void_p]
    lib.ttb_scale_rows_f64.restype = c_int
    lib.ttb_scale_rows_f64.argtypes = [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int32, c_void_p]
    lib.ttb_diag_f64.restype = c_int
    lib.ttb_diag_f64.argtypes = [c_void_p, c_int64, c_void_p, c_void_p]
    lib.ttb_pack_rounded_cores_f64.restype = c_int
    lib.ttb_pack_rounded_cores_f64.argtypes = [
        c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int32, c_int32, c_int64, c_int64, c_void_p, c_void_p,
    ]
    lib.ttb_pack_rounded_cores_scatter_f64.restype = c_int
    lib.ttb_pack_rounded_cores_scatter_f64.argtypes = [
        c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int32, c_int32, c_int64, c_int64, P(c_void_p), c_int32, c_int64,
        c_void_p,
    ]
    lib.ttb_axpby_f64.restype = c_int
    lib.ttb_axpby_f64.argtypes = [c_int64, c_double, c_void_p, c_double, c_void_p, c_void_p]


def lib():
    """Load (once) and return the ctypes handle; raises if the .so is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C tensor_networks_b200/csrc` (there is no CPU fallback)"
            )
        handle = ctypes.CDLL(LIB_PATH)
        _declare(handle)
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != 0:
        msg = lib().ttb_last_error()
        raise TTBError(status, msg.decode() if msg else "unknown error")


class TTDescriptor:
    """Owns the host arrays behind a `ttb_tt` for the duration of a call."""

    def __init__(self, shape, ranks, core_ptrs):
        d = len(shape)
        assert len(ranks) == d + 1 and len(core_ptrs) == d
        self._n = (c_int64 * d)(*[int(x) for x in shape])
        self._r = (c_int64 * (d + 1))(*[int(x) for x in ranks])
        self._c = (c_void_p * d)(*[int(p) for p in core_ptrs])
        self.struct = ttb_tt(
            d,
            ctypes.cast(self._n, POINTER(c_int64)),
            ctypes.cast(self._r, POINTER(c_int64)),
            ctypes.cast(self._c, POINTER(c_void_p)),
        )

    def ref(self):
        return ctypes.byref(self.struct)


class TTBatchDescriptor:
    """Owns the host arrays behind a `ttb_tt_batch` for the duration of a call."""

    def __init__(self, batch, shape, ranks, core_ptrs):
        d = len(shape)
        assert len(ranks) == d + 1 and len(core_ptrs) == d
        self._n = (c_int64 * d)(*[int(x) for x in shape])
        self._r = (c_int64 * (d + 1))(*[int(x) for x in ranks])
        self._c = (c_void_p * d)(*[int(p) for p in core_ptrs])
        self.struct = ttb_tt_batch(
            d,
            int(batch),
            ctypes.cast(self._n, POINTER(c_int64)),
            ctypes.cast(self._r, POINTER(c_int64)),
            ctypes.cast(self._c, POINTER(c_void_p)),
        )

    def ref(self):
        return ctypes.byref(self.struct)

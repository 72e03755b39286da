"""ctypes binding of the C ABI in include/slsb200.h — the same symbols the Julia shim `ccall`s
(INTEGRATION.md).  The only library the package ever loads by itself is the CUDA build
`lib/libslsb200.so`; if it is missing and cannot be built, importing a compute entry point raises.
There is no CPU fallback.  (tests/emu builds a debugging copy of the kernels on a SIMT emulator and
binds it by passing an explicit path to `CApi`; nothing in the package does that.)"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libslsb200.so")

ABI_VERSION = 9
STATUS_OK, STATUS_RESIDUAL, STATUS_RANK_SKIPPED, STATUS_NO_IMPULSE = 0, 1, 2, 4

i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)
i64pp = C.POINTER(i64p)
f64pp = C.POINTER(f64p)
u8pp = C.POINTER(u8p)


class SlsOptions(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_int32), ("rank_tol", C.c_double), ("feas_tol", C.c_double)]


class SlsTiming(C.Structure):
    _fields_ = [("gather_ms", C.c_float), ("solve_ms", C.c_float), ("scatter_ms", C.c_float),
                ("h2d_ms", C.c_float), ("d2h_ms", C.c_float),
                ("solve_launches", C.c_int64), ("gather_launches", C.c_int64), ("scatter_launches", C.c_int64),
                ("ncols", C.c_int64), ("nnz_out", C.c_int64),
                ("alg_flops", C.c_double), ("alg_bytes", C.c_double),
                ("sim_ms", C.c_float), ("sim_window", C.c_int32), ("sim_uwindow", C.c_int32),
                ("gather_gpu_ms", C.c_float), ("sim_pack_ms", C.c_float), ("refined", C.c_int64), ("sim_bytes", C.c_double),
                ("host_copied_bytes", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# name -> (restype, argtypes); every symbol include/slsb200.h declares
SIGNATURES = {
    "sls_abi_version": (C.c_int, []),
    "sls_device_count": (C.c_int, []),
    "sls_create": (C.c_int, [C.POINTER(SlsOptions), C.POINTER(C.c_void_p)]),
    "sls_destroy": (None, [C.c_void_p]),
    "sls_last_error": (C.c_char_p, [C.c_void_p]),
    "sls_set_plant": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, i64p, i64p, f64p, i64p, i64p, f64p]),
    "sls_set_supports": (C.c_int, [C.c_void_p, C.c_int64, i64pp, i64pp, u8pp, i64pp, i64pp, u8pp]),
    "sls_make_supports": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_double, i64p, i64p]),
    "sls_get_supports": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, i64p, i64p, u8p]),
    "sls_plan": (C.c_int, [C.c_void_p, i64p, C.c_int64, i64p, i64p, i32p, i32p]),
    "sls_reduce": (C.c_int, [C.c_void_p, i64p, C.c_int64, i64p, C.c_int64, i64p, i64p, C.c_int64, i64p, u8p]),
    "sls_reduce_of": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(i64p), C.POINTER(i64p), i64p, C.c_int64, i64p, C.c_int64,
                                C.POINTER(i64p), i64p, i64p, u8p]),
    "sls_local_problem": (C.c_int, [C.c_void_p, C.c_int64, f64p, f64p, u8p, u8p]),
    "sls_solve": (C.c_int, [C.c_void_p, i64pp, i64pp, f64pp, i64pp, i64pp, f64pp, i32p, f64p]),
    "sls_solve_device": (C.c_int, [C.c_void_p]),
    "sls_set_weights": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, f64p, f64p]),
    "sls_set_cost": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, i64p, i64p, f64p, i64p, i64p, f64p, i64p, i64p, f64p, i64p, i64p, f64p]),
    "sls_set_groups": (C.c_int, [C.c_void_p, C.c_int64, i64p]),
    "sls_synthesize": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(i64p), C.POINTER(i64p), C.POINTER(u8p),
                                 C.POINTER(i64p), C.POINTER(i64p), C.POINTER(u8p), C.c_int64, C.c_int64, i64p, i64p,
                                 C.POINTER(i64p), C.POINTER(i64p), C.POINTER(f64p),
                                 C.POINTER(i64p), C.POINTER(i64p), C.POINTER(f64p), i64p, i64p, i32p, f64p]),
    "sls_synthesize_window": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                        C.POINTER(i64p), C.POINTER(i64p), C.POINTER(u8p),
                                        C.POINTER(i64p), C.POINTER(i64p), C.POINTER(u8p), C.c_int64, C.c_int64, i64p, i64p,
                                        C.POINTER(i64p), C.POINTER(i64p), C.POINTER(f64p),
                                        C.POINTER(i64p), C.POINTER(i64p), C.POINTER(f64p), i64p, i64p, i32p, f64p]),
    "sls_host_register": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "sls_host_unregister": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sls_device_output": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_void_p), i64p]),
    "sls_device_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "sls_get_timing": (C.c_int, [C.c_void_p, C.POINTER(SlsTiming)]),
    "sls_simulate": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, i64p, i64p, f64p, f64p, f64p, f64p]),
    "sls_mark": (C.c_int, [C.c_void_p, C.c_int]),
    "sls_elapsed_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "sls_fp64_peak": (C.c_int, [C.c_int, C.c_int, f64p, f64p]),
}


def ptr(a, ty):
    """ctypes pointer of type `ty` to the data of a contiguous numpy array (None -> NULL)."""
    if a is None:
        return C.cast(None, ty)
    assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
    return C.cast(a.ctypes.data, ty)


class SlsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libslsb200 error {code}: {msg}")
        self.code = code


class CApi:
    """A loaded library with typed entry points.  Symbols missing from an older/partial build are
    reported by `missing`; calling one raises AttributeError."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise OSError(f"{path} not found: build it with `python systemlevelcontrol.jl_b200/build.py` "
                          "(needs nvcc; there is no CPU fallback)")
        self.path = path
        self.lib = C.CDLL(path)
        self.missing = []
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(self.lib, name)
            except AttributeError:
                self.missing.append(name)
                continue
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)

    def check(self, handle, rc):
        if rc != 0:
            msg = self.sls_last_error(handle)
            raise SlsError(rc, msg.decode() if msg else "")


_default = None


def default_api() -> CApi:
    """The CUDA library (built on first use when nvcc is present).  Never an emulator, never a CPU path."""
    global _default
    if _default is None:
        if not os.path.exists(LIB_PATH):
            from . import build as _b
            _b.build()
        _default = CApi(LIB_PATH)
    return _default


def ptr_array(arrays, elem_ptr_type):
    """(T pointers) array for a list of numpy arrays; returns (ctypes array, keepalive)."""
    T = len(arrays)
    arr = (elem_ptr_type * T)()
    for t, a in enumerate(arrays):
        arr[t] = ptr(a, elem_ptr_type)
    return arr

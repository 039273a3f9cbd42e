"""ctypes binding of libjalla_b200.so — the C ABI declared in include/jalla_b200.h.

The library is the product: there is no Python or CPU fallback.  Importing this module
without the built .so raises; calling into it without a B200 makes every entry point
report JB_ERR_CUDA.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjalla_b200.so")

JB_ERR_CUDA = -2000
JB_ERR_ARG = -2001
JB_MEM_DEVICE, JB_MEM_MANAGED, JB_MEM_PINNED = 0, 1, 2
ROW_MAJOR, COL_MAJOR = 101, 102
NO_TRANS, TRANS, CONJ_TRANS = 111, 112, 113

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m jalla_b200.build` "
        "(jalla_b200 has no fallback implementation)")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_LOCAL)

_i, _p, _d, _f, _ll, _c = C.c_int, C.c_void_p, C.c_double, C.c_float, C.c_longlong, C.c_char


def _sig(name, res, *args):
    fn = getattr(lib, name)
    fn.restype = res
    fn.argtypes = list(args)
    return fn


_sig("jb_init", _i, _i)
_sig("jb_shutdown", None)
_sig("jb_device_count", _i)
_sig("jb_last_error", _i)
_sig("jb_last_error_string", C.c_char_p)
_sig("jb_version", C.c_char_p)
_sig("jb_malloc", _p, C.c_size_t, _i)
_sig("jb_free", None, _p)
_sig("jb_memcpy_h2d", _i, _p, _p, C.c_size_t)
_sig("jb_memcpy_d2h", _i, _p, _p, C.c_size_t)
_sig("jb_memcpy_d2d", _i, _p, _p, C.c_size_t)
_sig("jb_memcpy_async", _i, _p, _p, C.c_size_t, _p)
_sig("jb_memcpy2d_async", _i, _p, C.c_size_t, _p, C.c_size_t, C.c_size_t, C.c_size_t, _p)
_sig("jb_ipc_export", _i, _p, _p)
_sig("jb_ipc_open", _p, _p)
_sig("jb_ipc_close", _i, _p)
_sig("jb_sync", _i)
_sig("jb_set_stream", None, _p)
_sig("jb_get_stream", _p)
_sig("jb_kernel_launches", C.c_uint64)
_sig("jb_set_option", _i, C.c_char_p, _i)

for _pfx in ("", "jb_"):
    _sig(_pfx + "cblas_sgemm", None, _i, _i, _i, _i, _i, _i, _f, _p, _i, _p, _i, _f, _p, _i)
    _sig(_pfx + "cblas_dgemm", None, _i, _i, _i, _i, _i, _i, _d, _p, _i, _p, _i, _d, _p, _i)
    _sig(_pfx + "cblas_cgemm", None, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _p, _p, _i)
    _sig(_pfx + "cblas_zgemm", None, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _p, _p, _i)
    _sig(_pfx + "cblas_ssyrk", None, _i, _i, _i, _i, _i, _f, _p, _i, _f, _p, _i)
    _sig(_pfx + "cblas_dsyrk", None, _i, _i, _i, _i, _i, _d, _p, _i, _d, _p, _i)
    _sig(_pfx + "cblas_csyrk", None, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p, _i)
    _sig(_pfx + "cblas_zsyrk", None, _i, _i, _i, _i, _i, _p, _p, _i, _p, _p, _i)
    _sig(_pfx + "cblas_cherk", None, _i, _i, _i, _i, _i, _f, _p, _i, _f, _p, _i)
    _sig(_pfx + "cblas_zherk", None, _i, _i, _i, _i, _i, _d, _p, _i, _d, _p, _i)
    _sig(_pfx + "cblas_strsm", None, _i, _i, _i, _i, _i, _i, _i, _f, _p, _i, _p, _i)
    _sig(_pfx + "cblas_dtrsm", None, _i, _i, _i, _i, _i, _i, _i, _d, _p, _i, _p, _i)
    _sig(_pfx + "cblas_ctrsm", None, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i)
    _sig(_pfx + "cblas_ztrsm", None, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i)
    _sz = C.c_size_t
    for _t, _ct in (("s", _f), ("d", _d)):
        _sig(f"{_pfx}cblas_{_t}trmm", None, _i, _i, _i, _i, _i, _i, _i, _ct, _p, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_t}symm", None, _i, _i, _i, _i, _i, _ct, _p, _i, _p, _i, _ct, _p, _i)
        _sig(f"{_pfx}cblas_{_t}copy", None, _i, _p, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_t}swap", None, _i, _p, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_t}scal", None, _i, _ct, _p, _i)
        _sig(f"{_pfx}cblas_{_t}axpy", None, _i, _ct, _p, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_t}dot", _ct, _i, _p, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_t}nrm2", _ct, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_t}asum", _ct, _i, _p, _i)
        _sig(f"{_pfx}cblas_i{_t}amax", _sz, _i, _p, _i)
    _sig(f"{_pfx}cblas_dsdot", _d, _i, _p, _i, _p, _i)
    for _t, _ct, _rt in (("c", _f, "s"), ("z", _d, "d")):
        _sig(f"{_pfx}cblas_{_t}trmm", None, _i, _i, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_t}symm", None, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _p, _p, _i)
        _sig(f"{_pfx}cblas_{_t}hemm", None, _i, _i, _i, _i, _i, _p, _p, _i, _p, _i, _p, _p, _i)
        _sig(f"{_pfx}cblas_{_t}copy", None, _i, _p, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_t}swap", None, _i, _p, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_t}scal", None, _i, _p, _p, _i)
        _sig(f"{_pfx}cblas_{_t}{_rt}scal", None, _i, _ct, _p, _i)
        _sig(f"{_pfx}cblas_{_t}axpy", None, _i, _p, _p, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_t}dotu_sub", None, _i, _p, _i, _p, _i, _p)
        _sig(f"{_pfx}cblas_{_t}dotc_sub", None, _i, _p, _i, _p, _i, _p)
        _sig(f"{_pfx}cblas_{_rt}{_t}nrm2", _ct, _i, _p, _i)
        _sig(f"{_pfx}cblas_{_rt}{_t}asum", _ct, _i, _p, _i)
        _sig(f"{_pfx}cblas_i{_t}amax", _sz, _i, _p, _i)
    for _t in "sdcz":
        _sig(f"{_pfx}LAPACKE_{_t}potri", _i, _i, _c, _i, _p, _i)
        _sig(f"{_pfx}LAPACKE_{_t}gesv", _i, _i, _i, _i, _p, _i, _p, _p, _i)
        _sig(f"{_pfx}LAPACKE_{_t}getrf", _i, _i, _i, _i, _p, _i, _p)
        _sig(f"{_pfx}LAPACKE_{_t}getrs", _i, _i, _c, _i, _i, _p, _i, _p, _p, _i)
        _sig(f"{_pfx}LAPACKE_{_t}getri", _i, _i, _i, _p, _i, _p)
        _sig(f"{_pfx}LAPACKE_{_t}potrf", _i, _i, _c, _i, _p, _i)
        _sig(f"{_pfx}LAPACKE_{_t}potrs", _i, _i, _c, _i, _i, _p, _i, _p, _i)
        _sig(f"{_pfx}LAPACKE_{_t}posv", _i, _i, _c, _i, _i, _p, _i, _p, _i)

for _t in "sdcz":
    _sig(f"jb_{_t}getrf_batched", _i, _i, _p, _i, _ll, _p, _p, _i)
    _sig(f"jb_{_t}getrs_batched", _i, _i, _i, _p, _i, _ll, _p, _p, _i, _ll, _i)
    _sig(f"jb_{_t}gesv_batched", _i, _i, _i, _p, _i, _ll, _p, _p, _i, _ll, _p, _i)
    _sig(f"jb_{_t}potrf_batched", _i, _c, _i, _p, _i, _ll, _p, _i)
    _sig(f"jb_{_t}potrs_batched", _i, _c, _i, _i, _p, _i, _ll, _p, _i, _ll, _i)
    _sig(f"jb_{_t}lacpy", _i, _i, _i, _i, _p, _i, _p, _i)
_sig("jb_dist_create", _i, C.POINTER(_p), _i, _i, _i)
_sig("jb_dist_destroy", _i, _p)
_sig("jb_dist_ngpus", _i, _p)
_sig("jb_dist_last_ms", _d, _p, _i)
_sig("jb_dist_dgemm", _i, _p, _i, _i, _i, _i, _i, _d, _p, _i, _p, _i, _d, _p, _i)
_sig("jb_dist_dgetrf", _i, _p, _i, _p, _i, _p)
_sig("jb_dist_dgetrs", _i, _p, _c, _i, _i, _p, _i, _p, _p, _i)
_sig("jb_dist_dpotrf", _i, _p, _c, _i, _p, _i)
_sig("jb_dist_dpotrs", _i, _p, _c, _i, _i, _p, _i, _p, _i)
_sig("jb_dist_dgesv_batched", _i, _p, _i, _i, _p, _i, _ll, _p, _p, _i, _ll, _p, _i)
_sig("jb_slaset", _i, _i, _i, _f, _f, _p, _i)
_sig("jb_dlaset", _i, _i, _i, _d, _d, _p, _i)
_sig("jb_claset", _i, _i, _i, _p, _p, _p, _i)
_sig("jb_zlaset", _i, _i, _i, _p, _p, _p, _i)
for _t in "sdcz":
    _sig(f"jb_{_t}set_diag", _i, _i, _i, _i, _p, _i, _p, _i)
    _sig(f"jb_{_t}mult_diag", _i, _i, _i, _i, _p, _i, _p, _p, _i)
_sig("jb_dsolve_batched", _i, _i, _i, _p, _i, _ll, _p, _i, _ll, _p, _i, _ll, _p, _i)
_sig("jb_dfill_uniform", _i, _p, _i, _i, _i, _ll, _ll, C.c_uint64, _i, _d)
_sig("jb_sfill_uniform", _i, _p, _i, _i, _i, _ll, _ll, C.c_uint64, _i, _f)
_sig("jb_zfill_uniform", _i, _p, _i, _i, _i, _ll, _ll, C.c_uint64)
_sig("jb_cfill_uniform", _i, _p, _i, _i, _i, _ll, _ll, C.c_uint64)


class JallaB200Error(RuntimeError):
    pass


def check(rc: int, what: str = "") -> int:
    """Raise on shim-level failures (JB_ERR_*); LAPACK info codes pass through."""
    if rc in (JB_ERR_CUDA, JB_ERR_ARG):
        raise JallaB200Error(f"{what}: {lib.jb_last_error_string().decode()} (code {rc})")
    return rc


def check_last(what: str = "") -> None:
    """cblas entry points return void: surface the recorded error, if any."""
    e = lib.jb_last_error()
    if e:
        raise JallaB200Error(f"{what}: {lib.jb_last_error_string().decode()} (code {e})")

"""ctypes binding of include/halab200.h (libhalab200.so). No CPU fallback: a missing library is an ImportError."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhalab200.so")

HB_F32, HB_F64, HB_C32, HB_C64 = 0, 1, 2, 3
HB_OK = 0
HB_POINTER_HOST, HB_POINTER_DEVICE = 0, 1
HB_H2D, HB_D2H, HB_D2D = 0, 1, 2
HB_TRANS_SCATTER, HB_TRANS_CHECKED, HB_TRANS_FROZEN = 0, 1, 2

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(or `make -C hala_b200/csrc`). hala_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

_vp, _i, _sz, _d = C.c_void_p, C.c_int, C.c_size_t, C.c_double
_pi, _pvp = C.POINTER(C.c_int), C.POINTER(C.c_void_p)

# every symbol include/halab200.h declares, with its argument types (tests check this table against the header)
SIGNATURES = {
    "hb_version": (C.c_char_p, []),
    "hb_last_error": (C.c_char_p, []),
    "hb_device_count": (_i, [_pi]),
    "hb_ctx_create": (_i, [_i, _pvp]),
    "hb_ctx_destroy": (_i, [_vp]),
    "hb_ctx_device": (_i, [_vp, _pi]),
    "hb_ctx_set_stream": (_i, [_vp, _vp]),
    "hb_ctx_get_stream": (_i, [_vp, _pvp]),
    "hb_ctx_sync": (_i, [_vp]),
    "hb_ctx_set_pointer_mode": (_i, [_vp, _i]),
    "hb_ctx_get_pointer_mode": (_i, [_vp, _pi]),
    "hb_ctx_launch_count": (_i, [_vp, C.POINTER(C.c_longlong)]),
    "hb_ctx_trim": (_i, [_vp]),
    "hb_ctx_profile": (_i, [_vp, _i]),
    "hb_ctx_profile_read": (_i, [_vp, _i, C.POINTER(_d), C.POINTER(C.c_longlong)]),
    "hb_timer_start": (_i, [_vp]),
    "hb_timer_stop": (_i, [_vp, C.POINTER(C.c_float)]),
    "hb_malloc": (_i, [_vp, _sz, _pvp]),
    "hb_free": (_i, [_vp, _vp]),
    "hb_memcpy": (_i, [_vp, _vp, _vp, _sz, _i]),
    "hb_memcpy_async": (_i, [_vp, _vp, _vp, _sz, _i]),
    "hb_memset_zero": (_i, [_vp, _vp, _sz]),
    "hb_fill": (_i, [_vp, _i, _sz, _vp, _vp]),
    "hb_dev_malloc": (_i, [_i, _sz, _pvp]),
    "hb_dev_free": (_i, [_vp]),
    "hb_dev_memcpy": (_i, [_vp, _vp, _sz, _i]),
    "hb_dev_fill": (_i, [_i, _i, _sz, _vp, _vp]),
    "hb_host_alloc": (_i, [_sz, _pvp]),
    "hb_host_free": (_i, [_vp]),
    "hb_csr_create": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _pvp]),
    "hb_csr_destroy": (_i, [_vp]),
    "hb_csr_info": (_i, [_vp, _pi, _pi, _pi, _pi, _pi]),
    "hb_spmv_buffer_size": (_i, [_vp, C.c_char, C.POINTER(_sz)]),
    "hb_spmv": (_i, [_vp, _vp, C.c_char, _vp, _vp, _vp, _vp]),
    "hb_spmv_dot": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "hb_csr_set_variant": (_i, [_vp, _i]),
    "hb_csr_set_transpose_mode": (_i, [_vp, _i]),
    "hb_csr_values_changed": (_i, [_vp]),
    "hb_csr_transpose_info": (_i, [_vp, _pi, _pi, C.POINTER(_sz)]),
    "hb_tri_create": (_i, [_vp, _i, C.c_char, C.c_char, _i, _i, _vp, _vp, _vp, _pvp]),
    "hb_tri_destroy": (_i, [_vp]),
    "hb_tri_info": (_i, [_vp, _pi, _pi, _pi]),
    "hb_sptrsv": (_i, [_vp, _vp, C.c_char, _vp, _vp, _i, _vp, _i]),
    "hb_sptrsm": (_i, [_vp, _vp, C.c_char, C.c_char, _i, _vp, _vp, _i]),
    "hb_ilu0": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "hb_spmm": (_i, [_vp, _vp, C.c_char, C.c_char, _i, _i, _vp, _vp, _i, _vp, _vp, _i]),
    "hb_geam": (_i, [_vp, _i, C.c_char, C.c_char, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _i]),
    "hb_dgmm": (_i, [_vp, _i, C.c_char, _i, _i, _vp, _i, _vp, _i, _vp, _i]),
    "hb_tbsv": (_i, [_vp, _i, C.c_char, C.c_char, C.c_char, _i, _i, _vp, _i, _vp, _i]),
    "hb_copy": (_i, [_vp, _i, _i, _vp, _i, _vp, _i]),
    "hb_axpy": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _i]),
    "hb_scal": (_i, [_vp, _i, _i, _vp, _vp, _i]),
    "hb_dot": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp]),
    "hb_nrm2": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "hb_asum": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "hb_swap": (_i, [_vp, _i, _i, _vp, _i, _vp, _i]),
    "hb_iamax": (_i, [_vp, _i, _i, _vp, _i, _pi]),
    "hb_rot": (_i, [_vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _i]),
    "hb_rotm": (_i, [_vp, _i, _i, _vp, _i, _vp, _i, _vp]),
    "hb_rotg": (_i, [_vp, _i, _vp, _vp, _vp, _vp]),
    "hb_rotmg": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "hb_gemv": (_i, [_vp, _i, C.c_char, _i, _i, _vp, _vp, _i, _vp, _i, _vp, _vp, _i]),
    "hb_multi_dot": (_i, [_vp, _i, _i, _i, _i, _vp, _sz, _vp, _vp]),
    "hb_multi_axpy_nrm2": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp, _vp, _vp]),
    "hb_axpy2_nrm2": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hb_xpby": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "hb_cg": (_i, [_vp, _vp, _vp, _vp, _d, _i, _pi, C.POINTER(_d)]),
    "hb_gmres": (_i, [_vp, _vp, _vp, _vp, _d, _i, _i, _i, _pi, C.POINTER(_d)]),
    "hb_pcg": (_i, [_vp, _vp, _vp, _vp, _d, _i, _vp, _vp, _pi, C.POINTER(_d)]),
    "hb_pgmres": (_i, [_vp, _vp, _vp, _vp, _d, _i, _i, _i, _vp, _vp, _pi, C.POINTER(_d)]),
}
# hb_precon_fn: int (*)(void *user, const void *in_dev, void *out_dev); pass C.cast(PRECON_FN(py_callable), C.c_void_p) (keep the object alive)
PRECON_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p)

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)      # AttributeError here == header/library mismatch
    _f.restype = _res
    _f.argtypes = _args


class HalaB200Error(RuntimeError):
    """Non-zero status from libhalab200 (the C++ header layer throws std::runtime_error at the same places)."""


def check(status, what=""):
    if status != HB_OK:
        raise HalaB200Error(f"{what} failed (status {status}): {lib.hb_last_error().decode()}")

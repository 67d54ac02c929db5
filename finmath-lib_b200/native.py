"""ctypes binding of libfinmath_b200.so (the C ABI of include/finmath_b200.h).

This is the Python twin of the JNI shim in INTEGRATION.md: same symbols, same argument meaning.  There is no CPU
fallback — a missing library or a missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfinmath_b200.so")

try:
    from . import _fmbfast as _F          # C accelerator of this binding (csrc/host/fmbfast.c): storage types + fast paths of the hot operations
except ImportError as e:                  # pragma: no cover
    raise ImportError("finmath_b200: the host accelerator _fmbfast is not built: run `python -c 'import __graft_entry__ as g; g.build()'` (%s)" % e)

FMB_OK, FMB_EINVAL, FMB_ENODEVICE, FMB_ENOMEM, FMB_ECUDA, FMB_EHANDLE, FMB_EUNSUPPORTED = range(7)

# op codes (include/finmath_b200.h)
U_SQUARED, U_SQRT, U_EXP, U_LOG, U_SIN, U_COS, U_INVERT, U_ABS, U_ISNAN, U_EXPM1 = range(10)
U_ADD, U_SUB, U_BUS, U_MULT, U_DIV, U_VID, U_CAP, U_FLOOR, U_POW, U_ICDF_NORMAL = range(10, 20)
B_ADD, B_SUB, B_MULT, B_DIV, B_CAP, B_FLOOR = range(6)
T_ADD_PRODUCT, T_ADD_PRODUCT_D, T_ADD_RATIO, T_SUB_RATIO, T_ACCRUE, T_DISCOUNT, T_CHOOSE = range(7)
R_SUM, R_SUM_PRODUCT, R_CENTERED_M2, R_CENTERED_M2_W, R_MIN, R_MAX = range(6)

c_dp = C.POINTER(C.c_double)
c_hp = C.POINTER(C.c_uint64)
c_ip = C.POINTER(C.c_int32)
c_u32p = C.POINTER(C.c_uint32)


class FmbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("finmath_b200 error %d: %s" % (code, msg))
        self.code = code


class NoDeviceError(FmbError):
    pass


_lib = None

# every exported symbol of the header, with its prototype (tests check the .so exports all of them)
PROTOTYPES = {
    "fmb_init": [C.c_int],
    "fmb_shutdown": [],
    "fmb_is_initialized": [],
    "fmb_last_error": [],
    "fmb_device_count": [C.POINTER(C.c_int)],
    "fmb_device_name": [C.c_char_p, C.c_int],
    "fmb_synchronize": [],
    "fmb_set_fp_mode": [C.c_int],
    "fmb_get_fp_mode": [C.POINTER(C.c_int)],
    "fmb_timer_start": [],
    "fmb_timer_stop_ms": [C.POINTER(C.c_float)],
    "fmb_kernel_launch_count": [c_hp],
    "fmb_rv_create": [C.c_uint64, c_hp],
    "fmb_rv_upload": [c_dp, C.c_uint64, c_hp],
    "fmb_rv_fill": [C.c_double, C.c_uint64, c_hp],
    "fmb_rv_download": [C.c_uint64, c_dp, C.c_uint64],
    "fmb_rv_get": [C.c_uint64, C.c_uint64, c_dp],
    "fmb_rv_size": [C.c_uint64, c_hp],
    "fmb_rv_retain": [C.c_uint64],
    "fmb_rv_free": [C.c_uint64],
    "fmb_rv_free_many": [c_hp, C.c_uint64],
    "fmb_rv_device_ptr": [C.c_uint64, C.POINTER(C.c_void_p)],
    "fmb_pool_stats": [c_hp, c_hp, c_hp],
    "fmb_pool_trim": [],
    "fmb_rv_unary": [C.c_int, C.c_uint64, C.c_double, c_hp],
    "fmb_rv_binary": [C.c_int, C.c_uint64, C.c_double, C.c_uint64, C.c_double, c_hp],
    "fmb_rv_ternary": [C.c_int, C.c_uint64, C.c_double, C.c_uint64, C.c_double, C.c_uint64, C.c_double, C.c_double, c_hp],
    "fmb_rv_accrue_chain": [C.c_int, c_hp, c_dp, C.c_double, c_hp],
    "fmb_rv_accrue_prefix": [C.c_int, C.c_double, c_hp, c_dp, c_hp],
    "fmb_rv_eval_chain": [C.c_int, C.c_char_p, C.c_int, c_hp, C.c_int, c_dp, C.c_int, c_hp],
    "fmb_rv_reduce": [C.c_int, C.c_uint64, C.c_uint64, C.c_double, c_dp],
    "fmb_rv_reduce_many": [C.c_int, C.c_int, c_hp, C.c_double, c_dp],
    "fmb_rv_select": [C.c_uint64, C.c_uint64, c_dp],
    "fmb_rv_count_le": [C.c_uint64, c_dp, C.c_int, c_hp],
    "fmb_rv_range_sum": [C.c_uint64, C.c_double, C.c_double, c_dp],
    "fmb_mt_words": [C.c_int64, C.c_uint64, C.c_uint64, c_u32p],
    "fmb_mt_uniforms": [C.c_int64, C.c_uint64, C.c_uint64, c_dp],
    "fmb_icdf": [c_dp, C.c_uint64, c_dp],
    "fmb_bm_generate": [C.c_int32, C.c_int, C.c_int, C.c_uint64, C.c_uint64, c_dp, c_hp],
    "fmb_uniforms_generate": [C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, c_hp],
    "fmb_euler_black_scholes": [C.c_int, C.c_int, C.c_int, C.c_uint64, c_dp, c_hp, C.c_double, C.c_double, C.c_double, c_hp],
    "fmb_euler_heston": [C.c_int, C.c_int, C.c_int, C.c_uint64, c_dp, c_hp, C.c_double, c_dp, C.c_double, C.c_double, C.c_double,
                         C.c_double, C.c_double, c_hp],
    "fmb_euler_lmm": [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_uint64, c_dp, c_hp, c_dp, c_dp, c_dp, c_dp,
                      c_ip, c_hp],
    "fmb_euler_hull_white": [C.c_int, C.c_uint64, c_dp, c_hp, c_dp, c_dp, c_dp, c_hp],
    "fmb_regression_moments": [C.c_int, c_hp, c_dp, C.c_uint64, c_dp, c_dp, c_dp, c_dp],
    "fmb_regression_solve_svd": [C.c_int, c_dp, c_dp, c_dp, c_dp],
    "fmb_regression_predict": [C.c_int, c_hp, c_dp, c_dp, c_hp],
    "fmb_regression_fit": [C.c_int, c_hp, c_dp, C.c_uint64, C.c_uint64, C.c_uint64, c_hp],
    "fmb_regression_fit_get": [C.c_uint64, C.c_int, c_dp, c_dp, c_dp, c_dp],
    "fmb_regression_predict_fit": [C.c_int, c_hp, c_dp, C.c_uint64, c_hp],
    "fmb_regression_conditional_expectation": [C.c_int, c_hp, c_dp, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, c_hp, c_dp, c_hp, c_hp],
    "fmb_comm_unique_id": [C.c_char_p, C.c_int],
    "fmb_comm_init": [C.c_char_p, C.c_int, C.c_int, C.c_int],
    "fmb_comm_shutdown": [],
    "fmb_comm_info": [C.POINTER(C.c_int), C.POINTER(C.c_int), c_hp],
    "fmb_comm_peer_handle": [C.c_char_p, C.c_int],
    "fmb_comm_peer_open": [C.c_char_p, C.c_int],
    "fmb_bench_dfma_tflops": [c_dp],
    "fmb_bench_copy_gbs": [C.c_uint64, c_dp],
}


def load():
    """Load the shared library (no device needed for this step)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FmbError(FMB_ENODEVICE, "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                          "(there is no CPU fallback)" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_char_p if name == "fmb_last_error" else C.c_int
        _lib = lib
        _bind_fast()
    return _lib


def _bind_fast():
    """Hand the accelerator the addresses of the four entry points it calls (the very library ctypes loaded) and the classes it creates."""
    if _lib is None:
        return
    try:
        from .stochastic import RandomVariableCuda
    except ImportError:                   # stochastic.py is still being imported: it calls _bind_fast() again when it is done
        return
    addr = lambda name: C.cast(getattr(_lib, name), C.c_void_p).value
    _F.bind(addr("fmb_rv_unary"), addr("fmb_rv_binary"), addr("fmb_rv_ternary"), addr("fmb_rv_free"), addr("fmb_rv_eval_chain"), check,
            RandomVariableCuda, DeviceVector, LazyVector)
    _F.set_lazy_min_n(_lazy_min_n if _lazy else 2 ** 64 - 1)


def check(rc):
    if rc != FMB_OK:
        msg = load().fmb_last_error().decode("utf-8", "replace")
        if rc == FMB_ENODEVICE:
            raise NoDeviceError(rc, msg)
        if rc == FMB_EINVAL:
            raise ValueError("finmath_b200: " + msg)            # IllegalArgumentException
        if rc == FMB_EUNSUPPORTED:
            raise NotImplementedError("finmath_b200: " + msg)   # UnsupportedOperationException
        if rc == FMB_ENOMEM:
            raise MemoryError("finmath_b200: " + msg)
        raise FmbError(rc, msg)


def init(device=None):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("FMB_DEVICE", "0")))
    check(load().fmb_init(device))


def dptr(a):
    return a.ctypes.data_as(c_dp)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def handles(hs):
    return np.ascontiguousarray(hs, dtype=np.uint64)


def hptr(a):
    return a.ctypes.data_as(c_hp)


class DeviceVector(_F.DV):
    """Owner of one native handle (h, n live in the C base type); released when the object is collected (the Java side uses a Cleaner)."""
    __slots__ = ()

    @staticmethod
    def upload(values):
        a = as_f64(values)
        out = C.c_uint64()
        check(load().fmb_rv_upload(dptr(a), a.size, C.byref(out)))
        return DeviceVector(out.value, a.size)

    def download(self):
        out = np.empty(self.n, dtype=np.float64)
        check(load().fmb_rv_download(self.h, dptr(out), self.n))
        return out

    def get(self, i):
        v = C.c_double()
        check(load().fmb_rv_get(self.h, int(i), C.byref(v)))
        return v.value


# ---- deferred element-wise arithmetic --------------------------------------------------------------------------------------
# unary / binary / ternary return a LazyVector: the operation is recorded, not launched.  When the result is consumed by the next
# element-wise operation the chain grows; when anything else needs it (a reduction, a kernel argument, a download: every access to
# `.h`), the whole chain is evaluated in ONE pass by fmb_rv_eval_chain.  Same device functions, order and roundings as the one-op
# kernels: results are bit-identical to eager evaluation (tests/test_gpu_rv.py compares both modes).  FMB_LAZY=0 or
# set_lazy(False) switches the deferral off.  The recording itself is C code (csrc/host/fmbfast.c: ~0.3 us per operation), so it pays at
# every vector length: measured with deferral for all sizes against deferral from 2 M elements (round 1, when the recording was Python):
# Bermudan 1 M paths 11.86 -> 11.57 ms (567 -> 313 launches), calibration evaluation (100 k paths, 16 swaptions) 7.4 -> 4.7 ms.
CHAIN_MAX_INSTR, CHAIN_MAX_LEAVES, CHAIN_MAX_SCALARS = 16, 8, 24
_lazy = os.environ.get("FMB_LAZY", "1") != "0"
try:
    _lazy_min_n = int(os.environ.get("FMB_LAZY_MIN_N", "0"))
except ValueError:
    _lazy_min_n = 0


def set_lazy(on, min_n=None):
    """Switch deferred evaluation on / off; min_n = shortest vector that is deferred (None keeps the current threshold)."""
    global _lazy, _lazy_min_n
    _lazy = bool(on)
    if min_n is not None:
        _lazy_min_n = int(min_n)
    _F.set_lazy_min_n(_lazy_min_n if _lazy else 2 ** 64 - 1)


def lazy_enabled():
    return _lazy


def lazy_min_n():
    return _lazy_min_n


class LazyVector(_F.LV):
    """Result of element-wise operations that has not been evaluated yet (the recorded chain and its evaluation live in the C base type,
    csrc/host/fmbfast.c).  Duck-types DeviceVector: h (evaluates on first access), n, download, get."""
    __slots__ = ()

    def download(self):
        out = np.empty(self.n, dtype=np.float64)
        check(load().fmb_rv_download(self.h, dptr(out), self.n))
        return out

    def get(self, i):
        v = C.c_double()
        check(load().fmb_rv_get(self.h, int(i), C.byref(v)))
        return v.value


def _lazy_op(kind, op, operands, a):
    """operands: positional tuple of DeviceVector / LazyVector / float (scalar broadcast).  Returns a LazyVector."""
    return _F.lazy_op(kind, op, operands, a)


def unary(op, x, a=0.0):
    if _lazy and x.n >= _lazy_min_n:
        return _lazy_op(0, op, (x,), float(a))
    out = C.c_uint64()
    check(load().fmb_rv_unary(op, x.h, float(a), C.byref(out)))
    return DeviceVector(out.value, x.n)


def binary(op, x, sx, y, sy):
    if _lazy and (x if x is not None else y).n >= _lazy_min_n:
        return _lazy_op(1, op, (x if x is not None else float(sx), y if y is not None else float(sy)), 0.0)
    out = C.c_uint64()
    check(load().fmb_rv_binary(op, x.h if x is not None else 0, float(sx), y.h if y is not None else 0, float(sy), C.byref(out)))
    n = x.n if x is not None else y.n
    return DeviceVector(out.value, n)


def ternary(op, x, sx, y, sy, z, sz, a=0.0):
    if _lazy and next(v for v in (x, y, z) if v is not None).n >= _lazy_min_n:
        return _lazy_op(2, op, (x if x is not None else float(sx), y if y is not None else float(sy), z if z is not None else float(sz)), float(a))
    out = C.c_uint64()
    check(load().fmb_rv_ternary(op, x.h if x is not None else 0, float(sx), y.h if y is not None else 0, float(sy),
                                z.h if z is not None else 0, float(sz), float(a), C.byref(out)))
    n = next(v.n for v in (x, y, z) if v is not None)
    return DeviceVector(out.value, n)


def reduce(op, x, w=None, a=0.0):
    out = (C.c_double * 2)()
    check(load().fmb_rv_reduce(op, x.h, w.h if w is not None else 0, float(a), out))
    return out[0], out[1]


RM_SUM, RM_SUM_INVERT_MULT = 0, 1


def reduce_many(op, vectors, a=0.0):
    """(hi, lo) arrays of sum_p f(x_i[p]) for a list of DeviceVector / LazyVector of one length: one launch, one synchronisation."""
    hs = np.array([v.h for v in vectors], dtype=np.uint64)
    out = np.zeros(2 * len(vectors))
    check(load().fmb_rv_reduce_many(op, len(vectors), hptr(hs), float(a), dptr(out)))
    return out[0::2].copy(), out[1::2].copy()


def launch_count():
    c = C.c_uint64()
    check(load().fmb_kernel_launch_count(C.byref(c)))
    return c.value


def synchronize():
    check(load().fmb_synchronize())


def timer_start():
    check(load().fmb_timer_start())


def timer_stop_ms():
    ms = C.c_float()
    check(load().fmb_timer_stop_ms(C.byref(ms)))
    return ms.value


def set_fp_mode(mode):
    """0 = STRICT (default: reference operation order, no FMA contraction), 1 = FAST (FMA contraction; functional schemes carry the
    log-state instead of re-deriving it as log(exp(y)) each step) — both within 1e-12 of the reference's paths."""
    check(load().fmb_set_fp_mode(int(mode)))

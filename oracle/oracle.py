"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end of oracle/liboracle.so (C++ restatement of the reference's CPU path, see orc_core.h).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product path (finmath-lib_b200/) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)
c_u32p = C.POINTER(C.c_uint32)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("orc_capi.cpp", "orc_core.h", "orc_models.h", "orc_products.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_rv_reduce.restype = C.c_double
        L.orc_bs_european.restype = C.c_double
        L.orc_heston_european.restype = C.c_double
        L.orc_lmm_create.restype = C.c_void_p
        L.orc_lmm_swaption.restype = C.c_double
        L.orc_lmm_caplet.restype = C.c_double
        L.orc_lmm_bermudan.restype = C.c_double
        L.orc_time_lmm_reference_shaped.restype = C.c_double
        L.orc_time_lmm_fused.restype = C.c_double
        L.orc_hull_white_caplet.restype = C.c_double
        L.orc_bs_bermudan_option.restype = C.c_double
    return _LIB


def set_math(mode):
    """0: libm exp / log (default).  1 (diagnostic): the device kernels' exp / log compiled for the host, to attribute deviations."""
    lib().orc_set_math(C.c_int(mode))


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_dp)


def _dn(a):
    if a is None:
        return None, None
    return _d(a)


def mt_words(seed, offset, n):
    out = np.empty(n, dtype=np.uint32)
    lib().orc_mt_words(C.c_int64(seed), C.c_uint64(offset), C.c_uint64(n), out.ctypes.data_as(c_u32p))
    return out


def mt_words_key(key, n):
    key = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.empty(n, dtype=np.uint32)
    lib().orc_mt_words_key(key.ctypes.data_as(c_u32p), C.c_int(len(key)), C.c_uint64(n), out.ctypes.data_as(c_u32p))
    return out


def mt_raw_sequence(seed, n):
    out = np.empty(n, dtype=np.uint32)
    lib().orc_mt_raw_sequence(C.c_int64(seed), C.c_uint64(n), out.ctypes.data_as(c_u32p))
    return out


def mt_uniforms(seed, offset, n):
    out = np.empty(n, dtype=np.float64)
    lib().orc_mt_uniforms(C.c_int64(seed), C.c_uint64(offset), C.c_uint64(n), out.ctypes.data_as(c_dp))
    return out


def icdf(p):
    p, pp = _d(p)
    out = np.empty_like(p)
    lib().orc_icdf(pp, C.c_uint64(p.size), out.ctypes.data_as(c_dp))
    return out


def time_discretization(initial, n_steps, dt):
    out = np.empty(n_steps + 1, dtype=np.float64)
    n = lib().orc_time_discretization(C.c_double(initial), C.c_int(n_steps), C.c_double(dt), out.ctypes.data_as(c_dp))
    return out[:n].copy()


def time_discretization_from_array(times):
    t, tp = _d(times)
    out = np.empty(t.size, dtype=np.float64)
    n = lib().orc_time_discretization_from_array(tp, C.c_int(t.size), out.ctypes.data_as(c_dp))
    return out[:n].copy()


def time_index(times, t):
    a, ap = _d(times)
    return lib().orc_time_index(ap, C.c_int(a.size), C.c_double(t))


def brownian(seed, times, F, paths, path_offset=0):
    t, tp = _d(times)
    out = np.empty((t.size - 1, F, paths), dtype=np.float64)
    lib().orc_brownian(C.c_int(seed), tp, C.c_int(t.size), C.c_int(F), C.c_int(paths), C.c_int64(path_offset), out.ctypes.data_as(c_dp))
    return out


def rv_unary(op, x, a=0.0):
    x, xp = _d(x)
    out = np.empty_like(x)
    assert lib().orc_rv_unary(C.c_int(op), xp, C.c_uint64(x.size), C.c_double(a), out.ctypes.data_as(c_dp)) == 0
    return out


def rv_binary(op, x, y):
    x, xp = _d(x)
    y, yp = _d(y)
    out = np.empty_like(x)
    assert lib().orc_rv_binary(C.c_int(op), xp, yp, C.c_uint64(x.size), out.ctypes.data_as(c_dp)) == 0
    return out


def rv_ternary(op, x, y, z=None, a=0.0):
    x, xp = _d(x)
    y, yp = _d(y)
    z, zp = _dn(z)
    out = np.empty_like(x)
    assert lib().orc_rv_ternary(C.c_int(op), xp, yp, zp, C.c_uint64(x.size), C.c_double(a), out.ctypes.data_as(c_dp)) == 0
    return out


def rv_reduce(op, x, w=None, a=0.0, b=0.0):
    x, xp = _d(x)
    w, wp = _dn(w)
    return lib().orc_rv_reduce(C.c_int(op), xp, wp, C.c_uint64(x.size), C.c_double(a), C.c_double(b))


def rv_histogram(x, pts):
    x, xp = _d(x)
    pts, pp = _d(pts)
    out = np.empty(pts.size + 1)
    lib().orc_rv_histogram(xp, C.c_uint64(x.size), pp, C.c_int(pts.size), out.ctypes.data_as(c_dp))
    return out


def solve_pinv(A, b):
    A, Ap = _d(A)
    b, bp = _d(b)
    K = b.size
    x = np.empty(K)
    cond = C.c_double()
    lib().orc_solve_pinv(Ap, bp, C.c_int(K), x.ctypes.data_as(c_dp), C.byref(cond))
    return x, cond.value


def bs_european(seed, times, paths, s0, r, sigma, scheme, maturity, strike, call_put=1, path_offset=0, want_process=True):
    t, tp = _d(times)
    proc = np.empty((t.size, 1, paths)) if want_process else None
    vals = np.empty(paths)
    price = lib().orc_bs_european(C.c_int(seed), tp, C.c_int(t.size), C.c_int(paths), C.c_int64(path_offset), C.c_double(s0),
                                  C.c_double(r), C.c_double(sigma), C.c_int(scheme), C.c_double(maturity), C.c_double(strike),
                                  C.c_int(call_put), proc.ctypes.data_as(c_dp) if want_process else None, vals.ctypes.data_as(c_dp))
    return price, proc, vals


def heston_european(seed, times, paths, s0, r, sigma, discount_rate, theta, kappa, xi, rho, heston_scheme, scheme, maturity, strike,
                    call_put=1, path_offset=0, want_process=True):
    t, tp = _d(times)
    proc = np.empty((t.size, 2, paths)) if want_process else None
    vals = np.empty(paths)
    price = lib().orc_heston_european(C.c_int(seed), tp, C.c_int(t.size), C.c_int(paths), C.c_int64(path_offset), C.c_double(s0),
                                      C.c_double(r), C.c_double(sigma), C.c_double(discount_rate), C.c_double(theta), C.c_double(kappa),
                                      C.c_double(xi), C.c_double(rho), C.c_int(heston_scheme), C.c_int(scheme), C.c_double(maturity),
                                      C.c_double(strike), C.c_int(call_put), proc.ctypes.data_as(c_dp) if want_process else None,
                                      vals.ctypes.data_as(c_dp))
    return price, proc, vals


class LMM:
    """Reference-shaped LMM simulation + products (oracle side)."""

    def __init__(self, seed, sim_times, tenor_times, F, paths, L0, sigma, factor_matrix, discount_factors=None, measure=0,
                 state_space=1, libor_cap=1e5, scheme=2, path_offset=0):
        self.sim_times, sp = _d(sim_times)
        self.tenor_times, tp = _d(tenor_times)
        self.paths, self.F = paths, F
        self.N = self.tenor_times.size - 1
        self.T = self.sim_times.size - 1
        L0, lp = _d(L0)
        sigma, sgp = _d(sigma)
        fm, fp = _d(factor_matrix)
        df, dfp = _dn(discount_factors)
        assert sigma.size == self.T * self.N and fm.size == self.N * F
        self.h = C.c_void_p(lib().orc_lmm_create(C.c_int(seed), sp, C.c_int(self.sim_times.size), tp, C.c_int(self.tenor_times.size),
                                                 C.c_int(F), C.c_int(paths), C.c_int64(path_offset), lp, dfp, sgp, fp, C.c_int(measure),
                                                 C.c_int(state_space), C.c_double(libor_cap), C.c_int(scheme)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_lmm_free(self.h)
            self.h = None

    def process(self):
        out = np.empty((self.T + 1, self.N, self.paths))
        lib().orc_lmm_process(self.h, out.ctypes.data_as(c_dp))
        return out

    def brownian(self):
        out = np.empty((self.T, self.F, self.paths))
        lib().orc_lmm_brownian(self.h, out.ctypes.data_as(c_dp))
        return out

    def set_interpolation(self, method):
        """0 LINEAR, 1 LOG_LINEAR_UNCORRECTED (default) — interpolation of forward rates on fractional tenor points."""
        lib().orc_lmm_set_interpolation(self.h, C.c_int(method))

    def numeraire(self, time):
        out = np.empty(self.paths)
        lib().orc_lmm_numeraire(self.h, C.c_double(time), out.ctypes.data_as(c_dp))
        return out

    def forward_rate(self, time, start, end):
        out = np.empty(self.paths)
        lib().orc_lmm_forward_rate(self.h, C.c_double(time), C.c_double(start), C.c_double(end), out.ctypes.data_as(c_dp))
        return out

    def swaption(self, exercise_date, fixing_dates, payment_dates, swaprates, notional=1.0, discounting_adjustments=None):
        f, fp = _d(fixing_dates)
        p, pp = _d(payment_dates)
        s, sp = _d(swaprates)
        vals = np.empty(self.paths)
        se = C.c_double()
        if discounting_adjustments is not None:
            a, ap = _d(discounting_adjustments)
            lib().orc_lmm_swaption_adj.restype = C.c_double
            price = lib().orc_lmm_swaption_adj(self.h, C.c_double(exercise_date), fp, pp, sp, C.c_int(f.size), C.c_double(notional), ap,
                                               vals.ctypes.data_as(c_dp), C.byref(se))
            return price, vals, se.value
        price = lib().orc_lmm_swaption(self.h, C.c_double(exercise_date), fp, pp, sp, C.c_int(f.size), C.c_double(notional),
                                       vals.ctypes.data_as(c_dp), C.byref(se))
        return price, vals, se.value

    def caplet(self, maturity, period_length, strike, daycount_fraction=None, is_floorlet=False):
        vals = np.empty(self.paths)
        dcf = period_length if daycount_fraction is None else daycount_fraction
        price = lib().orc_lmm_caplet(self.h, C.c_double(maturity), C.c_double(period_length), C.c_double(strike), C.c_double(dcf),
                                     C.c_int(1 if is_floorlet else 0), vals.ctypes.data_as(c_dp))
        return price, vals

    def bermudan_given(self, is_exercise, fixing_dates, period_lengths, payment_dates, notionals, swaprates, coefficients, weight, is_callable=True):
        """Replay of the backward induction on this (window of) paths with GIVEN regression coefficients [nExercise][6] and Monte-Carlo
        weight: per-path values and exercise times."""
        ex = np.ascontiguousarray(is_exercise, dtype=np.int32)
        f, fp = _d(fixing_dates)
        pl, plp = _d(period_lengths)
        p, pp = _d(payment_dates)
        nt, ntp = _d(notionals)
        s, sp = _d(swaprates)
        co, cop = _d(np.asarray(coefficients, dtype=np.float64).reshape(-1))
        vals, ext = np.empty(self.paths), np.empty(self.paths)
        lib().orc_lmm_bermudan_given(self.h, ex.ctypes.data_as(C.POINTER(C.c_int)), fp, plp, pp, ntp, sp, C.c_int(f.size), C.c_int(1 if is_callable else 0),
                                     cop, C.c_int(co.size // 6), C.c_double(weight), vals.ctypes.data_as(c_dp), ext.ctypes.data_as(c_dp))
        return vals, ext

    def bermudan_basis(self, fixing_date, fixing_dates, payment_dates):
        f, fp = _d(fixing_dates)
        p, pp = _d(payment_dates)
        out = np.empty((6, self.paths))
        lib().orc_lmm_bermudan_basis(self.h, C.c_double(fixing_date), fp, pp, C.c_int(f.size), out.ctypes.data_as(c_dp))
        return out

    def bermudan(self, is_exercise, fixing_dates, period_lengths, payment_dates, notionals, swaprates, is_callable=True):
        ex = np.ascontiguousarray(is_exercise, dtype=np.int32)
        f, fp = _d(fixing_dates)
        pl, plp = _d(period_lengths)
        p, pp = _d(payment_dates)
        nt, ntp = _d(notionals)
        s, sp = _d(swaprates)
        n_ex = int(ex.sum())
        vals = np.empty(self.paths)
        ext = np.empty(self.paths)
        reg = np.zeros((n_ex, 6))
        cond = np.zeros(n_ex)
        se = C.c_double()
        price = lib().orc_lmm_bermudan(self.h, ex.ctypes.data_as(c_ip), fp, plp, pp, ntp, sp, C.c_int(f.size), C.c_int(1 if is_callable else 0),
                                       vals.ctypes.data_as(c_dp), ext.ctypes.data_as(c_dp), reg.ctypes.data_as(c_dp),
                                       cond.ctypes.data_as(c_dp), C.byref(se))
        return dict(price=price, values=vals, exercise_time=ext, regression=reg, cond=cond, std_error=se.value)


def time_lmm_reference_shaped(seed, sim_times, tenor_times, F, paths, L0, sigma, factor_matrix, scheme=2):
    st, sp = _d(sim_times)
    tt, tp = _d(tenor_times)
    L0, lp = _d(L0)
    sigma, sgp = _d(sigma)
    fm, fp = _d(factor_matrix)
    chk = C.c_double()
    sec = lib().orc_time_lmm_reference_shaped(C.c_int(seed), sp, C.c_int(st.size), tp, C.c_int(tt.size), C.c_int(F), C.c_int(paths), lp,
                                              sgp, fp, C.c_int(scheme), C.byref(chk))
    return sec, chk.value


def time_lmm_fused(seed, sim_times, tenor_times, F, paths, L0, sigma, factor_matrix, scheme=2, threads=1, want_process=False):
    st, sp = _d(sim_times)
    tt, tp = _d(tenor_times)
    L0, lp = _d(L0)
    sigma, sgp = _d(sigma)
    fm, fp = _d(factor_matrix)
    proc = np.empty((st.size, tt.size - 1, paths)) if want_process else None
    sec = lib().orc_time_lmm_fused(C.c_int(seed), sp, C.c_int(st.size), tp, C.c_int(tt.size), C.c_int(F), C.c_int(paths), lp, sgp, fp,
                                   C.c_int(scheme), C.c_int(threads), proc.ctypes.data_as(c_dp) if want_process else None)
    return sec, proc


def hull_white_process(seed, times, paths, vol_times, vol, mr, scheme=0, path_offset=0):
    t, tp = _d(times)
    vt, vtp = _d(vol_times)
    v, vp = _d(vol)
    m, mp_ = _d(mr)
    proc = np.empty((t.size, 2, paths))
    coef = np.empty((t.size - 1, 6))
    lib().orc_hull_white_process(C.c_int(seed), tp, C.c_int(t.size), C.c_int(paths), C.c_int64(path_offset), vtp, C.c_int(vt.size), vp, mp_,
                                 C.c_int(scheme), proc.ctypes.data_as(c_dp), coef.ctypes.data_as(c_dp))
    return proc, coef


def hull_white_caplet(seed, times, paths, vol_times, vol, mr, curve_times, df_discount, df_forward, scheme, maturity, period_length, strike):
    t, tp = _d(times)
    vt, vtp = _d(vol_times)
    v, vp = _d(vol)
    m, mp_ = _d(mr)
    ct, ctp = _d(curve_times)
    dd, ddp = _dn(df_discount)
    dfw, dfwp = _d(df_forward)
    vals, num, fr = np.empty(paths), np.empty(paths), np.empty(paths)
    price = lib().orc_hull_white_caplet(C.c_int(seed), tp, C.c_int(t.size), C.c_int(paths), vtp, C.c_int(vt.size), vp, mp_, ctp, C.c_int(ct.size), ddp, dfwp,
                                        C.c_int(scheme), C.c_double(maturity), C.c_double(period_length), C.c_double(strike),
                                        vals.ctypes.data_as(c_dp), num.ctypes.data_as(c_dp), fr.ctypes.data_as(c_dp))
    return price, vals, num, fr


def bs_bermudan_option(seed, times, paths, s0, r, sigma, scheme, exercise_dates, notionals, strikes, n_basis=5, intrinsic=False, binning=False):
    t, tp = _d(times)
    e, ep = _d(exercise_dates)
    nt, ntp = _d(notionals)
    k, kp = _d(strikes)
    vals, ext, reg = np.empty(paths), np.empty(paths), np.zeros((e.size, n_basis))
    price = lib().orc_bs_bermudan_option(C.c_int(seed), tp, C.c_int(t.size), C.c_int(paths), C.c_double(s0), C.c_double(r), C.c_double(sigma),
                                         C.c_int(scheme), ep, ntp, kp, C.c_int(e.size), C.c_int(n_basis), C.c_int(1 if intrinsic else 0),
                                         C.c_int(1 if binning else 0), vals.ctypes.data_as(c_dp), ext.ctypes.data_as(c_dp), reg.ctypes.data_as(c_dp))
    return dict(price=price, values=vals, exercise_time=ext, regression=reg)


def regression_localized(basis, y, standard_deviations):
    b, bp = _d(basis)
    y, yp = _d(y)
    K = b.shape[0]
    x, ce = np.empty(K), np.empty(y.size)
    lib().orc_regression_localized(bp, C.c_int(K), yp, C.c_uint64(y.size), C.c_double(standard_deviations), x.ctypes.data_as(c_dp), ce.ctypes.data_as(c_dp))
    return x, ce

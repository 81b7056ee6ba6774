"""ctypes wrapper around oracle/libee_oracle.so -- the CPU restatement of the reference (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
LIB = ORACLE_DIR / "libee_oracle.so"

_dp = C.POINTER(C.c_double)


def build(force=False):
    src = ORACLE_DIR / "ee_oracle.cpp"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(ORACLE_DIR), "-B", "libee_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB


def _load():
    build()
    lib = C.CDLL(str(LIB))
    lib.ora_nbody_create.restype = C.c_void_p
    lib.ora_nbody_create.argtypes = [C.c_int64, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int32]
    lib.ora_nbody_destroy.argtypes = [C.c_void_p]
    lib.ora_nbody_create_compensated.restype = C.c_void_p
    lib.ora_nbody_create_compensated.argtypes = [C.c_int64, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int32]
    lib.ora_inbody_step.restype = C.c_int32
    lib.ora_inbody_step.argtypes = [C.c_void_p, C.c_int64]
    lib.ora_inbody_state.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.ora_inbody_destroy.argtypes = [C.c_void_p]
    lib.ora_nbody_set_solout.argtypes = [C.c_void_p, C.c_double, _dp, C.POINTER(C.c_int32), C.c_int32]
    lib.ora_nbody_step.restype = C.c_int32
    lib.ora_nbody_step.argtypes = [C.c_void_p, C.c_int64]
    lib.ora_nbody_state.argtypes = [C.c_void_p, _dp, _dp, _dp, _dp]
    lib.ora_nbody_evals.restype = C.c_uint64
    lib.ora_nbody_evals.argtypes = [C.c_void_p]
    lib.ora_nbody_spline_len.restype = C.c_int64
    lib.ora_nbody_spline_len.argtypes = [C.c_void_p, C.c_int64]
    lib.ora_nbody_spline_get.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, _dp, C.POINTER(C.c_int32)]
    lib.ora_nbody_take_solution.argtypes = [C.c_void_p]
    lib.ora_nbody_solution_time.restype = C.c_double
    lib.ora_nbody_solution_time.argtypes = [C.c_void_p]
    lib.ora_gravity_eval.argtypes = [C.c_int64, _dp, _dp, _dp]
    lib.ora_gravity_eval_rows.argtypes = [C.c_int64, _dp, _dp, C.c_int64, C.c_int64, _dp]
    lib.ora_lsq_fit.restype = C.c_int32
    lib.ora_lsq_fit.argtypes = [C.c_int32, _dp, _dp, C.c_int32, _dp]
    lib.ora_ephem_create.restype = C.c_void_p
    lib.ora_ephem_create.argtypes = [C.c_int64, _dp, _dp, _dp, C.POINTER(C.c_int64), _dp, C.POINTER(C.c_int32)]
    lib.ora_ephem_destroy.argtypes = [C.c_void_p]
    lib.ora_ephem_state_vector.restype = C.c_int32
    lib.ora_ephem_state_vector.argtypes = [C.c_void_p, C.c_int64, C.c_double, _dp, _dp]
    lib.ora_ephem_position.restype = C.c_int32
    lib.ora_ephem_position.argtypes = [C.c_void_p, C.c_int64, C.c_double, _dp]
    lib.ora_ship_create.restype = C.c_void_p
    lib.ora_ship_create.argtypes = [C.c_void_p, C.c_double, _dp, _dp, C.c_uint32, C.c_int32, _dp, _dp, _dp,
                                    C.POINTER(C.c_int32)]
    lib.ora_ship_destroy.argtypes = [C.c_void_p]
    lib.ora_ship_step.restype = C.c_int32
    lib.ora_ship_step.argtypes = [C.c_void_p, C.c_int64]
    lib.ora_ship_step_to.restype = C.c_int32
    lib.ora_ship_step_to.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.POINTER(C.c_int64)]
    lib.ora_ship_knot_count.restype = C.c_int64
    lib.ora_ship_knot_count.argtypes = [C.c_void_p]
    lib.ora_ship_knots.argtypes = [C.c_void_p, _dp]
    lib.ora_ship_info.argtypes = [C.c_void_p, _dp, _dp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    lib.ora_relative_state_vector.restype = C.c_int32
    lib.ora_relative_state_vector.argtypes = [C.c_void_p, _dp, C.c_int64, C.c_int32, C.c_int32, C.c_double, _dp, _dp]
    lib.ora_ship_set_method.restype = C.c_int32
    lib.ora_ship_set_method.argtypes = [C.c_void_p, C.c_int32]
    lib.ora_ship_enable_analytics.argtypes = [C.c_void_p, _dp]
    lib.ora_ship_analytics_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    _i32 = C.POINTER(C.c_int32)
    lib.ora_ship_analytics.argtypes = [C.c_void_p, _dp, _i32, _dp, _dp, _i32, _i32]
    lib.ora_hermite_eval.argtypes = [_dp, _dp, C.c_double, _dp, _dp]
    lib.ora_set_pow_mode.argtypes = [C.c_int32]
    lib.ora_set_pair_variant.argtypes = [C.c_int32]
    lib.ora_pow_portable.restype = C.c_double
    lib.ora_pow_portable.argtypes = [C.c_double, C.c_double]
    return lib


lib = _load()


def p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


POW_LIBM = 0      # std::pow: what Rust's f64::powf resolves to on this platform (the reference as built here)
POW_PORTABLE = 1  # the engine's bit-reproducible double-double pow (ee_pow.cuh / ee_oracle_pow.h)


def set_pow_mode(mode):
    """Selects the pow used by the ship step-size controller (runge_kutta/mod.rs:238)."""
    lib.ora_set_pow_mode(int(mode))


def set_pair_variant(v):
    """0 = dir * (mu / (n*sqrt(n))) (particular's published scalar form, default); 1 = one reciprocal, two products."""
    lib.ora_set_pair_variant(int(v))


def pow_portable(x, y):
    return lib.ora_pow_portable(float(x), float(y))


def gravity_eval(pos, mu):
    pos = f64(pos).reshape(-1, 3)
    mu = f64(mu)
    out = np.zeros_like(pos)
    lib.ora_gravity_eval(len(mu), p(pos), p(mu), p(out))
    return out


def lsq_fit(degree, ts, xs):
    ts = f64(ts)
    xs = f64(xs).reshape(-1, 3)
    out = np.zeros((9, 3))
    n = lib.ora_lsq_fit(int(degree), p(ts), p(xs), len(ts), p(out))
    return out[: max(n, 0)].copy(), n


class NBody:
    """Oracle twin of NBodyPropagator<QT12 | Stormer13> with the spline solout."""

    def __init__(self, pos, vel, mu, t0, h_signed, method=12):
        self.pos0 = f64(pos).reshape(-1, 3)
        vel = f64(vel).reshape(-1, 3)
        mu = f64(mu)
        self.n = len(mu)
        self.h = lib.ora_nbody_create(self.n, p(self.pos0), p(vel), p(mu), float(t0), float(h_signed), int(method))
        self.backward = h_signed < 0

    def set_solout(self, delta, periods, degrees):
        periods = f64(periods)
        degrees = np.ascontiguousarray(np.asarray(degrees, dtype=np.int32))
        lib.ora_nbody_set_solout(self.h, float(delta), p(periods), degrees.ctypes.data_as(C.POINTER(C.c_int32)),
                                 1 if self.backward else 0)

    def step(self, n=1):
        return lib.ora_nbody_step(self.h, int(n))

    def state(self):
        t = C.c_double()
        pos = np.zeros((self.n, 3))
        vel = np.zeros((self.n, 3))
        acc = np.zeros((self.n, 3))
        lib.ora_nbody_state(self.h, C.byref(t), p(pos), p(vel), p(acc))
        return t.value, pos, vel, acc

    def evals(self):
        return lib.ora_nbody_evals(self.h)

    def solution_time(self):
        return lib.ora_nbody_solution_time(self.h)

    def splines(self):
        """list of (start, interval, [coeff arrays])"""
        out = []
        for b in range(self.n):
            m = lib.ora_nbody_spline_len(self.h, b)
            st, iv = C.c_double(), C.c_double()
            co = np.zeros((max(m, 1), 9, 3))
            nc = np.zeros(max(m, 1), dtype=np.int32)
            lib.ora_nbody_spline_get(self.h, b, C.byref(st), C.byref(iv), p(co), nc.ctypes.data_as(C.POINTER(C.c_int32)))
            out.append((st.value, iv.value, [co[k, : nc[k]].copy() for k in range(m)]))
        return out

    def take_solution(self):
        s = self.splines()
        lib.ora_nbody_take_solution(self.h)
        return s

    def __del__(self):
        if getattr(self, "h", None):
            lib.ora_nbody_destroy(self.h)
            self.h = None


class NBodyCompensated:
    """Same integrator over the compensated Double<DVec3> state of the reference's convergence test
    (ephemeris/tests/solar_system_convergence.rs:12-110)."""

    def __init__(self, pos, vel, mu, t0, h_signed, method=12):
        pos = f64(pos).reshape(-1, 3)
        vel = f64(vel).reshape(-1, 3)
        mu = f64(mu)
        self.n = len(mu)
        self.h = lib.ora_nbody_create_compensated(self.n, p(pos), p(vel), p(mu), float(t0), float(h_signed), int(method))

    def step(self, n=1):
        return lib.ora_inbody_step(self.h, int(n))

    def state(self):
        t = C.c_double()
        pos = np.zeros((self.n, 3))
        vel = np.zeros((self.n, 3))
        lib.ora_inbody_state(self.h, C.byref(t), p(pos), p(vel), None)
        return t.value, pos, vel

    def __del__(self):
        if getattr(self, "h", None):
            lib.ora_inbody_destroy(self.h)
            self.h = None


class Ephem:
    def __init__(self, mu, splines):
        """splines: list of (start, interval, [coeff arrays (nc,3)])"""
        nb = len(splines)
        self.mu = f64(mu)
        start = f64([s[0] for s in splines])
        interval = f64([s[1] for s in splines])
        npoly = np.array([len(s[2]) for s in splines], dtype=np.int64)
        tot = int(npoly.sum())
        co = np.zeros((max(tot, 1), 9, 3))
        nc = np.zeros(max(tot, 1), dtype=np.int32)
        k = 0
        for s in splines:
            for c in s[2]:
                co[k, : len(c)] = c
                nc[k] = len(c)
                k += 1
        self.nb = nb
        self.h = lib.ora_ephem_create(nb, p(self.mu), p(start), p(interval), npoly.ctypes.data_as(C.POINTER(C.c_int64)), p(co),
                                      nc.ctypes.data_as(C.POINTER(C.c_int32)))

    def state_vector(self, b, t):
        pos = np.zeros(3)
        vel = np.zeros(3)
        ok = lib.ora_ephem_state_vector(self.h, b, float(t), p(pos), p(vel))
        return (pos, vel) if ok else None

    def position(self, b, t):
        pos = np.zeros(3)
        ok = lib.ora_ephem_position(self.h, b, float(t), p(pos))
        return pos if ok else None

    def __del__(self):
        if getattr(self, "h", None):
            lib.ora_ephem_destroy(self.h)
            self.h = None


class Ship:
    def __init__(self, ephem, t0, state6, params7, n_max, burns=(), method=0):
        """params7 = (h_init, h_max, tol_pos, tol_vel, fac_min, fac_max, fac); burns = [(start, end, acc3, ref)];
        method = IntegrationMethod id (0 Verner87 .. 7 Fine45, include/ee_b200.h EE_SHIP_*)"""
        self.ephem = ephem
        st = f64(state6)
        pr = f64(params7)
        nb = len(burns)
        bs = f64([b[0] for b in burns] or [0.0])
        be = f64([b[1] for b in burns] or [0.0])
        ba = f64([list(b[2]) for b in burns] or [[0.0, 0.0, 0.0]])
        br = np.ascontiguousarray(np.array([b[3] for b in burns] or [-1], dtype=np.int32))
        self.h = lib.ora_ship_create(ephem.h, float(t0), p(st), p(pr), int(n_max), nb, p(bs), p(be), p(ba),
                                     br.ctypes.data_as(C.POINTER(C.c_int32)))
        if method:
            assert lib.ora_ship_set_method(self.h, int(method)) == 0

    def enable_analytics(self, soi_radius):
        r = f64(soi_radius)
        lib.ora_ship_enable_analytics(self.h, p(r))

    def analytics(self):
        """(transitions [(time, body)], apsides [(time, distance, body, kind)])"""
        ntr, nap = C.c_int64(), C.c_int64()
        lib.ora_ship_analytics_counts(self.h, C.byref(ntr), C.byref(nap))
        tt, tb = np.zeros(max(ntr.value, 1)), np.zeros(max(ntr.value, 1), dtype=np.int32)
        at, ad = np.zeros(max(nap.value, 1)), np.zeros(max(nap.value, 1))
        ab, ak = np.zeros(max(nap.value, 1), dtype=np.int32), np.zeros(max(nap.value, 1), dtype=np.int32)
        i32 = C.POINTER(C.c_int32)
        lib.ora_ship_analytics(self.h, p(tt), tb.ctypes.data_as(i32), p(at), p(ad), ab.ctypes.data_as(i32), ak.ctypes.data_as(i32))
        return ([(float(tt[k]), int(tb[k])) for k in range(ntr.value)],
                [(float(at[k]), float(ad[k]), int(ab[k]), int(ak[k])) for k in range(nap.value)])

    def step(self, n=1):
        return lib.ora_ship_step(self.h, int(n))

    def step_to(self, t_end, max_steps=1 << 30):
        taken = C.c_int64()
        st = lib.ora_ship_step_to(self.h, float(t_end), int(max_steps), C.byref(taken))
        return st, taken.value

    def knots(self):
        n = lib.ora_ship_knot_count(self.h)
        out = np.zeros((n, 7))
        lib.ora_ship_knots(self.h, p(out))
        return out

    def info(self):
        t, nh = C.c_double(), C.c_double()
        na = C.c_uint32()
        ev = C.c_uint64()
        lib.ora_ship_info(self.h, C.byref(t), C.byref(nh), C.byref(na), C.byref(ev))
        return dict(time=t.value, next_h=nh.value, n_attempts=na.value, rhs_evals=ev.value)

    def __del__(self):
        if getattr(self, "h", None):
            lib.ora_ship_destroy(self.h)
            self.h = None


def relative_state_vector(ephem, at, body=0, reference=None, knots=None):
    """RelativeTrajectory::state_vector: (pos, vel) or None."""
    pos, vel = np.zeros(3), np.zeros(3)
    kn = f64(knots) if knots is not None else None
    ok = lib.ora_relative_state_vector(ephem.h, p(kn) if kn is not None else None, len(kn) if kn is not None else 0, int(body),
                                       -1 if reference is None else int(reference), float(at), p(pos), p(vel))
    return (pos, vel) if ok else None


def hermite_eval(k0, k1, t):
    k0, k1 = f64(k0), f64(k1)
    pos, vel = np.zeros(3), np.zeros(3)
    lib.ora_hermite_eval(p(k0), p(k1), float(t), p(pos), p(vel))
    return pos, vel


def spline_position_from_knots(knots, t):
    """CubicHermiteSpline::position (trajectory.rs:779-786)."""
    ts = knots[:, 0]
    i = int(np.searchsorted(ts, t))
    if i < len(ts) and ts[i] == t:
        return knots[i, 1:4].copy()
    if i == 0 or i >= len(ts):
        return None
    return hermite_eval(knots[i - 1], knots[i], t)[0]

"""Host-side mirror of the reference's propagator interface over the C ABI (plumbing, not the product).

Names follow the `ephemeris` crate (ephemeris/src/lib.rs:9-79, propagators/nbody.rs, propagators/spacecraft.rs,
trajectory.rs) so parity tests read like the reference's own: `NBodyPropagator.new(...)`, `.step()`, `.step_to()`,
`.propagate()`, `.take_solution()`, `.time()`, `.has_reached()`; `SpacecraftPropagator`; `UniformSpline`,
`CubicHermiteSpline`.  Every computation happens in libee_b200.so on the GPU.
"""
import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import AdaptiveParams, check, lib

QUINLAN_TREMAINE_12 = 12
STORMER_13 = 13
BLANES_MOAN_14A = 14  # symplectic Runge-Kutta-Nystrom, 15 stages (integration/src/methods.rs:1730-1774)
MODE_PARITY = 0
MODE_THROUGHPUT = 1
EXCHANGE_ALLREDUCE = 0
EXCHANGE_ALLGATHER = 1


def _dp(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_lib.c_double_p)


def _f64(a, shape=None) -> np.ndarray:
    out = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        out = out.reshape(shape)
    return out


@dataclass
class Forward:
    """propagators/mod.rs:23-57"""
    delta: float

    def signed_delta(self) -> float:
        return abs(self.delta)


@dataclass
class Backward:
    """propagators/mod.rs:59-93"""
    delta: float

    def signed_delta(self) -> float:
        return -abs(self.delta)


@dataclass
class LeastSquaresFit:
    """ephemeris_explorer/src/dynamics/celestial.rs:19-22"""
    degree: int


def _as_usize(x: float) -> int:
    """Rust's saturating `f64 as usize`."""
    if x != x or x <= 0.0:
        return 0
    return int(x) if x < 18446744073709551616.0 else 18446744073709551615


def _sign_negative(x: float) -> bool:
    return math.copysign(1.0, x) < 0.0  # f64::is_sign_negative: true for -0.0 (ftime/src/duration.rs:78-85)


@dataclass
class UniformSpline:
    """trajectory.rs:412-417: start epoch, interval, polynomials (each (n_coef, 3), lowest order first).

    The container operations are the ones the Prediction Planner applies to every solution it takes from a propagator
    (`PredictionTarget::merge`, dynamics/celestial.rs:194-235: forward = `clear_after(start)` + `append`, backward =
    `clear_before(end)` + `prepend`), with the reference's index rules (trajectory.rs:600-617).  Evaluation stays on the
    device (`Ephemeris.evaluate`)."""
    start: float
    interval: float
    polynomials: List[np.ndarray]

    def span(self) -> float:
        return self.interval * float(len(self.polynomials))

    def end(self) -> float:
        return self.start + self.span()

    # BoundedTrajectory (trajectory.rs:427-447)
    def contains(self, time: float) -> bool:
        local = time - self.start
        return (not _sign_negative(local)) and local <= self.span()

    def segment_count(self) -> int:
        return len(self.polynomials)

    # index rules (trajectory.rs:571-617)
    def _index_local(self, local: float) -> int:
        return _as_usize(local / self.interval)

    def _index_local_exclusive(self, local: float) -> int:
        return max(_as_usize(math.ceil(local / self.interval)) - 1, 0)

    def get_index(self, at: float) -> Optional[int]:
        local = at - self.start
        if _sign_negative(local) or local >= self.span():
            return None
        return self._index_local(local)

    def get_index_exclusive(self, at: float) -> Optional[int]:
        local = at - self.start
        if _sign_negative(local) or local > self.span():
            return None
        return self._index_local_exclusive(local)

    # container operations (trajectory.rs:474-540)
    def between(self, start: float, end: float) -> Optional["UniformSpline"]:
        if not self.polynomials:
            return None
        a = self.get_index_exclusive(start)
        b = self.get_index_exclusive(end)
        if a is None or b is None:
            return None
        return UniformSpline(self.start + self.interval * float(a), self.interval,
                             [self.polynomials[i] for i in range(a, b + 1) if i < len(self.polynomials)])

    def push_front(self, polynomial: np.ndarray) -> None:
        self.polynomials.insert(0, polynomial)
        self.start -= self.interval

    def push_back(self, polynomial: np.ndarray) -> None:
        self.polynomials.append(polynomial)

    def prepend(self, trajectory: "UniformSpline") -> None:
        assert self.start == trajectory.start + trajectory.span(), "prepend: the pieces do not meet"
        assert self.interval == trajectory.interval
        self.start = trajectory.start
        self.polynomials[:0] = trajectory.polynomials

    def append(self, trajectory: "UniformSpline") -> None:
        assert self.start + self.span() == trajectory.start, "append: the pieces do not meet"
        assert self.interval == trajectory.interval
        self.polynomials.extend(trajectory.polynomials)

    def clear_before(self, at: float) -> None:
        idx = self.get_index_exclusive(at + self.interval)
        if idx is not None:
            self.start += self.interval * float(idx)
            del self.polynomials[:idx]

    def clear_after(self, at: float) -> None:
        idx = self.get_index(at)
        if idx is not None:
            del self.polynomials[idx:]

    # PredictionTarget::merge for CelestialTrajectory<Forward | Backward> (dynamics/celestial.rs:194-235)
    def merge_forward(self, propagated: "UniformSpline") -> None:
        self.clear_after(propagated.start)
        self.append(propagated)

    def merge_backward(self, propagated: "UniformSpline") -> None:
        self.clear_before(propagated.end())
        self.prepend(propagated)


def set_pair_variant(variant: int) -> None:
    """Reading of `particular`'s pair kernel used by parity-mode handles created afterwards (include/ee_b200.h)."""
    check(lib.ee_set_pair_variant(int(variant)), "ee_set_pair_variant")


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(lib.ee_nccl_unique_id(buf), "ee_nccl_unique_id")
    return buf.raw


class Ephemeris:
    """Device-resident Vec<UniformSpline<DVec3>> + gravitational parameters (the ships' AccelerationModel context)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_splines(cls, mus, splines: Sequence[UniformSpline], device: int = 0) -> "Ephemeris":
        nb = len(splines)
        mus = _f64(mus)
        start = _f64([s.start for s in splines])
        interval = _f64([s.interval for s in splines])
        n_poly = np.array([len(s.polynomials) for s in splines], dtype=np.int64)
        total = int(n_poly.sum())
        coeffs = np.zeros((max(total, 1), 9, 3), dtype=np.float64)
        n_coef = np.zeros(max(total, 1), dtype=np.int32)
        k = 0
        for s in splines:
            for p in s.polynomials:
                coeffs[k, : len(p)] = p
                n_coef[k] = len(p)
                k += 1
        h = C.c_void_p()
        check(lib.ee_ephem_create(nb, _dp(mus), _dp(start), _dp(interval), n_poly.ctypes.data_as(_lib.c_i64_p), _dp(coeffs),
                                  n_coef.ctypes.data_as(_lib.c_i32_p), device, C.byref(h)), "ee_ephem_create")
        return cls(h)

    def sizes(self):
        nb = C.c_int64()
        check(lib.ee_ephem_sizes(self._h, C.byref(nb), None), "ee_ephem_sizes")
        n_poly = np.zeros(nb.value, dtype=np.int64)
        check(lib.ee_ephem_sizes(self._h, C.byref(nb), n_poly.ctypes.data_as(_lib.c_i64_p)), "ee_ephem_sizes")
        return nb.value, n_poly

    def splines(self):
        nb, n_poly = self.sizes()
        total = int(n_poly.sum())
        mus = np.zeros(nb)
        start = np.zeros(nb)
        interval = np.zeros(nb)
        coeffs = np.zeros((max(total, 1), 9, 3))
        n_coef = np.zeros(max(total, 1), dtype=np.int32)
        check(lib.ee_ephem_get(self._h, _dp(mus), _dp(start), _dp(interval), _dp(coeffs), n_coef.ctypes.data_as(_lib.c_i32_p)),
              "ee_ephem_get")
        return mus, _assemble(start, interval, n_poly, coeffs, n_coef)

    def evaluate(self, times, velocities: bool = True):
        """UniformSpline::position / state_vector for every body at every time -> (pos, vel, ok)."""
        nb, _ = self.sizes()
        times = _f64(times).reshape(-1)
        nt = len(times)
        pos = np.zeros((nt, nb, 3))
        vel = np.zeros((nt, nb, 3)) if velocities else None
        ok = np.zeros((nt, nb), dtype=np.int32)
        check(lib.ee_ephem_evaluate(self._h, nt, _dp(times), _dp(pos), _dp(vel), ok.ctypes.data_as(_lib.c_i32_p)),
              "ee_ephem_evaluate")
        return pos, vel, ok.astype(bool)

    def evaluate_relative(self, body: int, reference: Optional[int], times):
        """RelativeTrajectory::state_vector (trajectory.rs:315-335) of body `body` w.r.t. `reference` (None = no reference)
        at every time -> (pos, vel, ok): what the plot sampler evaluates per candidate point (ui/world/plot.rs:326-334)."""
        times = _f64(times).reshape(-1)
        nt = len(times)
        pos, vel = np.zeros((nt, 3)), np.zeros((nt, 3))
        ok = np.zeros(nt, dtype=np.int32)
        check(lib.ee_ephem_evaluate_relative(self._h, int(body), -1 if reference is None else int(reference), nt, _dp(times), _dp(pos),
                                             _dp(vel), ok.ctypes.data_as(_lib.c_i32_p)), "ee_ephem_evaluate_relative")
        return pos, vel, ok.astype(bool)

    def close(self):
        if self._h:
            lib.ee_ephem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _assemble(start, interval, n_poly, coeffs, n_coef) -> List[UniformSpline]:
    out = []
    k = 0
    for b in range(len(n_poly)):
        polys = []
        for _ in range(int(n_poly[b])):
            polys.append(np.array(coeffs[k, : n_coef[k]], dtype=np.float64))
            k += 1
        out.append(UniformSpline(float(start[b]), float(interval[b]), polys))
    return out


class NBodyPropagator:
    """ephemeris::NBodyPropagator<D, DVec3, M, SplineInterpolators<D, DVec3, LeastSquaresFit>> (nbody.rs:65-235)."""

    def __init__(self, handle, n: int):
        self._h = handle
        self.n = n

    @classmethod
    def new(cls, direction, initial_time: float, positions, velocities, gravitational_parameters,
            solout: Optional[tuple] = None, *, method: int = QUINLAN_TREMAINE_12, mode: int = MODE_PARITY,
            device: int = 0, rank: int = 0, world: int = 1, unique_id: Optional[bytes] = None,
            exchange: int = EXCHANGE_ALLGATHER) -> "NBodyPropagator":
        """NBodyPropagator::new (nbody.rs:93-121).  solout = (delta, sample_periods, [LeastSquaresFit | degree])
        mirrors SplineInterpolators::new + SplineInterpolator{..} (dynamics/celestial.rs:156-186)."""
        pos = _f64(positions).reshape(-1, 3)
        vel = _f64(velocities).reshape(-1, 3)
        mu = _f64(gravitational_parameters).reshape(-1)
        n = len(mu)
        assert pos.shape == (n, 3) and vel.shape == (n, 3)
        h = C.c_void_p()
        if world > 1:
            check(lib.ee_nbody_create_sharded(n, _dp(pos), _dp(vel), _dp(mu), float(initial_time), direction.signed_delta(),
                                              method, mode, device, rank, world, unique_id, exchange, C.byref(h)),
                  "ee_nbody_create_sharded")
        else:
            check(lib.ee_nbody_create(n, _dp(pos), _dp(vel), _dp(mu), float(initial_time), direction.signed_delta(), method,
                                      mode, device, C.byref(h)), "ee_nbody_create")
        self = cls(h, n)
        if solout is not None:
            delta, periods, algos = solout
            periods = _f64(periods).reshape(-1)
            degrees = np.array([a.degree if isinstance(a, LeastSquaresFit) else int(a) for a in algos], dtype=np.int32)
            assert len(periods) == n and len(degrees) == n
            check(lib.ee_nbody_set_solout(h, float(delta), _dp(periods), degrees.ctypes.data_as(_lib.c_i32_p)),
                  "ee_nbody_set_solout")
        return self

    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(512)
        check(lib.ee_nbody_p2p_export(self._h, buf), "ee_nbody_p2p_export")
        return buf.raw

    def p2p_connect(self, all_blobs: bytes) -> None:
        check(lib.ee_nbody_p2p_connect(self._h, all_blobs), "ee_nbody_p2p_connect")

    def p2p_trace(self, enable: bool):
        """Switch the per-launch event trace of the peer step; returns (mean ms of the 4 phases, steps) traced so far."""
        ms = np.zeros(4)
        k = C.c_int64()
        check(lib.ee_nbody_p2p_trace(self._h, 1 if enable else 0, _dp(ms), C.byref(k)), "ee_nbody_p2p_trace")
        return ms, k.value

    # IncrementalPropagator
    def step(self, n_steps: int = 1) -> None:
        check(lib.ee_nbody_step(self._h, int(n_steps)), "NBodyPropagator.step")

    def try_step(self, n_steps: int = 1) -> int:
        return lib.ee_nbody_step(self._h, int(n_steps))

    def step_to(self, time: float) -> None:
        check(lib.ee_nbody_step_to(self._h, float(time)), "NBodyPropagator.step_to")

    # BoundedPropagator (ephemeris/src/lib.rs:60-79)
    def propagate(self, to: float) -> List[UniformSpline]:
        self.step_to(to)
        return self.take_solution()

    def sync(self) -> None:
        check(lib.ee_nbody_sync(self._h), "ee_nbody_sync")

    # DirectionalPropagator
    def time(self) -> float:
        t = C.c_double()
        check(lib.ee_nbody_solution_time(self._h, C.byref(t)), "ee_nbody_solution_time")
        return t.value

    def has_reached(self, time: float) -> bool:
        r = C.c_int32()
        check(lib.ee_nbody_has_reached(self._h, float(time), C.byref(r)), "ee_nbody_has_reached")
        return bool(r.value)

    def delta(self) -> float:
        return lib.ee_nbody_delta(self._h)

    def step_count(self) -> int:
        return lib.ee_nbody_step_count(self._h)

    def state(self, accelerations: bool = False):
        """(problem.time, problem.state.y, problem.state.dy[, current_ddy])"""
        t = C.c_double()
        pos = np.zeros((self.n, 3))
        vel = np.zeros((self.n, 3))
        acc = np.zeros((self.n, 3)) if accelerations else None
        check(lib.ee_nbody_state(self._h, C.byref(t), _dp(pos), _dp(vel), _dp(acc)), "ee_nbody_state")
        return (t.value, pos, vel, acc) if accelerations else (t.value, pos, vel)

    def state_async(self, positions: np.ndarray, velocities: Optional[np.ndarray] = None) -> float:
        """Non-blocking state read into caller arrays ((n, 3) float64, ideally page-locked); pair with state_wait()."""
        t = C.c_double()
        check(lib.ee_nbody_state_async(self._h, C.byref(t), _dp(positions), _dp(velocities)), "ee_nbody_state_async")
        return t.value

    def state_wait(self) -> None:
        check(lib.ee_nbody_state_wait(self._h), "ee_nbody_state_wait")

    def _sizes(self) -> np.ndarray:
        n_poly = np.zeros(self.n, dtype=np.int64)
        check(lib.ee_nbody_solution_sizes(self._h, n_poly.ctypes.data_as(_lib.c_i64_p)), "ee_nbody_solution_sizes")
        return n_poly

    # Propagator
    def take_solution(self) -> List[UniformSpline]:
        n_poly = self._sizes()
        total = int(n_poly.sum())
        start = np.zeros(self.n)
        interval = np.zeros(self.n)
        coeffs = np.zeros((max(total, 1), 9, 3))
        n_coef = np.zeros(max(total, 1), dtype=np.int32)
        check(lib.ee_nbody_take_solution(self._h, _dp(start), _dp(interval), _dp(coeffs), n_coef.ctypes.data_as(_lib.c_i32_p)),
              "ee_nbody_take_solution")
        return _assemble(start, interval, n_poly, coeffs, n_coef)

    def take_solution_ephemeris(self) -> Ephemeris:
        h = C.c_void_p()
        check(lib.ee_nbody_take_solution_ephem(self._h, C.byref(h)), "ee_nbody_take_solution_ephem")
        return Ephemeris(h)

    def clone(self) -> "NBodyPropagator":
        h = C.c_void_p()
        check(lib.ee_nbody_clone(self._h, C.byref(h)), "ee_nbody_clone")
        return NBodyPropagator(h, self.n)

    def snapshot_size(self) -> int:
        b = C.c_int64()
        check(lib.ee_nbody_snapshot_size(self._h, C.byref(b)), "ee_nbody_snapshot_size")
        return b.value

    def snapshot(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Full multistep state -> host bytes (the reference's `propagator.clone()` snapshot, prediction.rs:224-229)."""
        if out is None:
            out = np.empty(self.snapshot_size(), dtype=np.uint8)
        check(lib.ee_nbody_snapshot(self._h, out.ctypes.data_as(C.c_void_p)), "ee_nbody_snapshot")
        return out

    def restore(self, blob: np.ndarray) -> None:
        blob = np.ascontiguousarray(blob)
        check(lib.ee_nbody_restore(self._h, blob.ctypes.data_as(C.c_void_p), blob.nbytes), "ee_nbody_restore")

    def step_timed(self, n_steps: int, flush_bytes: int = 0) -> float:
        ms = C.c_double()
        check(lib.ee_nbody_step_timed(self._h, int(n_steps), int(flush_bytes), C.byref(ms)), "ee_nbody_step_timed")
        return ms.value

    def last_timing(self):
        ms = C.c_double()
        k = C.c_int64()
        check(lib.ee_nbody_last_timing(self._h, C.byref(ms), C.byref(k)), "ee_nbody_last_timing")
        return ms.value, k.value

    def close(self):
        if self._h:
            lib.ee_nbody_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fp64_fma_peak(device: int = 0) -> float:
    t = C.c_double()
    check(lib.ee_fp64_fma_peak(device, C.byref(t)), "ee_fp64_fma_peak")
    return t.value


def gravity_eval(positions, mus, mode: int = MODE_PARITY, device: int = 0) -> np.ndarray:
    """NewtonianGravity::eval (nbody.rs:16-39) on the GPU."""
    pos = _f64(positions).reshape(-1, 3)
    mu = _f64(mus).reshape(-1)
    out = np.zeros_like(pos)
    check(lib.ee_gravity_eval(len(mu), _dp(pos), _dp(mu), mode, device, _dp(out)), "ee_gravity_eval")
    return out


def lsq_fit(degrees, ts, samples, device: int = 0):
    """Batched LeastSquaresFit::interpolate: samples (n, 9, 3) -> list of (n_coef, 3) arrays."""
    samples = _f64(samples).reshape(-1, 9, 3)
    nf = samples.shape[0]
    degrees = np.ascontiguousarray(np.broadcast_to(np.asarray(degrees, dtype=np.int32), (nf,)))
    ts = _f64(ts).reshape(9)
    coeffs = np.zeros((nf, 9, 3))
    n_coef = np.zeros(nf, dtype=np.int32)
    check(lib.ee_lsq_fit(nf, degrees.ctypes.data_as(_lib.c_i32_p), _dp(ts), _dp(samples), device, _dp(coeffs),
                         n_coef.ctypes.data_as(_lib.c_i32_p)), "ee_lsq_fit")
    return [coeffs[i, : n_coef[i]].copy() for i in range(nf)], n_coef


@dataclass
class ConstantThrust:
    """spacecraft.rs:30-57 with ReferenceFrame (dynamics/spacecraft.rs:254-293): reference = body index or -1 (inertial)."""
    acceleration: Sequence[float]
    reference: int = -1


@dataclass
class CubicHermiteSpline:
    """trajectory.rs:745-746: knots (t, position, velocity) -> array (k, 7).  `join` is what the Planner does with every
    ship solution it takes (`SpacecraftPropagator::join`, spacecraft.rs:558-561: `clear_after(rhs.start())` + `extend`)."""
    knots: np.ndarray

    def start(self) -> float:
        return float(self.knots[0, 0]) if len(self.knots) else -math.inf  # Epoch::MIN

    def end(self) -> float:
        return float(self.knots[-1, 0]) if len(self.knots) else math.inf  # Epoch::MAX

    def segment_count(self) -> int:
        return max(len(self.knots) - 1, 0)

    def binary_search(self, at: float):
        """(True, i) when knot i is at `at`, else (False, insertion index) -- Result<usize, usize> of trajectory.rs:811-813."""
        i = int(np.searchsorted(self.knots[:, 0], at, side="left"))
        return (i < len(self.knots) and float(self.knots[i, 0]) == at), i

    def get(self, at: float) -> Optional[np.ndarray]:
        found, i = self.binary_search(at)
        return self.knots[i, 1:] if found else None

    def clear_after(self, at: float) -> None:
        self.knots = self.knots[self.knots[:, 0] < at]  # retain(|k| at > k), trajectory.rs:835-838

    def extend(self, rhs: "CubicHermiteSpline") -> None:
        self.knots = np.concatenate([self.knots, rhs.knots], axis=0)

    def join(self, rhs: "CubicHermiteSpline") -> None:
        self.clear_after(rhs.start())
        self.extend(rhs)


POW_GLIBC = 0               # err.powf(-1/k) as glibc's pow evaluates it (the reference as built on Linux): default
POW_CORRECTLY_ROUNDED = 1   # the engine's libm-independent double-double pow


# IntegrationMethod (ephemeris_explorer/src/flight_plan.rs:175-184): the adaptive method a ship is integrated with
VERNER87, CASH_KARP45, DORMAND_PRINCE54, DORMAND_PRINCE87, FEHLBERG45, TSITOURAS75, VERNER98, FINE45 = range(8)
SHIP_METHOD_NAMES = ("Verner87", "CashKarp45", "DormandPrince54", "DormandPrince87", "Fehlberg45", "Tsitouras75", "Verner98",
                     "Fine45")


def default_adaptive_params(tol_position=1e-3, tol_velocity=1e-3, h_init=60.0, n_max=1_000_000,
                            pow_mode=POW_GLIBC, method=VERNER87) -> AdaptiveParams:
    """INITIAL_ADAPTIVE_PARAMS (ephemeris_explorer/src/load/mod.rs:472-486)."""
    import sys
    return AdaptiveParams(h_init, sys.float_info.max, tol_position, tol_velocity, 1.0 / 5.0, 5.0 / 1.0, 9.0 / 10.0, n_max,
                          pow_mode, method)


class ShipStepError(RuntimeError):
    """A ship's integrator returned a StepError (integration/src/lib.rs:312-337) during propagate()."""

    def __init__(self, code: int, ship: int, count: int):
        self.code, self.ship, self.count = code, ship, count
        super().__init__("ship %d: %s (%d ship(s) failed)" % (ship, _lib.STATUS_NAMES.get(code, "status %d" % code), count))


class SpacecraftPropagator:
    """A batch of ephemeris::SpacecraftPropagator<T, ReferenceFrame, Bodies, M, CubicHermiteSplineSolout> (spacecraft.rs:415-643),
    M = params.method (any IntegrationMethod of flight_plan.rs:175-184; default Verner87); enable_analytics() switches the
    solution to the app's SpacecraftSolout.  timelines[i] = list of (start, end, ConstantThrust)."""

    def __init__(self, handle, n, ephem):
        self._h = handle
        self.n = n
        self._ephem = ephem  # keep the context alive

    @classmethod
    def new(cls, initial_time, initial_state, params: AdaptiveParams, timelines, context: Ephemeris) -> "SpacecraftPropagator":
        st = _f64(initial_state).reshape(-1, 6)
        n = st.shape[0]
        t0 = _f64(np.broadcast_to(np.asarray(initial_time, dtype=np.float64), (n,)))
        if timelines is None:
            timelines = [[] for _ in range(n)]
        off = np.zeros(n + 1, dtype=np.int64)
        bs, be, ba, br = [], [], [], []
        for i, tl in enumerate(timelines):
            for (s, e, thrust) in tl:
                bs.append(s)
                be.append(e)
                ba.append(list(thrust.acceleration))
                br.append(thrust.reference)
            off[i + 1] = len(bs)
        bs_a, be_a = _f64(bs if bs else [0.0]), _f64(be if be else [0.0])
        ba_a = _f64(ba if ba else [[0.0, 0.0, 0.0]])
        br_a = np.ascontiguousarray(np.array(br if br else [-1], dtype=np.int32))
        h = C.c_void_p()
        check(lib.ee_ships_create(context._h, n, _dp(t0), _dp(st), C.byref(params), off.ctypes.data_as(_lib.c_i64_p), _dp(bs_a),
                                  _dp(be_a), _dp(ba_a), br_a.ctypes.data_as(_lib.c_i32_p), C.byref(h)), "ee_ships_create")
        return cls(h, n, context)

    def step_to(self, time: float, max_steps: int = 1 << 14) -> None:
        check(lib.ee_ships_step_to(self._h, float(time), int(max_steps)), "ee_ships_step_to")

    def info(self):
        status = np.zeros(self.n, dtype=np.int32)
        time = np.zeros(self.n)
        nk = np.zeros(self.n, dtype=np.int64)
        natt = np.zeros(self.n, dtype=np.uint32)
        evals = np.zeros(self.n, dtype=np.uint64)
        check(lib.ee_ships_info(self._h, status.ctypes.data_as(_lib.c_i32_p), _dp(time), nk.ctypes.data_as(_lib.c_i64_p),
                                natt.ctypes.data_as(_lib.c_u32_p), evals.ctypes.data_as(_lib.c_u64_p)), "ee_ships_info")
        return dict(status=status, time=time, n_knots=nk, n_attempts=natt, rhs_evals=evals)

    def take_solution(self) -> List[CubicHermiteSpline]:
        nk = self.info()["n_knots"]
        off = np.zeros(self.n + 1, dtype=np.int64)
        off[1:] = np.cumsum(nk)
        out = np.zeros((int(off[-1]), 7))
        check(lib.ee_ships_take_knots(self._h, off.ctypes.data_as(_lib.c_i64_p), _dp(out)), "ee_ships_take_knots")
        return [CubicHermiteSpline(out[off[i]: off[i + 1]].copy()) for i in range(self.n)]

    # SpacecraftSolout (ephemeris_explorer/src/dynamics/spacecraft.rs:448-586): trajectory + SOI transitions + apsides
    def enable_analytics(self, soi_radius) -> None:
        r = _f64(soi_radius)
        check(lib.ee_ships_enable_analytics(self._h, _dp(r)), "ee_ships_enable_analytics")

    def analytics(self):
        """Per ship: (transitions [(time, body)], apsides [(time, distance, body, kind)]), kind 0 = periapsis, 1 = apoapsis.
        Read before take_solution(), which starts the next solution."""
        ntr = np.zeros(self.n, dtype=np.int32)
        nap = np.zeros(self.n, dtype=np.int32)
        check(lib.ee_ships_analytics_counts(self._h, ntr.ctypes.data_as(_lib.c_i32_p), nap.ctypes.data_as(_lib.c_i32_p)),
              "ee_ships_analytics_counts")
        to = np.zeros(self.n + 1, dtype=np.int64)
        ao = np.zeros(self.n + 1, dtype=np.int64)
        to[1:] = np.cumsum(ntr)
        ao[1:] = np.cumsum(nap)
        tt, tb = np.zeros(max(int(to[-1]), 1)), np.zeros(max(int(to[-1]), 1), dtype=np.int32)
        at, ad = np.zeros(max(int(ao[-1]), 1)), np.zeros(max(int(ao[-1]), 1))
        ab, ak = np.zeros(max(int(ao[-1]), 1), dtype=np.int32), np.zeros(max(int(ao[-1]), 1), dtype=np.int32)
        i32 = _lib.c_i32_p
        check(lib.ee_ships_read_analytics(self._h, to.ctypes.data_as(_lib.c_i64_p), _dp(tt), tb.ctypes.data_as(i32),
                                          ao.ctypes.data_as(_lib.c_i64_p), _dp(at), _dp(ad), ab.ctypes.data_as(i32),
                                          ak.ctypes.data_as(i32)), "ee_ships_read_analytics")
        out = []
        for i in range(self.n):
            tr = [(float(tt[k]), int(tb[k])) for k in range(to[i], to[i + 1])]
            ap = [(float(at[k]), float(ad[k]), int(ab[k]), int(ak[k])) for k in range(ao[i], ao[i + 1])]
            out.append((tr, ap))
        return out

    def evaluate_relative(self, ship: int, reference: Optional[int], times):
        """RelativeTrajectory::state_vector of ship `ship`'s CubicHermiteSpline (the knots held since the last take_solution)
        w.r.t. body `reference` (None = no reference) -> (pos, vel, ok)."""
        times = _f64(times).reshape(-1)
        nt = len(times)
        pos, vel = np.zeros((nt, 3)), np.zeros((nt, 3))
        ok = np.zeros(nt, dtype=np.int32)
        check(lib.ee_ships_evaluate_relative(self._h, int(ship), -1 if reference is None else int(reference), nt, _dp(times), _dp(pos),
                                             _dp(vel), ok.ctypes.data_as(_lib.c_i32_p)), "ee_ships_evaluate_relative")
        return pos, vel, ok.astype(bool)

    def propagate(self, to: float, max_steps: int = 1 << 14) -> List[CubicHermiteSpline]:
        """BoundedPropagator::propagate (ephemeris/src/lib.rs:60-79): step every ship until its solution reaches `to`, then
        take the solutions.  Like the reference, a ship whose integrator returns an error (EvalFailed, MaxIterationsReached,
        BoundReached, StepSizeUnderflow) makes the call fail instead of handing back a silently truncated trajectory;
        `max_steps` only bounds one kernel launch, the call keeps launching until every ship has arrived."""
        while True:
            self.step_to(to, max_steps)
            info = self.info()
            bad = np.nonzero(info["status"] != 0)[0]
            if len(bad):
                i = int(bad[0])
                raise ShipStepError(int(info["status"][i]), i, len(bad))
            if np.all(info["time"] >= to):
                break
        return self.take_solution()

    def last_ms(self) -> float:
        return lib.ee_ships_last_ms(self._h)

    def close(self):
        if self._h:
            lib.ee_ships_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

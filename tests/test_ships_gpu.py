"""Massless-ship propagator on the GPU against the CPU oracle (GPU tests)."""
import sys

import numpy as np
import pytest

import oracle
from helpers import SHIP_TEST_DEGREES, SHIP_TEST_PERIOD_HOURS, load_system
import ephemeris_explorer_b200 as ee
from ephemeris_explorer_b200 import formats

pytestmark = pytest.mark.gpu

STATE = [-27204249.668775786, 132947582.43848978, 57641619.74241204, -22.253599106181895, -5.189518219791726, -2.2515617105336263]


def build_ephemeris():
    s = load_system("simple_solar_system_2433282.5")
    h = 6 * 3600.0
    periods = np.array(SHIP_TEST_PERIOD_HOURS) * 3600.0
    prop = ee.NBodyPropagator.new(ee.Forward(h), s.epoch, s.position, s.velocity, s.mu, solout=(h, periods, SHIP_TEST_DEGREES))
    prop.step_to(formats.parse_epoch("1951-03-01 00:00:00"))
    eph = prop.take_solution_ephemeris()
    mus, spl = eph.splines()
    ora = oracle.Ephem(mus, [(x.start, x.interval, x.polynomials) for x in spl])
    return s, eph, ora


def burns_for(s):
    E, D = formats.parse_epoch, formats.parse_duration
    n = s.names
    return [
        (E("1950-01-01 00:15:15"), E("1950-01-01 00:15:15") + D("5 min 15 s"), np.array([0.0, 0.0, 10.0]) / 1e3, n.index("Earth")),
        (E("1950-01-01 00:43:10"), E("1950-01-01 00:43:10") + D("6 min 30 s"), np.array([9.97, -2.31, 0.3]) / 1e3, n.index("Sun")),
        (E("1950-02-28 04:12:25"), E("1950-02-28 04:12:25") + D("1 min"), np.array([0.51, -0.1, -6.53]) / 1e3, n.index("Mars")),
        (E("1950-07-27 15:44:05"), E("1950-07-27 15:44:05") + D("5 min 10 s"), np.array([-10.0, 0.0, 0.0]) / 1e3, -1),
    ]


@pytest.mark.parametrize("pow_mode", ["glibc", "correctly_rounded"])
def test_ships_match_oracle_knot_for_knot(pow_mode):
    """Default (glibc) mode: the engine against the oracle in ITS default mode, where the controller calls the platform
    libm's pow -- the reference as built on Linux.  Correctly-rounded mode: both sides on the engine's libm-independent pow."""
    s, eph, ora = build_ephemeris()
    t0 = s.epoch
    end = formats.parse_epoch("1950-08-20 00:00:00")
    burns = burns_for(s)
    rng = np.random.default_rng(11)
    n = 8
    states = np.tile(np.array(STATE), (n, 1))
    states[1:, :3] += rng.uniform(-10, 10, (n - 1, 3))
    states[1:, 3:] += rng.uniform(-0.01, 0.01, (n - 1, 3))
    timelines = [[(b[0], b[1], ee.ConstantThrust(b[2], b[3])) for b in burns] if i % 2 == 0 else [] for i in range(n)]
    glibc = pow_mode == "glibc"
    params = ee.default_adaptive_params(pow_mode=ee.POW_GLIBC if glibc else ee.POW_CORRECTLY_ROUNDED)
    ships = ee.SpacecraftPropagator.new(t0, states, params, timelines, eph)
    ships.step_to(end, max_steps=100000)
    info = ships.info()
    sol = ships.take_solution()
    pr = (60.0, sys.float_info.max, 1e-3, 1e-3, 1 / 5, 5 / 1, 9 / 10)
    # The controller's `err.powf(-1/7)` feeds back into every later step size, and the embedded error estimate cancels ~8
    # digits, so the whole accepted-step sequence is only reproducible if pow is reproduced to the last bit: every knot
    # (time, position, velocity) must be bit-identical.
    oracle.set_pow_mode(oracle.POW_LIBM if glibc else oracle.POW_PORTABLE)
    try:
        for i in range(n):
            o = oracle.Ship(ora, t0, states[i], pr, 1_000_000, burns if i % 2 == 0 else ())
            st, _ = o.step_to(end)
            kn = o.knots()
            assert info["status"][i] == st == 0
            got = sol[i].knots
            assert got.shape == kn.shape, (i, got.shape, kn.shape)
            assert np.array_equal(got.view(np.uint64), kn.view(np.uint64)), i
            oi = o.info()
            assert info["n_attempts"][i] == oi["n_attempts"] and info["rhs_evals"][i] == oi["rhs_evals"]
            assert info["time"][i] == oi["time"]
    finally:
        oracle.set_pow_mode(oracle.POW_LIBM)


def test_ship_leaving_the_ephemeris_reports_eval_failed():
    s, eph, ora = build_ephemeris()
    params = ee.default_adaptive_params()
    t0 = formats.parse_epoch("1951-02-20 00:00:00")
    ships = ee.SpacecraftPropagator.new(t0, np.array([STATE, STATE]), params, None, eph)
    ships.step_to(formats.parse_epoch("1952-01-01 00:00:00"), max_steps=5000)
    info = ships.info()
    assert list(info["status"]) == [4, 4]  # StepError::EvalFailed
    o = oracle.Ship(ora, t0, STATE, (60.0, sys.float_info.max, 1e-3, 1e-3, 0.2, 5.0, 0.9), 1_000_000)  # libm pow on both sides
    st, _ = o.step_to(formats.parse_epoch("1952-01-01 00:00:00"))
    assert st == 4
    assert info["n_knots"][0] == len(o.knots())


def test_ship_take_solution_restarts_spline():
    s, eph, _ = build_ephemeris()
    params = ee.default_adaptive_params()
    ships = ee.SpacecraftPropagator.new(s.epoch, np.array([STATE]), params, None, eph)
    a = ships.propagate(s.epoch + 5 * 86400.0, max_steps=4000)[0]
    b = ships.propagate(s.epoch + 10 * 86400.0, max_steps=4000)[0]
    assert a.start() == s.epoch and a.end() >= s.epoch + 5 * 86400.0
    assert b.start() == a.end() and np.array_equal(b.knots[0], a.knots[-1])


def test_propagate_raises_on_ship_errors_and_loops_past_the_launch_cap():
    """propagate() mirrors BoundedPropagator::propagate: a ship that leaves the ephemeris is an error, not a truncated
    spline; a small max_steps only means more launches."""
    s, eph, _ = build_ephemeris()
    params = ee.default_adaptive_params()
    ships = ee.SpacecraftPropagator.new(s.epoch, np.array([STATE]), params, None, eph)
    sol = ships.propagate(s.epoch + 3 * 86400.0, max_steps=16)[0]  # a launch is capped at 16 accepted steps: several launches
    assert sol.end() >= s.epoch + 3 * 86400.0 and len(sol.knots) > 40
    whole = ee.SpacecraftPropagator.new(s.epoch, np.array([STATE]), params, None, eph).propagate(s.epoch + 3 * 86400.0)[0]
    assert np.array_equal(whole.knots.view(np.uint64), sol.knots.view(np.uint64))
    late = ee.SpacecraftPropagator.new(formats.parse_epoch("1951-02-20 00:00:00"), np.array([STATE]), params, None, eph)
    with pytest.raises(ee.ShipStepError) as err:
        late.propagate(formats.parse_epoch("1952-01-01 00:00:00"), max_steps=5000)
    assert err.value.code == 4  # StepError::EvalFailed


@pytest.mark.parametrize("method", range(1, 8))
def test_every_adaptive_method_matches_oracle_knot_for_knot(method):
    """The other seven IntegrationMethods a flight plan can select (flight_plan.rs:175-184): ERK tableaux incl. the FSAL
    DormandPrince54, and the second-order ERKNG form Fine45.  Ships with and without burns, several launches (the FSAL
    slope has to survive from launch to launch), a start step large enough to be rejected: every knot bit-identical to the
    oracle, attempt and evaluation counts equal."""
    s, eph, ora = build_ephemeris()
    t0 = s.epoch
    end = formats.parse_epoch("1950-01-09 00:00:00")
    burns = burns_for(s)[:2]
    rng = np.random.default_rng(100 + method)
    n = 4
    states = np.tile(np.array(STATE), (n, 1))
    states[1:, :3] += rng.uniform(-10, 10, (n - 1, 3))
    states[1:, 3:] += rng.uniform(-0.01, 0.01, (n - 1, 3))
    timelines = [[(b[0], b[1], ee.ConstantThrust(b[2], b[3])) for b in burns] if i % 2 == 0 else [] for i in range(n)]
    tol, h0 = 1e-5, 900.0
    params = ee.default_adaptive_params(tol, tol, h_init=h0, method=method)
    ships = ee.SpacecraftPropagator.new(t0, states, params, timelines, eph)
    sol = ships.propagate(end, max_steps=700)  # several launches
    info = ships.info()
    pr = (h0, sys.float_info.max, tol, tol, 1 / 5, 5 / 1, 9 / 10)
    for i in range(n):
        o = oracle.Ship(ora, t0, states[i], pr, 1_000_000, burns if i % 2 == 0 else (), method=method)
        st, _ = o.step_to(end)
        kn = o.knots()
        assert st == 0 and sol[i].knots.shape == kn.shape, (i, sol[i].knots.shape, kn.shape)
        assert np.array_equal(sol[i].knots.view(np.uint64), kn.view(np.uint64)), i
        oi = o.info()
        assert info["n_attempts"][i] == oi["n_attempts"] and info["rhs_evals"][i] == oi["rhs_evals"], i
        if i % 2 == 1:  # no burns: the attempt counter is never reset, so attempts > accepted steps means rejections happened
            assert oi["n_attempts"] > len(kn) - 1


def _analytics_equal(got, exp):
    (gtr, gap), (etr, eap) = got, exp
    assert len(gtr) == len(etr) and len(gap) == len(eap), (len(gtr), len(etr), len(gap), len(eap))
    for a, b in zip(gtr, etr):
        assert a[1] == b[1] and np.float64(a[0]).view(np.uint64) == np.float64(b[0]).view(np.uint64), (a, b)
    for a, b in zip(gap, eap):
        assert a[2:] == b[2:], (a, b)
        assert np.float64(a[0]).view(np.uint64) == np.float64(b[0]).view(np.uint64), (a, b)
        assert np.float64(a[1]).view(np.uint64) == np.float64(b[1]).view(np.uint64), (a, b)


def test_soi_transitions_and_apsides_event_for_event():
    """SpacecraftSolout on the device (dynamics/spacecraft.rs:536-586): the reference's Mars-transfer scenario (Earth ->
    Sun -> Mars with four burns) plus perturbed coasting copies.  Every SOI transition (time, body) and every apsis (time,
    distance, body, kind) equals the oracle's bit for bit, the knots are unchanged by the analytics, several launches, and
    take_solution() starts the next solution from soi_at(now)."""
    s, eph, ora = build_ephemeris()
    radii = formats.soi_radii(s)
    t0 = s.epoch
    end = formats.parse_epoch("1950-08-20 00:00:00")
    burns = burns_for(s)
    rng = np.random.default_rng(5)
    n = 6
    states = np.tile(np.array(STATE), (n, 1))
    states[1:, :3] += rng.uniform(-10, 10, (n - 1, 3))
    states[1:, 3:] += rng.uniform(-0.01, 0.01, (n - 1, 3))
    timelines = [[(b[0], b[1], ee.ConstantThrust(b[2], b[3])) for b in burns] if i % 2 == 0 else [] for i in range(n)]
    params = ee.default_adaptive_params()
    ships = ee.SpacecraftPropagator.new(t0, states, params, timelines, eph)
    ships.enable_analytics(radii)
    start = ships.analytics()
    earth = s.names.index("Earth")
    assert all(tr == [(t0, earth)] and ap == [] for tr, ap in start)
    while True:
        ships.step_to(end, max_steps=900)
        info = ships.info()
        assert np.all(info["status"] == 0)
        if np.all(info["time"] >= end):
            break
    got = ships.analytics()
    sol = ships.take_solution()
    plain = ee.SpacecraftPropagator.new(t0, states, params, timelines, eph).propagate(end, max_steps=100000)
    pr = (60.0, sys.float_info.max, 1e-3, 1e-3, 1 / 5, 5 / 1, 9 / 10)
    seen_bodies = set()
    for i in range(n):
        o = oracle.Ship(ora, t0, states[i], pr, 1_000_000, burns if i % 2 == 0 else ())
        o.enable_analytics(radii)
        st, _ = o.step_to(end)
        assert st == 0
        assert np.array_equal(sol[i].knots.view(np.uint64), o.knots().view(np.uint64)), i
        assert np.array_equal(sol[i].knots.view(np.uint64), plain[i].knots.view(np.uint64)), i
        _analytics_equal(got[i], o.analytics())
        seen_bodies.update(b for _, b in got[i][0])
    assert {s.names.index("Earth"), s.names.index("Sun"), s.names.index("Mars")} <= seen_bodies
    assert max(len(ap) for _, ap in got) > 300  # the captured ship circles Mars: the apsis lists had to grow on the device
    # the new solution starts in the sphere the ship is in now
    after = ships.analytics()
    now = ships.info()["time"]
    assert all(ap == [] and len(tr) <= 1 for tr, ap in after)
    assert after[1][0] == [(float(now[1]), earth)]  # the coasting copies never left the parking orbit


def test_relative_trajectory_sampling_bit_exact():
    """RelativeTrajectory::state_vector batched on the device (trajectory.rs:315-335; the plotter's per-point evaluation,
    ui/world/plot.rs:326-334): body w.r.t. body, body w.r.t. nothing, a ship's Hermite spline w.r.t. a body -- at knots,
    between knots, before the start and after the end; bit-identical to the oracle, same None cases."""
    s, eph, ora = build_ephemeris()
    earth, moon, sun = s.names.index("Earth"), s.names.index("Moon"), s.names.index("Sun")
    rng = np.random.default_rng(8)
    t_lo, t_hi = s.epoch, formats.parse_epoch("1951-03-01 00:00:00")
    times = np.concatenate([rng.uniform(t_lo, t_hi, 500), [t_lo, t_lo + 6 * 3600.0 * 12 * 8], t_lo - rng.uniform(1.0, 1e6, 10),
                            t_hi + 86400.0 * rng.uniform(400.0, 900.0, 10)])
    for body, ref in ((moon, earth), (earth, sun), (moon, None)):
        pos, vel, ok = eph.evaluate_relative(body, ref, times)
        for k, t in enumerate(times):
            r = oracle.relative_state_vector(ora, t, body=body, reference=ref)
            assert (r is not None) == bool(ok[k]), (body, ref, t)
            if r is not None:
                assert np.array_equal(pos[k].view(np.uint64), r[0].view(np.uint64))
                assert np.array_equal(vel[k].view(np.uint64), r[1].view(np.uint64))
        assert ok.sum() > 400 and (~ok).sum() >= 20
    ships = ee.SpacecraftPropagator.new(s.epoch, np.array([STATE, STATE]), ee.default_adaptive_params(), None, eph)
    end = s.epoch + 2 * 86400.0
    ships.step_to(end, max_steps=5000)
    nk = int(ships.info()["n_knots"][1])
    probe = np.concatenate([rng.uniform(s.epoch, end, 400), [s.epoch], s.epoch - rng.uniform(0.5, 1e4, 5), end + rng.uniform(7200.0, 1e5, 5)])
    pos, vel, ok = ships.evaluate_relative(1, earth, probe)
    pos0, vel0, ok0 = ships.evaluate_relative(1, None, probe)
    knots = ships.take_solution()[1].knots
    assert len(knots) == nk
    probe2 = np.concatenate([probe, knots[5:9, 0]])  # exact knot times return the knot itself
    ships2 = ee.SpacecraftPropagator.new(s.epoch, np.array([STATE, STATE]), ee.default_adaptive_params(), None, eph)
    ships2.step_to(end, max_steps=5000)
    pos, vel, ok = ships2.evaluate_relative(1, earth, probe2)
    pos0, vel0, ok0 = ships2.evaluate_relative(1, None, probe2)
    for k, t in enumerate(probe2):
        r = oracle.relative_state_vector(ora, t, reference=earth, knots=knots)
        r0 = oracle.relative_state_vector(ora, t, reference=None, knots=knots)
        assert (r is not None) == bool(ok[k]) and (r0 is not None) == bool(ok0[k]), t
        if r is not None:
            assert np.array_equal(pos[k].view(np.uint64), r[0].view(np.uint64)) and np.array_equal(vel[k].view(np.uint64), r[1].view(np.uint64))
        if r0 is not None:
            assert np.array_equal(pos0[k].view(np.uint64), r0[0].view(np.uint64)) and np.array_equal(vel0[k].view(np.uint64), r0[1].view(np.uint64))
    for k in range(4):  # at a knot: exactly the knot's state
        assert np.array_equal(pos0[len(probe) + k], knots[5 + k, 1:4]) and np.array_equal(vel0[len(probe) + k], knots[5 + k, 4:7])
    assert ok.sum() > 300 and (~ok).sum() >= 10


@pytest.mark.parametrize("nb", [3, 43, 140])
def test_ship_kernel_with_few_and_with_more_than_32_bodies(nb):
    """The ship kernel maps one lane to one body: systems that fill a fraction of a warp (3 bodies), systems that need
    a second pass over the lanes (43 bodies: the 10 real ones plus 33 light copies of them on the same splines, so the second
    group of lanes, its polynomial cache and its part of the ordered sum all carry weight), and systems whose position and
    polynomial caches no longer fit in shared memory and live in global memory instead (140 bodies).  Knots and analytics
    bit-identical to the oracle for an ERK, the FSAL ERK and the ERKNG method."""
    s, eph10, _ = build_ephemeris()
    mus10, spl10 = eph10.splines()
    rng = np.random.default_rng(nb)
    if nb <= 10:
        idx = [s.names.index(n) for n in ("Sun", "Earth", "Moon")][:nb]
        mus = [mus10[i] for i in idx]
    else:
        idx = list(range(10)) + [int(rng.integers(0, 10)) for _ in range(nb - 10)]
        mus = list(mus10) + [mus10[i] * 10.0 ** rng.uniform(-4.0, -2.0) for i in idx[10:]]
    spl = [spl10[i] for i in idx]
    eph = ee.Ephemeris.from_splines(mus, spl)
    ora = oracle.Ephem(mus, [(x.start, x.interval, x.polynomials) for x in spl])
    radii = np.full(nb, 1.0e5)
    earth = idx.index(s.names.index("Earth"))
    radii[idx.index(s.names.index("Sun"))] = np.inf
    radii[earth] = 9.2e5
    t0 = s.epoch
    end = t0 + 4 * 86400.0
    states = np.tile(np.array(STATE), (3, 1))
    states[1:, :3] += rng.uniform(-10, 10, (2, 3))
    for method in (ee.VERNER87, ee.DORMAND_PRINCE54, ee.FINE45):
        params = ee.default_adaptive_params(1e-4, 1e-4, method=method)
        ships = ee.SpacecraftPropagator.new(t0, states, params, None, eph)
        ships.enable_analytics(radii)
        while True:
            ships.step_to(end, max_steps=500)
            info = ships.info()
            assert np.all(info["status"] == 0)
            if np.all(info["time"] >= end):
                break
        got = ships.analytics()
        sol = ships.take_solution()
        pr = (60.0, sys.float_info.max, 1e-4, 1e-4, 1 / 5, 5 / 1, 9 / 10)
        for i in range(3):
            o = oracle.Ship(ora, t0, states[i], pr, 1_000_000, (), method=method)
            o.enable_analytics(radii)
            st, _ = o.step_to(end)
            assert st == 0
            assert np.array_equal(sol[i].knots.view(np.uint64), o.knots().view(np.uint64)), (nb, method, i)
            _analytics_equal(got[i], o.analytics())
            oi = o.info()
            assert info["n_attempts"][i] == oi["n_attempts"] and info["rhs_evals"][i] == oi["rhs_evals"]
        assert max(len(ap) for _, ap in got) >= 50  # ~60 revolutions in four days

"""CPU tests of the oracle's other adaptive methods (flight_plan.rs:175-184) and of SpacecraftSolout's analytics
(dynamics/spacecraft.rs:91-161, :536-586): behavioural pins of the restatement, no GPU."""
import sys

import numpy as np
import pytest

import oracle
from ephemeris_explorer_b200 import formats
from helpers import load_system
from test_oracle_cpu import _two_year_ephemeris

STATE = [-27204249.668775786, 132947582.43848978, 57641619.74241204, -22.253599106181895, -5.189518219791726, -2.2515617105336263]
NAMES = ("Verner87", "CashKarp45", "DormandPrince54", "DormandPrince87", "Fehlberg45", "Tsitouras75", "Verner98", "Fine45")
STAGES = (13, 6, 7, 13, 6, 9, 16, 7)
FSAL = (False, False, True, False, False, False, False, True)


def _params(tol):
    return (60.0, sys.float_info.max, tol, tol, 1 / 5, 5 / 1, 9 / 10)


@pytest.fixture(scope="module")
def eph10():
    return _two_year_ephemeris()


def test_every_method_reaches_the_same_place(eph10):
    """Three days of coast from the parking orbit at a tight tolerance: the eight methods are different discretisations of
    the same ODE, so their end points agree to well under a kilometre (measured 0.02-0.4 km after ~45 revolutions, shrinking
    with the tolerance) while their knot sequences differ; the evaluation counts follow the stage counts (an FSAL method
    evaluates its first stage only on the very first attempt)."""
    s, eph = eph10
    end = formats.parse_epoch("1950-01-04 00:00:00")
    probe = formats.parse_epoch("1950-01-03 12:00:00")
    ref = None
    seen = set()
    for m, name in enumerate(NAMES):
        ship = oracle.Ship(eph, s.epoch, STATE, _params(1e-6), 10_000_000, method=m)
        st, _ = ship.step_to(end)
        assert st == 0, name
        kn = ship.knots()
        info = ship.info()
        if FSAL[m]:
            # stage 0 is evaluated only while i == 0: the first attempt and any re-attempts after its rejection
            base = info["n_attempts"] * (STAGES[m] - 1)
            rejected = info["n_attempts"] - (len(kn) - 1)
            assert base + 1 <= info["rhs_evals"] <= base + 1 + rejected, name
        else:
            assert info["rhs_evals"] == info["n_attempts"] * STAGES[m], name
        pos = oracle.spline_position_from_knots(kn, probe)
        if ref is None:
            ref = pos
        assert np.linalg.norm(pos - ref) < 1.0, (name, np.linalg.norm(pos - ref))  # km
        seen.add(len(kn))
    assert len(seen) >= 6  # different methods take different numbers of steps


def test_higher_order_methods_take_fewer_steps(eph10):
    s, eph = eph10
    end = formats.parse_epoch("1950-01-03 00:00:00")
    steps = {}
    for m, name in enumerate(NAMES):
        ship = oracle.Ship(eph, s.epoch, STATE, _params(1e-6), 10_000_000, method=m)
        assert ship.step_to(end)[0] == 0
        steps[name] = len(ship.knots()) - 1
    assert steps["Verner98"] < steps["Verner87"] < steps["DormandPrince54"]
    assert steps["DormandPrince87"] < steps["CashKarp45"]
    assert steps["Tsitouras75"] < steps["Fehlberg45"]


def test_fsal_rejections_restore_the_saved_slope(eph10):
    """With a loose start step the first attempts are rejected; an FSAL method must then re-use the slope stored before the
    step (PreviousStep/undo_step), i.e. give exactly the knots of a run that starts from a step small enough to be accepted
    ... it cannot in general, so the check is self-consistency: stepping one step at a time equals one step_to call."""
    s, eph = eph10
    end = formats.parse_epoch("1950-01-02 00:00:00")
    for m in (2, 7):
        a = oracle.Ship(eph, s.epoch, STATE, (3600.0,) + _params(1e-4)[1:], 1_000_000, method=m)
        b = oracle.Ship(eph, s.epoch, STATE, (3600.0,) + _params(1e-4)[1:], 1_000_000, method=m)
        assert a.step_to(end)[0] == 0
        while b.knots()[-1, 0] < end:
            assert b.step(1) == 0
        assert np.array_equal(a.knots(), b.knots())
        assert a.info()["n_attempts"] > len(a.knots()) - 1  # some attempts were rejected


def test_soi_radii_follow_the_loader_rule():
    s = load_system("full_solar_system_2433282.5")
    r = formats.soi_radii(s)
    i = s.names.index
    assert np.isinf(r[i("Sun")])
    assert 0.9e6 < r[i("Earth")] < 0.95e6      # the textbook ~0.92e6 km
    assert 5.2e5 < r[i("Mars")] < 6.6e5       # instantaneous distance, not the semi-major axis
    assert 6.0e4 < r[i("Moon")] < 7.0e4         # child of Earth, not of the Sun
    assert np.all(r > 0)


def test_mars_transfer_ship_transitions_and_apsides():
    """The reference's own scenario (spacecraft_propagation.rs:401-483) through SpacecraftSolout: the ship starts inside
    Earth's sphere, leaves it for the Sun's, enters Mars' -- in that order; inside Earth's sphere before the departure burn
    the apsides alternate and bracket the parking orbit; every recorded distance equals the trajectory's own."""
    s = load_system("full_solar_system_2433282.5")
    nb = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
    nb.set_solout(s.dt, s.sample_period, s.degree)
    end_eph = s.epoch + 400 * 86400.0
    while nb.solution_time() < end_eph:
        assert nb.step(2000) == 0
    eph = oracle.Ephem(s.mu, nb.splines())
    from helpers import SYSTEMS
    ship = formats.load_ship(SYSTEMS / "full_solar_system_2433282.5.json", s.names, name="Mars Transfer Ship")
    burns = [(b.start, b.end, b.acceleration, b.reference) for b in ship.burns]
    o = oracle.Ship(eph, ship.start, np.concatenate([ship.position, ship.velocity]), _params(ship.tolerance), 1_000_000, burns)
    radii = formats.soi_radii(s)
    o.enable_analytics(radii)
    tr0, ap0 = o.analytics()
    assert tr0 == [(ship.start, s.names.index("Earth"))] and ap0 == []
    st, _ = o.step_to(formats.parse_epoch("1950-09-01 00:00:00"))
    assert st == 0
    tr, ap = o.analytics()
    kn = o.knots()
    bodies = [s.names[b] for _, b in tr]
    assert bodies[0] == "Earth" and "Sun" in bodies and "Mars" in bodies
    assert bodies.index("Sun") < bodies.index("Mars")
    assert all(t1 > t0 for (t0, _), (t1, _) in zip(tr, tr[1:]))
    assert all(b0 != b1 for (_, b0), (_, b1) in zip(tr, tr[1:]))
    # leaving Earth's sphere: the recorded time is where the distance to Earth equals the sphere radius (bisected to 1 ms)
    t_leave = next(t for t, b in tr if s.names[b] == "Sun")
    d = np.linalg.norm(oracle.spline_position_from_knots(kn, t_leave) - eph.position(s.names.index("Earth"), t_leave))
    assert abs(d - radii[s.names.index("Earth")]) < 1.0
    assert len(ap) >= 2 and all(a1[0] > a0[0] for a0, a1 in zip(ap, ap[1:]))
    earth = [a for a in ap if s.names[a[2]] == "Earth"]
    assert earth and earth[0][3] in (0, 1)
    for t, dist, b, kind in ap[:20]:
        dd = np.linalg.norm(oracle.spline_position_from_knots(kn, t) - eph.position(b, t))
        assert abs(dd - dist) <= 1e-9 * dist
    for (t0, d0, b0, k0), (t1, d1, b1, k1) in zip(earth, earth[1:]):
        if k0 == 0 and k1 == 1:
            assert d1 >= d0  # apoapsis above the preceding periapsis
    # a periapsis is a local minimum of the distance along the trajectory
    t, dist, b, kind = next(a for a in ap if a[3] == 0)
    for dt in (-5.0, 5.0):
        p = oracle.spline_position_from_knots(kn, t + dt)
        if p is not None:
            assert np.linalg.norm(p - eph.position(b, t + dt)) >= dist - 1e-6


def test_relative_state_vector_semantics(eph10):
    """RelativeTrajectory::state_vector (trajectory.rs:315-335): difference of the two state vectors, None when either side
    is outside its span; CubicHermiteSpline::state_vector returns a knot verbatim at a knot time."""
    s, eph = eph10
    earth, moon = s.names.index("Earth"), s.names.index("Moon")
    t = s.epoch + 86400.0 * 10.25
    r = oracle.relative_state_vector(eph, t, body=moon, reference=earth)
    pm, pe = eph.position(moon, t), eph.position(earth, t)
    assert np.array_equal(r[0], pm - pe) and 3.5e5 < np.linalg.norm(r[0]) < 4.1e5
    assert np.array_equal(oracle.relative_state_vector(eph, t, body=moon)[0], pm)
    assert oracle.relative_state_vector(eph, s.epoch - 1.0, body=moon, reference=earth) is None
    ship = oracle.Ship(eph, s.epoch, STATE, _params(1e-3), 1_000_000)
    assert ship.step_to(s.epoch + 86400.0)[0] == 0
    kn = ship.knots()
    at_knot = oracle.relative_state_vector(eph, kn[7, 0], knots=kn)
    assert np.array_equal(at_knot[0], kn[7, 1:4]) and np.array_equal(at_knot[1], kn[7, 4:7])
    mid = 0.5 * (kn[7, 0] + kn[8, 0])
    between = oracle.relative_state_vector(eph, mid, reference=earth, knots=kn)
    assert np.array_equal(between[0], oracle.hermite_eval(kn[7], kn[8], mid)[0] - eph.position(earth, mid))
    assert oracle.relative_state_vector(eph, kn[0, 0] - 1.0, knots=kn) is None
    assert oracle.relative_state_vector(eph, kn[-1, 0] + 1.0, knots=kn) is None

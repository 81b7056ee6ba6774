"""Host-side container operations of the trajectory mirrors: what the Prediction Planner does with every solution it takes
(`PredictionTarget::merge`, ephemeris_explorer/src/dynamics/celestial.rs:194-235; `SpacecraftPropagator::join`,
ephemeris/src/propagators/spacecraft.rs:558-561).  Checked on solutions produced by the CPU oracle: merging the pieces taken
at arbitrary ticks must rebuild, bit for bit, the spline a single take at the end returns."""
import numpy as np
import pytest

import oracle
from helpers import load_system, bits_equal
import ephemeris_explorer_b200 as ee


def as_splines(sol):
    return [ee.UniformSpline(st, iv, list(polys)) for st, iv, polys in sol]


def same_spline(a, b):
    return (a.start == b.start and a.interval == b.interval and len(a.polynomials) == len(b.polynomials)
            and all(bits_equal(x, y) for x, y in zip(a.polynomials, b.polynomials)))


@pytest.mark.parametrize("system,backward", [("sun_earth_moon_2433282.5", False), ("sun_earth_moon_2433282.5", True),
                                             ("full_solar_system_2433282.5", False), ("full_solar_system_2433282.5", True)])
def test_merging_the_pieces_taken_at_every_tick_rebuilds_the_whole_spline(system, backward):
    s = load_system(system)
    h = -s.dt if backward else s.dt
    total = 8 * int(s.count.max()) * 3 + 5
    whole = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, h)
    whole.set_solout(s.dt, s.sample_period, s.degree)
    assert whole.step(total) == 0
    want = as_splines(whole.take_solution())

    ticked = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, h)
    ticked.set_solout(s.dt, s.sample_period, s.degree)
    rng = np.random.default_rng(11)
    world = None
    done = 0
    while done < total:
        k = min(int(rng.integers(1, 4 * int(s.count.max()))), total - done)
        assert ticked.step(k) == 0
        done += k
        piece = as_splines(ticked.take_solution())
        if world is None:
            world = piece  # PredictionTarget::overwrite
            continue
        for w, p in zip(world, piece):
            if backward:
                w.merge_backward(p)
            else:
                w.merge_forward(p)
    assert len(world) == len(want)
    for w, x in zip(world, want):
        assert same_spline(w, x)
        assert w.contains(w.start) and w.contains(w.end()) and not w.contains(w.end() + w.interval)


def make_spline(n=6, start=100.0, interval=8.0):
    return ee.UniformSpline(start, interval, [np.full((3, 3), float(i)) for i in range(n)])


def test_uniform_spline_index_rules():
    """trajectory.rs:571-617: get_index is half-open [start, end), the exclusive form maps a knot to the polynomial before
    it (so that end() can be evaluated) and saturates at 0."""
    u = make_spline()
    assert u.span() == 48.0 and u.end() == 148.0 and u.segment_count() == 6
    assert u.get_index(100.0) == 0 and u.get_index(107.999) == 0 and u.get_index(108.0) == 1 and u.get_index(147.9) == 5
    assert u.get_index(148.0) is None and u.get_index(99.999) is None
    assert u.get_index_exclusive(100.0) == 0 and u.get_index_exclusive(108.0) == 0 and u.get_index_exclusive(108.001) == 1
    assert u.get_index_exclusive(148.0) == 5 and u.get_index_exclusive(148.001) is None
    assert u.contains(100.0) and u.contains(148.0) and not u.contains(99.0)
    # -0.0 has its sign bit set: Duration::is_negative (duration.rs:83-85) says it is negative
    z = ee.UniformSpline(0.0, 1.0, [np.zeros((1, 3))])
    assert z.get_index(-0.0) is None and z.get_index(0.0) == 0  # -0.0 - 0.0 = -0.0: negative by the reference's rule
    assert (ee.propagators._sign_negative(-0.0), ee.propagators._sign_negative(0.0)) == (True, False)


def test_uniform_spline_container_operations():
    u = make_spline()
    u.clear_after(116.0)  # index 2 -> keeps polynomials 0, 1
    assert [p[0, 0] for p in u.polynomials] == [0.0, 1.0] and u.start == 100.0
    u.clear_after(500.0)  # outside: nothing happens
    assert u.segment_count() == 2
    u = make_spline()
    u.clear_before(116.0)  # exclusive index of 124 = 2: drops polynomials 0, 1
    assert [p[0, 0] for p in u.polynomials] == [2.0, 3.0, 4.0, 5.0] and u.start == 116.0
    u.clear_before(50.0)
    assert u.segment_count() == 4
    u.push_front(np.full((3, 3), 9.0))
    assert u.start == 108.0 and u.polynomials[0][0, 0] == 9.0
    u.push_back(np.full((3, 3), 7.0))
    assert u.end() == 108.0 + 6 * 8.0
    a, b = make_spline(2, 100.0), make_spline(3, 116.0)
    a.append(b)
    assert a.segment_count() == 5 and a.end() == 140.0
    c = make_spline(2, 84.0)
    a.prepend(c)
    assert a.start == 84.0 and a.segment_count() == 7 and [p[0, 0] for p in a.polynomials[:3]] == [0.0, 1.0, 0.0]
    with pytest.raises(AssertionError):
        a.append(make_spline(1, 999.0))
    with pytest.raises(AssertionError):
        a.prepend(make_spline(1, 0.0))
    w = make_spline().between(108.0, 125.0)  # exclusive indices 0 and 3
    assert w.start == 100.0 and [p[0, 0] for p in w.polynomials] == [0.0, 1.0, 2.0, 3.0]
    assert make_spline().between(0.0, 125.0) is None and ee.UniformSpline(0.0, 1.0, []).between(0.0, 0.0) is None


def test_cubic_hermite_spline_join_is_the_planners_merge():
    k = np.zeros((5, 7))
    k[:, 0] = [0.0, 1.0, 2.5, 4.0, 7.0]
    k[:, 1] = np.arange(5.0)
    a = ee.CubicHermiteSpline(k.copy())
    assert a.start() == 0.0 and a.end() == 7.0 and a.segment_count() == 4
    assert a.binary_search(2.5) == (True, 2) and a.binary_search(3.0) == (False, 3) and a.binary_search(-1.0) == (False, 0)
    assert a.get(4.0)[0] == 3.0 and a.get(4.1) is None
    r = np.zeros((3, 7))
    r[:, 0] = [2.5, 3.0, 9.0]
    r[:, 1] = [20.0, 21.0, 22.0]
    a.join(ee.CubicHermiteSpline(r))  # clear_after(2.5) keeps the knots strictly before 2.5
    assert list(a.knots[:, 0]) == [0.0, 1.0, 2.5, 3.0, 9.0] and list(a.knots[:, 1]) == [0.0, 1.0, 20.0, 21.0, 22.0]
    empty = ee.CubicHermiteSpline(np.zeros((0, 7)))
    assert empty.segment_count() == 0 and empty.start() == -np.inf and empty.end() == np.inf


def test_ship_solutions_joined_at_every_launch_equal_one_solution():
    """The oracle's ship advanced in pieces: joining the knot lists as the Planner does gives the list of one long run."""
    s = load_system("simple_solar_system_2433282.5")
    from helpers import SHIP_TEST_PERIOD_HOURS, SHIP_TEST_DEGREES
    periods = np.array(SHIP_TEST_PERIOD_HOURS) * 3600.0
    nb = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
    nb.set_solout(s.dt, periods, np.array(SHIP_TEST_DEGREES, dtype=np.int32))
    t_end = s.epoch + 20 * 86400.0
    while nb.solution_time() < t_end:
        assert nb.step(256) == 0
    eph = oracle.Ephem(s.mu, nb.take_solution())
    earth = s.names.index("Earth")
    state = np.concatenate([s.position[earth] + np.array([7000.0, 0.0, 0.0]), s.velocity[earth] + np.array([0.0, 7.5, 0.0])])
    params = [60.0, np.inf, 1e-3, 1e-3, 0.2, 5.0, 0.9]
    whole = oracle.Ship(eph, s.epoch, state, params, 1_000_000)
    whole.step_to(s.epoch + 10 * 86400.0)
    want = whole.knots()
    pieces = oracle.Ship(eph, s.epoch, state, params, 1_000_000)
    world = None
    taken = 0
    for n in (7, 1, 50, 300, 10**9):
        pieces.step_to(s.epoch + 10 * 86400.0, max_steps=n)
        kn = pieces.knots()
        # a solution taken mid-way starts at the last knot of the previous one (CubicHermiteSplineSolout::new_solution)
        piece = ee.CubicHermiteSpline(kn[max(taken - 1, 0):].copy())
        taken = len(kn)
        if world is None:
            world = piece
        else:
            world.join(piece)
    assert bits_equal(world.knots, want)

"""The CUDA path against the COMMITTED golden fixtures (tests/golden/oracle_outputs.json), through the C ABI.

The other GPU tests compare with the oracle built on the box; these compare with the hex-float outputs recorded in the
repository (tests/golden/make_oracle_outputs.py wrote them; tests/test_oracle_cpu.py keeps the oracle equal to them), so a
parity result does not depend on the oracle binary of the day.  Bit-exact: parity mode.  The file sorts last on purpose.
"""
import json

import numpy as np
import pytest

from helpers import ROOT, load_system
import ephemeris_explorer_b200 as ee

pytestmark = pytest.mark.gpu

GOLDEN = json.loads((ROOT / "tests" / "golden" / "oracle_outputs.json").read_text())


def unhex(lst, shape=None):
    a = np.array([float.fromhex(x) for x in lst], dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


def same_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    return a.shape == b.shape and bool(np.all(a.view(np.uint64) == b.view(np.uint64)))


def test_pair_kernel_against_the_recorded_accelerations_both_readings():
    g = GOLDEN["gravity_eval_7"]
    pos, mu = unhex(g["pos"], (7, 3)), unhex(g["mu"])
    try:
        assert same_bits(ee.gravity_eval(pos, mu, ee.MODE_PARITY), unhex(g["acc"]))
        ee.set_pair_variant(1)
        assert same_bits(ee.gravity_eval(pos, mu, ee.MODE_PARITY), unhex(g["acc_variant1"]))
    finally:
        ee.set_pair_variant(0)


@pytest.mark.parametrize("method,steps", [(12, 40), (13, 40), (14, 10)])
def test_fixed_step_methods_against_the_recorded_states(method, steps):
    """QuinlanTremaine12 / Stormer13 / BlanesMoan14A on the 3-body system: time, positions and velocities after the recorded
    step count (start-up included) equal the recorded bits."""
    g = GOLDEN["sun_earth_moon_method%d_%dsteps" % (method, steps)]
    s = load_system("sun_earth_moon_2433282.5")
    prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY, method=method)
    prop.step(steps)
    t, pos, vel = prop.state()
    assert t == float.fromhex(g["t"])
    assert same_bits(pos, unhex(g["pos"])) and same_bits(vel, unhex(g["vel"]))


def test_backward_spline_solution_against_the_recorded_polynomials():
    g = GOLDEN["sun_earth_moon_backward_splines"]
    s = load_system("sun_earth_moon_2433282.5")
    prop = ee.NBodyPropagator.new(ee.Backward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY,
                                  solout=(s.dt, s.sample_period, s.degree))
    prop.step(8 * 12 * 2 + 5)
    assert prop.time() == float.fromhex(g["solution_time"])
    sol = prop.take_solution()
    assert len(sol) == len(g["bodies"])
    for u, rec in zip(sol, g["bodies"]):
        assert u.start == float.fromhex(rec["start"]) and u.interval == float.fromhex(rec["interval"])
        assert len(u.polynomials) == rec["n_poly"]
        if rec["n_poly"]:
            assert same_bits(u.polynomials[0], unhex(rec["first"])) and same_bits(u.polynomials[-1], unhex(rec["last"]))


def test_least_squares_fit_against_the_recorded_coefficients():
    g = GOLDEN["lsq_fit_degree6"]
    xs = unhex(g["xs"], (9, 3))
    got, n_coef = ee.lsq_fit(6, np.arange(9) / 8.0, xs)
    assert int(n_coef[0]) == g["n_coef"] and same_bits(got[0], unhex(g["coeffs"]))


def test_32_body_system_against_the_recorded_state_and_splines():
    """C2's system, 2 000 steps (12 start-up + 1 988 through the persistent small-system kernel) with the spline solout."""
    g = GOLDEN["full_solar_system_2000steps"]
    s = load_system("full_solar_system_2433282.5")
    prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY,
                                  solout=(s.dt, s.sample_period, s.degree))
    prop.step(2000)
    t, pos, vel = prop.state()
    assert t == float.fromhex(g["t"])
    assert same_bits(pos, unhex(g["pos"])) and same_bits(vel, unhex(g["vel"]))
    sol = prop.take_solution()
    assert [len(u.polynomials) for u in sol] == g["n_poly"]
    for u, rec in zip(sol, g["last_poly_first_coeff"]):
        if u.polynomials:
            assert same_bits(u.polynomials[-1][0], unhex(rec))


SHIP_STATE = [-27204249.668775786, 132947582.43848978, 57641619.74241204, -22.253599106181895, -5.189518219791726, -2.2515617105336263]


def ten_body_ephemeris(days):
    from helpers import SHIP_TEST_DEGREES, SHIP_TEST_PERIOD_HOURS
    s = load_system("simple_solar_system_2433282.5")
    h = 6 * 3600.0
    periods = np.array(SHIP_TEST_PERIOD_HOURS) * 3600.0
    prop = ee.NBodyPropagator.new(ee.Forward(h), s.epoch, s.position, s.velocity, s.mu, solout=(h, periods, SHIP_TEST_DEGREES))
    prop.step_to(s.epoch + days * 86400.0)
    return s, prop.take_solution_ephemeris()


def test_ship_with_a_burn_against_the_recorded_knots():
    """A ship through the 10-body spline ephemeris built on the device, one burn in Earth's TNB frame, Verner 8(7) with the
    libm-independent pow (what the fixture was recorded with): knot count, attempts, right-hand-side evaluations, knot 100 and
    the last knot equal the recorded bits."""
    g = GOLDEN["ship_10body_20days"]
    s, eph = ten_body_ephemeris(40)
    burn = (s.epoch + 915.0, s.epoch + 915.0 + 315.0, ee.ConstantThrust(np.array([0.0, 0.0, 10.0]) / 1e3, s.names.index("Earth")))
    params = ee.default_adaptive_params(pow_mode=ee.POW_CORRECTLY_ROUNDED)
    ships = ee.SpacecraftPropagator.new(s.epoch, np.array([SHIP_STATE]), params, [[burn]], eph)
    ships.step_to(s.epoch + 20 * 86400.0, max_steps=100000)
    info = ships.info()
    kn = ships.take_solution()[0].knots
    assert int(info["status"][0]) == g["status"] == 0
    assert len(kn) == g["n_knots"]
    assert int(info["n_attempts"][0]) == g["n_attempts"] and int(info["rhs_evals"][0]) == g["rhs_evals"]
    assert same_bits(kn[100], unhex(g["knot_100"])) and same_bits(kn[-1], unhex(g["last_knot"]))

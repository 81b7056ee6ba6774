"""Parity on the BASELINE.json configs themselves (C2, C4, C5), through the C ABI, in the driver-run GPU suite.

C4 (65 536 bodies) is too big for the CPU oracle to step (one evaluation = 2.1e9 pairs), so the check is transitive: the
parity-mode kernel is bit-exact to the oracle at any N (test_nbody_gpu.py), hence a parity-mode handle at 65 536 is a
valid stand-in for the oracle there; the throughput kernels must stay within the north-star tolerance of it.
"""
import os
import sys

import numpy as np
import pytest

import oracle
from helpers import bits_equal, load_system, rel_err
import ephemeris_explorer_b200 as ee
from ephemeris_explorer_b200 import formats

pytestmark = pytest.mark.gpu

H = 2.0 ** -10


class dev_env:
    """Developer switches of the engine (EE_SYM_VARIANT, EE_SYM_RANGE, ...) only act when EE_DEV_AIDS=1."""

    def __init__(self, **kv):
        self.kv = dict(kv, EE_DEV_AIDS="1")

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_c4_headline_config_throughput_vs_parity_kernel_65536():
    """12 start-up + 8 steady-state steps at the bench size: the pair-symmetric kernel + reduce + QT12 epilogue (what
    bench.py times) against the bit-exact parity kernel stepping the same system."""
    n = 65536
    p0, v0, mu = ee.synthetic.plummer(n)
    fast = ee.NBodyPropagator.new(ee.Forward(H), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    ref = ee.NBodyPropagator.new(ee.Forward(H), 0.0, p0, v0, mu, mode=ee.MODE_PARITY)
    for steps in (12, 8):  # checked at the start-up boundary and after 8 multistep steps
        fast.step(steps)
        ref.step(steps)
        t, pos, vel = fast.state()
        rt, rpos, rvel = ref.state()
        assert t == rt
        assert rel_err(pos, rpos) <= 1e-12  # north-star tolerance
        assert rel_err(vel, rvel) <= 1e-10
    assert fast.step_count() == 20
    # the same 20 steps again give the same bits: dynamic item scheduling does not leak into the sums
    again = ee.NBodyPropagator.new(ee.Forward(H), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    again.step(20)
    assert bits_equal(again.state()[1], pos)


@pytest.mark.parametrize("variant", ["4,256,2,16", "4,128,3,16"])
def test_c4_rank_shares_of_every_kernel_variant_add_up_65536(variant):
    """EE_SYM_RANGE=a/b makes a one-GPU engine evaluate rank a's share of a b-way sharded run (partial sums of its pair
    units only): the 8 shares must add up to the full evaluation, and the full evaluation must match the parity kernel."""
    n = 65536
    p0, _, mu = ee.synthetic.plummer(n, seed=9)
    exact = ee.gravity_eval(p0, mu, ee.MODE_PARITY)
    with dev_env(EE_SYM_VARIANT=variant):
        full = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
        assert rel_err(full, exact) < 1e-12
        parts = np.zeros_like(full)
        for a in range(8):
            with dev_env(EE_SYM_RANGE="%d/8" % a):
                share = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
            assert not bits_equal(share, full)
            parts += share
    assert rel_err(parts, exact) < 1e-12


def test_c2_full_solar_system_million_steps_bitwise():
    """BASELINE.json configs[1]: 32 bodies, dt = 600 s, 10^6 steps, spline solout on -- positions, velocities and every
    fitted polynomial bit-identical to the oracle at 10^3 / 10^4 / 10^5 / 10^6 steps."""
    s = load_system("full_solar_system_2433282.5")
    prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY,
                                  solout=(s.dt, s.sample_period, s.degree))
    ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
    ref.set_solout(s.dt, s.sample_period, s.degree)
    done = 0
    for mark in (1_000, 10_000, 100_000, 1_000_000):
        prop.step(mark - done)
        assert ref.step(mark - done) == 0
        done = mark
        t, pos, vel = prop.state()
        rt, rpos, rvel, _ = ref.state()
        assert t == rt, mark
        assert bits_equal(pos, rpos) and bits_equal(vel, rvel), mark
    assert prop.time() == ref.solution_time()
    got, exp = prop.take_solution(), ref.splines()
    assert len(got) == len(exp) == 32
    total = 0
    for g, e in zip(got, exp):
        assert g.start == e[0] and g.interval == e[1] and len(g.polynomials) == len(e[2])
        if g.polynomials:
            a = np.concatenate([np.asarray(p).reshape(-1) for p in g.polynomials])
            b = np.concatenate([np.asarray(q).reshape(-1) for q in e[2]])
            assert bits_equal(a, b)
        total += len(g.polynomials)
    assert total > 400_000  # Phobos alone contributes 125 000 polynomials


def test_c5_1024_ships_against_the_32_body_spline_ephemeris():
    """BASELINE.json configs[4]: 1 024 perturbed copies of the reference's "Mars Transfer Ship" coasting 1950-01-01 ->
    1950-08-20 against the 2-year spline ephemeris of the 32-body system (built on the device by the n-body path).
    Eight ships spread over the batch are compared knot for knot with the oracle IN ITS DEFAULT MODE (the step-size
    controller calls the platform libm's pow, as the reference does on Linux); two more are re-run alone to show that a
    ship's result does not depend on the batch it is in."""
    s = load_system("full_solar_system_2433282.5")
    eph_prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY,
                                      solout=(s.dt, s.sample_period, s.degree))
    eph_prop.step_to(s.epoch + 2 * 365 * 86400.0)
    eph = eph_prop.take_solution_ephemeris()
    nb, n_poly = eph.sizes()
    assert nb == 32 and n_poly.sum() > 40_000
    from helpers import SYSTEMS
    ship = formats.load_ship(SYSTEMS / "full_solar_system_2433282.5.json", s.names, name="Mars Transfer Ship")
    ns = 1024
    rng = np.random.default_rng(20260924)
    states = np.tile(np.concatenate([ship.position, ship.velocity]), (ns, 1))
    states[:, :3] += 10.0 * rng.uniform(-1, 1, (ns, 3))
    states[:, 3:] += 0.010 * rng.uniform(-1, 1, (ns, 3))
    params = ee.default_adaptive_params(ship.tolerance, ship.tolerance)
    ships = ee.SpacecraftPropagator.new(ship.start, states, params, None, eph)
    ships.step_to(ship.end, max_steps=200000)
    info = ships.info()
    assert np.all(info["status"] == 0) and np.all(info["time"] >= ship.end)
    sol = ships.take_solution()
    mus, spl = eph.splines()
    ora = oracle.Ephem(mus, [(x.start, x.interval, x.polynomials) for x in spl])
    prm = (60.0, sys.float_info.max, ship.tolerance, ship.tolerance, 0.2, 5.0, 0.9)
    assert params.pow_mode == ee.POW_GLIBC
    for i in (0, 1, 127, 300, 511, 512, 800, 1023):
        o = oracle.Ship(ora, ship.start, states[i], prm, 1_000_000)
        st, _ = o.step_to(ship.end)
        kn = o.knots()
        assert st == 0 and kn.shape == sol[i].knots.shape, i
        assert np.array_equal(kn.view(np.uint64), sol[i].knots.view(np.uint64)), i
        oi = o.info()
        assert info["n_attempts"][i] == oi["n_attempts"] and info["rhs_evals"][i] == oi["rhs_evals"], i
    for i in (5, 900):
        alone = ee.SpacecraftPropagator.new(ship.start, states[i:i + 1], params, None, eph)
        alone.step_to(ship.end, max_steps=200000)
        assert np.array_equal(alone.take_solution()[0].knots.view(np.uint64), sol[i].knots.view(np.uint64)), i


@pytest.mark.parametrize("n,steps", [(2048, 16), (2176, 16), (4096, 20), (16384, 14)])
def test_c3_mid_size_systems_run_the_pair_symmetric_kernel(n, steps):
    """BASELINE.json configs[2] (4 096 bodies) and its neighbours: from 2 048 bodies up throughput mode runs the
    pair-symmetric kernel with one-warp CTAs and 128-body tiles (any multiple of 128, not only of 1 024).  Checked against
    the bit-exact parity kernel through the start-up boundary into steady state, against the plain all-pairs kernel
    (different bits, same physics), for run-to-run determinism, and the 8 rank shares must add up."""
    p0, v0, mu = ee.synthetic.plummer(n, seed=n)
    fast = ee.NBodyPropagator.new(ee.Forward(H), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    ref = ee.NBodyPropagator.new(ee.Forward(H), 0.0, p0, v0, mu, mode=ee.MODE_PARITY)
    fast.step(steps)
    ref.step(steps)
    t, pos, vel = fast.state()
    rt, rpos, rvel = ref.state()
    assert t == rt and rel_err(pos, rpos) <= 1e-12 and rel_err(vel, rvel) <= 1e-10
    again = ee.NBodyPropagator.new(ee.Forward(H), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    again.step(steps)
    assert bits_equal(again.state()[1], pos)
    exact = ee.gravity_eval(p0, mu, ee.MODE_PARITY)
    full = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    with dev_env(EE_SYM="0"):
        plain = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    assert rel_err(full, exact) < 1e-12 and rel_err(plain, exact) < 1e-12
    assert not bits_equal(full, plain)  # really two different kernels
    parts = np.zeros_like(full)
    for a in range(8):
        with dev_env(EE_SYM_RANGE="%d/8" % a):
            parts += ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    assert rel_err(parts, exact) < 1e-12

"""Sizes of the persistent single-CTA kernel other than the reference systems' 3/10/32 bodies.  Kept in a file that
sorts last so that `pytest -x` runs every other GPU test first."""
import numpy as np
import pytest

import oracle
from helpers import bits_equal
import ephemeris_explorer_b200 as ee

pytestmark = pytest.mark.gpu


def rand_system(n, seed):
    rng = np.random.default_rng(seed)
    pos = rng.normal(size=(n, 3)) * 10.0
    vel = rng.normal(size=(n, 3)) * 0.1
    mu = rng.uniform(0.01, 1.0, n)
    return pos, vel, mu


@pytest.mark.parametrize("n", [2, 17, 50, 64])
def test_persistent_small_kernel_sizes_bit_exact(n):
    """n <= 64 in parity mode runs the persistent single-CTA kernel (more than one pair per thread above 32 bodies,
    two load batches in the row sums); positions, velocities, accelerations and the spline solution stay bit-exact."""
    pos, vel, mu = rand_system(n, 1000 + n)
    h = 2.0 ** -10  # exact in binary: the `last_sample_time == sample_period` test of the solout is hit on schedule
    periods = np.array([h * (1 + (b % 5)) for b in range(n)])
    degrees = np.array([5 + (b % 4) for b in range(n)], dtype=np.int32)
    prop = ee.NBodyPropagator.new(ee.Forward(h), 0.0, pos, vel, mu, mode=ee.MODE_PARITY, solout=(h, periods, degrees))
    ref = oracle.NBody(pos, vel, mu, 0.0, h)
    ref.set_solout(h, periods, degrees)
    for chunk in (12, 1, 50, 137):
        prop.step(chunk)
        assert ref.step(chunk) == 0
        t, p, v, a = prop.state(accelerations=True)
        rt, rp, rv, ra = ref.state()
        assert t == rt and bits_equal(p, rp) and bits_equal(v, rv) and bits_equal(a, ra), (n, chunk)
    got, exp = prop.take_solution(), ref.take_solution()
    for g, e in zip(got, exp):
        assert g.start == e[0] and len(g.polynomials) == len(e[2])
        for x, y in zip(g.polynomials, e[2]):
            assert bits_equal(x, y)


def test_rank_shares_of_the_pair_items_add_up_on_one_gpu():
    """The sharded runs split the pair-symmetric kernel's item list by rank.  EE_SYM_RANGE=a/b makes a single-GPU engine
    behave like rank a of b (partial accelerations of its items only), so the N>1 item-range logic is covered on a
    one-GPU box: the 8 shares must add up to the full evaluation."""
    import os
    from helpers import rel_err
    n = 32768
    p0, _, mu = ee.synthetic.plummer(n, seed=9)
    full = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    parts = np.zeros_like(full)
    os.environ["EE_DEV_AIDS"] = "1"  # developer switches are ignored without it
    try:
        for a in range(8):
            os.environ["EE_SYM_RANGE"] = "%d/8" % a
            share = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
            assert not bits_equal(share, full)
            parts += share
    finally:
        os.environ.pop("EE_SYM_RANGE", None)
        os.environ.pop("EE_DEV_AIDS", None)
    assert rel_err(parts, full) < 1e-12


def test_export_state_at_an_epoch_is_the_oracle_state_vector_of_every_body():
    """`ExportType::State { epoch }` (ui/windows/export.rs:222-257): every body's `Trajectory::state_vector(epoch)` from
    the device-resident splines, written in the state.json schema; None when a trajectory does not cover the epoch."""
    from helpers import load_system
    s = load_system("sun_earth_moon_2433282.5")
    prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu,
                                  solout=(s.dt, s.sample_period, s.degree))
    prop.step(8 * 12 * 4)
    eph = prop.take_solution_ephemeris()
    mus, spl = eph.splines()
    ora = oracle.Ephem(mus, [(x.start, x.interval, x.polynomials) for x in spl])
    t = s.epoch + s.dt * 100.5
    doc = ee.formats.export_state(eph, s.names, mus, t, name=s.name)
    assert doc is not None and doc["name"] == s.name and doc["epoch"] == ee.formats.format_epoch(t)
    assert [b["name"] for b in doc["bodies"]] == list(s.names)
    for b in range(3):
        r = ora.state_vector(b, t)
        assert r is not None
        assert doc["bodies"][b]["mu"] == float(mus[b])
        assert bits_equal(np.array(doc["bodies"][b]["position"]), r[0])
        assert bits_equal(np.array(doc["bodies"][b]["velocity"]), r[1])
    assert ee.formats.export_state(eph, s.names, mus, s.epoch - 1.0) is None

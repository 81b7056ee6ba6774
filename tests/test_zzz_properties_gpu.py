"""Size-independent properties of the n-body path at the bench size (BASELINE.json configs[3], 65 536 bodies), where the CPU
oracle is too slow to be the checker: Newton's third law, exact power-of-two scaling laws, conservation of linear momentum
through the stepper.  Through the C ABI; the file sorts last on purpose."""
import numpy as np
import pytest

import ephemeris_explorer_b200 as ee

pytestmark = pytest.mark.gpu

N_FULL = 65536
H = 2.0 ** -10


def same_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    return a.shape == b.shape and bool(np.all(a.view(np.uint64) == b.view(np.uint64)))


def rel(a, b):
    return float(np.max(np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)))


def test_newtons_third_law_at_the_bench_size():
    """Every pair force acts on both bodies with opposite sign (the reference's loop applies one pair result to both,
    nbody.rs:27-33), so sum_i mu_i a_i vanishes up to rounding whatever the summation order."""
    p0, _, mu = ee.synthetic.plummer(N_FULL)
    for mode in (ee.MODE_THROUGHPUT, ee.MODE_PARITY):
        acc = ee.gravity_eval(p0, mu, mode)
        net = np.linalg.norm(np.sum(mu[:, None] * acc, axis=0))
        scale = float(np.sum(mu * np.linalg.norm(acc, axis=1)))
        assert scale > 0.0 and net / scale <= 1e-12, (mode, net / scale)


def test_power_of_two_scaling_laws_at_the_bench_size():
    """a(2 x, mu) = a(x, mu) / 4 and a(x, 2 mu) = 2 a(x, mu).  Scaling by a power of two commutes with every IEEE operation
    (no overflow or subnormals here), so in parity mode the laws hold BIT FOR BIT at any size -- a check of all 2.1e9 pair
    evaluations that needs no reference; the throughput kernel (FMA, rsqrt seed + refinement) must hold them to rounding."""
    p0, _, mu = ee.synthetic.plummer(N_FULL)
    base = ee.gravity_eval(p0, mu, ee.MODE_PARITY)
    assert same_bits(ee.gravity_eval(2.0 * p0, mu, ee.MODE_PARITY) * 4.0, base)
    assert same_bits(ee.gravity_eval(p0, 2.0 * mu, ee.MODE_PARITY), 2.0 * base)
    fast = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    assert rel(fast, base) <= 1e-12
    assert rel(ee.gravity_eval(2.0 * p0, mu, ee.MODE_THROUGHPUT) * 4.0, fast) <= 1e-12
    assert rel(ee.gravity_eval(p0, 2.0 * mu, ee.MODE_THROUGHPUT), 2.0 * fast) <= 1e-12


def test_linear_momentum_is_conserved_through_start_up_and_steady_steps_at_the_bench_size():
    """sum_i mu_i a_i = 0 at every evaluation and both the Blanes-Moan start-up and the QT12 / Cowell formulas are linear
    in the accelerations, so the total momentum of the Plummer sphere stays where it started (zero up to rounding)."""
    p0, v0, mu = ee.synthetic.plummer(N_FULL)
    prop = ee.NBodyPropagator.new(ee.Forward(H), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    scale = float(np.sum(mu * np.linalg.norm(v0, axis=1)))
    before = np.sum(mu[:, None] * v0, axis=0)
    for steps in (12, 8):
        prop.step(steps)
        _, pos, vel = prop.state()
        drift = np.linalg.norm(np.sum(mu[:, None] * vel, axis=0) - before)
        assert drift / scale <= 1e-12, (steps, drift / scale)
    # and the centre of mass moves with that (zero) momentum
    com0 = np.sum(mu[:, None] * p0, axis=0)
    com1 = np.sum(mu[:, None] * pos, axis=0)
    assert np.linalg.norm(com1 - com0) <= 1e-12 * float(np.sum(mu * np.linalg.norm(p0, axis=1)))

"""Parity of the CUDA n-body path against the CPU oracle, through the C ABI (GPU tests)."""
import numpy as np
import pytest

import oracle
from helpers import bits_equal, load_system, rel_err
import ephemeris_explorer_b200 as ee

pytestmark = pytest.mark.gpu


def rand_system(n, seed):
    rng = np.random.default_rng(seed)
    pos = rng.normal(size=(n, 3)) * 10.0
    vel = rng.normal(size=(n, 3)) * 0.1
    mu = rng.uniform(0.01, 1.0, n)
    return pos, vel, mu


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 127, 128, 129, 300])
def test_gravity_eval_parity_bit_exact(n):
    pos, _, mu = rand_system(n, n)
    got = ee.gravity_eval(pos, mu, ee.MODE_PARITY)
    assert bits_equal(got, oracle.gravity_eval(pos, mu))


@pytest.mark.parametrize("n", [2, 33, 256, 1000, 4096])
def test_gravity_eval_throughput_close(n):
    pos, _, mu = rand_system(n, 100 + n)
    got = ee.gravity_eval(pos, mu, ee.MODE_THROUGHPUT)
    ref = oracle.gravity_eval(pos, mu)
    assert rel_err(got, ref) < 2e-13  # both sides round ~sqrt(n) ulp in different orders


def test_gravity_eval_coincident_bodies_match_reference_nan():
    pos = np.array([[0.0, 0, 0], [0.0, 0, 0], [1.0, 0, 0]])
    mu = np.ones(3)
    got = ee.gravity_eval(pos, mu, ee.MODE_PARITY)
    ref = oracle.gravity_eval(pos, mu)
    assert np.array_equal(np.isnan(got), np.isnan(ref))


@pytest.mark.parametrize("name,steps", [("sun_earth_moon_2433282.5", 1000), ("simple_solar_system_2433282.5", 300),
                                        ("full_solar_system_2433282.5", 500)])
def test_parity_mode_bit_exact_on_reference_systems(name, steps):
    s = load_system(name)
    prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY)
    ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
    done = 0
    for chunk in (1, 5, 6, 1, steps - 13):  # crosses the start-up / steady-state boundary at step 12
        prop.step(chunk)
        assert ref.step(chunk) == 0
        done += chunk
        t, pos, vel, acc = prop.state(accelerations=True)
        rt, rpos, rvel, racc = ref.state()
        assert t == rt, done
        assert bits_equal(pos, rpos), done
        assert bits_equal(vel, rvel), done
        assert bits_equal(acc, racc), done
    assert prop.step_count() == steps


def test_parity_mode_backward_and_stormer13():
    s = load_system("sun_earth_moon_2433282.5")
    for method in (ee.QUINLAN_TREMAINE_12, ee.STORMER_13):
        prop = ee.NBodyPropagator.new(ee.Backward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY,
                                      method=method)
        ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, -s.dt, method)
        prop.step(100)
        ref.step(100)
        t, pos, vel = prop.state()
        rt, rpos, rvel, _ = ref.state()
        assert t == rt and bits_equal(pos, rpos) and bits_equal(vel, rvel)


def test_parity_mode_bit_exact_many_bodies():
    pos, vel, mu = rand_system(300, 7)
    h = 1e-3
    prop = ee.NBodyPropagator.new(ee.Forward(h), 0.0, pos, vel, mu, mode=ee.MODE_PARITY)
    ref = oracle.NBody(pos, vel, mu, 0.0, h)
    prop.step(30)
    ref.step(30)
    _, p, v = prop.state()
    _, rp, rv, _ = ref.state()
    assert bits_equal(p, rp) and bits_equal(v, rv)


@pytest.mark.parametrize("n,steps", [(256, 64), (1024, 44), (4096, 24)])
def test_throughput_mode_within_1e12_on_plummer(n, steps):
    p0, v0, mu = ee.synthetic.plummer(n)
    h = 2.0 ** -10
    prop = ee.NBodyPropagator.new(ee.Forward(h), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    ref = oracle.NBody(p0, v0, mu, 0.0, h)
    prop.step(steps)
    ref.step(steps)
    t, pos, vel = prop.state()
    rt, rpos, rvel, _ = ref.state()
    assert t == rt
    assert rel_err(pos, rpos) <= 1e-12  # north-star tolerance, positions
    assert rel_err(vel, rvel) <= 1e-10


def test_throughput_mode_solar_system_short_run():
    s = load_system("full_solar_system_2433282.5")
    prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_THROUGHPUT)
    ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
    prop.step(100)
    ref.step(100)
    _, pos, _ = prop.state()
    _, rpos, _, _ = ref.state()
    assert rel_err(pos, rpos) <= 1e-12


def test_throughput_is_deterministic_run_to_run():
    p0, v0, mu = ee.synthetic.plummer(2048)
    outs = []
    for _ in range(2):
        prop = ee.NBodyPropagator.new(ee.Forward(2.0 ** -10), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
        prop.step(20)
        outs.append(prop.state()[1])
    assert bits_equal(outs[0], outs[1])


def test_step_errors_mirror_reference():
    pos, vel, mu = rand_system(4, 3)
    prop = ee.NBodyPropagator.new(ee.Forward(1e-30), 1.0e6, pos, vel, mu)  # t + h == t
    assert prop.try_step(1) == 1  # StepError::StepSizeUnderflow (multistep/mod.rs:207-209)
    with pytest.raises(ee.EngineError):
        ee.NBodyPropagator.new(ee.Forward(1.0), 0.0, pos, vel, mu, method=7)


def test_clone_continues_identically():
    s = load_system("simple_solar_system_2433282.5")
    a = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu,
                               solout=(s.dt, s.sample_period, s.degree))
    a.step(40)
    b = a.clone()
    a.step(60)
    b.step(60)
    assert bits_equal(a.state()[1], b.state()[1])
    sa, sb = a.take_solution(), b.take_solution()
    for x, y in zip(sa, sb):
        assert x.start == y.start and len(x.polynomials) == len(y.polynomials)
        for p, q in zip(x.polynomials, y.polynomials):
            assert bits_equal(p, q)


@pytest.mark.parametrize("name,nsteps,backward", [("sun_earth_moon_2433282.5", 500, False),
                                                  ("full_solar_system_2433282.5", 3700, False),
                                                  ("sun_earth_moon_2433282.5", 300, True)])
def test_spline_solution_bit_exact(name, nsteps, backward):
    s = load_system(name)
    d = ee.Backward(s.dt) if backward else ee.Forward(s.dt)
    prop = ee.NBodyPropagator.new(d, s.epoch, s.position, s.velocity, s.mu, solout=(s.dt, s.sample_period, s.degree))
    ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, -s.dt if backward else s.dt)
    ref.set_solout(s.dt, s.sample_period, s.degree)
    prop.step(nsteps)
    assert ref.step(nsteps) == 0
    assert prop.time() == ref.solution_time()
    got = prop.take_solution()
    exp = ref.take_solution()
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        assert g.start == e[0] and g.interval == e[1]
        assert len(g.polynomials) == len(e[2])
        for p, q in zip(g.polynomials, e[2]):
            assert bits_equal(p, q)
    # a second take continues where the first left off (take_solution swaps in a fresh solution, nbody.rs:181-189)
    prop.step(200)
    ref.step(200)
    got2 = prop.take_solution()
    exp2 = ref.take_solution()
    for g, e in zip(got2, exp2):
        assert g.start == e[0] and len(g.polynomials) == len(e[2])
        for p, q in zip(g.polynomials, e[2]):
            assert bits_equal(p, q)


def test_step_to_and_propagate():
    s = load_system("sun_earth_moon_2433282.5")
    prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu,
                                  solout=(s.dt, s.sample_period, s.degree))
    end = s.epoch + 30 * 86400.0
    sol = prop.propagate(end)
    assert all(sp.end() >= end for sp in sol)
    ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
    ref.set_solout(s.dt, s.sample_period, s.degree)
    while ref.solution_time() < end:
        ref.step(1)
    assert prop.state()[0] == ref.state()[0]
    assert [len(x.polynomials) for x in sol] == [len(e[2]) for e in ref.splines()]


def test_lsq_fit_kernel_bit_exact():
    rng = np.random.default_rng(5)
    ts = np.arange(9) / 8.0
    samples = rng.normal(size=(64, 9, 3)) * 1e6
    samples[3] = 0.0  # exact zeros -> empty polynomial after trim
    degs = np.array([(i % 9) for i in range(64)], dtype=np.int32)
    got, nc = ee.lsq_fit(degs, ts, samples)
    for i in range(64):
        exp, n = oracle.lsq_fit(int(degs[i]), ts, samples[i])
        assert nc[i] == n, i
        assert bits_equal(got[i], exp), i


def test_ephemeris_evaluate_bit_exact():
    s = load_system("sun_earth_moon_2433282.5")
    prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu,
                                  solout=(s.dt, s.sample_period, s.degree))
    prop.step(8 * 12 * 4)
    eph = prop.take_solution_ephemeris()
    mus, spl = eph.splines()
    ora = oracle.Ephem(mus, [(x.start, x.interval, x.polynomials) for x in spl])
    end = min(x.end() for x in spl)
    times = np.concatenate([np.linspace(s.epoch, end, 97), [s.epoch - 1.0, end + 1e7, s.epoch + s.dt * 8, end]])
    pos, vel, ok = eph.evaluate(times)
    for i, t in enumerate(times):
        for b in range(3):
            r = ora.state_vector(b, t)
            assert ok[i, b] == (r is not None)
            if r is not None:
                assert bits_equal(pos[i, b], r[0]) and bits_equal(vel[i, b], r[1])


def test_pair_symmetric_kernel_matches_oracle_and_plain_kernel():
    """n >= 32768 takes the Newton's-third-law kernel (ee_sym.cuh); EE_SYM=0 forces the plain all-pairs kernel."""
    import os
    n = 32768
    p0, _, mu = ee.synthetic.plummer(n, seed=5)
    got = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    os.environ["EE_SYM"] = "0"
    os.environ["EE_DEV_AIDS"] = "1"  # developer switches are ignored without it
    try:
        plain = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    finally:
        del os.environ["EE_SYM"]
        del os.environ["EE_DEV_AIDS"]
    ref = oracle.gravity_eval(p0, mu)
    assert rel_err(got, ref) < 1e-12
    assert rel_err(plain, ref) < 1e-12
    assert rel_err(got, plain) < 1e-12
    assert not bits_equal(got, plain)  # different summation order: proves the two paths really are different kernels
    again = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    assert bits_equal(got, again)  # deterministic although work items are scheduled dynamically


def test_pair_symmetric_stepping_short_run():
    n = 32768
    p0, v0, mu = ee.synthetic.plummer(n, seed=6)
    h = 2.0 ** -10
    prop = ee.NBodyPropagator.new(ee.Forward(h), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    ref = oracle.NBody(p0, v0, mu, 0.0, h)
    prop.step(1)   # one start-up call = 25 evaluations; the oracle needs ~2 s per evaluation at this size
    ref.step(1)
    _, pos, vel = prop.state()
    _, rpos, rvel, _ = ref.state()
    assert rel_err(pos, rpos) <= 1e-12
    assert rel_err(vel, rvel) <= 1e-10


def test_snapshot_restore_resumes_bit_exactly():
    """ee_nbody_snapshot / ee_nbody_restore: the host blob is the reference's `propagator.clone()` checkpoint."""
    s = load_system("full_solar_system_2433282.5")
    a = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu)
    a.step(30)
    blob = a.snapshot()
    assert blob.nbytes == a.snapshot_size()
    a.step(50)
    ta, pa, va = a.state()
    b = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu)
    b.restore(blob)
    assert b.step_count() == 30
    b.step(50)
    tb, pb, vb = b.state()
    assert ta == tb and bits_equal(pa, pb) and bits_equal(va, vb)
    ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
    ref.step(80)
    assert bits_equal(pb, ref.state()[1])
    # snapshots taken during the start-up phase resume too
    c = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu)
    c.step(5)
    blob5 = c.snapshot()
    d = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu)
    d.restore(blob5)
    d.step(75)
    assert bits_equal(d.state()[1], pb)
    with pytest.raises(ee.EngineError):
        other = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position[:3], s.velocity[:3], s.mu[:3])
        other.restore(blob)


def test_fp64_peak_probe_is_sane():
    peak = ee.fp64_fma_peak()
    assert 20.0 < peak < 45.0  # B200: 148 SMs x 64 DFMA/clk x 2 flop x ~1.9 GHz = 37 TFLOP/s nominal


def _concat_polys(chunks, n):
    """Per body: the polynomials of consecutive take_solution() results, in order."""
    out = [[] for _ in range(n)]
    for sol in chunks:
        for b, sp in enumerate(sol):
            out[b].extend(sp.polynomials if hasattr(sp, "polynomials") else sp[2])
    return out


def test_planner_loop_shape_single_steps_with_snapshots_is_bit_exact():
    """The Prediction Planner's call shape (prediction.rs:408-446): step() one at a time, has_reached() after every step,
    take_solution() + clone() at every synchronisation tick.  The engine runs ahead (steps are launched in batches behind
    the C ABI); nothing observable may change: splines, times and the cloned propagators' continuation are bit-identical to
    the oracle stepping one by one."""
    s = load_system("full_solar_system_2433282.5")
    prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu,
                                  solout=(s.dt, s.sample_period, s.degree))
    ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
    ref.set_solout(s.dt, s.sample_period, s.degree)
    end = s.epoch + 40 * 86400.0
    got, exp, clones = [], [], []
    k = 0
    while True:
        prop.step(1)
        assert ref.step(1) == 0
        k += 1
        reached = prop.has_reached(end)
        assert reached == (ref.solution_time() >= end)
        if k % 977 == 0 or reached:  # a "sync tick"
            assert prop.time() == ref.solution_time()
            got.append(prop.take_solution())
            exp.append(ref.take_solution())
            clones.append((k, prop.clone()))
        if reached:
            break
    assert k > 5000 and len(got) >= 6
    a, b = _concat_polys(got, 32), _concat_polys(exp, 32)
    for body in range(32):
        assert len(a[body]) == len(b[body])
        for p, q in zip(a[body], b[body]):
            assert bits_equal(p, q)
    assert prop.state()[0] == ref.state()[0] and bits_equal(prop.state()[1], ref.state()[1])
    # a snapshot clone taken mid-run continues exactly like the original did
    k0, c = clones[2]
    c.step(k - k0)
    assert bits_equal(c.state()[1], prop.state()[1]) and bits_equal(c.state()[2], prop.state()[2])


def test_run_ahead_is_invisible_single_steps_equal_one_batched_call():
    s = load_system("sun_earth_moon_2433282.5")
    a = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, solout=(s.dt, s.sample_period, s.degree))
    b = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, solout=(s.dt, s.sample_period, s.degree))
    for _ in range(9000):
        a.step(1)
    b.step(9000)
    assert a.step_count() == b.step_count() == 9000 and a.time() == b.time()
    ta, pa, va = a.state()
    tb, pb, vb = b.state()
    assert ta == tb and bits_equal(pa, pb) and bits_equal(va, vb)
    for x, y in zip(a.take_solution(), b.take_solution()):
        assert x.start == y.start and len(x.polynomials) == len(y.polynomials)
        assert all(bits_equal(p, q) for p, q in zip(x.polynomials, y.polynomials))


@pytest.mark.parametrize("steps_before", [5, 12, 700])
def test_snapshot_restore_with_solout_attached(steps_before):
    """Host checkpoint of a propagator WITH its dense output: during start-up, at the start-up boundary and in steady state
    with samples pending in the buffers and polynomials already fitted."""
    s = load_system("simple_solar_system_2433282.5")
    mk = lambda: ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu,
                                        solout=(s.dt, s.sample_period, s.degree))
    a = mk()
    a.step(steps_before)
    blob = a.snapshot()
    assert blob.nbytes == a.snapshot_size()
    a.step(400)
    b = mk()
    b.step(3)  # restore replaces whatever the handle held, solout included
    b.restore(blob)
    assert b.step_count() == steps_before
    b.step(400)
    assert a.time() == b.time()
    ta, pa, va = a.state()
    tb, pb, vb = b.state()
    assert ta == tb and bits_equal(pa, pb) and bits_equal(va, vb)
    sa, sb = a.take_solution(), b.take_solution()
    for x, y in zip(sa, sb):
        assert x.start == y.start and x.interval == y.interval and len(x.polynomials) == len(y.polynomials)
        assert all(bits_equal(p, q) for p, q in zip(x.polynomials, y.polynomials))
    ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
    ref.set_solout(s.dt, s.sample_period, s.degree)
    ref.step(steps_before + 400)
    for x, e in zip(sb, ref.take_solution()):
        assert x.start == e[0] and len(x.polynomials) == len(e[2])
        assert all(bits_equal(p, q) for p, q in zip(x.polynomials, e[2]))
    with pytest.raises(ee.EngineError):
        b.restore(blob[: blob.nbytes // 2].copy())  # truncated blob is rejected, not read past its end


def test_clone_during_startup_and_with_pending_samples():
    s = load_system("sun_earth_moon_2433282.5")
    a = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, solout=(s.dt, s.sample_period, s.degree))
    a.step(7)  # inside the Blanes-Moan start-up
    c1 = a.clone()
    for _ in range(100):
        a.step(1)  # run-ahead pending when the next clone is taken
    c2 = a.clone()
    a.step(500)
    c1.step(600)
    c2.step(500)
    for c in (c1, c2):
        assert c.time() == a.time() and bits_equal(c.state()[1], a.state()[1]) and bits_equal(c.state()[2], a.state()[2])
        for x, y in zip(c.take_solution(), a.clone().take_solution()):
            assert x.start == y.start and len(x.polynomials) == len(y.polynomials)
            assert all(bits_equal(p, q) for p, q in zip(x.polynomials, y.polynomials))


def test_state_async_matches_state_and_overlaps_steps():
    p0, v0, mu = ee.synthetic.plummer(4096)
    prop = ee.NBodyPropagator.new(ee.Forward(2.0 ** -10), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    prop.step(14)
    bufs = [(np.zeros((4096, 3)), np.zeros((4096, 3))) for _ in range(3)]
    times = []
    for k in range(3):  # three reads enqueued back to back with steps in between: none may see a later state
        times.append(prop.state_async(*bufs[k]))
        prop.step(1)
    prop.state_wait()
    chk = ee.NBodyPropagator.new(ee.Forward(2.0 ** -10), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    chk.step(14)
    for k in range(3):
        t, pos, vel = chk.state()
        assert t == times[k] and bits_equal(pos, bufs[k][0]) and bits_equal(vel, bufs[k][1])
        chk.step(1)


@pytest.mark.parametrize("name,steps,backward", [("sun_earth_moon_2433282.5", 200, False), ("full_solar_system_2433282.5", 60, False),
                                                 ("sun_earth_moon_2433282.5", 120, True)])
def test_blanes_moan_14a_parity_bit_exact(name, steps, backward):
    """BlanesMoan14A as the fixed-step method (integration/src/methods.rs:1730-1774; one of the three methods the reference's
    convergence test asserts, solar_system_convergence.rs:354-357): 15 kick/drift stages per step, FSAL with B[0] = 0."""
    s = load_system(name)
    d = ee.Backward(s.dt) if backward else ee.Forward(s.dt)
    prop = ee.NBodyPropagator.new(d, s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY, method=ee.BLANES_MOAN_14A,
                                  solout=(s.dt, s.sample_period, s.degree))
    ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, -s.dt if backward else s.dt, 14)
    ref.set_solout(s.dt, s.sample_period, s.degree)
    for chunk in (1, 2, steps - 3):
        prop.step(chunk)
        assert ref.step(chunk) == 0
        t, pos, vel = prop.state()
        rt, rpos, rvel, _ = ref.state()
        assert t == rt and bits_equal(pos, rpos) and bits_equal(vel, rvel)
    assert ref.evals() == 1 + 14 * steps  # stage 0 is only evaluated on the very first step
    for g, e in zip(prop.take_solution(), ref.take_solution()):
        assert g.start == e[0] and len(g.polynomials) == len(e[2])
        assert all(bits_equal(p, q) for p, q in zip(g.polynomials, e[2]))


def test_blanes_moan_14a_throughput_mode():
    p0, v0, mu = ee.synthetic.plummer(1024)
    h = 2.0 ** -10
    prop = ee.NBodyPropagator.new(ee.Forward(h), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT, method=ee.BLANES_MOAN_14A)
    ref = oracle.NBody(p0, v0, mu, 0.0, h, 14)
    prop.step(6)
    ref.step(6)
    t, pos, vel = prop.state()
    rt, rpos, rvel, _ = ref.state()
    assert t == rt and rel_err(pos, rpos) <= 1e-12 and rel_err(vel, rvel) <= 1e-10


def test_both_readings_of_the_pair_kernel_are_bit_exact():
    """`particular`'s source is not in the reference tree, so the scalar factor of the pair force has two plausible
    readings (mu / mag and mu * (1 / mag)).  Both exist as bit-exact kernels (generic parity kernel and the persistent
    small-system kernel) against the oracle's twin switch; they differ from each other in the last bits."""
    s = load_system("full_solar_system_2433282.5")
    pos, vel, mu = rand_system(200, 21)
    results = []
    try:
        for variant in (0, 1):
            ee.set_pair_variant(variant)
            oracle.set_pair_variant(variant)
            assert bits_equal(ee.gravity_eval(pos, mu, ee.MODE_PARITY), oracle.gravity_eval(pos, mu))
            prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY)
            ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
            prop.step(300)  # start-up through the generic kernels, steady state through the persistent kernel
            ref.step(300)
            t, p, v, a = prop.state(accelerations=True)
            rt, rp, rv, ra = ref.state()
            assert t == rt and bits_equal(p, rp) and bits_equal(v, rv) and bits_equal(a, ra)
            results.append(a)
    finally:
        ee.set_pair_variant(0)
        oracle.set_pair_variant(0)
    # the two readings differ in the last bit of some accelerations (at h = 600 s that is far below one ulp of a position,
    # which is why km-level tests of the reference cannot tell them apart)
    assert not bits_equal(results[0], results[1])
    assert rel_err(results[0], results[1]) < 1e-14

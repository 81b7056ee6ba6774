#!/usr/bin/env python3
"""Records outputs of the CPU oracle (oracle/ee_oracle.cpp) as hex-float fixtures: tests/golden/oracle_outputs.json.

The reference holds no golden vectors for this path and cannot be built here (no Rust toolchain), so these are NOT
reference outputs: they freeze what the oracle computes today, so that a change to the oracle (compiler flags, a new
method, a refactor) that moves a single bit is caught by tests/test_oracle_cpu.py without re-deriving anything.
Regenerate (and review the diff) only when the oracle is changed on purpose:  python tests/golden/make_oracle_outputs.py
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
sys.path.insert(0, str(HERE.parent.parent))


def hexa(a):
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).reshape(-1)]


def compute():
    import oracle
    from helpers import SHIP_TEST_DEGREES, SHIP_TEST_PERIOD_HOURS, load_system
    out = {}
    rng = np.random.default_rng(20260925)
    pos = rng.normal(size=(7, 3)) * 3.0
    mu = rng.uniform(0.1, 2.0, 7)
    out["gravity_eval_7"] = {"pos": hexa(pos), "mu": hexa(mu), "acc": hexa(oracle.gravity_eval(pos, mu))}
    oracle.set_pair_variant(1)
    out["gravity_eval_7"]["acc_variant1"] = hexa(oracle.gravity_eval(pos, mu))
    oracle.set_pair_variant(0)
    s = load_system("sun_earth_moon_2433282.5")
    for method, steps in ((12, 40), (13, 40), (14, 10)):
        nb = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt, method)
        assert nb.step(steps) == 0
        t, p, v, a = nb.state()
        out["sun_earth_moon_method%d_%dsteps" % (method, steps)] = {"t": float(t).hex(), "pos": hexa(p), "vel": hexa(v), "acc": hexa(a),
                                                                   "evals": int(nb.evals())}
    nb = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, -s.dt)
    nb.set_solout(s.dt, s.sample_period, s.degree)
    assert nb.step(8 * 12 * 2 + 5) == 0
    spl = nb.splines()
    out["sun_earth_moon_backward_splines"] = {"solution_time": float(nb.solution_time()).hex(),
                                              "bodies": [{"start": float(a).hex(), "interval": float(b).hex(), "n_poly": len(c),
                                                          "first": hexa(c[0]) if c else [], "last": hexa(c[-1]) if c else []}
                                                         for a, b, c in spl]}
    ts = np.arange(9) / 8.0
    xs = rng.normal(size=(9, 3)) * 1e5
    co, n = oracle.lsq_fit(6, ts, xs)
    out["lsq_fit_degree6"] = {"xs": hexa(xs), "n_coef": int(n), "coeffs": hexa(co)}
    out["pow_portable"] = {"%r,%r" % (x, y): float(oracle.pow_portable(x, y)).hex()
                           for x, y in ((0.37, -1.0 / 7.0), (12.5, -1.0 / 7.0), (1e-9, -1.0 / 7.0), (3.0, 0.5))}
    # a ship through the 10-body system (coast + one burn), engine-independent pow so that the fixture does not depend on libm
    s10 = load_system("simple_solar_system_2433282.5")
    h = 6 * 3600.0
    periods = np.array(SHIP_TEST_PERIOD_HOURS) * 3600.0
    nb = oracle.NBody(s10.position, s10.velocity, s10.mu, s10.epoch, h)
    nb.set_solout(h, periods, SHIP_TEST_DEGREES)
    while nb.solution_time() < s10.epoch + 40 * 86400.0:
        nb.step(1)
    eph = oracle.Ephem(s10.mu, nb.splines())
    state = [-27204249.668775786, 132947582.43848978, 57641619.74241204, -22.253599106181895, -5.189518219791726, -2.2515617105336263]
    burn = (s10.epoch + 915.0, s10.epoch + 915.0 + 315.0, np.array([0.0, 0.0, 10.0]) / 1e3, s10.names.index("Earth"))
    oracle.set_pow_mode(oracle.POW_PORTABLE)
    try:
        sh = oracle.Ship(eph, s10.epoch, state, (60.0, sys.float_info.max, 1e-3, 1e-3, 0.2, 5.0, 0.9), 1_000_000, [burn])
        st, _ = sh.step_to(s10.epoch + 20 * 86400.0)
        kn = sh.knots()
        info = sh.info()
    finally:
        oracle.set_pow_mode(oracle.POW_LIBM)
    out["ship_10body_20days"] = {"status": int(st), "n_knots": int(len(kn)), "last_knot": hexa(kn[-1]), "knot_100": hexa(kn[100]),
                                 "n_attempts": int(info["n_attempts"]), "rhs_evals": int(info["rhs_evals"])}
    # round 2: the other adaptive methods (flight_plan.rs:175-184), two days of coast + the burn, same ephemeris
    names = ("Verner87", "CashKarp45", "DormandPrince54", "DormandPrince87", "Fehlberg45", "Tsitouras75", "Verner98", "Fine45")
    oracle.set_pow_mode(oracle.POW_PORTABLE)
    try:
        for m, name in enumerate(names):
            sh = oracle.Ship(eph, s10.epoch, state, (60.0, sys.float_info.max, 1e-3, 1e-3, 0.2, 5.0, 0.9), 1_000_000, [burn], method=m)
            st, _ = sh.step_to(s10.epoch + 2 * 86400.0)
            kn = sh.knots()
            info = sh.info()
            out["ship_method_%s_2days" % name] = {"status": int(st), "n_knots": int(len(kn)), "last_knot": hexa(kn[-1]),
                                                  "n_attempts": int(info["n_attempts"]), "rhs_evals": int(info["rhs_evals"])}
        # SpacecraftSolout analytics (dynamics/spacecraft.rs:91-161, :536-586): SOI transitions and apsides of the same ship
        from ephemeris_explorer_b200 import formats
        sh = oracle.Ship(eph, s10.epoch, state, (60.0, sys.float_info.max, 1e-3, 1e-3, 0.2, 5.0, 0.9), 1_000_000, [burn])
        sh.enable_analytics(formats.soi_radii(s10))
        st, _ = sh.step_to(s10.epoch + 20 * 86400.0)
        tr, ap = sh.analytics()
        kn = sh.knots()
        out["ship_analytics_20days"] = {"status": int(st), "n_knots": int(len(kn)), "n_transitions": len(tr), "n_apsides": len(ap),
                                        "transitions": [[float(t).hex(), int(b)] for t, b in tr[:8]],
                                        "first_apsides": [[float(t).hex(), float(d).hex(), int(b), int(k)] for t, d, b, k in ap[:6]],
                                        "last_apsis": [float(ap[-1][0]).hex(), float(ap[-1][1]).hex(), int(ap[-1][2]), int(ap[-1][3])] if ap else []}
        # RelativeTrajectory::state_vector (trajectory.rs:315-335): a body w.r.t. a body, the ship w.r.t. a body, no reference
        at = s10.epoch + 7.25 * 86400.0
        rel = {}
        for label, kw in (("moon_wrt_earth", dict(body=s10.names.index("Moon"), reference=s10.names.index("Earth"))),
                          ("ship_wrt_earth", dict(reference=s10.names.index("Earth"), knots=kn)),
                          ("ship_inertial", dict(reference=None, knots=kn))):
            r = oracle.relative_state_vector(eph, at, **kw)
            rel[label] = hexa(np.concatenate(r)) if r is not None else None
        out["relative_state_vector"] = rel
    finally:
        oracle.set_pow_mode(oracle.POW_LIBM)
    # C2's system: 2000 steps of QT12 with the spline solout (12 start-up + 1988 steady), positions + one polynomial per body
    s32 = load_system("full_solar_system_2433282.5")
    nb = oracle.NBody(s32.position, s32.velocity, s32.mu, s32.epoch, s32.dt)
    nb.set_solout(s32.dt, s32.sample_period, s32.degree)
    assert nb.step(2000) == 0
    t, pp, vv, _ = nb.state()
    spl = nb.splines()
    out["full_solar_system_2000steps"] = {"t": float(t).hex(), "pos": hexa(pp), "vel": hexa(vv), "n_poly": [len(c) for _, _, c in spl],
                                          "last_poly_first_coeff": [hexa(c[-1][0]) if c else [] for _, _, c in spl]}
    return out


if __name__ == "__main__":
    res = compute()
    (HERE / "oracle_outputs.json").write_text(json.dumps(res, indent=1) + "\n")
    print("wrote", HERE / "oracle_outputs.json", len(res), "entries")

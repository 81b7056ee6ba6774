#!/usr/bin/env python3
"""The other BASELINE.json configs, measured on one B200 next to the CPU oracle (not the headline bench):

  C2  full_solar_system_2433282.5 (32 bodies), dt = 600 s, 10^6 steps, parity mode with the spline solout on:
      steps/s, body-steps/s, max relative position error vs the oracle at 10^3/10^4/10^5/10^6 steps (must be 0)
  C3  Plummer 4096, throughput mode
  C5  1024 massless ships (perturbed "Mars Transfer Ship", coasting) against the 2-year 32-body spline ephemeris,
      1950-01-01 -> 1950-08-20: ship-steps/s, RHS evaluations/s; oracle timed on a few ships
Writes one JSON object to stdout.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def main():
    import ephemeris_explorer_b200 as ee
    import oracle
    from ephemeris_explorer_b200 import formats
    quick = "--quick" in sys.argv
    out = {}
    sysdir = ROOT / "tests" / "golden" / "systems" / "full_solar_system_2433282.5.json"
    s = formats.load_system(sysdir)

    # ---------------- C2
    total = 100_000 if quick else 1_000_000
    marks = [m for m in (1_000, 10_000, 100_000, 1_000_000) if m <= total]
    prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY,
                                  solout=(s.dt, s.sample_period, s.degree))
    ref = oracle.NBody(s.position, s.velocity, s.mu, s.epoch, s.dt)
    ref.set_solout(s.dt, s.sample_period, s.degree)
    errs = {}
    done = 0
    gpu_s = 0.0
    cpu_s = 0.0
    for m in marks:
        t0 = time.perf_counter()
        prop.step(m - done)
        prop.sync()
        gpu_s += time.perf_counter() - t0
        t0 = time.perf_counter()
        assert ref.step(m - done) == 0
        cpu_s += time.perf_counter() - t0
        done = m
        _, pos, vel = prop.state()
        _, rpos, rvel, _ = ref.state()
        errs[str(m)] = {"max_rel_pos": float(np.max(np.linalg.norm(pos - rpos, axis=1) / np.linalg.norm(rpos, axis=1))),
                        "bitwise": bool(np.array_equal(pos.view(np.uint64), rpos.view(np.uint64)) and
                                        np.array_equal(vel.view(np.uint64), rvel.view(np.uint64)))}
    sol = prop.take_solution()
    rsol = ref.splines()
    same = all(len(a.polynomials) == len(b[2]) and all(np.array_equal(p.view(np.uint64), np.asarray(q).view(np.uint64))
                                                      for p, q in zip(a.polynomials, b[2])) for a, b in zip(sol, rsol))
    out["C2_full_solar_system"] = {
        "bodies": 32, "steps": total, "gpu_steps_per_s": total / gpu_s, "gpu_body_steps_per_s": 32 * total / gpu_s,
        "cpu_oracle_steps_per_s": total / cpu_s, "cpu_cores": 1, "parity": errs, "splines_bitwise_equal": bool(same),
        "polynomials": int(sum(len(a.polynomials) for a in sol)),
        "note": "wall clock incl. the 12 start-up steps, spline solout (sampling + LSQ fits) on in both arms"}

    # ---------------- C3
    p0, v0, mu = ee.synthetic.plummer(4096)
    pr = ee.NBodyPropagator.new(ee.Forward(2.0 ** -10), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    pr.step(15)
    ms = pr.step_timed(64, 0)
    n = 4096
    out["C3_plummer_4096"] = {"body_steps_per_s": n * 64 / (ms * 1e-3), "ms_per_step": ms / 64,
                              "tflops_20flop_convention": n * 64 / (ms * 1e-3) * (20.0 * (n - 1) + 236) / 1e12}

    # ---------------- C5
    eph_prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY,
                                      solout=(s.dt, s.sample_period, s.degree))
    t0 = time.perf_counter()
    eph_prop.step_to(s.epoch + 2 * 365 * 86400.0)
    eph = eph_prop.take_solution_ephemeris()
    t_eph = time.perf_counter() - t0
    nb, n_poly = eph.sizes()
    ship = formats.load_ship(sysdir, s.names, name="Mars Transfer Ship")
    ns = 256 if quick else 1024
    rng = np.random.default_rng(20260924)
    states = np.tile(np.concatenate([ship.position, ship.velocity]), (ns, 1))
    states[:, :3] += 10.0 * rng.uniform(-1, 1, (ns, 3))
    states[:, 3:] += 0.010 * rng.uniform(-1, 1, (ns, 3))
    end = ship.end
    ships = ee.SpacecraftPropagator.new(ship.start, states, ee.default_adaptive_params(ship.tolerance, ship.tolerance), None, eph)
    t0 = time.perf_counter()
    ships.step_to(end, max_steps=200000)
    t_gpu = time.perf_counter() - t0
    info = ships.info()
    steps = int(info["n_knots"].sum() - ns)
    evals = int(info["rhs_evals"].sum())
    # oracle on a few ships
    mus, spl = eph.splines()
    ora = oracle.Ephem(mus, [(x.start, x.interval, x.polynomials) for x in spl])
    k = 2 if quick else 4
    import sys as _s
    prm = (60.0, _s.float_info.max, ship.tolerance, ship.tolerance, 0.2, 5.0, 0.9)
    oracle.set_pow_mode(oracle.POW_PORTABLE)
    t0 = time.perf_counter()
    osteps, oevals, same_knots = 0, 0, True
    sol = ships.take_solution()
    for i in range(k):
        o = oracle.Ship(ora, ship.start, states[i], prm, 1_000_000)
        o.step_to(end)
        kn = o.knots()
        osteps += len(kn) - 1
        oevals += o.info()["rhs_evals"]
        same_knots = same_knots and kn.shape == sol[i].knots.shape and np.array_equal(kn.view(np.uint64), sol[i].knots.view(np.uint64))
    t_cpu = time.perf_counter() - t0
    oracle.set_pow_mode(oracle.POW_LIBM)
    out["C5_ships"] = {
        "ships": ns, "ephemeris_bodies": int(nb), "ephemeris_polynomials": int(n_poly.sum()), "ephemeris_build_s": t_eph,
        "status_ok": int((info["status"] == 0).sum()), "accepted_steps": steps, "rhs_evals": evals, "gpu_wall_s": t_gpu,
        "gpu_kernel_ms": ships.last_ms(), "gpu_ship_steps_per_s": steps / t_gpu, "gpu_rhs_evals_per_s": evals / t_gpu,
        "cpu_oracle_ships": k, "cpu_ship_steps_per_s": osteps / t_cpu, "cpu_rhs_evals_per_s": oevals / t_cpu, "cpu_cores": 1,
        "knots_bitwise_equal_on_checked_ships": bool(same_knots)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

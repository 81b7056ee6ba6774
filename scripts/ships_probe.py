#!/usr/bin/env python3
"""Developer probe (one GPU): C5 -- 1 024 ships against the 2-year 32-body spline ephemeris -- for every adaptive method:
accepted steps, RHS evaluations, kernel time, RHS/s.  One JSON line per method."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ephemeris_explorer_b200 as ee  # noqa: E402

s = ee.formats.load_system(ROOT / "tests" / "golden" / "systems" / "full_solar_system_2433282.5.json")
eph_prop = ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, mode=ee.MODE_PARITY,
                                  solout=(s.dt, s.sample_period, s.degree))
eph_prop.step_to(s.epoch + 2 * 365 * 86400.0)
eph = eph_prop.take_solution_ephemeris()
ship = ee.formats.load_ship(ROOT / "tests" / "golden" / "systems" / "full_solar_system_2433282.5.json", s.names, name="Mars Transfer Ship")
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
methods = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(8))
rng = np.random.default_rng(20260924)
states = np.tile(np.concatenate([ship.position, ship.velocity]), (ns, 1))
states[:, :3] += 10.0 * rng.uniform(-1, 1, (ns, 3))
states[:, 3:] += 0.010 * rng.uniform(-1, 1, (ns, 3))
for m in methods:
    for analytics in (False, True) if m == 0 else (False,):
        ships = ee.SpacecraftPropagator.new(ship.start, states, ee.default_adaptive_params(ship.tolerance, ship.tolerance, method=m), None, eph)
        if analytics:
            ships.enable_analytics(ee.formats.soi_radii(s))
        end = ship.end if m in (0, 3, 6) else ship.start + 20 * 86400.0  # low-order methods: 20 days are enough to time
        if len(sys.argv) > 3:
            end = ship.start + float(sys.argv[3]) * 86400.0
        kms = 0.0
        while True:
            ships.step_to(end, max_steps=200000)
            kms += ships.last_ms()
            info = ships.info()
            if np.all((info["time"] >= end) | (info["status"] != 0)):
                break
        steps = int(info["n_knots"].sum() - ns)
        evals = int(info["rhs_evals"].sum())
        print(json.dumps({"method": ee.SHIP_METHOD_NAMES[m], "analytics": analytics, "ships": ns, "days": (end - ship.start) / 86400.0,
                          "status_ok": int((info["status"] == 0).sum()), "accepted_steps": steps, "rhs_evals": evals, "kernel_ms": kms,
                          "ship_steps_per_s": steps / (kms * 1e-3), "rhs_per_s": evals / (kms * 1e-3)}), flush=True)
        ships.close()

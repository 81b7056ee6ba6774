#!/usr/bin/env python3
"""Tiny ship run for compute-sanitizer (memcheck / racecheck): a 10-body ephemeris over a few days, 3 ships, every method for
a few hundred steps, analytics on, relative sampling, and a mid-size n-body step."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import ephemeris_explorer_b200 as ee  # noqa: E402
from helpers import SHIP_TEST_DEGREES, SHIP_TEST_PERIOD_HOURS, load_system  # noqa: E402

s = load_system("simple_solar_system_2433282.5")
h = 6 * 3600.0
prop = ee.NBodyPropagator.new(ee.Forward(h), s.epoch, s.position, s.velocity, s.mu,
                              solout=(h, np.array(SHIP_TEST_PERIOD_HOURS) * 3600.0, SHIP_TEST_DEGREES))
prop.step_to(s.epoch + 40 * 86400.0)
eph = prop.take_solution_ephemeris()
state = [-27204249.668775786, 132947582.43848978, 57641619.74241204, -22.253599106181895, -5.189518219791726, -2.2515617105336263]
states = np.tile(np.array(state), (3, 1))
states[1:, :3] += 5.0
radii = ee.formats.soi_radii(s)
for m in range(8):
    ships = ee.SpacecraftPropagator.new(s.epoch, states, ee.default_adaptive_params(1e-3, 1e-3, method=m), None, eph)
    ships.enable_analytics(radii)
    for _ in range(3):
        ships.step_to(s.epoch + 0.5 * 86400.0, max_steps=40)
    ships.analytics()
    ships.evaluate_relative(1, 3, np.linspace(s.epoch, s.epoch + 3000.0, 50))
    ships.take_solution()
    ships.close()
    plain = ee.SpacecraftPropagator.new(s.epoch, states, ee.default_adaptive_params(1e-3, 1e-3, method=m), None, eph)
    plain.step_to(s.epoch + 0.2 * 86400.0, max_steps=60)
    plain.close()
eph.evaluate_relative(2, 1, np.linspace(s.epoch, s.epoch + 1e6, 100))
for n in (2048, 2176):
    p0, v0, mu = ee.synthetic.plummer(n)
    p = ee.NBodyPropagator.new(ee.Forward(2.0 ** -10), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    p.step(14)
    p.state()
    p.close()
print("tiny ok")

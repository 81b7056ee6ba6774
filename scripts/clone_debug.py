#!/usr/bin/env python3
"""Developer probe: minimal scenarios around clone() + run-ahead + take_solution (each in its own process)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import ephemeris_explorer_b200 as ee  # noqa: E402
from helpers import load_system  # noqa: E402

sc = sys.argv[1]
s = load_system("sun_earth_moon_2433282.5")
mk = lambda: ee.NBodyPropagator.new(ee.Forward(s.dt), s.epoch, s.position, s.velocity, s.mu, solout=(s.dt, s.sample_period, s.degree))
a = mk()
if sc == "startup_clone":        # clone during start-up, step the clone alone, take
    a.step(7)
    c = a.clone()
    c.step(600)
    c.state()
    c.take_solution()
elif sc == "startup_clone_sync":  # same, device-wide sync before take
    a.step(7)
    c = a.clone()
    c.step(600)
    c.sync()
    import torch
    torch.cuda.synchronize()
    c.take_solution()
elif sc == "steady_clone":       # clone in steady state with run-ahead pending
    a.step(107)
    c = a.clone()
    c.step(500)
    c.state()
    c.take_solution()
elif sc == "orig_after_clone":   # the original keeps going after a start-up clone
    a.step(7)
    c = a.clone()
    a.step(600)
    a.state()
    a.take_solution()
elif sc == "no_clone":           # control
    a.step(607)
    a.state()
    a.take_solution()
elif sc == "startup_clone_generic_only":  # clone during start-up, finish start-up only, take
    a.step(7)
    c = a.clone()
    c.step(5)
    c.state()
    c.take_solution()
elif sc == "startup_clone_two_takes":
    a.step(7)
    c = a.clone()
    c.step(100)
    c.take_solution()
    c.step(500)
    c.take_solution()
print("scenario", sc, "ok")

#!/usr/bin/env python3
"""Developer probe: a few steady-state steps at N = 65 536 (for ncu captures).  EE_SYM_VARIANT etc. via the environment."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ephemeris_explorer_b200 as ee  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
pos, vel, mu = ee.synthetic.plummer(n)
p = ee.NBodyPropagator.new(ee.Forward(2.0 ** -10), 0.0, pos, vel, mu, mode=ee.MODE_THROUGHPUT)
p.step(12 + 4)
p.sync()
print("ok", p.step_count())

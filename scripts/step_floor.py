#!/usr/bin/env python3
"""Developer probe: per-step cost of throughput-mode stepping as a function of N (back-to-back steps inside one event pair,
and wall clock of the same call) -- separates the fixed per-step cost (launches, kernel boundaries) from the pair work."""
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ephemeris_explorer_b200 as ee  # noqa: E402

os.environ["EE_DEV_AIDS"] = "1"
for n in (128, 512, 1024, 2048, 4096):
    pos, vel, mu = ee.synthetic.plummer(n)
    for kind in ("default", "plain"):
        if kind == "plain":
            os.environ["EE_SYM"] = "0"
        else:
            os.environ.pop("EE_SYM", None)
        p = ee.NBodyPropagator.new(ee.Forward(2.0 ** -10), 0.0, pos, vel, mu, mode=ee.MODE_THROUGHPUT)
        p.step(15)
        p.sync()
        best_ev, best_wall = 1e9, 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            p.step(256)
            t1 = time.perf_counter()
            p.sync()
            t2 = time.perf_counter()
            best_ev = min(best_ev, p.last_timing()[0] / 256)
            best_wall = min(best_wall, (t2 - t0) / 256 * 1e3)
            enq = (t1 - t0) / 256 * 1e3
        print(json.dumps({"n": n, "kernel": kind, "ms_per_step_events": best_ev, "ms_per_step_wall": best_wall,
                          "ms_per_step_host_enqueue": enq}), flush=True)
        p.close()

#!/usr/bin/env python3
"""Developer probe (one GPU): mid-size systems (2 048 .. 16 384 bodies) on the pair-symmetric kernel variants vs the plain
all-pairs kernel.  Per (N, variant): ms/step for 64 back-to-back steps inside one event pair (the engine's own step(k)
timing) and with a host synchronisation per step (step_timed), fraction of the DFMA peak, and the acceleration's relative
difference from the plain kernel.  One JSON line per measurement."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ephemeris_explorer_b200 as ee  # noqa: E402

H = 2.0 ** -10
os.environ["EE_DEV_AIDS"] = "1"
sizes = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "2048,4096,8192,16384".split(","))]
variants = sys.argv[2:] or ["plain", "4,32,16,4", "4,128,4,4", "4,256,2,16"]
peak = ee.fp64_fma_peak(0)
for n in sizes:
    pos, vel, mu = ee.synthetic.plummer(n)
    base = None
    for v in variants:
        os.environ.pop("EE_SYM", None)
        os.environ.pop("EE_SYM_VARIANT", None)
        if v == "plain":
            os.environ["EE_SYM"] = "0"
        else:
            os.environ["EE_SYM_VARIANT"] = v
            ti, nt = int(v.split(",")[0]), int(v.split(",")[1])
            if n % (ti * nt):
                continue
        row = {"n": n, "variant": v}
        try:
            p = ee.NBodyPropagator.new(ee.Forward(H), 0.0, pos, vel, mu, mode=ee.MODE_THROUGHPUT)
            p.step(12 + 3)
            p.sync()
            best = 1e9
            for _ in range(3):
                p.step(64)
                p.sync()
                best = min(best, p.last_timing()[0] / 64)
            synced = p.step_timed(64, 0) / 64
            fl = n * (20.0 * (n - 1) + 236)
            row.update({"ms_per_step_back_to_back": best, "ms_per_step_host_sync_each": synced,
                        "frac_of_dfma_peak": fl / (best * 1e-3) / 1e12 / peak, "peak_tflops": peak})
            acc = ee.gravity_eval(pos, mu, ee.MODE_THROUGHPUT)
            if base is None:
                base = acc
            row["accel_rel_vs_first"] = float(np.max(np.linalg.norm(acc - base, axis=1) / np.linalg.norm(base, axis=1)))
            p.close()
        except Exception as exc:
            row["error"] = repr(exc)
        print(json.dumps(row), flush=True)

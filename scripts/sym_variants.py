#!/usr/bin/env python3
"""Developer probe (one GPU): time the pair-symmetric kernel variants at N = 65 536, whole list and one rank's 1/8 share,
and check each against variant 4,256,2,16.  Prints one JSON line per measurement."""
import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ephemeris_explorer_b200 as ee  # noqa: E402

N = 65536
H = 2.0 ** -10
FLUSH = 256 << 20
os.environ["EE_DEV_AIDS"] = "1"
variants = sys.argv[1:] or ["4,256,2,16", "4,128,3,16"]
pos, vel, mu = ee.synthetic.plummer(N)
base = None
for v in variants:
    os.environ["EE_SYM_VARIANT"] = v
    for maxc, guided in (("16", "1"), ("16", "2"), ("64", "1")):
        os.environ["EE_SYM_MAXC"] = maxc
        os.environ["EE_SYM_GUIDED"] = guided
        for share in (None, "3/8"):
            if share:
                os.environ["EE_SYM_RANGE"] = share
            else:
                os.environ.pop("EE_SYM_RANGE", None)
            try:
                p = ee.NBodyPropagator.new(ee.Forward(H), 0.0, pos, vel, mu, mode=ee.MODE_THROUGHPUT)
                p.step(12 + 3)
                ms = p.step_timed(16, FLUSH) / 16
                row = {"variant": v, "maxc": int(maxc), "guided": int(guided), "share": share or "1/1", "ms_per_step": ms}
                if not share:
                    row["tflops"] = N * (20.0 * (N - 1) + 236) / (ms * 1e-3) / 1e12
                    if (maxc, guided) == ("16", "1"):
                        acc = ee.gravity_eval(pos, mu, ee.MODE_THROUGHPUT)
                        if base is None:
                            base = acc
                        row["rel_vs_first"] = float(np.max(np.linalg.norm(acc - base, axis=1) / np.linalg.norm(base, axis=1)))
                p.close()
            except Exception as exc:  # a variant that cannot launch must not stop the sweep
                row = {"variant": v, "maxc": int(maxc), "guided": int(guided), "share": share or "1/1", "error": repr(exc)}
            print(json.dumps(row), flush=True)

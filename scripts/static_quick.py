import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import ephemeris_explorer_b200 as ee
from helpers import rel_err
p0, v0, mu = ee.synthetic.plummer(65536)
dyn = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
os.environ["EE_SYM_STATIC"] = "1"
st = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
print("static(vs dynamic) rel err %.2e" % rel_err(st, dyn), flush=True)
for static, share in (("1", None), ("1", "8"), ("0", "8")):
    os.environ["EE_SYM_STATIC"] = static
    if share: os.environ["EE_SYM_SHARE"] = share
    elif "EE_SYM_SHARE" in os.environ: del os.environ["EE_SYM_SHARE"]
    pr = ee.NBodyPropagator.new(ee.Forward(2.0**-10), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
    pr.step(12); pr.step_timed(3, 0)
    print("static=%s share=1/%s: %.4f ms/step" % (static, share or "1", pr.step_timed(10, 0) / 10), flush=True)
    pr.close()

"""Developer check of the opt-in static ("stream-K") split of the pair-symmetric kernel: EE_SYM_STATIC=1."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import ephemeris_explorer_b200 as ee
from helpers import rel_err
n = 65536
p0, v0, mu = ee.synthetic.plummer(n)
dyn = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
os.environ["EE_SYM_STATIC"] = "1"
ok = True
for js in ("512", "256"):
    os.environ["EE_SYM_JS"] = js
    full = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    again = ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    e_dyn = rel_err(full, dyn)
    det = bool(np.array_equal(full.view(np.uint64), again.view(np.uint64)))
    parts = np.zeros_like(full)
    for a in range(8):
        os.environ["EE_SYM_RANGE"] = "%d/8" % a
        parts += ee.gravity_eval(p0, mu, ee.MODE_THROUGHPUT)
    del os.environ["EE_SYM_RANGE"]
    e_parts = rel_err(parts, dyn)
    good = e_dyn < 1e-12 and det and e_parts < 1e-12
    ok = ok and good
    print("static js%s: vs dynamic %.2e, deterministic %s, sum of 8 rank shares vs dynamic %.2e -> %s" % (js, e_dyn, det, e_parts, "OK" if good else "FAIL"), flush=True)
del os.environ["EE_SYM_JS"]
for static in ("0", "1"):
    for share, js in ((None, "512"), ("8", "512"), ("8", "256")):
        os.environ["EE_SYM_STATIC"] = static; os.environ["EE_SYM_JS"] = js
        if share: os.environ["EE_SYM_SHARE"] = share
        elif "EE_SYM_SHARE" in os.environ: del os.environ["EE_SYM_SHARE"]
        pr = ee.NBodyPropagator.new(ee.Forward(2.0**-10), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
        pr.step(12); pr.step_timed(3, 0)
        ms = pr.step_timed(12, 0) / 12
        print("static=%s share=1/%s js=%s: %.4f ms/step" % (static, share or "1", js, ms), flush=True)
        pr.close()
print("ALL OK" if ok else "SOME FAILED", flush=True)

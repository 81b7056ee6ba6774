import os, sys
sys.path.insert(0, ".")
import ephemeris_explorer_b200 as ee
p0, v0, mu = ee.synthetic.plummer(65536)
for share in (8, 4):
    for js in (512, 256, 128):
        os.environ["EE_SYM_SHARE"] = str(share); os.environ["EE_SYM_JS"] = str(js)
        pr = ee.NBodyPropagator.new(ee.Forward(2.0**-10), 0.0, p0, v0, mu, mode=ee.MODE_THROUGHPUT)
        pr.step(12 + 0)
        pr.step_timed(3, 0)
        ms = pr.step_timed(20, 0)
        print("share 1/%d js %d: %.4f ms/step (ideal %.4f)" % (share, js, ms / 20, 3.18 / share), flush=True)
        pr.close()

#!/bin/bash
# round 2, session X (1 GPU): one rank's share of the 8-way sharded pair phase, every share, timed on one GPU
mkdir -p gpurun_out
python - > gpurun_out/x_shares.jsonl 2> gpurun_out/x_shares.err <<'PY'
import json, os, sys
sys.path.insert(0, '.')
os.environ["EE_DEV_AIDS"] = "1"
import ephemeris_explorer_b200 as ee
N = 65536
pos, vel, mu = ee.synthetic.plummer(N)
for world in (8, 4, 2, 1):
    for a in range(world):
        os.environ["EE_SYM_RANGE"] = "%d/%d" % (a, world)
        p = ee.NBodyPropagator.new(ee.Forward(2.0 ** -10), 0.0, pos, vel, mu, mode=ee.MODE_THROUGHPUT)
        p.step(12 + 3)
        ms = p.step_timed(16, 256 << 20) / 16
        print(json.dumps({"share": "%d/%d" % (a, world), "ms_per_step": ms}), flush=True)
        p.close()
PY
cat gpurun_out/x_shares.jsonl; tail -n 2 gpurun_out/x_shares.err

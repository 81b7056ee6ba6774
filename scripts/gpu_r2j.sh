#!/bin/bash
# round 2, session J (1 GPU): ncu of the ship kernel v3 on a 20-day run (plain and with analytics)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ships_step_to -c 2 -f -o gpurun_out/j_prof_ships \
    python scripts/ships_probe.py 1024 0 20 > gpurun_out/j_ncu_ships.log 2>&1
tail -n 4 gpurun_out/j_ncu_ships.log | cut -c1-300

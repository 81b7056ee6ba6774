#!/bin/bash
# round 2, session Y (1 GPU): ncu of the ship kernel v6 (plain, 20-day run)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ships_step_to -c 1 -f -o gpurun_out/y_prof_ships \
    python scripts/ships_probe.py 1024 0 20 > gpurun_out/y_ncu_ships.log 2>&1
tail -n 3 gpurun_out/y_ncu_ships.log | cut -c1-200

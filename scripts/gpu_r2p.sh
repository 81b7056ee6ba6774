#!/bin/bash
# round 2, session P (1 GPU): ncu of the ship kernel with analytics (20-day run)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ships_step_to -s 1 -c 3 -f -o gpurun_out/p_prof_ships_ana \
    python scripts/ships_probe.py 1024 0 20 > gpurun_out/p_ncu_ships.log 2>&1
tail -n 4 gpurun_out/p_ncu_ships.log | cut -c1-300

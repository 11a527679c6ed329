#!/bin/bash
# sweep of the first-order KMC launch shapes at low occupancy (one process per shape); "auto" = the library's own choice
for cfg in "1 20000 40" "148 4096 8" "296 2048 8" "512 2048 8" "1024 2048 8" "2048 2048 8"; do
  for lanes in 0 8 16 32 auto; do
    if [ $lanes = auto ]; then unset LMC_KMC_TEAM_LANES; else export LMC_KMC_TEAM_LANES=$lanes; fi
    python tools/kmc_team_probe.py $cfg 2>&1 | tail -1
  done
done

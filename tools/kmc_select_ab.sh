#!/bin/bash
# A/B of two library builds on one box for the first-order KMC kernels:
#   [TEAM_ONLY=1] tools/kmc_select_ab.sh ab/old.so ab/new.so
# per build: the bench shape (8192 walkers x 2048 hops x 8 launches), one GPU's share of the 8-GPU job (1024 walkers), a single trajectory
for rep in 1 2; do
  for v in "$@"; do
    echo "== $v (pass $rep)"
    [ -z "$TEAM_ONLY" ] && LMC_B200_LIB=$v python tools/kmc_age_once.py 8192 2048 8
    LMC_B200_LIB=$v python tools/kmc_latency_once.py 1024 2048
    LMC_B200_LIB=$v LMC_KMC_TEAM_SMEM=0 python tools/kmc_latency_once.py 1024 2048
    LMC_B200_LIB=$v python tools/kmc_latency_once.py 1 20000
    LMC_B200_LIB=$v python tools/kmc_latency_once.py 1 20000 63
  done
done

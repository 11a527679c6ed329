#!/bin/bash
# the tail of a hybrid KMC launch: resident blocks per SM of the latency kernel x fraction handed over (bench shape, 6 launches)
for per_sm in 7 5 4 3; do
  for f in 0.15 0.2 0.3; do
    echo -n "tail_per_sm $per_sm "
    LMC_KMC_TAIL_PER_SM=$per_sm LMC_KMC_HANDOFF=$f python tools/kmc_age_once.py ${1:-8192} ${2:-2048} ${3:-6}
  done
done
LMC_KMC_HANDOFF=0 python tools/kmc_age_once.py ${1:-8192} ${2:-2048} ${3:-6}

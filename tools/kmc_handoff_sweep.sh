#!/bin/bash
# fraction of the walkers handed from kmc_run_kernel to the latency kernel (Engine::kmc_run): kernel times of 8 launches at
# the bench shape per setting; the state digest must not depend on the setting
for f in ${FRACTIONS:-0 0.05 0.1 0.15 0.2 0.25 0.3 0.4 0.5}; do
  LMC_KMC_HANDOFF=$f python tools/kmc_age_once.py ${1:-8192} ${2:-2048} ${3:-8}
done

import sys, os, time, tempfile
sys.path.insert(0, '.')
import numpy as np
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
for f, nw in ((20, 1), (20, 16), (63, 1)):
    e = capi.Engine(f, n_walkers=nw, device=0); e.load_coefficients(js)
    occ = synth.random_alloy(f, 0.02, 0.02, seed=42)
    for w in range(nw): e.set_occupancy(occ, walker=w)
    e.kmc_reset(); e.kmc_run(2000, temperature=500.0, seed=1)
    t = time.perf_counter(); e.kmc_run(100000, temperature=500.0, seed=1); dt = time.perf_counter() - t
    print('f', f, 'walkers', nw, 'steps/s per walker', 100000/dt, 'kernel ms', e.last_kernel_ms())

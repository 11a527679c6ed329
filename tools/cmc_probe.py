import sys, os, time, tempfile
sys.path.insert(0, '.')
import numpy as np
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
f = int(sys.argv[1]) if len(sys.argv) > 1 else 40
sizes = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 256, 64]
e = capi.Engine(f, n_walkers=1, device=0); e.load_coefficients(js)
occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)
e.set_occupancy(occ); e.cmc_reset()
for bs in sizes:
    e.cmc_run(5000, temperature=800.0, seed=5, batch_size=bs)
    s0 = e.cmc_state(); e.cmc_run(20000, temperature=800.0, seed=5, batch_size=bs); ms = e.last_kernel_ms(); s1 = e.cmc_state()
    print('batch', bs, 'ms', ms, 'trials', s1['steps'][0]-s0['steps'][0], 'rate', (s1['steps'][0]-s0['steps'][0])/ms*1e3)

import sys, os, time, tempfile
sys.path.insert(0, '.')
import numpy as np
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
for f, trials in ((40, 200000), (100, 2000000)):
    e = capi.Engine(f, n_walkers=1, device=0); e.load_coefficients(js)
    occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)
    for it in range(3):
        t = [time.perf_counter()]
        e.set_occupancy_all(occ[None]); t.append(time.perf_counter())
        e.cmc_reset(); t.append(time.perf_counter())
        e.cmc_run(trials, temperature=800.0, seed=5); t.append(time.perf_counter())
        st = e.cmc_state(); t.append(time.perf_counter())
        e.get_occupancy_all(); t.append(time.perf_counter())
        print(f, it, ['%.2f ms' % ((b - a) * 1e3) for a, b in zip(t, t[1:])], 'kernel', round(e.last_kernel_ms(), 3), int(st['steps'][0]), flush=True)

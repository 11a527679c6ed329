import sys, os, tempfile
sys.path.insert(0, '.')
import numpy as np
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
for f, bs in ((6, 0), (6, 32), (6, 64), (8, 0), (12, 0), (12, 32)):
    e = capi.Engine(f, n_walkers=1, device=0); e.load_coefficients(js)
    occ = synth.random_alloy(f, 0.06, 0.06, seed=31, vacancy_site=None)
    e.set_occupancy(occ); e0 = e.total_energy(); e.cmc_reset()
    bad = None
    for it in range(30):
        e.cmc_run(100, temperature=800.0, seed=5, batch_size=bs)
        st = e.cmc_state(); e1 = e.total_energy()
        diff = (e1 - e0) - st['energy'][0]
        if abs(diff) > 1e-8 and bad is None: bad = (it, diff, int(st['steps'][0]))
    print('f', f, 'batch', bs, 'steps', int(st['steps'][0]), 'final diff', diff, 'first bad', bad, 'conserved', np.array_equal(np.sort(e.get_occupancy()), np.sort(occ)))

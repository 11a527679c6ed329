"""Lanes-per-trial-side sweep of the whole-GPU single-lattice CMC kernel:
python tools/cmc_grid_lanes_probe.py <factor> <trials> <proposals:lanes,...>"""
import sys, os, tempfile
sys.path.insert(0, '.')
import numpy as np
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
f = int(sys.argv[1]); trials = int(sys.argv[2])
combos = [tuple(int(x) for x in c.split(':')) for c in sys.argv[3].split(',')]
e = capi.Engine(f, n_walkers=1, device=0); e.load_coefficients(js)
occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)
for bs, lanes in combos:
    os.environ['LMC_CMC_GRID_LANES'] = str(lanes)
    e.set_occupancy(occ); e0 = e.total_energy(); e.cmc_reset()
    e.cmc_grid_run(trials // 10, temperature=800.0, seed=5, batch_size=bs)
    out = []
    for rep in range(3):
        s0 = e.cmc_state(); e.cmc_grid_run(trials, temperature=800.0, seed=5, batch_size=bs); ms = e.last_kernel_ms(); s1 = e.cmc_state()
        n = int(s1['steps'][0] - s0['steps'][0])
        out.append('%.3f ms %.3e/s acc %.3f' % (ms, n / ms * 1e3, float((s1['accepted'][0] - s0['accepted'][0]) / n)))
    final = e.get_occupancy(0)
    print('f', f, 'proposals/CTA', bs, 'lanes', lanes, '|', ' | '.join(out), '| bookkeeping diff %.2e' % abs((e.total_energy() - e0) - e.cmc_state()['energy'][0]),
          'composition ok', bool(np.array_equal(np.sort(final), np.sort(occ))), flush=True)

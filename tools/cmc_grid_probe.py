"""Whole-GPU single-lattice CMC probe: python tools/cmc_grid_probe.py <factor> <batch sizes> [trials]
Compares lmc_cmc_grid_run (cooperative grid) with lmc_cmc_run (one cluster) and checks the energy bookkeeping."""
import sys, os, tempfile
sys.path.insert(0, '.')
import numpy as np
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
f = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sizes = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0]
trials = int(sys.argv[3]) if len(sys.argv) > 3 else 200000
e = capi.Engine(f, n_walkers=1, device=0); e.load_coefficients(js)
occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)
for mode in ("grid", "cluster"):
    for bs in sizes:
        e.set_occupancy(occ); e0 = e.total_energy(); e.cmc_reset()
        run = e.cmc_grid_run if mode == "grid" else e.cmc_run
        run(trials // 10, temperature=800.0, seed=5, batch_size=bs)
        for rep in range(2):
            s0 = e.cmc_state(); run(trials, temperature=800.0, seed=5, batch_size=bs); ms = e.last_kernel_ms(); s1 = e.cmc_state()
            n = int(s1['steps'][0] - s0['steps'][0])
            print(mode, 'f', f, 'batch', bs, 'ms', round(ms, 3), 'trials', n, 'rate', round(n / ms * 1e3),
                  'acc', round(float((s1['accepted'][0] - s0['accepted'][0]) / n), 3), flush=True)
        final = e.get_occupancy(0)
        print('   bookkeeping diff', abs((e.total_energy() - e0) - e.cmc_state()['energy'][0]),
              'composition ok', bool(np.array_equal(np.sort(final), np.sort(occ))), flush=True)

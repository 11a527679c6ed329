"""One short lmc_cmc_grid_run launch (for ncu): python tools/cmc_grid_once.py <factor> <trials> [batch]"""
import sys, os, tempfile
sys.path.insert(0, '.')
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
f, trials = int(sys.argv[1]), int(sys.argv[2])
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 0
e = capi.Engine(f, n_walkers=1, device=0); e.load_coefficients(js)
e.set_occupancy(synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)); e.cmc_reset()
e.cmc_grid_run(trials, temperature=800.0, seed=5, batch_size=bs)
print(e.cmc_state(), e.last_kernel_ms())

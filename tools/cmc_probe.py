"""CMC launch-shape probe: python tools/cmc_probe.py <factor> <batch sizes> [trials] [order] [replicas]"""
import sys, os, tempfile
sys.path.insert(0, '.')
import numpy as np
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
f = int(sys.argv[1]) if len(sys.argv) > 1 else 40
sizes = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [0, 256, 64]
trials = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
order = int(sys.argv[4]) if len(sys.argv) > 4 else capi.ORDER_REASSIGNED
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
e = capi.Engine(f, id_order=order, n_walkers=reps, device=0); e.load_coefficients(js)
occ = np.stack([synth.random_alloy(f, 0.02, 0.02, seed=1000 + r, vacancy_site=None) for r in range(reps)])
e.set_occupancy_all(occ); e.cmc_reset()
for bs in sizes:
    e.cmc_run(5000, temperature=800.0, seed=5, batch_size=bs)
    for rep in range(3):
        s0 = e.cmc_state(); e.cmc_run(trials, temperature=800.0, seed=5, batch_size=bs); ms = e.last_kernel_ms(); s1 = e.cmc_state()
        n = int(s1['steps'].sum() - s0['steps'].sum())
        print('f', f, 'order', order, 'reps', reps, 'batch', bs, 'ms', round(ms, 4), 'trials', n, 'rate', round(n / ms * 1e3),
              'acc', round(float((s1['accepted'].sum() - s0['accepted'].sum()) / n), 3), flush=True)

"""Launch-shape sweep of the whole-GPU CMC kernel: python tools/cmc_grid_shape_probe.py <factor> <trials>"""
import sys, os, tempfile
sys.path.insert(0, '.')
import numpy as np
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
f = int(sys.argv[1]); trials = int(sys.argv[2])
e = capi.Engine(f, n_walkers=1, device=0); e.load_coefficients(js)
occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)
h = e.cmc_exchange_handle()
for ctas in (24, 32, 48, 64, 96, 148):
    for bs in (64, 128, 256, 512):
        if ctas * bs // 2 > 65535: continue
        e.cmc_attach_peers(0, 1, [h], ctas)
        e.set_occupancy(occ); e.cmc_reset()
        e.cmc_grid_run(trials // 4, temperature=800.0, seed=5, batch_size=bs)
        s0 = e.cmc_state(); e.cmc_grid_run(trials, temperature=800.0, seed=5, batch_size=bs); ms = e.last_kernel_ms(); s1 = e.cmc_state()
        n = int(s1['steps'][0] - s0['steps'][0])
        print('f', f, 'ctas', ctas, 'threads', bs, 'rate %.3g' % (n / ms * 1e3), flush=True)

"""One low-occupancy first-order KMC launch (for ncu): python tools/kmc_latency_once.py <walkers> <hops> [factor]"""
import sys, os, tempfile
import numpy as np
sys.path.insert(0, '.')
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
nw, hops = int(sys.argv[1]), int(sys.argv[2])
f = int(sys.argv[3]) if len(sys.argv) > 3 else 8
e = capi.Engine(f, n_walkers=nw, device=0); e.load_coefficients(js)
occ = np.stack([synth.random_alloy(f, 0.02, 0.02, seed=42 + w) for w in range(nw)])
e.set_occupancy_all(occ); e.kmc_reset()
e.kmc_run(512, temperatures=np.linspace(400, 600, nw), seed=1)
e.kmc_run(hops, temperatures=np.linspace(400, 600, nw), seed=1)
print(nw, hops, e.last_kernel_ms(), "ms ->", nw * hops / (e.last_kernel_ms() * 1e-3), "hops/s,", e.last_kernel_ms() * 1e3 / hops, "us/step")

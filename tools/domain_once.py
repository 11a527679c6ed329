"""One short lmc_cmc_domain_run launch (for ncu): python tools/domain_once.py <factor> <trials> [lanes] [edge] [replicas]"""
import sys, os, tempfile
import numpy as np
sys.path.insert(0, '.')
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
f, trials = int(sys.argv[1]), int(sys.argv[2])
lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
edge = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 1
e = capi.Engine(f, n_walkers=reps, device=0); e.load_coefficients(js)
e.set_occupancy_all(np.stack([synth.random_alloy(f, 0.02, 0.02, seed=1000 + r, vacancy_site=None) for r in range(reps)])); e.cmc_reset()
warm = int(os.environ.get("WARM_TRIALS", "0"))        # aged state: a first launch of this many trials (profile the second: ncu -s 1)
if warm:
    e.cmc_domain_run(warm, temperature=800.0, seed=5, lanes=lanes, domain_edge=edge)
e.cmc_domain_run(trials, temperature=800.0, seed=5, lanes=lanes, domain_edge=edge)
st = e.cmc_state()
import hashlib
ms = e.last_kernel_ms()
digest = hashlib.sha1(np.ascontiguousarray(e.get_occupancy_all()).tobytes()).hexdigest()[:16]
print(st["steps"].sum(), ms, e.cmc_domain_last_shape())
print("digest", digest, "energy %.12f" % float(np.sum(st["energy"])), "accepted", int(np.sum(st["accepted"])))

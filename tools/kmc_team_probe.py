"""First-order KMC at low occupancy, one launch shape per process (the library reads LMC_KMC_TEAM_LANES once):
    LMC_KMC_TEAM_LANES=16 python tools/kmc_team_probe.py <walkers> <hops> [cells per axis]
prints the kernel time and a digest of the final walker state (vacancy sites, step counts, clock sum)."""
import sys, os, tempfile, hashlib
import numpy as np
sys.path.insert(0, '.')
from latticemontecarlo_b200 import capi, synth
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
nw, hops = int(sys.argv[1]), int(sys.argv[2])
f = int(sys.argv[3]) if len(sys.argv) > 3 else 8
e = capi.Engine(f, n_walkers=nw, device=0); e.load_coefficients(js)
occ = np.stack([synth.random_alloy(f, 0.02, 0.02, seed=42 + w) for w in range(nw)])
e.set_occupancy_all(occ); e.kmc_reset()
temps = np.linspace(400, 600, nw)
e.kmc_run(512, temperatures=temps, seed=1)
e.kmc_run(hops, temperatures=temps, seed=1)
ms = e.last_kernel_ms()
st = e.kmc_state(); t, en, steps, vac = st['time'], st['energy'], st['steps'], st['vacancy']
digest = hashlib.sha1(np.ascontiguousarray(vac).tobytes() + np.ascontiguousarray(steps).tobytes()).hexdigest()[:12]
print(f"lanes={os.environ.get('LMC_KMC_TEAM_LANES', 'auto')} walkers={nw} f={f} hops={hops}: {ms:.3f} ms -> {nw * hops / (ms * 1e-3):.4g} hops/s, "
      f"{ms * 1e3 / hops:.3f} us/step; digest {digest} time_sum {t.sum():.12e} energy_sum {en.sum():.9f}")

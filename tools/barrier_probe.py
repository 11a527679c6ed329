"""Time lmc_eval_barriers_dev (resident inputs) on arbitrary (vacancy, neighbour) events of the walker workload:
python tools/barrier_probe.py [walkers] [reps ...]"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latticemontecarlo_b200 import capi, synth

js = os.path.join(tempfile.mkdtemp(), "c.json")
synth.write_synthetic_json(js)
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
f = 8
e = capi.Engine(f, n_walkers=W, device=0)
e.load_coefficients(js)
vac = np.empty(W, dtype=np.int64)
for w in range(W):
    occ = synth.random_alloy(f, 0.02, 0.02, seed=42 + w)
    e.set_occupancy(occ, walker=w)
    vac[w] = int(np.nonzero(occ == 0)[0][0])
nb = np.stack([e.neighbors(1, int(v)) for v in vac])
rng = np.random.default_rng(3)
for reps in [int(v) for v in sys.argv[2:]] or [1, 16]:
    n = W * 12 * reps
    perm = rng.permutation(n)                                      # events in arbitrary order
    w_all = np.tile(np.repeat(np.arange(W, dtype=np.int32), 12), reps)[perm]
    i_all = np.tile(np.repeat(vac, 12), reps)[perm]
    j_all = np.tile(nb.reshape(-1), reps)[perm]
    d_w, d_i, d_j = torch.from_numpy(w_all).cuda(), torch.from_numpy(i_all).cuda(), torch.from_numpy(j_all).cuda()
    d_ea = torch.empty(n, dtype=torch.float64, device="cuda"); d_de = torch.empty(n, dtype=torch.float64, device="cuda")
    ms = []
    for k in range(8):
        e.eval_barriers_dev(n, d_w.data_ptr(), d_i.data_ptr(), d_j.data_ptr(), d_ea.data_ptr(), d_de.data_ptr())
        ms.append(e.last_kernel_ms())
    t = float(np.mean(ms[3:]))
    print("events %d: %.4f ms  %.3e events/s  frac %.3f  checksum %.9f %.9f" % (n, t, n / t * 1e3, n * 316 / t * 1e3 / 1e9 / 6557.4,
          float(d_ea.sum().item()), float(d_de.sum().item())), flush=True)

"""One launch each of the batched evaluators at the bench shapes (for ncu):
    python tools/eval_once.py events|barriers|swap|cmc_replicas
events: vacancy_events_kernel, 131072 (walker, vacancy) items; barriers: barrier_kernel, 1.57M events grouped by vacancy;
swap: swap_de_rows_kernel, 4.2M random unlike pairs of a 40^3 lattice; cmc_replicas: cmc_run_kernel, 148 replicas of 20^3."""
import sys, os, tempfile
import numpy as np
sys.path.insert(0, '.')
import torch
import bench
from latticemontecarlo_b200 import capi, synth, sharding
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
what = sys.argv[1]
dev = torch.device("cuda")
if what in ("events", "barriers"):
    W = 8192
    e = capi.Engine(bench.FACTOR, n_walkers=W, device=0); e.load_coefficients(js)
    e.set_occupancy_all(bench.walker_occupancy(0, W)); e.kmc_reset()
    e.kmc_run(2048, temperatures=sharding.walker_temperatures(0, W, W), seed=20260101)
    vac = e.kmc_state()["vacancy"]
    reps = 16
    if what == "events":
        d_v = torch.from_numpy(np.tile(vac, reps)).to(dev); d_w = torch.from_numpy(np.tile(np.arange(W, dtype=np.int32), reps)).to(dev)
        n = W * reps
        d_nb = torch.empty(n * 12, dtype=torch.int64, device=dev); d_ea = torch.empty(n * 12, dtype=torch.float64, device=dev); d_de = torch.empty_like(d_ea)
        for _ in range(2):
            e.eval_vacancy_events_dev(n, d_w.data_ptr(), d_v.data_ptr(), d_nb.data_ptr(), d_ea.data_ptr(), d_de.data_ptr())
        print("events", n * 12, e.last_kernel_ms())
    else:
        nb = np.stack([e.neighbors(1, int(v)) for v in vac])
        d_w = torch.from_numpy(np.tile(np.repeat(np.arange(W, dtype=np.int32), 12), reps)).to(dev)
        d_i = torch.from_numpy(np.tile(np.repeat(vac, 12), reps)).to(dev); d_j = torch.from_numpy(np.tile(nb.reshape(-1), reps)).to(dev)
        n = W * 12 * reps
        d_ea = torch.empty(n, dtype=torch.float64, device=dev); d_de = torch.empty_like(d_ea)
        for _ in range(2):
            e.eval_barriers_dev(n, d_w.data_ptr(), d_i.data_ptr(), d_j.data_ptr(), d_ea.data_ptr(), d_de.data_ptr())
        print("barriers", n, e.last_kernel_ms())
elif what == "swap":
    f = 40
    e = capi.Engine(f, device=0); e.load_coefficients(js)
    occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None); e.set_occupancy(occ)
    rng = np.random.default_rng(11); n = 1 << 22
    a = rng.integers(0, occ.size, n); b = rng.integers(0, occ.size, n)
    same = np.nonzero(occ[a] == occ[b])[0]
    while same.size:
        b[same] = rng.integers(0, occ.size, same.size); same = same[occ[a[same]] == occ[b[same]]]
    d_a, d_b = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev); d_de = torch.empty(n, dtype=torch.float64, device=dev)
    for _ in range(2):
        e.eval_swap_de_dev(n, 0, d_a.data_ptr(), d_b.data_ptr(), d_de.data_ptr())
    print("swap", n, e.last_kernel_ms())
elif what == "cmc_replicas":
    f, R = 20, 148
    e = capi.Engine(f, n_walkers=R, device=0); e.load_coefficients(js)
    e.set_occupancy_all(np.stack([synth.random_alloy(f, 0.02, 0.02, seed=1000 + r, vacancy_site=None) for r in range(R)])); e.cmc_reset()
    for _ in range(2):
        e.cmc_run(20000, temperatures=np.linspace(600.0, 1000.0, R), seed=5)
    print("cmc_replicas", e.last_kernel_ms())

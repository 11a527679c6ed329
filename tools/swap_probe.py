"""Time lmc_eval_swap_de_dev (resident inputs) on random unlike-species pairs: python tools/swap_probe.py [factor ...]"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latticemontecarlo_b200 import capi, synth

js = os.path.join(tempfile.mkdtemp(), "c.json")
synth.write_synthetic_json(js)
for f in [int(v) for v in sys.argv[1:]] or [40]:
    e = capi.Engine(f, device=0)
    e.load_coefficients(js)
    occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)
    e.set_occupancy(occ)
    rng = np.random.default_rng(11)
    n = 1 << 22
    a = rng.integers(0, occ.size, n); b = rng.integers(0, occ.size, n)
    same = np.nonzero(occ[a] == occ[b])[0]
    while same.size:
        b[same] = rng.integers(0, occ.size, same.size)
        same = same[occ[a[same]] == occ[b[same]]]
    if os.environ.get("PROBE_PRESORT"):                       # the order grouping would produce, made on the host (no sort, no indirection)
        order = np.argsort(a, kind="stable"); a, b = a[order], b[order]
    d_a, d_b = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    d_de = torch.empty(n, dtype=torch.float64, device="cuda")
    ms = []
    for k in range(8):
        e.eval_swap_de_dev(n, 0, d_a.data_ptr(), d_b.data_ptr(), d_de.data_ptr())
        ms.append(e.last_kernel_ms())
    t = float(np.mean(ms[3:]))
    print("f=%d occ_variant=%s general=%s: %.3f ms  %.3e pairs/s  frac %.3f  checksum %.9f" % (
        f, os.environ.get("LMC_SWAP_OCC", "-"), os.environ.get("LMC_SWAP_GENERAL_KERNEL", "-"), t, n / t * 1e3, n * 440 / t * 1e3 / 1e9 / 6557.4,
        float(d_de.sum().item())))
    e.close()

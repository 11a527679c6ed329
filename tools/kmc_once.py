"""One lmc_kmc_run launch at the bench shape (for ncu): python tools/kmc_once.py [walkers] [hops]"""
import sys, os, tempfile
sys.path.insert(0, '.')
import bench
from latticemontecarlo_b200 import capi, synth, sharding
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
e = capi.Engine(bench.FACTOR, n_walkers=W, device=0); e.load_coefficients(js)
e.set_occupancy_all(bench.walker_occupancy(0, W)); e.kmc_reset()
e.kmc_run(H, temperatures=sharding.walker_temperatures(0, W, W), seed=20260101)
print(e.last_kernel_ms(), W * H / e.last_kernel_ms() * 1e3)

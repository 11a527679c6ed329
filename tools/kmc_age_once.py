"""Several lmc_kmc_run launches at the bench shape, printing the kernel time of each (walkers age from launch to launch):
    LMC_B200_LIB=ab/x.so python tools/kmc_age_once.py [walkers] [hops] [launches]"""
import sys, os, tempfile
sys.path.insert(0, '.')
import bench
from latticemontecarlo_b200 import capi, synth, sharding
W = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
L = int(sys.argv[3]) if len(sys.argv) > 3 else 8
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
e = capi.Engine(bench.FACTOR, n_walkers=W, device=0); e.load_coefficients(js)
e.set_occupancy_all(bench.walker_occupancy(0, W)); e.kmc_reset()
ms = []
for _ in range(L):
    e.kmc_run(H, temperatures=sharding.walker_temperatures(0, W, W), seed=20260101)
    ms.append(e.last_kernel_ms())
st = e.kmc_state()
import hashlib
digest = hashlib.sha1(st["time"].tobytes() + st["energy"].tobytes() + st["vacancy"].tobytes() + st["steps"].tobytes()).hexdigest()[:12]
print(os.environ.get("LMC_B200_LIB", "in-tree"), "handoff", os.environ.get("LMC_KMC_HANDOFF", "default"), " ".join("%.3f" % m for m in ms), "sum %.3f ms" % sum(ms),
      "vacancy xor", int((st["vacancy"] * 7919 % 1000003).sum()), "state sha1", digest, "hybrid" if e.kmc_last_launch_handoff() else "plain")

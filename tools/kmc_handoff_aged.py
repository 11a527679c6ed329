"""Hand-off fraction against walker age: age the bench's walkers, then time launches at several fractions (interleaved).
    python tools/kmc_handoff_aged.py [age_launches] [p_mg] [p_zn]"""
import sys, os, tempfile
sys.path.insert(0, '.')
import bench
from latticemontecarlo_b200 import capi, synth, sharding
age = int(sys.argv[1]) if len(sys.argv) > 1 else 100
p_mg = float(sys.argv[2]) if len(sys.argv) > 2 else bench.P_MG
p_zn = float(sys.argv[3]) if len(sys.argv) > 3 else bench.P_ZN
W, H = 8192, 2048
d = tempfile.mkdtemp(); js = os.path.join(d, 'c.json'); synth.write_synthetic_json(js)
e = capi.Engine(bench.FACTOR, n_walkers=W, device=0); e.load_coefficients(js)
e.set_occupancy_all(bench.walker_occupancy(0, W, p_mg=p_mg, p_zn=p_zn)); e.kmc_reset()
temps = sharding.walker_temperatures(0, W, W)
os.environ.pop("LMC_KMC_HANDOFF", None)
for _ in range(age):
    e.kmc_run(H, temperatures=temps, seed=20260101)
fractions = ["0", "0.2", "0.35", "0.5", "0.65", "0.8"]
ms = {f: [] for f in fractions}
for rep in range(3):
    for f in fractions:
        os.environ["LMC_KMC_HANDOFF"] = f
        e.kmc_run(H, temperatures=temps, seed=20260101)
        ms[f].append(e.last_kernel_ms())
print("age %d launches x %d hops, alloy %.2f/%.2f:" % (age, H, p_mg, p_zn), "  ".join("%s: %s" % (f, " ".join("%.2f" % m for m in ms[f])) for f in fractions))

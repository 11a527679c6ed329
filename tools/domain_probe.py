"""GPU probe of the domain-decomposed CMC driver: correctness invariants + rate per launch shape.
    python tools/domain_probe.py [f ...]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latticemontecarlo_b200 import capi, synth


def check(f, replicas, trials, **kw):
    js = "/tmp/coef_probe.json"
    if not os.path.exists(js):
        synth.write_synthetic_json(js)
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=replicas, device=0)
    e.load_coefficients(js)
    occ = np.stack([synth.random_alloy(f, 0.05, 0.05, seed=10 + r, vacancy_site=None) for r in range(replicas)])
    e.set_occupancy_all(occ)
    e0 = np.array([e.total_energy(w) for w in range(replicas)])
    e.cmc_reset()
    temps = np.linspace(600.0, 1200.0, replicas) if replicas > 1 else None
    e.cmc_domain_run(trials, temperature=800.0, temperatures=temps, seed=3, **kw)
    e.cmc_domain_run(trials, temperature=800.0, temperatures=temps, seed=3, **kw)
    st = e.cmc_state()
    e1 = np.array([e.total_energy(w) for w in range(replicas)])
    fin = e.get_occupancy_all()
    ok_e = np.max(np.abs((e1 - e0) - st["energy"]))
    cons = all(np.array_equal(np.sort(fin[w]), np.sort(occ[w])) for w in range(replicas))
    print("f=%d replicas=%d %s: steps %s acc ratio %.3f  |dE bookkeeping| %.2e  conserved %s  shape %s" % (
        f, replicas, kw, st["steps"][:3], st["accepted"].sum() / max(1, st["steps"].sum()), ok_e, cons, e.cmc_domain_last_shape()), flush=True)
    # the A + B table form (LMC_CMC_DOMAIN_TABLES=0) and another lane count must give the same trajectory
    ref_state = (fin.copy(), st["energy"].copy(), st["steps"].copy())
    for env, extra in (({"LMC_CMC_DOMAIN_TABLES": "0"}, {}), ({}, {"lanes": 8, "speculate": 1}), ({}, {"lanes": 8, "speculate": 4}), ({}, {"lanes": 16, "speculate": 2}),
                       ({}, {"lanes": 32, "speculate": 1}), ({"LMC_CMC_DOMAIN_PASSES": "3"}, {"lanes": 8, "speculate": 2})):
        os.environ.update(env)
        e.set_occupancy_all(occ); e.cmc_reset()
        kw2 = dict(kw); kw2.update(extra)
        e.cmc_domain_run(trials, temperature=800.0, temperatures=temps, seed=3, **kw2)
        e.cmc_domain_run(trials, temperature=800.0, temperatures=temps, seed=3, **kw2)
        for k in env: del os.environ[k]
        st3 = e.cmc_state(); fin3 = e.get_occupancy_all()
        print("   variant %s %s: same occupancy %s, same steps %s, max |dE_total diff| %.2e" % (env, extra, np.array_equal(fin3, ref_state[0]),
              np.array_equal(st3["steps"], ref_state[2]), np.max(np.abs(st3["energy"] - ref_state[1]))), flush=True)
    # the batched driver keeps working on the result
    e.cmc_run(2000, temperature=800.0, temperatures=temps, seed=1)
    e2 = np.array([e.total_energy(w) for w in range(replicas)])
    st2 = e.cmc_state()
    print("   after batched cmc_run: |bookkeeping| %.2e" % np.max(np.abs((e2 - e0) - st2["energy"])), flush=True)
    e.close()


def rate(f, replicas, trials, p=0.02, sa=None, **kw):
    js = "/tmp/coef_probe.json"
    if not os.path.exists(js):
        synth.write_synthetic_json(js)
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=replicas, device=0)
    e.load_coefficients(js)
    occ = np.stack([synth.random_alloy(f, p, p, seed=1000 + r, vacancy_site=None) for r in range(replicas)])
    e.set_occupancy_all(occ)
    e.cmc_reset(*(sa or ()))
    temps = np.linspace(600.0, 1000.0, replicas) if replicas > 1 else None
    e.cmc_domain_run(trials // 4, temperatures=temps, seed=5, **kw)
    ms, done = [], []
    for _ in range(3):
        s0 = e.cmc_state()["steps"].sum()
        e.cmc_domain_run(trials, temperatures=temps, seed=5, **kw)
        ms.append(e.last_kernel_ms())
        done.append(int(e.cmc_state()["steps"].sum() - s0))
    st = e.cmc_state()
    print("rate f=%d x%d p=%.2f %s: %.3e trials/s  (%.2f ms, %d trials/launch, acc %.3f) shape %s" % (
        f, replicas, p, kw, sum(done) / (sum(ms) * 1e-3), np.mean(ms), np.mean(done), st["accepted"].sum() / max(1, st["steps"].sum()),
        e.cmc_domain_last_shape()), flush=True)
    e.close()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    if mode in ("all", "check"):
        check(4, 1, 2000)
        check(6, 3, 5000)
        check(12, 4, 20000)
        check(12, 4, 20000, domain_edge=6, lanes=8)
        check(12, 4, 20000, domain_edge=12, lanes=16)
        check(10, 2, 20000, domain_edge=7, lanes=4)
        check(24, 1, 200000)
        check(24, 1, 200000, lanes=2)
    if mode in ("all", "rate"):
        for f in (100, 50, 40):
            for edge, rounds in ((8, 0), (6, 0), (6, 128), (6, 216), (10, 0)):
                rate(f, 1, 8 * 4 * f ** 3, domain_edge=edge, rounds_per_sweep=rounds)

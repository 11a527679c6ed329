"""Reference ensemble for the batched SimulatedAnnealing drivers: the unmodified mc::SimulatedAnnealing (oracle/_ref,
/root/reference/lmc/mc/src/SimulatedAnnealing.cpp:99-185) on the 6x6x6 cell of make_golden_cmc_stat.py, 16 seeds x 3e5
trials from T0 = 1500 K.  Per seed: final energy (relative to the common initial configuration), final temperature and
the final first-neighbour Warren-Cowley parameters.  Output: tests/golden/golden_sa_stat_v1.npz.

    python tests/golden/make_golden_sa_stat.py            (about 5 minutes on 8 cores)
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_cmc_stat import FACTOR, P_MG, P_ZN, OCC_SEED, JSON_SEED, warren_cowley  # noqa: E402

T0, MAX_STEPS, N_SEEDS = 1500.0, 300000, 16


def run_seed(seed):
    from oracle import ref_lib as R
    from latticemontecarlo_b200 import synth
    js = "/tmp/golden_sa_stat_%d.json" % os.getpid()
    synth.write_synthetic_json(js, seed=JSON_SEED)
    occ = synth.random_alloy(FACTOR, P_MG, P_ZN, seed=OCC_SEED, vacancy_site=None)
    counts = {"Mg": int(np.count_nonzero(occ == 2)), "Zn": int(np.count_nonzero(occ == 3))}
    r = R.simulated_annealing(FACTOR, "Al", counts, occ, js, initial_temperature=T0, maximum_steps=MAX_STEPS, seed=500 + seed, trace=False)
    nn1 = R.RefConfig.fcc(FACTOR, occ, reassign=False).neighbors(1)
    return seed, r["final_energy"] - r["energy0"], r["final_temperature"], warren_cowley(r["final_occ"], nn1)


def main():
    from oracle import ref_lib as R
    if not R.build():
        raise SystemExit("oracle/_ref not available")
    t0 = time.time()
    with mp.Pool(min(N_SEEDS, os.cpu_count() or 1)) as pool:
        res = sorted(pool.map(run_seed, range(N_SEEDS)))
    out = os.path.join(ROOT, "tests", "golden", "golden_sa_stat_v1.npz")
    np.savez(out, params=np.array([FACTOR, P_MG, P_ZN, OCC_SEED, T0, MAX_STEPS, N_SEEDS, JSON_SEED], dtype=np.float64),
             final_energy=np.array([r[1] for r in res]), final_temperature=np.array([r[2] for r in res]), sro=np.array([r[3] for r in res]))
    print("wrote %s in %.0f s" % (out, time.time() - t0))
    for r in res:
        print(r)


if __name__ == "__main__":
    main()

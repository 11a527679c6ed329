"""Reference ensemble for the statistical parity of the CMC drivers whose chains are NOT replays of the reference's.

Runs the unmodified mc::CanonicalMcSerial (oracle/_ref, /root/reference/lmc/mc/src/CanonicalMcSerial.cpp:40-51) on a
6x6x6 Al-Mg-Zn cell at fixed temperature: 16 seeds x 2e6 trials, in chunks of 50 000 trials (each chunk continues from the
previous chunk's final occupancy with a fresh seed; the reference reseeds from the clock anyway).  Per seed, after a
burn-in of 2e5 trials: mean and variance of the energy (per-trial series, relative to the common initial configuration),
acceptance ratio, and the first-neighbour Warren-Cowley parameters of Mg-Mg, Zn-Zn and Mg-Zn pairs averaged over the
chunk-end snapshots.  Output: tests/golden/golden_cmc_stat_v1.npz (the GPU tests compare ensemble means within 3 sigma).

    python tests/golden/make_golden_cmc_stat.py            (about 25 minutes on 8 cores)
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

FACTOR, P_MG, P_ZN, OCC_SEED = 6, 0.06, 0.06, 77
TEMPERATURE = 1500.0
N_SEEDS, N_TRIALS, CHUNK, BURN_IN = 16, 2000000, 50000, 200000
JSON_SEED = 20240611


def warren_cowley(occ, nn1):
    """alpha_ij = 1 - P(j | neighbour of i) / c_j over first neighbours, for (Mg,Mg), (Zn,Zn), (Mg,Zn); codes Mg=2, Zn=3."""
    out = []
    for i, j in ((2, 2), (3, 3), (2, 3)):
        sites = np.nonzero(occ == i)[0]
        p = np.mean(occ[nn1[sites]] == j)
        out.append(1.0 - p / np.mean(occ == j))
    return out


def run_seed(seed):
    from oracle import ref_lib as R
    from latticemontecarlo_b200 import synth
    js = "/tmp/golden_cmc_stat_%d.json" % os.getpid()
    synth.write_synthetic_json(js, seed=JSON_SEED)
    occ = synth.random_alloy(FACTOR, P_MG, P_ZN, seed=OCC_SEED, vacancy_site=None)
    cfg = R.RefConfig.fcc(FACTOR, occ, reassign=False)
    nn1 = cfg.neighbors(1)
    offset, e_sum, e_sq, n, acc, sro = 0.0, 0.0, 0.0, 0, 0, []
    done = 0
    while done < N_TRIALS:
        cfg = R.RefConfig.fcc(FACTOR, occ, reassign=False)
        r = R.cmc_serial(cfg, js, temperature=TEMPERATURE, maximum_steps=CHUNK - 1, seed=100000 * (seed + 1) + done // CHUNK)
        series = offset + r["energy_before"][:CHUNK]
        if done >= BURN_IN:
            e_sum += series.sum(); e_sq += (series ** 2).sum(); n += CHUNK
            acc += int(np.count_nonzero(np.diff(np.append(series, offset + r["final_energy"]))))
            sro.append(warren_cowley(r["final_occ"], nn1))
        offset += r["final_energy"]
        occ = r["final_occ"]
        done += CHUNK
    mean = e_sum / n
    return seed, mean, e_sq / n - mean * mean, acc / n, np.mean(np.array(sro), axis=0)


def main():
    from oracle import ref_lib as R
    if not R.build():
        raise SystemExit("oracle/_ref not available")
    t0 = time.time()
    with mp.Pool(min(N_SEEDS, os.cpu_count() or 1)) as pool:
        res = sorted(pool.map(run_seed, range(N_SEEDS)))
    out = os.path.join(ROOT, "tests", "golden", "golden_cmc_stat_v1.npz")
    np.savez(out, params=np.array([FACTOR, P_MG, P_ZN, OCC_SEED, TEMPERATURE, N_SEEDS, N_TRIALS, CHUNK, BURN_IN, JSON_SEED], dtype=np.float64),
             mean_energy=np.array([r[1] for r in res]), var_energy=np.array([r[2] for r in res]),
             accept_ratio=np.array([r[3] for r in res]), sro=np.array([r[4] for r in res]))
    print("wrote %s in %.0f s" % (out, time.time() - t0))
    for r in res:
        print(r)


if __name__ == "__main__":
    main()

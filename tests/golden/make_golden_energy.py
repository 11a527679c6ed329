"""Golden values of the EnergyPredictor entry points beyond GetEnergy (pred/include/EnergyPredictor.h:19-25), from the
reference itself (oracle/_ref) with the coefficient file of golden_v1.npz: GetEncode, GetEnergyOfCluster /
GetEncodeOfCluster for several atom-id lists, GetChemicalPotential(Al).  Output: tests/golden/golden_energy_v1.npz.
    python tests/golden/make_golden_energy.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from latticemontecarlo_b200 import synth  # noqa: E402
from oracle import ref_lib as R  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    assert R.build()
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), allow_pickle=False)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        js = H.golden_json(golden, d)
        pred = R.RefEnergy(js)
        rng = np.random.default_rng(5)
        for tag, f, reassign in (("A", 5, True), ("B", 6, False)):
            occ = synth.random_alloy(f, 0.07, 0.07, seed=500 + f)              # one vacancy
            cfg = R.RefConfig.fcc(f, occ, reassign=reassign)
            l2a, a2l, _ = cfg.maps()
            out[tag + "_params"] = np.array([f, int(reassign)])
            out[tag + "_occ"] = cfg.occupancy()
            out[tag + "_lattice_of_atom"] = a2l
            e, enc = pred.energy(cfg)
            out[tag + "_energy"] = np.array([e]); out[tag + "_encode"] = enc
            n = cfg.num_sites
            lists = [np.array([7]), rng.choice(n, 5, replace=False), rng.choice(n, 40, replace=False), np.arange(n), np.array([int(l2a[cfg.vacancy()])])]
            out[tag + "_n_lists"] = np.array([len(lists)])
            for k, ids in enumerate(lists):
                e, enc = pred.energy_of_cluster(cfg, ids)
                out["%s_list%d" % (tag, k)] = ids.astype(np.int64)
                out["%s_cluster_energy%d" % (tag, k)] = np.array([e])
                out["%s_cluster_encode%d" % (tag, k)] = enc
        # SimulatedAnnealing's reference energy (mc/src/SimulatedAnnealing.cpp:58-70): energy_ at construction for the
        # occupancy of golden_v1's SA case = [E(config) - E(pure solvent)] - sum_e mu_e * count_e
        occ_sa = golden["SA_occ"]
        f_sa = int(golden["SA_params"][0])
        cnt = {"Mg": int((occ_sa == 2).sum()), "Zn": int((occ_sa == 3).sum())}
        tr = R.simulated_annealing(f_sa, "Al", cnt, occ_sa, js, initial_temperature=700.0, maximum_steps=1, seed=9, trace=False)
        out["SA_initial_energy"] = np.array([tr["energy0"]])
        print("SA initial_energy", tr["energy0"])
        mu = pred.chemical_potential("Al")
        out["mu_elements"] = np.array(list(mu.keys()), dtype=np.int32)
        out["mu_values"] = np.array(list(mu.values()))
        print("chemical potential", mu)
    path = os.path.join(ROOT, "tests", "golden", "golden_energy_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()

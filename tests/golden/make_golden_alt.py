"""Generate tests/golden/golden_alt_v1.npz: outputs of the reference's ALTERNATIVE predictors --
pred::VacancyMigrationPredictorE0 / E0Lru, pred::EnergyChangePredictorPair, pred::EnergyChangePredictorSite (unmodified
sources compiled by oracle/Makefile) -- on the two small cases of golden_v1.npz (same start configurations, same candidate
events), with synthetic E0-format coefficient files (latticemontecarlo_b200.synth.synthetic_coefficients_e0; the second
file has a low mu_e0 so that the max(0, .) clamp of VacancyMigrationPredictorE0.cpp:157 is exercised).
Run from the repo root where /root/reference exists:   python tests/golden/make_golden_alt.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from latticemontecarlo_b200 import synth  # noqa: E402
from oracle import ref_lib as R  # noqa: E402

ELEMENTS = ("Al", "Mg", "Zn")


def main():
    assert R.build(), "needs /root/reference to compile oracle/_ref"
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    tmp = tempfile.mkdtemp()
    out = {}
    files = {}
    for name, kw in (("e0", {}), ("e0low", dict(seed=77, mu_e0=float(np.log(0.08))))):
        co = synth.synthetic_coefficients_e0(k_mmm=6, **kw)
        files[name] = os.path.join(tmp, name + ".json")
        with open(files[name], "w") as f:
            json.dump(co, f)
        for top, block in co.items():
            for key, val in block.items():
                out["coef__%s__%s__%s" % (name, top, key)] = np.asarray(val, dtype=np.float64)
    rng = np.random.default_rng(2024)
    for tag in ("A", "B"):
        f, reassign = (int(v) for v in golden[tag + "_factor"])
        vac, I, J = golden[tag + "_ev_vac"], golden[tag + "_ev_i"], golden[tag + "_ev_j"]
        base = golden[tag + "_ev_base_occ"]
        cfg = R.RefConfig.fcc(f, golden[tag + "_occ_generate_order"], reassign=bool(reassign))
        for k in range(len(base)):
            cfg.set_element(k, int(base[k]))
        for name in files:
            pred = R.RefE0(files[name], cfg, ELEMENTS)
            lru = R.RefE0(files[name], cfg, ELEMENTS, lru_size=64)
            ea, de, e0 = (np.empty(len(I)) for _ in range(3))
            for k in range(len(I)):                       # one vacancy at a time, like golden_v1's event set
                cfg.set_element(int(vac[k]), 0)
                a, d, e = pred.eval(cfg, [I[k]], [J[k]])
                a2, d2, _ = lru.eval(cfg, [I[k]], [J[k]])
                a3, _, _ = lru.eval(cfg, [I[k]], [J[k]])  # second call: served from the cache
                assert a2[0] == a[0] and d2[0] == d[0] and a3[0] == a[0]
                ea[k], de[k], e0[k] = a[0], d[0], e[0]
                cfg.set_element(int(vac[k]), int(base[vac[k]]))
            out["%s_%s_Ea" % (tag, name)], out["%s_%s_dE" % (tag, name)], out["%s_%s_e0" % (tag, name)] = ea, de, e0
        # EnergyChangePredictorPair: first-neighbour pairs of the vacancy-free base occupancy, and of the configuration with a vacancy
        n = cfg.num_sites
        nn1 = cfg.neighbors(1)
        pair = R.RefPair(files["e0"], cfg, ELEMENTS)
        site = R.RefSite(files["e0"], cfg, ELEMENTS)
        a = rng.integers(0, n, 96)
        b = nn1[a, rng.integers(0, 12, 96)]
        out[tag + "_pair_a"], out[tag + "_pair_b"] = a, b
        out[tag + "_pair_dE"] = pair.de_pair(cfg, a, b)
        v = int(vac[0])
        cfg.set_element(v, 0)
        av = np.full(12, v)
        out[tag + "_pairvac_site"] = np.int64(v)
        out[tag + "_pairvac_dE"] = pair.de_pair(cfg, av, nn1[v])          # vacancy <-> each first neighbour
        out[tag + "_pairvac_dE_reversed"] = pair.de_pair(cfg, nn1[v], av)
        cfg.set_element(v, int(base[v]))
        s = rng.integers(0, n, 64)
        c = rng.integers(0, 4, 64).astype(np.uint8)                       # includes X (a site turned into a vacancy)
        out[tag + "_site_id"], out[tag + "_site_new"] = s, c
        out[tag + "_site_dE"] = site.de_site(cfg, s, c)
    path = os.path.join(ROOT, "tests", "golden", "golden_alt_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f kB" % (os.path.getsize(path) / 1e3), len(out), "arrays")
    for k in sorted(out):
        if not k.startswith("coef"):
            print(k, np.asarray(out[k]).shape, float(np.min(out[k])), float(np.max(out[k])))


if __name__ == "__main__":
    main()

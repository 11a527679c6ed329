"""Generate tests/golden/golden_v1.npz from the reference itself (oracle/_ref/liblmc_ref.so, i.e. the
unmodified sources under /root/reference compiled by oracle/Makefile).

The reference ships no tests, fixtures or known-answer vectors, so these are produced by
running its own predictors / drivers on small seeded inputs.  Run from the repo root where
/root/reference exists:   python tests/golden/make_golden.py
The file is committed; the GPU box (which has no /root/reference) only reads it.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from latticemontecarlo_b200 import synth  # noqa: E402
from oracle import ref_lib as R  # noqa: E402

ELEMENTS = ("Al", "Mg", "Zn")


def flatten_coefficients(co, out):
    for top, body in co.items():
        for key, val in body.items():
            out["coef__%s__%s" % (top, key)] = np.asarray(val, dtype=np.float64)


def flat_mapping(groups):
    """canonical (clusters sorted inside groups) flattening: G, then per group C, L, entries."""
    flat = [len(groups)]
    for g in groups:
        g = sorted(g)
        flat += [len(g), len(g[0])]
        for c in g:
            flat += list(c)
    return np.asarray(flat, dtype=np.int64)


def main():
    assert R.build(), "needs /root/reference to compile oracle/_ref"
    out = {}
    tmp = tempfile.mkdtemp()
    co = synth.synthetic_coefficients(seed=20240611, elements=ELEMENTS, k_mmm=4, k_mm2=5)
    flatten_coefficients(co, out)
    js = os.path.join(tmp, "coef.json")
    import json
    with open(js, "w") as f:
        json.dump(co, f)
    tt = os.path.join(tmp, "tt.dat")
    synth.write_time_temperature(tt)
    out["tt_points"] = np.array([[0.0, 300.0], [1e-3, 500.0], [1e-1, 700.0]])

    types = R.cluster_types(ELEMENTS)
    out["cluster_types"] = np.array([[t[0], len(t[1])] + list(t[1]) + [-1] * (3 - len(t[1])) for t in types], dtype=np.int32)

    rng = np.random.default_rng(1234)
    for tag, f, reassign, p in (("A", 4, True, 0.06), ("B", 5, False, 0.03)):
        occ_gen = synth.random_alloy(f, p, p, seed=100 + f)
        cfg = R.RefConfig.fcc(f, occ_gen, reassign=reassign)
        n = cfg.num_sites
        out[tag + "_factor"] = np.array([f, int(reassign)], dtype=np.int32)
        out[tag + "_occ_generate_order"] = occ_gen
        out[tag + "_occ"] = cfg.occupancy()
        out[tag + "_positions"] = cfg.positions()
        for s in (1, 2, 3):
            out["%s_nn%d" % (tag, s)] = cfg.neighbors(s).astype(np.int32)
        nn1 = cfg.neighbors(1)
        # ordered lists for a stride of all 12N pairs + every site
        pi = np.repeat(np.arange(n), 12)[::5]
        pj = nn1.ravel()[::5]
        lists = [cfg.pair_lists(int(a), int(b)) for a, b in zip(pi, pj)]
        out[tag + "_pair_i"] = pi.astype(np.int32)
        out[tag + "_pair_j"] = pj.astype(np.int32)
        out[tag + "_list_state"] = np.array([l[0] for l in lists], dtype=np.int32)
        out[tag + "_list_mmm"] = np.array([l[1] for l in lists], dtype=np.int32)
        out[tag + "_list_mm2"] = np.array([l[2] for l in lists], dtype=np.int32)
        out[tag + "_list_site"] = np.array([cfg.site_list(k) for k in range(n)], dtype=np.int32)
        for name in ("state_pair", "state_site", "mmm", "mm2"):
            out["%s_mapping_%s" % (tag, name)] = flat_mapping(cfg.mapping(name))

        # barrier events: move the single vacancy to 20 different sites, all 12 jumps each
        quartic = R.RefQuartic(js, cfg, ELEMENTS)
        base = cfg.occupancy()
        base[base == 0] = 1
        ev = {k: [] for k in ("vac", "i", "j", "Ea", "dE", "D", "Ks", "start", "end", "mmm", "mm2f", "mm2b")}
        for v in rng.choice(n, 20, replace=False):
            c2 = cfg.clone()
            for l in np.nonzero(cfg.occupancy() == 0)[0]:
                c2.set_element(int(l), 1)
            c2.set_element(int(v), 0)
            ea, de = quartic.eval(c2, np.full(12, v), nn1[v])
            for q in range(12):
                p_ = quartic.parts(c2, int(v), int(nn1[v][q]))
                assert p_["dE"] == de[q]
                ev["vac"].append(v); ev["i"].append(v); ev["j"].append(nn1[v][q])
                ev["Ea"].append(ea[q]); ev["dE"].append(de[q]); ev["D"].append(p_["D"]); ev["Ks"].append(p_["Ks"])
                ev["start"].append(p_["start_counts"]); ev["end"].append(p_["end_counts"])
                if q < 2:
                    ev["mmm"].append(p_["enc_mmm"]); ev["mm2f"].append(p_["enc_mm2_f"]); ev["mm2b"].append(p_["enc_mm2_b"])
        out[tag + "_ev_base_occ"] = base
        for k in ("vac", "i", "j"):
            out["%s_ev_%s" % (tag, k)] = np.asarray(ev[k], dtype=np.int32)
        for k in ("Ea", "dE", "D", "Ks"):
            out["%s_ev_%s" % (tag, k)] = np.asarray(ev[k], dtype=np.float64)
        out[tag + "_ev_start_counts"] = np.asarray(ev["start"], dtype=np.int16)
        out[tag + "_ev_end_counts"] = np.asarray(ev["end"], dtype=np.int16)
        for k in ("mmm", "mm2f", "mm2b"):   # rows 0,1 of every vacancy position
            out["%s_ev_enc_%s" % (tag, k)] = np.asarray(ev[k], dtype=np.float64)

        # swap / site energy changes on the config as it is (with its vacancy)
        ps = R.RefPairSite(js, cfg, ELEMENTS)
        a = np.concatenate([rng.integers(0, n, 300), np.arange(30), np.arange(30, 50), np.arange(50, 70)])
        b = np.concatenate([rng.integers(0, n, 300), nn1[np.arange(30), 4], cfg.neighbors(2)[np.arange(30, 50), 2],
                            cfg.neighbors(3)[np.arange(50, 70), 11]])
        no_vac = (cfg.occupancy()[a] != 0) | (cfg.occupancy()[b] != 0)
        a, b = a[no_vac], b[no_vac]
        out[tag + "_swap_a"] = a.astype(np.int32)
        out[tag + "_swap_b"] = b.astype(np.int32)
        out[tag + "_swap_dE"] = ps.de_pair(cfg, a, b)
        energy, enc = R.RefEnergy(js, ELEMENTS).energy(cfg)
        out[tag + "_energy"] = np.array([energy])
        out[tag + "_energy_encode"] = enc

        # driver traces
        for name, kw in (("kmc", dict(temperature=500.0, seed=11)),
                         ("kmc_tt", dict(temperature=500.0, seed=12, tt_file=tt, rate_corrector=True))):
            tr = R.kmc_first_omp(cfg, js, ELEMENTS, maximum_steps=80, threads=1, **kw)
            for k in ("u1", "u2", "from", "to", "slot", "dt", "time", "energy", "Ea", "dE", "temperature", "total_rate",
                      "final_occ"):
                out["%s_%s_%s" % (tag, name, k)] = tr[k]
        occ_novac = synth.random_alloy(f, 0.10, 0.10, seed=200 + f, vacancy_site=None)
        cfg2 = R.RefConfig.fcc(f, occ_novac, reassign=reassign)
        out[tag + "_cmc_occ"] = cfg2.occupancy()
        # single-site changes (incl. -> X, as used by lmc/ansys) on the vacancy-free config
        ps2 = R.RefPairSite(js, cfg2, ELEMENTS)
        sites = rng.integers(0, n, 200)
        new = rng.integers(0, 4, 200).astype(np.uint8)
        out[tag + "_site"] = sites.astype(np.int32)
        out[tag + "_site_new"] = new
        out[tag + "_site_dE"] = ps2.de_site(cfg2, sites, new)
        sc = [ps2.site_counts(cfg2, int(s), int(e)) for s, e in zip(sites[:40], new[:40])]
        out[tag + "_site_start_counts"] = np.asarray([c[1] for c in sc], dtype=np.int16)
        out[tag + "_site_end_counts"] = np.asarray([c[2] for c in sc], dtype=np.int16)
        tr = R.cmc_serial(cfg2, js, ELEMENTS, temperature=800.0, maximum_steps=400, seed=5)
        for k in ("a", "b", "u", "dE", "energy_before", "final_occ"):
            out["%s_cmc_%s" % (tag, k)] = tr[k]
    # simulated annealing (GenerateFCC order by construction)
    f = 5
    occ_sa = synth.random_alloy(f, 0.08, 0.08, seed=300, vacancy_site=None)
    cnt = {"Mg": int((occ_sa == 2).sum()), "Zn": int((occ_sa == 3).sum())}
    tr = R.simulated_annealing(f, "Al", cnt, occ_sa, js, initial_temperature=700.0, maximum_steps=4000, seed=9)
    out["SA_occ"] = occ_sa
    out["SA_params"] = np.array([f, 700.0, 4000])
    for k in ("a", "b", "u", "temperature_before", "final_occ"):
        out["SA_" + k] = tr[k]
    out["SA_energy_before"] = tr["energy_before"] - tr["energy0"]
    out["SA_final"] = np.array([tr["final_energy"] - tr["energy0"], tr["final_temperature"]])

    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f kB" % (os.path.getsize(path) / 1e3), len(out), "arrays")


if __name__ == "__main__":
    main()

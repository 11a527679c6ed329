"""the reference's alternative predictors -- VacancyMigrationPredictorE0[Lru], EnergyChangePredictorPair,
EnergyChangePredictorSite -- against tests/golden/golden_alt_v1.npz (written by the compiled reference,
tests/golden/make_golden_alt.py).  CPU tests pin the numpy oracle and the contracted host tables; the gpu tests are the
parity tests of the CUDA path through the C ABI."""
import json
import os

import numpy as np
import pytest

from tests import helpers as H
from latticemontecarlo_b200 import capi
from oracle import lmc_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE_OF_ENUM = {1: 0, 2: 1, 3: 2, 0: 3}
CASES = [("A", capi.ORDER_REASSIGNED), ("B", capi.ORDER_GENERATE)]


@pytest.fixture(scope="module")
def alt():
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_alt_v1.npz"), allow_pickle=False)


@pytest.fixture(scope="module")
def alt_json(alt, tmp_path_factory):
    d = tmp_path_factory.mktemp("alt")
    paths = {}
    for name in ("e0", "e0low"):
        co = {}
        for key in alt.files:
            if key.startswith("coef__" + name + "__"):
                _, _, top, k = key.split("__")
                co.setdefault(top, {})[k] = float(alt[key]) if alt[key].ndim == 0 else alt[key].tolist()
        paths[name] = str(d / (name + ".json"))
        with open(paths[name], "w") as f:
            json.dump(co, f)
    return paths


def _event_config(golden, tag, k):
    occ = golden[tag + "_ev_base_occ"].copy()
    occ[golden[tag + "_ev_vac"][k]] = 0
    return occ


# ------------------------------------------------------------------------------------------------ oracle (CPU)
@pytest.mark.parametrize("tag,order", CASES)
def test_oracle_e0_matches_reference(golden, alt, alt_json, tag, order):
    cfg = H.oracle_config(golden, tag)
    for name in ("e0", "e0low"):
        pred = O.VacancyMigrationPredictorE0(alt_json[name], cfg, H.CODES)
        I, J = golden[tag + "_ev_i"], golden[tag + "_ev_j"]
        for k in range(0, len(I), 5):
            cfg.occ = _event_config(golden, tag, k)
            ea, de = pred.barrier_and_diff(cfg, I[k], J[k])
            assert abs(de[0] - alt["%s_%s_dE" % (tag, name)][k]) < 1e-12
            assert abs(ea[0] - alt["%s_%s_Ea" % (tag, name)][k]) < 1e-12
            assert abs(pred.get_e0(cfg, I[k], J[k])[0] - alt["%s_%s_e0" % (tag, name)][k]) < 1e-12
    assert (alt[tag + "_e0low_Ea"] == 0).any() and (alt[tag + "_e0low_Ea"] > 0).any()      # the clamp is exercised


@pytest.mark.parametrize("tag,order", CASES)
def test_oracle_pair_and_site_match_reference(golden, alt, alt_json, tag, order):
    cfg = H.oracle_config(golden, tag, occ=golden[tag + "_ev_base_occ"])
    pair = O.EnergyChangePredictorPair(alt_json["e0"], cfg, H.CODES)
    assert np.max(np.abs(pair.de_pair(cfg, alt[tag + "_pair_a"], alt[tag + "_pair_b"]) - alt[tag + "_pair_dE"])) < 1e-12
    site = O.EnergyChangePredictorSite(alt_json["e0"], cfg, H.CODES)
    assert np.max(np.abs(site.de_site(cfg, alt[tag + "_site_id"], alt[tag + "_site_new"]) - alt[tag + "_site_dE"])) < 1e-12
    v = int(alt[tag + "_pairvac_site"])
    cfg.occ = cfg.occ.copy()
    cfg.occ[v] = 0
    assert np.max(np.abs(pair.de_pair(cfg, np.full(12, v), cfg.nn[0][v]) - alt[tag + "_pairvac_dE"])) < 1e-12
    assert np.array_equal(alt[tag + "_pairvac_dE"], alt[tag + "_pairvac_dE_reversed"])
    with pytest.raises(IndexError):
        pair.de_pair(cfg, [v], [cfg.nn[1][v][0]])


# ------------------------------------------------------------------------------------------------ host tables (CPU)
@pytest.mark.parametrize("tag,order", CASES)
def test_contracted_e0_tables_reproduce_reference(golden, alt, alt_json, tag, order):
    """E0 model in the contracted form: slot 0 = dE, slot 1 = 0, slot 2 = log e0; Ea = max(0, e0 + dE/2)."""
    e = capi.Engine(int(golden[tag + "_factor"][0]), id_order=order, device=-1)
    for name in ("e0", "e0low"):
        e.load_coefficients(alt_json[name], model=capi.BARRIER_E0)
        assert e.barrier_model() == capi.BARRIER_E0
        T, pairs = e.get_tables(), capi.tables_env_pairs("pair")
        I, J = golden[tag + "_ev_i"], golden[tag + "_ev_j"]
        for k in range(0, len(I), 3):
            occ = _event_config(golden, tag, k)
            s = e.pair_lists(I[k], J[k])[0]
            codes = np.array([CODE_OF_ENUM[c] for c in occ[np.delete(s, [21, 38])]])
            mig = CODE_OF_ENUM[occ[J[k]]]
            q = T["pair_C"][mig].copy()
            for t in np.nonzero(codes != 0)[0]:
                q += T["pair_A"][mig, t, codes[t]]
            for p, (t, u) in enumerate(pairs):
                if codes[t] != 0 and codes[u] != 0:
                    q += T["pair_B"][mig, p, codes[t], codes[u]]
            assert q[1] == 0.0
            assert abs(q[0] - alt["%s_%s_dE" % (tag, name)][k]) < 1e-12
            assert abs(np.exp(q[2]) - alt["%s_%s_e0" % (tag, name)][k]) < 1e-12
            assert abs(max(0.0, np.exp(q[2]) + q[0] / 2) - alt["%s_%s_Ea" % (tag, name)][k]) < 1e-9


def test_model_selection_errors(alt_json, coef_json):
    e = capi.Engine(4, device=-1)
    e.load_coefficients(alt_json["e0"])                       # quartic loader on an E0 file: dE tables only
    assert e.barrier_model() == -1
    e.load_coefficients(coef_json)
    assert e.barrier_model() == capi.BARRIER_QUARTIC
    e.load_coefficients(coef_json, model=capi.BARRIER_E0)     # a quartic file has no theta_e0
    assert e.barrier_model() == -1
    with pytest.raises(capi.LmcInvalidArgument):
        e.load_coefficients(coef_json, model=7)


# ------------------------------------------------------------------------------------------------ CUDA path
@pytest.mark.gpu
@pytest.mark.parametrize("tag,order", CASES)
def test_gpu_e0_barriers_match_reference(golden, alt, alt_json, tag, order):
    f = int(golden[tag + "_factor"][0])
    I, J, vac = golden[tag + "_ev_i"], golden[tag + "_ev_j"], golden[tag + "_ev_vac"]
    uniq = np.unique(vac)
    e = capi.Engine(f, id_order=order, n_walkers=len(uniq), device=0)
    occ = np.repeat(golden[tag + "_ev_base_occ"][None, :], len(uniq), axis=0).copy()
    occ[np.arange(len(uniq)), uniq] = 0                         # one replica per vacancy position
    e.set_occupancy_all(occ)
    walker = np.searchsorted(uniq, vac).astype(np.int32)
    for name in ("e0", "e0low"):
        e.load_coefficients(alt_json[name], model=capi.BARRIER_E0)
        ea, de, d, ks = e.eval_barriers(I, J, walker=walker, want_parts=True)
        assert np.max(np.abs(de - alt["%s_%s_dE" % (tag, name)])) < 1e-9
        assert np.max(np.abs(ea - alt["%s_%s_Ea" % (tag, name)])) < 1e-9
        assert np.max(np.abs(ks - alt["%s_%s_e0" % (tag, name)])) < 1e-9 and np.all(d == 1.0)
        if name == "e0low":
            assert np.array_equal(ea == 0, alt["%s_%s_Ea" % (tag, name)] == 0)
        # the event-list form (box scan kernel): same events grouped by vacancy
        nb, ea2, de2 = e.eval_vacancy_events(uniq, walker=np.arange(len(uniq), dtype=np.int32))
        for k in range(len(I)):
            w = walker[k]
            assert I[k] == uniq[w]
            slot = int(np.nonzero(nb[w] == J[k])[0][0])
            assert abs(ea2[w, slot] - alt["%s_%s_Ea" % (tag, name)][k]) < 1e-9
            assert abs(de2[w, slot] - alt["%s_%s_dE" % (tag, name)][k]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("tag,order", CASES)
def test_gpu_pair_and_site_predictors_match_reference(golden, alt, alt_json, tag, order):
    f = int(golden[tag + "_factor"][0])
    e = capi.Engine(f, id_order=order, device=0)
    e.load_coefficients(alt_json["e0"], model=capi.BARRIER_E0)
    base = golden[tag + "_ev_base_occ"].copy()
    e.set_occupancy(base)
    assert np.max(np.abs(e.eval_pair_de(alt[tag + "_pair_a"], alt[tag + "_pair_b"]) - alt[tag + "_pair_dE"])) < 1e-9
    assert np.max(np.abs(e.eval_site_de(alt[tag + "_site_id"], alt[tag + "_site_new"]) - alt[tag + "_site_dE"])) < 1e-9
    v = int(alt[tag + "_pairvac_site"])
    occ = base.copy()
    occ[v] = 0
    e.set_occupancy(occ)
    nb = e.neighbors(1, v)
    assert np.max(np.abs(e.eval_pair_de(np.full(12, v), nb) - alt[tag + "_pairvac_dE"])) < 1e-9
    assert np.max(np.abs(e.eval_pair_de(nb, np.full(12, v)) - alt[tag + "_pairvac_dE"])) < 1e-9
    # not a first-neighbour pair: std::out_of_range in the reference unless the species are equal (then 0)
    second = e.neighbors(2, v)
    with pytest.raises(capi.LmcOutOfRange):
        e.eval_pair_de([v], [second[0]])
    same = np.nonzero(occ == occ[second[0]])[0]
    far = [s for s in same if s not in set(e.neighbors(1, int(second[0])))][:4]
    assert np.all(e.eval_pair_de(np.full(len(far), second[0]), far) == 0.0)


@pytest.mark.gpu
def test_gpu_kmc_runs_on_e0_model(golden, alt_json):
    """The KMC drivers take either barrier model: first- and second-order steps on the E0 model replay the oracle's."""
    tag, order = "A", capi.ORDER_REASSIGNED
    f = int(golden[tag + "_factor"][0])
    cfg = H.oracle_config(golden, tag)
    pred = O.VacancyMigrationPredictorE0(alt_json["e0"], cfg, H.CODES)
    rng = np.random.default_rng(5)
    u = rng.random((40, 2))
    for second_order in (False, True):
        cfg = H.oracle_config(golden, tag)
        e = capi.Engine(f, id_order=order, device=0)
        e.load_coefficients(alt_json["e0"], model=capi.BARRIER_E0)
        e.set_occupancy(cfg.occ)
        e.kmc_reset()
        tr = e.kmc_run(len(u), temperature=600.0, replay_u1=None if second_order else u[:, 0], replay_u2=u[:, 1], trace=True,
                       second_order=second_order)
        want = O.kmc_chain(cfg, pred, 600.0, u[:, 1]) if second_order else O.kmc_first(cfg, pred, 600.0, u[:, 0], u[:, 1])
        assert np.array_equal(tr["to"][0], want["to"])
        assert np.allclose(tr["dt"][0], want["dt"], rtol=1e-9, atol=0)
        assert np.array_equal(e.get_occupancy(0), cfg.occ)

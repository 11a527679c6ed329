"""CPU suite, part 1: the numpy oracle (oracle/lmc_oracle.py) against the committed golden vectors that
tests/golden/make_golden.py produced by running the reference itself (the reference has no
tests of its own, so reference outputs are the only possible pin).  Bit-exact for ids / mappings / counts /
one-hot encodes; <= 1e-9 eV (in practice ~1e-15) for energies and barriers."""
import numpy as np
import pytest

from oracle import lmc_oracle as O
from tests import helpers as H

TOL = 1e-9  # eV, the tolerance BASELINE.json's north_star states for dE / Ea


@pytest.mark.parametrize("tag", ["A", "B"])
def test_lattice_and_neighbours(golden, tag):
    cfg = H.oracle_config(golden, tag)
    assert np.allclose(cfg.rel, golden[tag + "_positions"], atol=1e-12)
    for s in (1, 2, 3):
        assert np.array_equal(cfg.nn[s - 1], golden["%s_nn%d" % (tag, s)])


def test_reassigned_occupancy_follows_sites(golden):
    f = int(golden["A_factor"][0])
    cfg = O.Config.generate_fcc(f, golden["A_occ_generate_order"])
    cfg.reassign_lattice_vector()
    assert np.array_equal(cfg.occ, golden["A_occ"])


@pytest.mark.parametrize("tag", ["A", "B"])
def test_ordered_lists(golden, tag):
    cfg = H.oracle_config(golden, tag)
    state, mmm, mm2 = O.sorted_lists_of_pairs(cfg, golden[tag + "_pair_i"], golden[tag + "_pair_j"])
    assert np.array_equal(state, golden[tag + "_list_state"])
    assert np.array_equal(mmm, golden[tag + "_list_mmm"])
    assert np.array_equal(mm2, golden[tag + "_list_mm2"])
    assert np.array_equal(O.sorted_list_of_sites(cfg, np.arange(cfg.num_sites)), golden[tag + "_list_site"])


@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("name,sizes", [
    ("state_pair", [2, 23, 12, 48, 44, 68, 136, 140]), ("state_site", [1, 12, 6, 24, 24, 36, 72, 72]),
    ("mmm", None), ("mm2", None)])
def test_cluster_mappings(golden, tag, name, sizes):
    cfg = H.oracle_config(golden, tag)
    mine = getattr(O, "mapping_" + name)(cfg)
    want = H.unflatten_mapping(golden["%s_mapping_%s" % (tag, name)])
    assert O.canonical_mapping(mine) == want
    if sizes:
        assert [len(g) for g in mine] == sizes
    else:
        assert len(mine) == {"mmm": 91, "mm2": 175}[name] and sum(len(g) for g in mine) == 614


def test_cluster_types(golden):
    want = [(int(r[0]), tuple(int(v) for v in r[2:2 + r[1]])) for r in golden["cluster_types"]]
    assert O.cluster_types(H.CODES) == want
    assert len(want) == 95


@pytest.mark.parametrize("tag", ["A", "B"])
def test_barrier_events(golden, tag):
    co = H.golden_coefficients(golden)
    cfg = H.oracle_config(golden, tag)
    pred = O.VacancyMigrationPredictorQuartic(co, cfg, H.CODES)
    vac, i, j = golden[tag + "_ev_vac"], golden[tag + "_ev_i"], golden[tag + "_ev_j"]
    enc_row = 0
    for v in np.unique(vac):
        occ = golden[tag + "_ev_base_occ"].copy()
        occ[v] = 0
        c = H.oracle_config(golden, tag, occ)
        sel = np.nonzero(vac == v)[0]
        sc, ec = pred.de_counts(c, i[sel], j[sel])
        assert np.array_equal(sc, golden[tag + "_ev_start_counts"][sel])       # bit-exact integer counts
        assert np.array_equal(ec, golden[tag + "_ev_end_counts"][sel])
        ea, de = pred.barrier_and_diff(c, i[sel], j[sel])
        d, ks = pred.get_d_ks(c, i[sel], j[sel])
        assert np.max(np.abs(de - golden[tag + "_ev_dE"][sel])) < TOL
        assert np.max(np.abs(ea - golden[tag + "_ev_Ea"][sel])) < TOL
        assert np.allclose(d, golden[tag + "_ev_D"][sel], rtol=1e-12, atol=0)
        assert np.allclose(ks, golden[tag + "_ev_Ks"][sel], rtol=1e-12, atol=0)
    # one-hot encodes are stored for the first two jumps of each vacancy position, in generation order
    order = []
    seen = {}
    for k, v in enumerate(vac):
        seen[v] = seen.get(v, 0) + 1
        if seen[v] <= 2:
            order.append(k)
    for row, k in enumerate(order):
        occ = golden[tag + "_ev_base_occ"].copy()
        occ[vac[k]] = 0
        c = H.oracle_config(golden, tag, occ)
        (xm, _, _), (xf, _, _), (xb, _, _) = pred.encodes(c, i[k:k + 1], j[k:k + 1])
        assert np.array_equal(xm[0], golden[tag + "_ev_enc_mmm"][row])          # bit-exact encodes
        assert np.array_equal(xf[0], golden[tag + "_ev_enc_mm2f"][row])
        assert np.array_equal(xb[0], golden[tag + "_ev_enc_mm2b"][row])


@pytest.mark.parametrize("tag", ["A", "B"])
def test_swap_and_site_energy_changes(golden, tag):
    co = H.golden_coefficients(golden)
    cfg = H.oracle_config(golden, tag)
    pred = O.EnergyChangePredictorPairSite(co, cfg, H.CODES)
    de = pred.de_pair(cfg, golden[tag + "_swap_a"], golden[tag + "_swap_b"])
    assert np.max(np.abs(de - golden[tag + "_swap_dE"])) < TOL
    cfg2 = H.oracle_config(golden, tag, golden[tag + "_cmc_occ"])
    sites, new = golden[tag + "_site"], golden[tag + "_site_new"]
    assert np.max(np.abs(pred.de_site(cfg2, sites, new) - golden[tag + "_site_dE"])) < TOL
    sc, ec = pred.site_counts(cfg2, sites[:40], new[:40])
    changed = cfg2.occ[sites[:40]] != new[:40]        # unchanged sites return early in the reference (stale buffers)
    assert np.array_equal(sc[changed], golden[tag + "_site_start_counts"][changed])
    assert np.array_equal(ec[changed], golden[tag + "_site_end_counts"][changed])


@pytest.mark.parametrize("tag", ["A", "B"])
def test_total_energy(golden, tag):
    cfg = H.oracle_config(golden, tag)
    pred = O.EnergyPredictor(H.golden_coefficients(golden), H.CODES)
    assert np.array_equal(pred.encode(cfg), golden[tag + "_energy_encode"])       # counts/normaliser: exact
    assert abs(pred.energy(cfg) - float(golden[tag + "_energy"][0])) < 1e-9


@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("run", ["kmc", "kmc_tt"])
def test_kmc_replay(golden, tag, run):
    """Feeding the reference's own (u1,u2) stream reproduces its event sequence (north_star replay mode)."""
    co = H.golden_coefficients(golden)
    cfg = H.oracle_config(golden, tag)
    pred = O.VacancyMigrationPredictorQuartic(co, cfg, H.CODES)
    g = lambda k: golden["%s_%s_%s" % (tag, run, k)]
    tt = O.TimeTemperatureInterpolator(points=[tuple(p) for p in golden["tt_points"]]) if run == "kmc_tt" else None
    tr = O.kmc_first(cfg, pred, 500.0, g("u1"), g("u2"), tt=tt, rate_corrector=(run == "kmc_tt"))
    for k in ("from", "to", "slot"):
        assert np.array_equal(tr[k], g(k)), k
    for k in ("time", "total_rate"):
        assert np.allclose(tr[k], g(k), rtol=1e-10, atol=0), k
    # the traced dt is time_after - time_before (cancellation once time is large): absolute tolerance in ulps of time
    assert np.allclose(tr["dt"], g("dt"), rtol=1e-10, atol=4e-16 * float(np.abs(g("time")).max()))
    for k in ("energy", "Ea", "dE"):
        assert np.max(np.abs(tr[k] - g(k))) < TOL, k
    assert np.array_equal(tr["temperature"], g("temperature"))
    assert np.array_equal(cfg.occ, g("final_occ"))


@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("run", ["chain", "chain_tt"])
def test_kmc_chain_replay(golden, golden_chain, tag, run):
    """Second-order KMC (mc::KineticMcChainOmpi): the reference's selecting uniforms reproduce its trajectory."""
    co = H.golden_coefficients(golden)
    cfg = H.oracle_config(golden, tag)
    pred = O.VacancyMigrationPredictorQuartic(co, cfg, H.CODES)
    n = 30                                                 # 144 numpy barrier evaluations per step: keep the CPU suite short
    g = lambda k: golden_chain["%s_%s_%s" % (tag, run, k)][:n]
    tt = O.TimeTemperatureInterpolator(points=[tuple(p) for p in golden["tt_points"]]) if run == "chain_tt" else None
    tr = O.kmc_chain(cfg, pred, 500.0, g("u2"), tt=tt, rate_corrector=(run == "chain_tt"))
    for k in ("from", "to", "slot"):
        assert np.array_equal(tr[k], g(k)), k
    for k in ("time", "total_rate"):
        assert np.allclose(tr[k], g(k), rtol=1e-9, atol=0), k
    assert np.allclose(tr["dt"], g("dt"), rtol=1e-9, atol=4e-16 * float(np.abs(g("time")).max()))
    for k in ("energy", "Ea", "dE"):
        assert np.max(np.abs(tr[k] - g(k))) < TOL, k
    assert np.array_equal(tr["temperature"], g("temperature"))
    # (the full 121-step trajectories incl. the final occupancy are replayed by the CUDA path in tests/test_gpu_kmc.py)


@pytest.mark.parametrize("tag", ["A", "B"])
def test_cmc_replay(golden, tag):
    co = H.golden_coefficients(golden)
    cfg = H.oracle_config(golden, tag, golden[tag + "_cmc_occ"])
    pred = O.EnergyChangePredictorPairSite(co, cfg, H.CODES)
    g = lambda k: golden["%s_cmc_%s" % (tag, k)]
    tr = O.metropolis_trials(cfg, pred, g("a"), g("b"), g("u"), temperature=800.0)
    assert np.max(np.abs(tr["dE"] - g("dE"))) < TOL
    assert np.max(np.abs(tr["energy_before"] - g("energy_before"))) < 1e-9
    assert np.array_equal(cfg.occ, g("final_occ"))
    assert 0.05 < tr["accepted"].mean() < 0.95


@pytest.mark.parametrize("tag", ["P", "Q"])
def test_cmc_omp_batches(golden, tag):
    """mc::CanonicalMcOmp (mc/src/CanonicalMcOmp.cpp:40-92), golden_omp_v1.npz: every batch holds at most `threads` trials,
    no trial site lies within the 43-site neighbourhoods of an earlier trial of its batch (the unavailable_position_ rule),
    the dE the reference evaluated on the batch-START configuration equals the dE the serial chain sees, and the serial
    replay of the stream reproduces its energies and final occupancy."""
    import os
    omp = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_omp_v1.npz"), allow_pickle=False)
    g = lambda k: omp["%s_%s" % (tag, k)]
    f, reassign, threads, temperature, _ = g("params")
    co = H.golden_coefficients(golden)
    cfg = O.Config.generate_fcc(int(f), None)
    if int(reassign):
        cfg.reassign_lattice_vector()
    cfg.occ[:] = g("occ")
    pred = O.EnergyChangePredictorPairSite(co, cfg, H.CODES)
    a, b, batch = g("a"), g("b"), g("batch")
    for q in range(int(batch[-1]) + 1):
        idx = np.nonzero(batch == q)[0]
        assert 1 <= len(idx) <= int(threads)
        taken = set()
        for i in idx:
            assert int(a[i]) not in taken and int(b[i]) not in taken
            for s in (int(a[i]), int(b[i])):
                taken.update(int(x) for x in cfg.neighbors_set_of_site(s))
                taken.add(s)
    tr = O.metropolis_trials(cfg, pred, a, b, g("u"), temperature=float(temperature))
    assert np.max(np.abs(tr["dE"] - g("dE"))) < TOL
    assert np.max(np.abs(tr["energy_before"] - g("energy_before"))) < 1e-9
    assert np.array_equal(cfg.occ, g("final_occ"))


def test_simulated_annealing_replay(golden):
    co = H.golden_coefficients(golden)
    f, t0, steps = golden["SA_params"]
    cfg = O.Config.generate_fcc(int(f), golden["SA_occ"])
    pred = O.EnergyChangePredictorPairSite(co, cfg, H.CODES)
    sched = O.SaSchedule(float(t0), int(steps))
    tr = O.metropolis_trials(cfg, pred, golden["SA_a"], golden["SA_b"], golden["SA_u"], sa_schedule=sched)
    assert np.allclose(tr["temperature_before"], golden["SA_temperature_before"], rtol=1e-13, atol=0)
    assert np.max(np.abs(tr["energy_before"] - golden["SA_energy_before"])) < 1e-9
    assert np.array_equal(cfg.occ, golden["SA_final_occ"])
    assert abs(sched.temperature - golden["SA_final"][1]) < 1e-9


def test_helpers_rate_corrector_and_interpolator():
    tt = O.TimeTemperatureInterpolator(points=[(0.0, 300.0), (1e-3, 500.0), (1e-1, 700.0)])
    assert tt.temperature(-1.0) == 300.0 and tt.temperature(0.0) == 300.0 and tt.temperature(1.0) == 700.0
    assert abs(tt.temperature(5e-4) - 400.0) < 1e-12
    # pred/include/RateCorrector.hpp:17-24 evaluated by hand at 500 K
    import math
    want = 2e-4 / (1.64 * math.exp(-(0.66 / 8.617333262145e-5 / 500.0 - 0.7))) / (1 - 13 * 0.04)
    assert abs(O.rate_correction_factor(2e-4, 0.04, 500.0) - want) < 1e-9 * want

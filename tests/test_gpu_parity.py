"""GPU suite (-m gpu): the hand-written CUDA path, called through the C ABI (include/lmc_b200.h), against
 (1) the committed golden vectors produced by the reference itself,
 (2) the numpy oracle on fresh seeded inputs (sizes the oracle finishes in seconds),
 (3) size-independent identities at BASELINE.json's full sizes (f=40: 256k sites).
Bars (BASELINE.json north_star): neighbour ids, counts and encodes bit-exact; dE and Ea within 1e-9 eV."""
import numpy as np
import pytest

from latticemontecarlo_b200 import capi, synth
from oracle import lmc_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-9   # eV


def _engine(golden, tag, tmp_path, n_walkers=1):
    order = capi.ORDER_REASSIGNED if int(golden[tag + "_factor"][1]) else capi.ORDER_GENERATE
    e = capi.Engine(int(golden[tag + "_factor"][0]), id_order=order, n_walkers=n_walkers, device=0)
    e.load_coefficients(H.golden_json(golden, tmp_path))
    return e


@pytest.mark.parametrize("order", [capi.ORDER_GENERATE, capi.ORDER_REASSIGNED])
def test_occupancy_round_trip_and_jump(order):
    e = capi.Engine((4, 5, 6), id_order=order, n_walkers=3, device=0)
    rng = np.random.default_rng(1)
    occ = rng.integers(0, 4, (3, e.num_sites)).astype(np.uint8)
    e.set_occupancy_all(occ)
    assert np.array_equal(e.get_occupancy_all(), occ)
    e.set_occupancy(occ[2], walker=0)
    assert np.array_equal(e.get_occupancy(0), occ[2])
    a, b = 0, e.num_sites - 1          # corner sites: every halo image must follow
    e.lattice_jump(a, b, walker=1)
    want = occ[1].copy()
    want[[a, b]] = want[[b, a]]
    assert np.array_equal(e.get_occupancy(1), want)
    with pytest.raises(capi.LmcInvalidArgument):
        e.set_occupancy(np.full(e.num_sites, 5, np.uint8))      # Sn is not in the element set


@pytest.mark.parametrize("tag", ["A", "B"])
def test_barriers_against_golden(golden, tag, tmp_path):
    e = _engine(golden, tag, tmp_path)
    vac, i, j = golden[tag + "_ev_vac"], golden[tag + "_ev_i"], golden[tag + "_ev_j"]
    for v in np.unique(vac):
        occ = golden[tag + "_ev_base_occ"].copy()
        occ[v] = 0
        e.set_occupancy(occ)
        sel = np.nonzero(vac == v)[0]
        ea, de, d, ks = e.eval_barriers(i[sel], j[sel], want_parts=True)
        assert np.max(np.abs(de - golden[tag + "_ev_dE"][sel])) < TOL
        assert np.max(np.abs(ea - golden[tag + "_ev_Ea"][sel])) < TOL
        assert np.allclose(d, golden[tag + "_ev_D"][sel], rtol=1e-11, atol=0)
        assert np.allclose(ks, golden[tag + "_ev_Ks"][sel], rtol=1e-11, atol=0)
        for k in sel[:2]:                                         # integer artefacts: bit exact
            dbg = e.debug_pair(i[k], j[k])
            assert np.array_equal(dbg["start_counts"], golden[tag + "_ev_start_counts"][k])
            assert np.array_equal(dbg["end_counts"], golden[tag + "_ev_end_counts"][k])
    # one-hot encodes (rows: first two jumps of each vacancy position, generation order)
    seen, rows = {}, []
    for k, v in enumerate(vac):
        seen[v] = seen.get(v, 0) + 1
        if seen[v] <= 2:
            rows.append(k)
    s_mmm, s_mm2 = capi.tables_group_sizes("mmm", 3), capi.tables_group_sizes("mm2", 3)
    for row, k in enumerate(rows[:12]):
        occ = golden[tag + "_ev_base_occ"].copy()
        occ[vac[k]] = 0
        e.set_occupancy(occ)
        dbg = e.debug_pair(i[k], j[k])
        assert np.array_equal(dbg["enc_mmm"] / s_mmm, golden[tag + "_ev_enc_mmm"][row])
        assert np.array_equal(dbg["enc_mm2_f"] / s_mm2, golden[tag + "_ev_enc_mm2f"][row])
        assert np.array_equal(dbg["enc_mm2_b"] / s_mm2, golden[tag + "_ev_enc_mm2b"][row])


@pytest.mark.parametrize("tag", ["A", "B"])
def test_device_ordered_lists_bit_exact(golden, tag, tmp_path):
    e = _engine(golden, tag, tmp_path)
    e.set_occupancy(golden[tag + "_ev_base_occ"])
    cfg = H.oracle_config(golden, tag)
    pi, pj = golden[tag + "_pair_i"][::9], golden[tag + "_pair_j"][::9]
    _, _, backward = O.sorted_lists_of_pairs(cfg, pj, pi)
    e2 = capi.Engine(e.factors, id_order=capi.ORDER_REASSIGNED if tag == "A" else capi.ORDER_GENERATE, device=0)
    e2.set_occupancy(golden[tag + "_ev_base_occ"])
    for k in range(len(pi)):
        k9 = 9 * k
        import ctypes as C
        s = np.empty(60, np.int64); m = np.empty(58, np.int64); m2 = np.empty(58, np.int64); mb = np.empty(58, np.int64)
        capi._check(capi.lib().lmc_debug_pair(e2.h, 0, C.c_int64(int(pi[k])), C.c_int64(int(pj[k])), capi._p(s), capi._p(m),
                                              capi._p(m2), capi._p(mb), None, None, None, None, None))
        assert np.array_equal(s, golden[tag + "_list_state"][k9])
        assert np.array_equal(m, golden[tag + "_list_mmm"][k9])
        assert np.array_equal(m2, golden[tag + "_list_mm2"][k9])
        assert np.array_equal(mb, backward[k])
    for site in range(0, e.num_sites, 17):
        assert np.array_equal(e.debug_site(site, 1)["state"], golden[tag + "_list_site"][site])


@pytest.mark.parametrize("tag", ["A", "B"])
def test_swap_site_and_total_energy_against_golden(golden, tag, tmp_path):
    e = _engine(golden, tag, tmp_path)
    e.set_occupancy(golden[tag + "_occ"])
    de = e.eval_swap_de(golden[tag + "_swap_a"], golden[tag + "_swap_b"])
    assert np.max(np.abs(de - golden[tag + "_swap_dE"])) < TOL
    energy, counts = e.total_energy(want_counts=True)
    assert abs(energy - float(golden[tag + "_energy"][0])) < 1e-9
    norm = np.array([O.E_CLUSTER_COUNTER[t[0]] for t in O.cluster_types(H.CODES)], dtype=np.float64)
    assert np.array_equal(counts / norm, golden[tag + "_energy_encode"])          # exact integer counts
    e.set_occupancy(golden[tag + "_cmc_occ"])
    sites, new = golden[tag + "_site"], golden[tag + "_site_new"]
    assert np.max(np.abs(e.eval_site_de(sites, new) - golden[tag + "_site_dE"])) < TOL
    occ = golden[tag + "_cmc_occ"]
    for k in range(40):
        if occ[sites[k]] == new[k]:
            continue
        dbg = e.debug_site(sites[k], int(new[k]))
        assert np.array_equal(dbg["start_counts"], golden[tag + "_site_start_counts"][k])
        assert np.array_equal(dbg["end_counts"], golden[tag + "_site_end_counts"][k])


def test_error_paths_on_device(golden, tmp_path):
    e = _engine(golden, "A", tmp_path)
    occ = golden["A_ev_base_occ"].copy()
    v = 100
    nn = e.neighbors(1, v)
    occ[v] = 0
    e.set_occupancy(occ)
    with pytest.raises(capi.LmcOutOfRange):            # not first neighbours (GetPairFlatIndex throws in the reference)
        e.eval_barriers([v], [e.neighbors(2, v)[0]])
    with pytest.raises(capi.LmcOutOfRange):            # first site does not hold the vacancy
        e.eval_barriers([nn[0]], [v])
    occ2 = occ.copy()
    occ2[nn[3]] = 0                                    # second vacancy in range: "Cluster not found in ClusterIndexer"
    e.set_occupancy(occ2)
    with pytest.raises(capi.LmcOutOfRange):
        e.eval_barriers([v], [nn[0]])
    e.set_occupancy(occ)
    ea, de = e.eval_barriers([v] * 12, nn)             # and the engine still works afterwards
    assert np.all(np.isfinite(ea)) and np.all(np.isfinite(de))
    with pytest.raises(capi.LmcInvalidArgument):
        e.eval_swap_de([0], [e.num_sites])


def test_random_alloys_against_oracle(coef_json):
    """Fresh seeded inputs, full-size synthetic coefficients (K=24/32), both id orders, dilute and concentrated."""
    rng = np.random.default_rng(77)
    for f, order, p in ((6, capi.ORDER_REASSIGNED, 0.02), (5, capi.ORDER_GENERATE, 0.30), (4, capi.ORDER_REASSIGNED, 0.10)):
        occ_gen = synth.random_alloy(f, p, p, seed=int(rng.integers(1 << 30)))
        cfg = O.Config.generate_fcc(f, occ_gen)
        if order == capi.ORDER_REASSIGNED:
            cfg.reassign_lattice_vector()
        e = capi.Engine(f, id_order=order, device=0)
        e.load_coefficients(coef_json)
        e.set_occupancy(cfg.occ)
        quartic = O.VacancyMigrationPredictorQuartic(coef_json, cfg, H.CODES)
        vac = int(np.nonzero(cfg.occ == 0)[0][0])
        for _ in range(6):
            nbrs = cfg.nn[0][vac]
            ea_o, de_o = quartic.barrier_and_diff(cfg, np.full(12, vac), nbrs)
            ea, de = e.eval_barriers(np.full(12, vac), nbrs)
            assert np.max(np.abs(ea - ea_o)) < TOL and np.max(np.abs(de - de_o)) < TOL
            # reverse-jump identities on the post-jump configuration
            to = int(nbrs[int(rng.integers(12))])
            slot = list(nbrs).index(to)
            cfg.lattice_jump(vac, to)
            e.lattice_jump(vac, to)
            ea_r, de_r = e.eval_barriers([to], [vac])
            assert abs(de_r[0] + de[slot]) < 1e-12 and abs(ea_r[0] - (ea[slot] - de[slot])) < 1e-9
            vac = to
        assert np.array_equal(e.get_occupancy(), cfg.occ)
        pairsite = O.EnergyChangePredictorPairSite(coef_json, cfg, H.CODES)
        n = cfg.num_sites
        a = np.concatenate([rng.integers(0, n, 200), np.arange(40)])
        b = np.concatenate([rng.integers(0, n, 200), cfg.nn[int(rng.integers(3))][np.arange(40), 1]])
        keep = (cfg.occ[a] != 0) | (cfg.occ[b] != 0)
        # pairs that would put two vacancies in range are rejected by both implementations; skip the vacancy's shell
        near_vac = np.isin(a, cfg.neighbors_set_of_site(vac)) | np.isin(b, cfg.neighbors_set_of_site(vac))
        keep &= ~near_vac | (a == vac) | (b == vac)
        a, b = a[keep], b[keep]
        assert np.max(np.abs(e.eval_swap_de(a, b) - pairsite.de_pair(cfg, a, b))) < TOL
        energy = O.EnergyPredictor(coef_json, H.CODES)
        assert abs(e.total_energy() - energy.energy(cfg)) < 1e-8


def test_quaternary_element_set_against_oracle(tmp_path):
    elements = ("Al", "Cu", "Mg", "Zn")
    codes = [1, 4, 2, 3]
    js = tmp_path / "q.json"
    synth.write_synthetic_json(js, seed=5, elements=elements, k_mmm=3, k_mm2=4)
    f = 4
    rng = np.random.default_rng(3)
    occ = rng.choice(np.array([1, 1, 1, 1, 2, 3, 4], dtype=np.uint8), 4 * f ** 3)
    occ[17] = 0
    cfg = O.Config.generate_fcc(f, occ)
    cfg.reassign_lattice_vector()
    e = capi.Engine(f, element_set=codes, solvent=1, device=0)
    e.load_coefficients(js)
    e.set_occupancy(cfg.occ)
    vac = int(np.nonzero(cfg.occ == 0)[0][0])
    quartic = O.VacancyMigrationPredictorQuartic(str(js), cfg, codes)
    ea_o, de_o = quartic.barrier_and_diff(cfg, np.full(12, vac), cfg.nn[0][vac])
    ea, de = e.eval_barriers(np.full(12, vac), cfg.nn[0][vac])
    assert np.max(np.abs(ea - ea_o)) < TOL and np.max(np.abs(de - de_o)) < TOL
    dbg = e.debug_pair(vac, cfg.nn[0][vac][5])
    sc, ec = quartic.de_counts(cfg, [vac], [cfg.nn[0][vac][5]])
    assert np.array_equal(dbg["start_counts"], sc[0]) and np.array_equal(dbg["end_counts"], ec[0])
    (xm, cm, _), (xf, cf, _), (xb, cb, _) = quartic.encodes(cfg, [vac], [cfg.nn[0][vac][5]])
    assert np.array_equal(dbg["enc_mmm"], cm[0]) and np.array_equal(dbg["enc_mm2_f"], cf[0]) and np.array_equal(dbg["enc_mm2_b"], cb[0])
    pairsite = O.EnergyChangePredictorPairSite(str(js), cfg, codes)
    a = rng.integers(0, cfg.num_sites, 100); b = rng.integers(0, cfg.num_sites, 100)
    far = ~(np.isin(a, cfg.neighbors_set_of_site(vac)) | np.isin(b, cfg.neighbors_set_of_site(vac)))
    assert np.max(np.abs(e.eval_swap_de(a[far], b[far]) - pairsite.de_pair(cfg, a[far], b[far]))) < TOL


def test_full_size_identities(coef_json):
    """BASELINE configs[1] size (40x40x40, 256k sites): properties that need no oracle run.
    dE_swap == E(after) - E(before), reverse-jump antisymmetry, translation invariance."""
    f = 40
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=2, device=0)
    e.load_coefficients(coef_json)
    occ = synth.random_alloy(f, 0.02, 0.02, seed=42)
    e.set_occupancy(occ, walker=0)
    n = e.num_sites
    rng = np.random.default_rng(9)
    e0 = e.total_energy(0)
    a = rng.integers(0, n, 64); b = rng.integers(0, n, 64)
    de = e.eval_swap_de(a, b)
    for k in range(8):
        if occ[a[k]] == occ[b[k]]:
            assert de[k] == 0.0
            continue
        e.lattice_jump(a[k], b[k])
        e1 = e.total_energy(0)
        e.lattice_jump(a[k], b[k])
        assert abs((e1 - e0) - de[k]) < 5e-9, (e1 - e0, de[k])      # total energy ~1e3 eV: ulp-limited
    # translation invariance: shift the whole occupancy by one conventional cell along x (ids: +2*2*f*f sites)
    vac = int(np.nonzero(occ == 0)[0][0])
    nn = e.neighbors(1, vac)
    ea, de_v = e.eval_barriers(np.full(12, vac), nn)
    shift = 2 * (2 * f * f)
    occ2 = np.roll(occ, shift)
    e.set_occupancy(occ2, walker=1)
    vac2 = (vac + shift) % n
    ea2, de2 = e.eval_barriers(np.full(12, vac2), (nn + shift) % n, walker=np.ones(12, np.int32))
    order1, order2 = np.argsort(nn), np.argsort((nn + shift) % n)
    assert np.allclose(np.sort(ea), np.sort(ea2), rtol=0, atol=1e-12) and np.allclose(np.sort(de_v), np.sort(de2), rtol=0, atol=1e-12)


def test_vacancy_event_lists_match_per_event_evaluation(golden, coef_json, tmp_path):
    """lmc_eval_vacancy_events (one box scan per vacancy) against the golden Ea / dE of the reference for the same events,
    in the reference's event order, and against lmc_eval_barriers on a many-walker engine."""
    for tag in ("A", "B"):
        order = capi.ORDER_REASSIGNED if int(golden[tag + "_factor"][1]) else capi.ORDER_GENERATE
        e = capi.Engine(int(golden[tag + "_factor"][0]), id_order=order, n_walkers=1, device=0)
        e.load_coefficients(H.golden_json(golden, tmp_path))
        base = golden[tag + "_ev_base_occ"]
        vacs = golden[tag + "_ev_vac"][::12]
        for q, v in enumerate(vacs):
            occ = base.copy(); occ[v] = 0
            e.set_occupancy(occ)
            nb, ea, de = e.eval_vacancy_events([int(v)])
            sl = slice(12 * q, 12 * q + 12)
            assert np.array_equal(nb[0], golden[tag + "_ev_j"][sl])                       # ascending neighbour ids = reference order
            assert np.max(np.abs(ea[0] - golden[tag + "_ev_Ea"][sl])) < 1e-9 and np.max(np.abs(de[0] - golden[tag + "_ev_dE"][sl])) < 1e-9
    f, W = 6, 40
    e = capi.Engine(f, n_walkers=W, device=0)
    e.load_coefficients(coef_json)
    rng = np.random.default_rng(5)
    vac = rng.integers(0, 4 * f ** 3, W)
    for w in range(W):
        e.set_occupancy(synth.random_alloy(f, 0.1, 0.1, seed=50 + w, vacancy_site=int(vac[w])), walker=w)
    nb, ea, de = e.eval_vacancy_events(vac, walker=np.arange(W))
    ea2, de2 = e.eval_barriers(np.repeat(vac, 12), nb.reshape(-1), walker=np.repeat(np.arange(W), 12))
    assert np.array_equal(nb, np.stack([e.neighbors(1, int(v)) for v in vac]))
    assert np.max(np.abs(ea.reshape(-1) - ea2)) < 1e-12 and np.max(np.abs(de.reshape(-1) - de2)) < 1e-12
    with pytest.raises(capi.LmcOutOfRange):                                                   # not a vacancy there
        e.eval_vacancy_events([int((vac[0] + 1) % (4 * f ** 3))], walker=[0])


def test_row_gather_swap_kernel_equals_general_kernel(coef_json, tmp_path):
    """swap_de_rows_kernel (one aligned load per (dx, dy) row, codes in registers) against the byte-gather kernel it replaced
    (LMC_SWAP_GENERAL_KERNEL=1 in a second process): 300k random pairs over 3 replicas of a concentrated 9x9x9 alloy with a
    vacancy each (periodic wrap in every direction, coupled pairs, pairs that involve the vacancy), bit for bit, including
    the status code and the NaN pattern of pairs the reference has no cluster type for."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "swap_ab.py"
    script.write_text("""
import ctypes as C, sys, numpy as np
sys.path.insert(0, %r)
from latticemontecarlo_b200 import capi, synth
f = 9
e = capi.Engine(f, id_order=capi.ORDER_GENERATE, n_walkers=3, device=0)
e.load_coefficients(%r)
occ = np.stack([synth.random_alloy(f, 0.15, 0.2, seed=s) for s in (1, 2, 3)])
e.set_occupancy_all(occ)
rng = np.random.default_rng(4)
n = 300000
w = rng.integers(0, 3, n).astype(np.int32)
a = rng.integers(0, e.num_sites, n); b = rng.integers(0, e.num_sites, n)
vac = np.array([int(np.nonzero(o == 0)[0][0]) for o in occ])
a[:3000] = vac[w[:3000]]
out = np.empty(n)
rc = capi.lib().lmc_eval_swap_de(e.h, C.c_int64(n), capi._p(w), capi._p(a), capi._p(b), capi._p(out))
np.save(sys.argv[1], np.concatenate([[rc], out]))
""" % (root, coef_json))
    outs = []
    for k, env_extra in enumerate(({}, {"LMC_SWAP_GENERAL_KERNEL": "1"})):
        out = tmp_path / ("swap_%d.npy" % k)
        res = subprocess.run([sys.executable, str(script), str(out)], capture_output=True, text=True, env=dict(os.environ, **env_extra))
        assert res.returncode == 0, res.stderr
        outs.append(np.load(out))
    rows, general = outs
    assert rows[0] == general[0]
    assert np.array_equal(np.isnan(rows), np.isnan(general))
    ok = ~np.isnan(rows)
    assert np.array_equal(rows[ok], general[ok]) and np.count_nonzero(rows[ok]) > 150000


@pytest.mark.parametrize("tag", ["A", "B"])
def test_energy_predictor_encode_cluster_energy_and_chemical_potential(golden, tag, tmp_path):
    """EnergyPredictor::GetEncode, GetEnergyOfCluster / GetEncodeOfCluster and GetChemicalPotential(Al)
    (pred/src/EnergyPredictor.cpp:40-214) against the reference's own outputs (tests/golden/golden_energy_v1.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_energy_v1.npz"), allow_pickle=False)
    f, reassign = (int(v) for v in g[tag + "_params"])
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED if reassign else capi.ORDER_GENERATE, device=0)
    e.load_coefficients(H.golden_json(golden, tmp_path))
    occ = g[tag + "_occ"]
    e.set_occupancy(occ)
    assert abs(e.total_energy() - g[tag + "_energy"][0]) < TOL
    assert np.max(np.abs(e.energy_encode() - g[tag + "_encode"])) < 1e-15          # integer counts / fixed normalisers
    lattice_of_atom = g[tag + "_lattice_of_atom"]
    for k in range(int(g[tag + "_n_lists"][0])):
        ids = lattice_of_atom[g["%s_list%d" % (tag, k)]]
        assert abs(e.energy_of_cluster(ids) - g["%s_cluster_energy%d" % (tag, k)][0]) < TOL, k
        assert np.max(np.abs(e.energy_encode(ids) - g["%s_cluster_encode%d" % (tag, k)])) < 1e-15, k
    assert e.energy_of_cluster([]) == 0.0
    # the full list is the whole configuration
    assert abs(e.energy_of_cluster(np.arange(occ.size)) - e.total_energy()) < 1e-9
    mu = e.chemical_potential(1)
    assert list(mu) == [int(v) for v in g["mu_elements"]]                          # std::map order: by element name
    assert np.max(np.abs(np.array(list(mu.values())) - g["mu_values"])) < TOL
    # SimulatedAnnealing's reference energy (mc/src/SimulatedAnnealing.cpp:58-70): [E(config) - E(pure solvent)] - sum mu_e n_e
    # vanishes for isolated solutes -- the reference's constructor value for one Mg + one Zn placed >= 4NN apart is stored
    pure = np.ones_like(occ)
    iso = pure.copy()
    iso[0], iso[occ.size // 2 + 1] = 2, 3
    e.set_occupancy(pure); e_pure = e.total_energy()
    e.set_occupancy(iso); e_iso = e.total_energy()
    initial_energy = (e_iso - e_pure) - mu[2] - mu[3]
    assert abs(initial_energy) < 1e-9 and abs(initial_energy - g["SA_initial_energy"][0]) < 1e-9
    e.set_occupancy(occ)
    # Config read accessors
    sites = np.array([0, 5, occ.size - 1, occ.size // 2])
    assert np.array_equal(e.get_elements(sites), occ[sites])
    first, count = e.find_element(0)
    assert count == np.count_nonzero(occ == 0) and first == int(np.nonzero(occ == 0)[0][0])
    assert e.find_element(2)[1] == np.count_nonzero(occ == 2)
    with pytest.raises(capi.LmcInvalidArgument):
        e.get_elements([occ.size])


def test_grouped_requests_give_the_callers_order_bit_for_bit(coef_json, monkeypatch):
    """Device-side grouping of scattered requests (grouping.cu: requests ordered by the lattice region of their first
    site, kernels run over the permutation) must not change a single bit of any result nor their order; the automatic
    choice (barrier events only; swaps are grouped only on request) leaves an already local request order alone."""
    f, W = 8, 24
    e = capi.Engine(f, n_walkers=W, device=0)
    e.load_coefficients(coef_json)
    vac = np.empty(W, dtype=np.int64)
    occs = []
    for w in range(W):
        occ = synth.random_alloy(f, 0.05, 0.05, seed=900 + w)
        e.set_occupancy(occ, walker=w)
        occs.append(occ)
        vac[w] = int(np.nonzero(occ == 0)[0][0])
    nb = np.stack([e.neighbors(1, int(v)) for v in vac])
    rng = np.random.default_rng(5)
    n_ev = W * 12 * 40
    perm = rng.permutation(n_ev)
    w_ev = np.tile(np.repeat(np.arange(W, dtype=np.int32), 12), 40)[perm]
    i_ev = np.tile(np.repeat(vac, 12), 40)[perm]
    j_ev = np.tile(nb.reshape(-1), 40)[perm]
    n_sw = 20000
    w_sw = rng.integers(0, W, n_sw).astype(np.int32)
    a_sw = rng.integers(0, e.num_sites, n_sw)
    b_sw = rng.integers(0, e.num_sites, n_sw)
    vac_hit = (a_sw == vac[w_sw]) | (b_sw == vac[w_sw])             # keep the vacancy out of the swaps
    a_sw[vac_hit] = (vac[w_sw[vac_hit]] + 7) % e.num_sites
    b_sw[vac_hit] = (vac[w_sw[vac_hit]] + 11) % e.num_sites

    def run(mode):
        if mode is None:
            monkeypatch.delenv("LMC_GROUP_REQUESTS", raising=False)
            monkeypatch.setenv("LMC_GROUP_MIN", "1000")
        else:
            monkeypatch.setenv("LMC_GROUP_REQUESTS", mode)
        ea, de, d, ks = e.eval_barriers(i_ev, j_ev, walker=w_ev, want_parts=True)
        return ea, de, d, ks, e.eval_swap_de(a_sw, b_sw, walker=w_sw)

    plain, grouped, auto = run("0"), run("1"), run(None)
    for k in range(5):
        assert np.array_equal(plain[k], grouped[k]), k
        assert np.array_equal(plain[k], auto[k]), k
    assert np.all(np.isfinite(plain[0])) and np.all(np.isfinite(plain[4]))
    # requests that are already grouped by vacancy are left in place by the automatic choice (and give the same numbers)
    monkeypatch.delenv("LMC_GROUP_REQUESTS", raising=False)
    monkeypatch.setenv("LMC_GROUP_MIN", "100")
    ea_g, de_g = e.eval_barriers(np.repeat(vac, 12), nb.reshape(-1), walker=np.repeat(np.arange(W, dtype=np.int32), 12))
    back = np.argsort(perm)[: W * 12]                                 # where the first copy of every event went
    assert np.array_equal(ea_g, plain[0][back]) and np.array_equal(de_g, plain[1][back])

"""CPU suite, part 2: the product's host logic (integer-geometry tables, coefficient contraction, C-ABI surface)
against the golden vectors and the oracle.  No compute kernel is launched here; the contracted tables are read
back through lmc_engine_get_tables and evaluated by the test itself."""
import ctypes

import numpy as np
import pytest

from latticemontecarlo_b200 import capi
from oracle import lmc_oracle as O
from tests import helpers as H

CODE_OF_ENUM = {1: 0, 2: 1, 3: 2, 0: 3}   # compact species codes for element_set (Al, Mg, Zn): sorted by name, vacancy last


def test_library_exports_every_declared_symbol():
    lib = capi.lib()
    names = capi.declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), name
    assert lib.lmc_abi_version() == 1


def test_host_only_engine_refuses_compute(golden, tmp_path):
    e = capi.Engine(4, device=-1)
    e.load_coefficients(H.golden_json(golden, tmp_path))
    with pytest.raises(capi.LmcError) as err:
        e.eval_barriers([0], [1])
    assert err.value.code == capi.LMC_ERR_NO_DEVICE and "no CPU fallback" in str(err.value)
    with pytest.raises(capi.LmcError):
        e.set_occupancy(np.ones(e.num_sites, np.uint8))


def test_error_behaviour_matches_reference(tmp_path):
    with pytest.raises(capi.LmcInvalidArgument):
        capi.Engine(3, device=-1)                      # cell list of the reference needs >= 4 cells per axis
    e = capi.Engine(4, device=-1)
    with pytest.raises(capi.LmcError) as err:
        e.load_coefficients(tmp_path / "missing.json")
    assert "Cannot open" in str(err.value)             # same message as the reference predictors
    with pytest.raises(capi.LmcOutOfRange):
        e.pair_lists(0, 5)                             # not first neighbours: std::out_of_range in the reference
    bad = tmp_path / "bad.json"
    bad.write_text('{"Base": {"theta": [1.0, 2.0]}}')
    with pytest.raises(capi.LmcInvalidArgument):
        e.load_coefficients(bad)                       # wrong Base.theta length (the reference would read out of bounds)


@pytest.mark.parametrize("name", ["state_pair", "state_site", "mmm", "mm2"])
def test_cluster_mappings_bit_exact(golden, name):
    mine = capi.tables_mapping(name)
    assert O.canonical_mapping(mine) == H.unflatten_mapping(golden["A_mapping_" + name])
    assert O.canonical_mapping(mine) == H.unflatten_mapping(golden["B_mapping_" + name])


def test_cluster_types_and_encode_layout(golden):
    want = [(int(r[0]), tuple(int(v) for v in r[2:2 + r[1]])) for r in golden["cluster_types"]]
    assert capi.tables_cluster_types([1, 2, 3]) == want
    assert capi.tables_cluster_types([3, 1, 2]) == want                       # order of the input set is irrelevant
    assert len(capi.tables_cluster_types([1, 2, 3, 4])) == 167                # quaternary
    assert O.cluster_types([1, 2, 3, 4]) == capi.tables_cluster_types([1, 2, 3, 4])
    assert len(capi.tables_group_sizes("mmm", 3)) == 711 and len(capi.tables_group_sizes("mm2", 3)) == 1401
    assert len(capi.tables_group_sizes("mmm", 4)) == 1240 and len(capi.tables_group_sizes("mm2", 4)) == 2452   # EnergyUtility.cpp:751
    cfg = O.Config.generate_fcc(4)
    _, _, sizes = O.one_hot_encode(O.mapping_mmm(cfg), np.ones((1, 58), np.uint8), [1, 2, 3])
    assert np.array_equal(sizes, capi.tables_group_sizes("mmm", 3))


@pytest.mark.parametrize("tag,order", [("A", capi.ORDER_REASSIGNED), ("B", capi.ORDER_GENERATE)])
def test_neighbour_and_ordered_lists_bit_exact(golden, tag, order):
    f = int(golden[tag + "_factor"][0])
    e = capi.Engine(f, id_order=order, device=-1)
    for s in (1, 2, 3):
        nn = golden["%s_nn%d" % (tag, s)]
        for site in range(e.num_sites):
            assert np.array_equal(e.neighbors(s, site), nn[site])
    pos = golden[tag + "_positions"]
    for site in range(0, e.num_sites, 7):
        assert np.allclose(np.array(e.site_coords(site)) / (2.0 * f), pos[site], atol=1e-12)
    cfg = H.oracle_config(golden, tag)
    pi, pj = golden[tag + "_pair_i"], golden[tag + "_pair_j"]
    _, _, backward = O.sorted_lists_of_pairs(cfg, pj, pi)
    for k in range(len(pi)):
        s, m, m2, mb = e.pair_lists(pi[k], pj[k])
        assert np.array_equal(s, golden[tag + "_list_state"][k])
        assert np.array_equal(m, golden[tag + "_list_mmm"][k])
        assert np.array_equal(m2, golden[tag + "_list_mm2"][k])
        assert np.array_equal(mb, backward[k])
    for site in range(e.num_sites):
        assert np.array_equal(e.site_list(site), golden[tag + "_list_site"][site])


def test_non_cubic_cell_lists_match_oracle_geometry():
    """Neighbour shells of a non-cubic supercell (both id orders) against the oracle's distance-based search."""
    for order in (capi.ORDER_GENERATE, capi.ORDER_REASSIGNED):
        e = capi.Engine((4, 5, 6), id_order=order, device=-1)
        cfg = O.Config.generate_fcc((4, 5, 6))
        if order == capi.ORDER_REASSIGNED:
            cfg.reassign_lattice_vector()
        for site in range(0, e.num_sites, 13):
            for s in (1, 2, 3):
                assert np.array_equal(e.neighbors(s, site), cfg.nn[s - 1][site])
            assert np.array_equal(np.sort(e.site_list(site)), np.sort(cfg.neighbors_set_of_site(site)))


def _contracted_pair(T, pairs, mig, codes):
    q = T["pair_C"][mig].copy()
    for t in np.nonzero(codes != 0)[0]:
        q += T["pair_A"][mig, t, codes[t]]
    for p, (t, u) in enumerate(pairs):
        if codes[t] != 0 and codes[u] != 0:
            q += T["pair_B"][mig, p, codes[t], codes[u]]
    return q


@pytest.mark.parametrize("tag,order", [("A", capi.ORDER_REASSIGNED), ("B", capi.ORDER_GENERATE)])
def test_contracted_jump_tables_reproduce_reference(golden, tag, order, tmp_path):
    """Q = C + sum A + sum B over the solute sites equals the reference's dE / D / Ks / Ea (1e-9 eV bar)."""
    e = capi.Engine(int(golden[tag + "_factor"][0]), id_order=order, device=-1)
    e.load_coefficients(H.golden_json(golden, tmp_path))
    T, pairs = e.get_tables(), capi.tables_env_pairs("pair")
    assert pairs.shape == (556, 2)
    vac, I, J = golden[tag + "_ev_vac"], golden[tag + "_ev_i"], golden[tag + "_ev_j"]
    for k in range(0, len(I), 3):
        occ = golden[tag + "_ev_base_occ"].copy()
        occ[vac[k]] = 0
        s = e.pair_lists(I[k], J[k])[0]
        codes = np.array([CODE_OF_ENUM[c] for c in occ[np.delete(s, [21, 38])]])
        q = _contracted_pair(T, pairs, CODE_OF_ENUM[occ[J[k]]], codes)
        de, d, ks = q[0], np.exp(q[1]), np.exp(q[2])
        assert abs(de - golden[tag + "_ev_dE"][k]) < 1e-12
        assert abs(d / golden[tag + "_ev_D"][k] - 1) < 1e-12 and abs(ks / golden[tag + "_ev_Ks"][k] - 1) < 1e-12
        assert abs(O.quartic_barrier(de, d, ks) - golden[tag + "_ev_Ea"][k]) < 1e-9


def test_contracted_tables_independent_of_solvent_choice(golden, tmp_path):
    """Any expansion origin gives the same energies (only the speed differs)."""
    js = H.golden_json(golden, tmp_path)
    pairs = capi.tables_env_pairs("pair")
    rng = np.random.default_rng(0)
    codes = rng.integers(0, 3, 58)
    out = []
    for solvent in (1, 2, 3):
        e = capi.Engine(4, device=-1, solvent=solvent)
        e.load_coefficients(js)
        T = e.get_tables()
        s0 = CODE_OF_ENUM[solvent]
        q = T["pair_C"][1].copy()
        for t in range(58):
            if codes[t] != s0:
                q += T["pair_A"][1, t, codes[t]]
        for p, (t, u) in enumerate(pairs):
            if codes[t] != s0 and codes[u] != s0:
                q += T["pair_B"][1, p, codes[t], codes[u]]
        out.append(q)
    assert np.allclose(out[0], out[1], rtol=0, atol=1e-12) and np.allclose(out[0], out[2], rtol=0, atol=1e-12)


@pytest.mark.parametrize("tag,order", [("A", capi.ORDER_REASSIGNED), ("B", capi.ORDER_GENERATE)])
def test_contracted_site_tables_reproduce_reference(golden, tag, order, tmp_path):
    e = capi.Engine(int(golden[tag + "_factor"][0]), id_order=order, device=-1)
    e.load_coefficients(H.golden_json(golden, tmp_path))
    T, pairs = e.get_tables(), capi.tables_env_pairs("site")
    assert pairs.shape == (204, 2)
    occ = golden[tag + "_cmc_occ"]
    sites, new = golden[tag + "_site"], golden[tag + "_site_new"]
    for k in range(0, len(sites), 2):
        codes = np.array([CODE_OF_ENUM[c] for c in occ[np.delete(e.site_list(sites[k]), 21)]])

        def h(x):
            q = T["site_C"][x]
            for t in np.nonzero(codes != 0)[0]:
                q += T["site_A"][x, t, codes[t]]
            for p, (t, u) in enumerate(pairs):
                if codes[t] != 0 and codes[u] != 0:
                    q += T["site_B"][x, p, codes[t], codes[u]]
            return q

        old, nw = CODE_OF_ENUM[occ[sites[k]]], CODE_OF_ENUM[new[k]]
        de = 0.0 if old == nw else h(nw) - h(old)
        assert abs(de - golden[tag + "_site_dE"][k]) < 1e-12


def test_synthetic_inputs_are_reference_compatible(coef_json):
    """The SURVEY 8(d) synthetic JSON parses and has the lengths the reference expects."""
    import json
    co = json.load(open(coef_json))
    assert len(co["Base"]["theta"]) == 95 and len(co["Al"]["mu_x_mmm"]) == 711 and len(co["Zn"]["U_mm2"][0]) == 1401
    e = capi.Engine(4, device=-1)
    e.load_coefficients(coef_json)
    assert e.get_tables()["pair_B"].shape == (3, 556, 3, 3, 3)


@pytest.mark.parametrize("factors", [(4, 4, 4), (4, 5, 6), (6, 4, 5), (8, 8, 8)])
@pytest.mark.parametrize("order", [capi.ORDER_GENERATE, capi.ORDER_REASSIGNED])
def test_kmc_event_order_table_equals_sorted_neighbour_ids(factors, order):
    """The first-order KMC kernels take the event order of the 12 jumps (KineticMcFirstOmp::BuildEventList: ascending
    lattice id of the neighbour) from a 64 x 12 table by the vacancy's boundary / parity class instead of ranking ids per
    step.  The table must reproduce the sorted first-neighbour list for EVERY site: all boundary classes, non-cubic
    cells, both id orders."""
    e = capi.Engine(factors, id_order=order, device=-1)
    n = 4 * factors[0] * factors[1] * factors[2]
    for site in range(n):
        assert np.array_equal(e.kmc_event_order(site), e.neighbors(1, site)), site


@pytest.mark.parametrize("source", ["golden", "synthetic"])
def test_folded_kmc_tables_sit_on_a_binary_grid_with_exact_sums(golden, coef_json, tmp_path, source):
    """The (dE, log E0) tables of the KMC kernels are rounded to multiples of 2^-bits with the largest possible sum below
    2^(51 - bits): then every partial sum is exact in double, so the sequential sums of the half-warp kernel and the tree
    sums of the block-per-walker kernel are the same numbers (which is what makes the tail hand-off deterministic).
    Checks: grid membership, the bound, the rounding error against the unrounded fold, and order independence of random
    environment sums evaluated sequentially, in reverse and as a tree -- in float64, on the host."""
    e = capi.Engine(4, device=-1)
    e.load_coefficients(H.golden_json(golden, tmp_path) if source == "golden" else coef_json)
    F, T = e.kmc_folded_tables(), e.get_tables()
    n = F["C"].shape[0]
    for c, bits in enumerate(F["bits"]):
        scale = 2.0 ** bits
        unrounded = {"C": T["pair_C"][..., 0] if c == 0 else T["pair_C"][..., 2] + 2.0 * T["pair_C"][..., 1],
                     "A": T["pair_A"][..., 0] if c == 0 else T["pair_A"][..., 2] + 2.0 * T["pair_A"][..., 1],
                     "B": T["pair_B"][..., 0] if c == 0 else T["pair_B"][..., 2] + 2.0 * T["pair_B"][..., 1]}
        bound = 0.0
        for m in range(n):
            bound = max(bound, abs(F["C"][m, c]) + np.abs(F["A"][m, :, :, c]).max(axis=1).sum() + np.abs(F["B"][m, :, :, :, c]).max(axis=(1, 2)).sum())
        assert bound < 2.0 ** (52 - bits)                                    # sums of grid points stay exactly representable
        for key in ("C", "A", "B"):
            v = F[key][..., c]
            assert np.array_equal(v * scale, np.rint(v * scale))             # every entry is a multiple of 2^-bits
            assert np.max(np.abs(v - unrounded[key])) <= 0.5 / scale         # and the nearest one
        assert 0.5 / scale < 1e-12                                           # far inside the 1e-9 eV parity tolerance
    # order independence on random environments: species e_t per site, the pair terms of all 556 pairs
    pairs = capi.tables_env_pairs("pair")
    rng = np.random.default_rng(5)
    for trial in range(20):
        m = int(rng.integers(0, n))
        env = rng.integers(0, n, 58)
        for c in range(2):
            terms = [F["C"][m, c]] + [F["A"][m, t, env[t], c] for t in range(58)] + [F["B"][m, p, env[t], env[u], c] for p, (t, u) in enumerate(pairs)]
            terms = np.array(terms)
            forward = 0.0
            for x in terms:
                forward += x
            backward = 0.0
            for x in terms[::-1]:
                backward += x
            tree = terms.copy()
            while len(tree) > 1:
                if len(tree) % 2:
                    tree = np.append(tree, 0.0)
                tree = tree[0::2] + tree[1::2]
            shuffled = 0.0
            for x in rng.permutation(terms):
                shuffled += x
            assert forward == backward == tree[0] == shuffled

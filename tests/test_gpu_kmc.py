"""GPU suite: the batched KMC driver (lmc_kmc_run) against the reference's own KineticMcFirstOmp traces
(golden, generated with a seeded std::mt19937_64) in replay mode -- north_star: "a replay mode that feeds the
reference's random stream must reproduce its event sequence" -- plus determinism / chunking / multi-walker checks."""
import numpy as np
import pytest

from latticemontecarlo_b200 import capi, synth
from oracle import lmc_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(autouse=True, params=["0", "8", "16", "32"], ids=lambda v: "halfwarp" if v == "0" else "team%s" % v)
def kmc_launch_shape(request, monkeypatch):
    """Every test of this module runs on the throughput kernel (half-warp per walker, kmc_run_kernel) and on the three
    instantiations of the latency kernel (block per walker with 8 / 16 / 32 lanes per candidate jump,
    kmc_team_run_kernel); the library reads the switch at every launch.  The second-order driver ignores it."""
    monkeypatch.setenv("LMC_KMC_TEAM_LANES", request.param)
    return request.param


def _engine(golden, tag, tmp_path, n_walkers=1):
    order = capi.ORDER_REASSIGNED if int(golden[tag + "_factor"][1]) else capi.ORDER_GENERATE
    e = capi.Engine(int(golden[tag + "_factor"][0]), id_order=order, n_walkers=n_walkers, device=0)
    e.load_coefficients(H.golden_json(golden, tmp_path))
    return e


@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("run", ["kmc", "kmc_tt"])
def test_replay_reproduces_reference_event_sequence(golden, tag, run, tmp_path):
    e = _engine(golden, tag, tmp_path, n_walkers=3)
    g = lambda k: golden["%s_%s_%s" % (tag, run, k)]
    n = len(g("u1"))
    for w in range(3):
        e.set_occupancy(golden[tag + "_occ"], walker=w)
    e.kmc_reset()
    kw = dict(time_temperature=golden["tt_points"], rate_corrector=True) if run == "kmc_tt" else {}
    u1 = np.tile(g("u1"), (3, 1)); u2 = np.tile(g("u2"), (3, 1))
    tr = e.kmc_run(n, temperature=500.0, replay_u1=u1, replay_u2=u2, trace=True, **kw)
    for w in range(3):
        assert np.array_equal(tr["from"][w], g("from")) and np.array_equal(tr["to"][w], g("to"))
        assert np.array_equal(tr["slot"][w], g("slot"))
        assert np.max(np.abs(tr["Ea"][w] - g("Ea"))) < TOL and np.max(np.abs(tr["dE"][w] - g("dE"))) < TOL
        assert np.allclose(tr["total_rate"][w], g("total_rate"), rtol=1e-9, atol=0)
        assert np.allclose(tr["dt"][w], g("dt"), rtol=1e-9, atol=4e-16 * float(np.abs(g("time")).max()))
        assert np.allclose(tr["temperature"][w], g("temperature"), rtol=1e-12, atol=0)
        assert np.array_equal(e.get_occupancy(w), g("final_occ"))
    st = e.kmc_state()
    assert np.all(st["steps"] == n)
    assert np.allclose(st["time"], g("time")[-1], rtol=1e-9) and np.max(np.abs(st["energy"] - g("energy")[-1])) < TOL
    assert np.all(st["vacancy"] == g("to")[-1])


def test_philox_runs_are_deterministic_and_chunkable(golden, tmp_path):
    e = _engine(golden, "B", tmp_path, n_walkers=64)
    occ = golden["B_occ"]

    def run(chunks):
        for w in range(64):
            e.set_occupancy(occ, walker=w)
        e.kmc_reset()
        for c in chunks:
            e.kmc_run(c, temperatures=np.linspace(400.0, 600.0, 64), seed=1234)
        return e.kmc_state(), e.get_occupancy_all()

    s1, o1 = run([200])
    s2, o2 = run([200])
    s3, o3 = run([50, 150])
    assert np.array_equal(o1, o2) and np.array_equal(s1["time"], s2["time"]) and np.array_equal(s1["energy"], s2["energy"])
    assert np.array_equal(o1, o3) and np.array_equal(s1["time"], s3["time"])          # counter = step number
    assert len(set(s1["vacancy"].tolist())) > 8                                          # walkers decorrelate
    assert np.all(s1["steps"] == 200) and np.all(s1["time"] > 0)
    # every walker still holds exactly one vacancy and the composition is conserved
    assert np.all((o1 == 0).sum(axis=1) == 1)
    assert np.array_equal(np.sort(o1, axis=1), np.tile(np.sort(occ), (64, 1)))
    # hotter walkers run faster clocks per step on average
    assert s1["time"][:16].mean() > s1["time"][-16:].mean()


def test_production_and_instrumented_kernels_agree(golden, tmp_path):
    """kmc_run_kernel<false> (no trace, no replay: the benchmarked instantiation, random numbers drawn 16 steps ahead) and
    kmc_run_kernel<true> (per-step traces) advance the same Philox trajectories: identical states, and the traced hops
    lead from the start to the untraced run's final vacancy sites."""
    e = _engine(golden, "B", tmp_path, n_walkers=48)
    occ = golden["B_occ"]
    temps = np.linspace(420.0, 580.0, 48)

    def run(trace):
        for w in range(48):
            e.set_occupancy(occ, walker=w)
        e.kmc_reset()
        tr = e.kmc_run(37, temperatures=temps, seed=77, trace=trace)          # 37: not a multiple of the 16-step draw-ahead
        tr2 = e.kmc_run(90, temperatures=temps, seed=77, trace=trace)
        return e.kmc_state(), e.get_occupancy_all(), tr, tr2

    s0, o0, _, _ = run(False)
    s1, o1, tr, tr2 = run(True)
    assert np.array_equal(o0, o1)
    for key in ("time", "energy", "steps", "vacancy"):
        assert np.array_equal(s0[key], s1[key]), key
    assert np.array_equal(tr2["to"][:, -1], s0["vacancy"])
    assert np.allclose(tr["dt"].sum(axis=1) + tr2["dt"].sum(axis=1), s0["time"], rtol=1e-12)


def test_walkers_follow_oracle_with_device_random_numbers(coef_json):
    """Philox mode: recompute each step's decision on the CPU oracle from the traced state (events are a function of
    occupancy only, so the traced (from, to, Ea, dE, total_rate) of every step can be verified independently)."""
    f = 5
    occ = synth.random_alloy(f, 0.04, 0.04, seed=5)
    cfg = O.Config.generate_fcc(f, occ)
    cfg.reassign_lattice_vector()
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=2, device=0)
    e.load_coefficients(coef_json)
    for w in range(2):
        e.set_occupancy(cfg.occ, walker=w)
    e.kmc_reset()
    tr = e.kmc_run(12, temperature=450.0, seed=99, trace=True)
    quartic = O.VacancyMigrationPredictorQuartic(coef_json, cfg, H.CODES)
    beta = 1.0 / O.K_BOLTZMANN / 450.0
    for s in range(12):
        vac = int(tr["from"][0, s])
        assert cfg.occ[vac] == 0
        nbrs = cfg.nn[0][vac]
        ea, de = quartic.barrier_and_diff(cfg, np.full(12, vac), nbrs)
        slot = int(tr["slot"][0, s])
        assert int(nbrs[slot]) == int(tr["to"][0, s])
        assert abs(ea[slot] - tr["Ea"][0, s]) < TOL and abs(de[slot] - tr["dE"][0, s]) < TOL
        assert abs(np.exp(-ea * beta).sum() / tr["total_rate"][0, s] - 1) < 1e-10
        assert tr["dt"][0, s] > 0
        cfg.lattice_jump(vac, int(tr["to"][0, s]))
    assert np.array_equal(e.get_occupancy(0), cfg.occ)


def test_kmc_rejects_bad_walkers(golden, tmp_path):
    e = _engine(golden, "A", tmp_path, n_walkers=2)
    e.set_occupancy(golden["A_occ"], walker=0)
    e.set_occupancy(golden["A_cmc_occ"], walker=1)         # no vacancy at all
    with pytest.raises(capi.LmcOutOfRange):
        e.kmc_reset()


def test_large_cell_ramp_and_rate_corrector_follow_oracle(coef_json):
    """BASELINE configs[4] shape at the largest size where the reference's own predictor is valid (f=48, 442k sites;
    SURVEY 7: its mmm/mm2 comparators break at f >= 50): T(t) ramp + rate corrector, device RNG, every step re-derived
    by the oracle from the traced state."""
    f = 48
    occ = synth.random_alloy(f, 0.02, 0.02, seed=42)
    cfg = O.Config.generate_fcc(f, occ)
    cfg.reassign_lattice_vector()
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, device=0)
    e.load_coefficients(coef_json)
    e.set_occupancy(cfg.occ)
    e.kmc_reset()
    points = np.array([[0.0, 300.0], [1e-3, 500.0], [1e-1, 700.0]])
    n = 16
    tr = e.kmc_run(n, temperature=500.0, time_temperature=points, rate_corrector=True, seed=7, trace=True)
    quartic = O.VacancyMigrationPredictorQuartic(coef_json, cfg, H.CODES)
    tt = O.TimeTemperatureInterpolator(points=[tuple(p) for p in points])
    c_vac = float(np.mean(cfg.occ == 0)); c_sol = float(np.mean((cfg.occ != 1) & (cfg.occ != 0)))
    time = 0.0
    for s in range(n):
        vac = int(tr["from"][0, s])
        temperature = tt.temperature(time)
        assert abs(temperature - tr["temperature"][0, s]) < 1e-9 * temperature
        nbrs = cfg.nn[0][vac]
        ea, de = quartic.barrier_and_diff(cfg, np.full(12, vac), nbrs)
        slot = int(tr["slot"][0, s])
        assert int(nbrs[slot]) == int(tr["to"][0, s])
        assert abs(ea[slot] - tr["Ea"][0, s]) < TOL and abs(de[slot] - tr["dE"][0, s]) < TOL
        total = np.exp(-ea / O.K_BOLTZMANN / temperature).sum()
        assert abs(total / tr["total_rate"][0, s] - 1) < 1e-9
        # dt = -ln(u1) / total / 1e13 * correction: the correction factor is implied by dt * total (u1 is the device's)
        corr = O.rate_correction_factor(c_vac, c_sol, temperature)
        assert tr["dt"][0, s] > 0 and corr > 0
        time += float(tr["dt"][0, s])
        cfg.lattice_jump(vac, int(tr["to"][0, s]))
    st = e.kmc_state()
    assert abs(st["time"][0] - time) < 1e-12 * time
    assert np.array_equal(e.get_occupancy(), cfg.occ)


# ------------------------------------------------------------------------------------------------ second-order KMC
@pytest.mark.parametrize("tag", ["A", "B"])
@pytest.mark.parametrize("run", ["chain", "chain_tt"])
def test_chain_replay_reproduces_reference_event_sequence(golden, golden_chain, tag, run, tmp_path):
    """lmc_kmc_chain_run against the traces of the reference's own mc::KineticMcChainOmpi (12 ranks), fed with the
    selecting uniforms of its rank 0: same sites, same slots, same second-order residence times, same final occupancy."""
    e = _engine(golden, tag, tmp_path, n_walkers=2)
    g = lambda k: golden_chain["%s_%s_%s" % (tag, run, k)]
    n = len(g("u2"))
    for w in range(2):
        e.set_occupancy(golden[tag + "_occ"], walker=w)
    e.kmc_reset()
    kw = dict(time_temperature=golden["tt_points"], rate_corrector=True) if run == "chain_tt" else {}
    tr = e.kmc_chain_run(n, temperature=500.0, replay_u=np.tile(g("u2"), (2, 1)), trace=True, **kw)
    for w in range(2):
        assert np.array_equal(tr["from"][w], g("from")) and np.array_equal(tr["to"][w], g("to"))
        assert np.array_equal(tr["slot"][w], g("slot"))
        assert np.max(np.abs(tr["Ea"][w] - g("Ea"))) < TOL and np.max(np.abs(tr["dE"][w] - g("dE"))) < TOL
        assert np.allclose(tr["total_rate"][w], g("total_rate"), rtol=1e-9, atol=0)
        assert np.allclose(tr["dt"][w], g("dt"), rtol=1e-9, atol=4e-16 * float(np.abs(g("time")).max()))
        assert np.allclose(tr["temperature"][w], g("temperature"), rtol=1e-12, atol=0)
        assert np.array_equal(e.get_occupancy(w), g("final_occ"))
    st = e.kmc_state()
    assert np.all(st["steps"] == n)
    assert np.allclose(st["time"], g("time")[-1], rtol=1e-9) and np.max(np.abs(st["energy"] - g("energy")[-1])) < TOL
    assert np.all(st["vacancy"] == g("to")[-1])


def test_chain_runs_are_chunkable_and_suppress_flicker(golden, tmp_path):
    """Philox mode: chunked runs continue exactly (previous_j is carried between calls); and the point of the method:
    immediate returns to the previous site are rarer than in first-order KMC on the same start state."""
    e = _engine(golden, "B", tmp_path, n_walkers=32)
    occ = golden["B_occ"]

    def run(chunks, second_order=True):
        for w in range(32):
            e.set_occupancy(occ, walker=w)
        e.kmc_reset()
        trs = [e.kmc_run(c, temperature=450.0, seed=77, trace=True, second_order=second_order) for c in chunks]
        return e.kmc_state(), e.get_occupancy_all(), {k: np.concatenate([t[k] for t in trs], axis=1) for k in ("from", "to")}

    s1, o1, t1 = run([120])
    s2, o2, t2 = run([40, 80])
    assert np.array_equal(o1, o2) and np.array_equal(s1["time"], s2["time"]) and np.array_equal(t1["to"], t2["to"])
    assert np.all((o1 == 0).sum(axis=1) == 1)
    assert np.array_equal(np.sort(o1, axis=1), np.tile(np.sort(occ), (32, 1)))
    _, _, tf = run([120], second_order=False)
    back = lambda t: float(np.mean(t["to"][:, 1:] == t["from"][:, :-1]))
    assert back(t1) < back(tf)


def test_latency_kernel_walks_the_throughput_kernels_trajectories(golden, tmp_path, monkeypatch):
    """kmc_team_run_kernel adds the contracted table terms in a different order (tree over lanes) than kmc_run_kernel;
    the tables are multiples of a common power of two, so both sums are exact and equal.  Same Philox stream, same rate
    chain, same select: identical trajectories, clocks and energies.  Also the automatic choice (few walkers -> a block
    per walker)."""
    e = _engine(golden, "B", tmp_path, n_walkers=40)
    occ = golden["B_occ"]
    temps = np.linspace(420.0, 580.0, 40)

    def run(shape):
        if shape is None:
            monkeypatch.delenv("LMC_KMC_TEAM_LANES", raising=False)
        else:
            monkeypatch.setenv("LMC_KMC_TEAM_LANES", shape)
        for w in range(40):
            e.set_occupancy(occ, walker=w)
        e.kmc_reset()
        tr = e.kmc_run(300, temperatures=temps, seed=2024, trace=True)
        e.kmc_run(211, temperatures=temps, seed=2024)
        return e.kmc_state(), e.get_occupancy_all(), tr

    s0, o0, t0 = run("0")
    for shape in ("8", "16", "32", None):
        s1, o1, t1 = run(shape)
        assert np.array_equal(o0, o1), shape
        assert np.array_equal(s0["vacancy"], s1["vacancy"]) and np.array_equal(s0["steps"], s1["steps"])
        assert np.array_equal(t0["to"], t1["to"]) and np.array_equal(t0["slot"], t1["slot"])
        # the folded tables sit on a binary grid (exact sums in any order) and both kernels share one rate chain:
        # not just the same sites, the same bits
        assert np.array_equal(t0["Ea"], t1["Ea"]) and np.array_equal(t0["dE"], t1["dE"])
        assert np.array_equal(t0["total_rate"], t1["total_rate"]) and np.array_equal(t0["dt"], t1["dt"])
        assert np.array_equal(s0["time"], s1["time"]) and np.array_equal(s0["energy"], s1["energy"])


def test_one_pass_select_equals_the_sequential_select(golden, tmp_path, monkeypatch):
    """select_event_fast (one pass over the 12 rates, guarded by a rounding margin) must pick the slot the reference's
    total / division / running-sum form picks: LMC_KMC_SELECT_MARGIN=2 sends every step through the sequential form
    (select_event_sequential); trajectories, clocks and energies must be bit-identical on every launch shape."""
    e = _engine(golden, "B", tmp_path, n_walkers=40)
    occ = golden["B_occ"]
    temps = np.linspace(420.0, 580.0, 40)

    def run(margin):
        if margin is None:
            monkeypatch.delenv("LMC_KMC_SELECT_MARGIN", raising=False)
        else:
            monkeypatch.setenv("LMC_KMC_SELECT_MARGIN", margin)
        for w in range(40):
            e.set_occupancy(occ, walker=w)
        e.kmc_reset()
        tr = e.kmc_run(400, temperatures=temps, seed=77, trace=True)
        e.kmc_run(333, temperatures=temps, seed=77)
        return e.kmc_state(), e.get_occupancy_all(), tr

    s0, o0, t0 = run(None)
    s1, o1, t1 = run("2")
    assert np.array_equal(o0, o1)
    for key in ("vacancy", "steps", "time", "energy"):
        assert np.array_equal(s0[key], s1[key]), key
    for key in ("from", "to", "slot", "dt", "Ea", "dE", "total_rate"):
        assert np.array_equal(t0[key], t1[key]), key


@pytest.mark.parametrize("ramp", [False, True], ids=["constant_T", "T_of_t_and_corrector"])
def test_tail_handoff_is_bit_identical(coef_json, monkeypatch, kmc_launch_shape, ramp):
    """Thousands of walkers: once most are through, kmc_run_kernel hands the walkers still running to the latency kernel
    (Engine::kmc_run).  Every launch shape computes bit-identical (dE, log E0, rate) -- the folded tables sit on a binary
    grid, so the sums are exact in any order -- hence clocks, energies, sites and occupancies must not depend on whether,
    or when, the hand-off happens."""
    if kmc_launch_shape != "0":
        pytest.skip("the hand-off belongs to the half-warp kernel's launch")
    monkeypatch.delenv("LMC_KMC_TEAM_LANES", raising=False)
    nw = 2304
    e = capi.Engine(8, n_walkers=nw, device=0)
    e.load_coefficients(coef_json)
    occ = np.stack([synth.random_alloy(8, 0.03, 0.03, seed=900 + w % 37) for w in range(nw)])
    temps = np.linspace(400.0, 600.0, nw)

    # a T(t) ramp that is log-spaced over eight decades of time, so that the walkers' clocks are on it whatever their scale,
    # + the rate corrector
    ramp_table = np.column_stack([np.concatenate([[0.0], np.logspace(-6, 2, 32)]), np.linspace(450.0, 600.0, 33)])
    kw = dict(time_temperature=ramp_table, rate_corrector=True) if ramp else {}

    def run(setting):
        monkeypatch.setenv("LMC_KMC_HANDOFF", setting)
        e.set_occupancy_all(occ)
        e.kmc_reset()
        e.kmc_run(160, temperatures=temps, seed=5, **kw)
        used = e.kmc_last_launch_handoff()
        e.kmc_run(200, temperatures=temps, seed=5, **kw)
        return e.kmc_state(), e.get_occupancy_all(), used

    s0, o0, used0 = run("0")
    assert not used0 and e.kmc_last_launch_lanes() == 0
    for setting in ("0.2", "0.6", "0.97"):
        s1, o1, used1 = run(setting)
        assert used1, setting
        assert np.array_equal(o0, o1), setting
        for key in ("vacancy", "steps", "time", "energy"):
            assert np.array_equal(s0[key], s1[key]), (setting, key)
    assert np.all(s0["steps"] == 360)
    if ramp:
        assert len(np.unique(s0["temperature"])) > 16          # the walkers sit at different points of the ramp

"""GPU suite: the batched CMC / simulated-annealing driver against the reference's CanonicalMcSerial and
SimulatedAnnealing traces (golden; replay mode reproduces their accept/reject sequence) and, for the device-RNG
batched mode, against invariants that need no oracle run (energy bookkeeping == total-energy difference)."""
import numpy as np
import pytest

from latticemontecarlo_b200 import capi, synth
from tests import helpers as H

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _engine(golden, tag, tmp_path, n_walkers=1):
    order = capi.ORDER_REASSIGNED if int(golden[tag + "_factor"][1]) else capi.ORDER_GENERATE
    e = capi.Engine(int(golden[tag + "_factor"][0]), id_order=order, n_walkers=n_walkers, device=0)
    e.load_coefficients(H.golden_json(golden, tmp_path))
    return e


@pytest.mark.parametrize("tag", ["A", "B"])
def test_cmc_replay_reproduces_reference_chain(golden, tag, tmp_path):
    e = _engine(golden, tag, tmp_path, n_walkers=2)
    g = lambda k: golden["%s_cmc_%s" % (tag, k)]
    e.set_occupancy(golden[tag + "_cmc_occ"], walker=1)
    e.set_occupancy(golden[tag + "_cmc_occ"], walker=0)
    e.cmc_reset()
    out = e.cmc_replay(g("a"), g("b"), g("u"), temperature=800.0, walker=1)
    assert np.max(np.abs(out["dE"] - g("dE"))) < TOL
    assert np.max(np.abs(out["energy_before"] - g("energy_before"))) < 1e-9
    assert np.array_equal(e.get_occupancy(1), g("final_occ"))
    assert np.array_equal(e.get_occupancy(0), golden[tag + "_cmc_occ"])          # the other replica is untouched
    st = e.cmc_state()
    assert st["steps"][1] == len(g("a")) and st["steps"][0] == 0
    want_acc = np.diff(g("energy_before")) != 0                  # accept decisions of the reference (all but the last trial)
    assert np.array_equal(out["accepted"][:-1], want_acc)
    assert abs(st["energy"][1] - (g("energy_before")[-1] + (g("dE")[-1] if out["accepted"][-1] else 0.0))) < 1e-9


@pytest.mark.parametrize("tag", ["P", "Q"])
def test_cmc_replay_reproduces_reference_omp_chain(golden, tag, tmp_path):
    """The trace of mc::CanonicalMcOmp (batches of non-interfering trials, dE on the batch-start configuration:
    mc/src/CanonicalMcOmp.cpp:40-92; tests/golden/make_golden_omp.py) replayed through lmc_cmc_replay."""
    import os
    omp = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_omp_v1.npz"), allow_pickle=False)
    g = lambda k: omp["%s_%s" % (tag, k)]
    f, reassign, _, temperature, _ = g("params")
    e = capi.Engine(int(f), id_order=capi.ORDER_REASSIGNED if int(reassign) else capi.ORDER_GENERATE, device=0)
    e.load_coefficients(H.golden_json(golden, tmp_path))
    e.set_occupancy(g("occ"))
    e.cmc_reset()
    out = e.cmc_replay(g("a"), g("b"), g("u"), temperature=float(temperature))
    assert np.max(np.abs(out["dE"] - g("dE"))) < TOL
    assert np.max(np.abs(out["energy_before"] - g("energy_before"))) < 1e-9
    assert np.array_equal(e.get_occupancy(0), g("final_occ"))
    assert abs(e.cmc_state()["energy"][0] - g("final_energy")[0]) < 1e-9


def test_simulated_annealing_replay_reproduces_reference_schedule(golden, tmp_path):
    f, t0, steps = golden["SA_params"]
    e = capi.Engine(int(f), id_order=capi.ORDER_GENERATE, device=0)
    e.load_coefficients(H.golden_json(golden, tmp_path))
    e.set_occupancy(golden["SA_occ"])
    e.cmc_reset(sa_initial_temperature=float(t0), sa_maximum_steps=int(steps))
    out = e.cmc_replay(golden["SA_a"], golden["SA_b"], golden["SA_u"])
    assert np.allclose(out["temperature_before"], golden["SA_temperature_before"], rtol=1e-12, atol=0)
    assert np.max(np.abs(out["energy_before"] - golden["SA_energy_before"])) < 1e-9
    assert np.array_equal(e.get_occupancy(), golden["SA_final_occ"])
    st = e.cmc_state()
    assert abs(st["energy"][0] - golden["SA_final"][0]) < 1e-9 and abs(st["temperature"][0] - golden["SA_final"][1]) < 1e-9


def test_batched_cmc_energy_bookkeeping_and_conservation(coef_json):
    """Device RNG, 4 replicas at different temperatures on a 12x12x12 cell: the accumulated dE of all accepted swaps
    must equal E_total(final) - E_total(initial) (every dE and every applied swap is checked at once), the composition
    is conserved, colder replicas accept less, and runs are reproducible."""
    f = 12
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=4, device=0)
    e.load_coefficients(coef_json)
    occ = synth.random_alloy(f, 0.08, 0.08, seed=11, vacancy_site=None)
    temps = np.array([300.0, 600.0, 1200.0, 2400.0])

    def run():
        for w in range(4):
            e.set_occupancy(occ, walker=w)
        e0 = np.array([e.total_energy(w) for w in range(4)])
        e.cmc_reset()
        e.cmc_run(6000, temperatures=temps, seed=7)
        e.cmc_run(6000, temperatures=temps, seed=7)
        return e0, e.cmc_state(), e.get_occupancy_all()

    e0, st, final = run()
    assert np.all(st["steps"] >= 12000) and np.all(st["steps"] < 12000 + 2048)
    e1 = np.array([e.total_energy(w) for w in range(4)])
    assert np.max(np.abs((e1 - e0) - st["energy"])) < 5e-9
    assert np.array_equal(np.sort(final, axis=1), np.tile(np.sort(occ), (4, 1)))
    ratio = st["accepted"] / st["steps"]
    assert ratio[0] < ratio[1] < ratio[2] < ratio[3] and ratio[3] > ratio[0] + 0.1
    assert st["energy"][0] < st["energy"][3]                     # cold replica relaxes, hot one stays disordered
    _, st2, final2 = run()
    assert np.array_equal(final, final2) and np.array_equal(st["energy"], st2["energy"])


def test_batched_simulated_annealing_schedule(coef_json):
    f = 10
    e = capi.Engine(f, id_order=capi.ORDER_GENERATE, device=0)
    e.load_coefficients(coef_json)
    occ = synth.random_alloy(f, 0.05, 0.05, seed=2, vacancy_site=None)
    e.set_occupancy(occ)
    e0 = e.total_energy()
    max_steps = 40000
    e.cmc_reset(sa_initial_temperature=700.0, sa_maximum_steps=max_steps)
    e.cmc_run(max_steps)
    st = e.cmc_state()
    steps = int(st["steps"][0])
    assert steps >= max_steps
    # baseline cooling T0 * exp(-3 steps/max) (SimulatedAnnealing.cpp:134), possibly lowered by the 0.99 window factor
    base = 700.0 * np.exp(-3.0 * steps / max_steps)
    # ... and raised by at most five x1.10 reheats (:120-131)
    assert 0.3 * base < st["temperature"][0] <= base * 1.1 ** 5 * (1 + 1e-9), (st["temperature"][0], base)
    assert abs((e.total_energy() - e0) - st["energy"][0]) < 5e-9
    assert st["energy"][0] < 0.0                                 # annealing lowers the energy of a random alloy


def test_whole_gpu_single_lattice_driver(coef_json):
    """lmc_cmc_grid_run (one lattice spread over a cooperative grid): energy bookkeeping == total-energy difference,
    composition conserved, reproducible, chunkable; lmc_cmc_run dispatches to it for one large replica; the simulated
    annealing schedule runs through it."""
    f = 24
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=1, device=0)
    e.load_coefficients(coef_json)
    occ = synth.random_alloy(f, 0.05, 0.05, seed=21, vacancy_site=None)

    def run(fn, chunks, **reset):
        e.set_occupancy(occ)
        e0 = e.total_energy()
        e.cmc_reset(**reset)
        for n in chunks:
            fn(n, temperature=700.0, seed=3)
        st = e.cmc_state()
        final = e.get_occupancy(0)
        assert abs((e.total_energy() - e0) - st["energy"][0]) < 5e-9
        assert np.array_equal(np.sort(final), np.sort(occ))
        return st, final

    st1, o1 = run(e.cmc_grid_run, [30000])
    st2, o2 = run(e.cmc_grid_run, [30000])
    assert np.array_equal(o1, o2) and st1["energy"][0] == st2["energy"][0] and st1["steps"][0] == st2["steps"][0]
    assert 30000 <= st1["steps"][0] < 30000 + 148 * 256 and 0.02 < st1["accepted"][0] / st1["steps"][0] < 0.9
    st3, o3 = run(e.cmc_run, [30000])                                # 55k sites, one replica -> same kernel
    assert np.array_equal(o1, o3) and st1["energy"][0] == st3["energy"][0]
    st4, _ = run(e.cmc_grid_run, [40000], sa_initial_temperature=900.0, sa_maximum_steps=40000)
    base = 900.0 * np.exp(-3.0 * st4["steps"][0] / 40000)
    assert 0.3 * base < st4["temperature"][0] <= base * 1.1 ** 5 * (1 + 1e-9) and st4["energy"][0] < 0.0
    with pytest.raises(capi.LmcInvalidArgument):
        capi.Engine(6, n_walkers=2, device=0).cmc_grid_run(10)       # one lattice only


def test_grid_lane_groups_match_lane_pairs(coef_json, monkeypatch):
    """The lane-GROUP evaluation (2 L lanes per trial: dealt cell loads, REDUX mask / mark reduction, walk split by
    neighbour position) against the lane-pair evaluation (L = 1) on the same batches: same proposals, same conflict
    decisions, dE equal to rounding (another summation order) => the same trajectory; energies within 1e-9 eV."""
    f = 16
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=1, device=0)
    e.load_coefficients(coef_json)
    occ = synth.random_alloy(f, 0.08, 0.08, seed=77, vacancy_site=None)      # solute-rich: long partner lists
    ref = None
    for lanes in (1, 2, 4, 8):
        monkeypatch.setenv("LMC_CMC_GRID_LANES", str(lanes))
        e.set_occupancy(occ)
        e0 = e.total_energy()
        e.cmc_reset()
        for n in (4000, 9000):
            e.cmc_grid_run(n, temperature=650.0, seed=11, batch_size=64)
        st = e.cmc_state()
        final = e.get_occupancy(0)
        assert abs((e.total_energy() - e0) - st["energy"][0]) < 5e-9
        if ref is None:
            ref = (st, final)
            continue
        assert st["steps"][0] == ref[0]["steps"][0] and st["accepted"][0] == ref[0]["accepted"][0], lanes
        assert np.array_equal(final, ref[1]), lanes
        assert abs(st["energy"][0] - ref[0]["energy"][0]) < 1e-9, lanes


def test_full_size_lattices_bookkeeping(coef_json):
    """BASELINE configs[1] (40^3, 256k sites) and configs[3] (100^3, 4M sites, simulated annealing) at full size through a
    size-independent identity: the accumulated dE of the accepted swaps equals the total-energy difference, and the
    composition is conserved."""
    for f, trials, sa in ((40, 300000, {}), (100, 1500000, dict(sa_initial_temperature=900.0, sa_maximum_steps=6000000))):
        e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=1, device=0)
        e.load_coefficients(coef_json)
        occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)
        e.set_occupancy(occ)
        e0 = e.total_energy()
        e.cmc_reset(**sa)
        e.cmc_run(trials, temperature=800.0, seed=11)
        st = e.cmc_state()
        final = e.get_occupancy(0)
        assert st["steps"][0] >= trials and st["accepted"][0] > 0.05 * trials
        assert abs((e.total_energy() - e0) - st["energy"][0]) < 1e-7 * max(1.0, abs(st["energy"][0]))
        assert np.array_equal(np.bincount(final, minlength=4), np.bincount(occ, minlength=4))
        e.close()

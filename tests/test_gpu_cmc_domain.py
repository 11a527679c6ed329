"""GPU suite: the domain-decomposed ("sublattice") CMC / SA driver, lmc_cmc_domain_run (cmc_domain_kernels.cuh).

Its chain is not a replay of the reference's (pairs are drawn inside a domain core, see include/lmc_b200.h), so it is
checked by (i) identities that need no oracle run -- every applied swap and every dE at once through
E_total(final) - E_total(initial) == accumulated dE (energy_kernel, itself pinned to EnergyPredictor::GetEnergy in
test_gpu_parity.py), conservation of the composition; (ii) invariance of the trajectory under every launch shape
(lanes per trial, speculative rounds, table form, dynamic hand-out), which is what makes the multi-GPU run equal to the
single-GPU run; (iii) ensemble statistics against mc::CanonicalMcSerial in test_gpu_cmc_stat.py."""
import numpy as np
import pytest

from latticemontecarlo_b200 import capi, synth

pytestmark = pytest.mark.gpu


def _run(e, occ, chunks, temperatures=None, reset=(), **kw):
    e.set_occupancy_all(occ)
    e0 = np.array([e.total_energy(w) for w in range(e.n_walkers)])
    e.cmc_reset(*reset)
    for n in chunks:
        e.cmc_domain_run(n, temperature=800.0, temperatures=temperatures, seed=7, **kw)
    st = e.cmc_state()
    final = e.get_occupancy_all()
    e1 = np.array([e.total_energy(w) for w in range(e.n_walkers)])
    return st, final, e1 - e0


def test_domain_driver_bookkeeping_conservation_and_temperature_order(coef_json):
    f = 12
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=4, device=0)
    e.load_coefficients(coef_json)
    occ = np.stack([synth.random_alloy(f, 0.08, 0.08, seed=11 + w, vacancy_site=None) for w in range(4)])
    temps = np.array([300.0, 600.0, 1200.0, 2400.0])
    st, final, de_total = _run(e, occ, [20000, 20000], temperatures=temps)
    assert np.all(st["steps"] >= 40000)
    assert np.max(np.abs(de_total - st["energy"])) < 5e-9
    for w in range(4):
        assert np.array_equal(np.sort(final[w]), np.sort(occ[w]))
    ratio = st["accepted"] / st["steps"]
    assert ratio[0] < ratio[1] < ratio[2] < ratio[3] and ratio[3] > ratio[0] + 0.1
    assert st["energy"][0] < st["energy"][3]
    st2, final2, _ = _run(e, occ, [20000, 20000], temperatures=temps)             # reproducible
    assert np.array_equal(final, final2) and np.array_equal(st["energy"], st2["energy"]) and np.array_equal(st["steps"], st2["steps"])
    # the batched driver continues on the result (its cell arrays are rebuilt from the occupancy)
    e.cmc_run(3000, temperatures=temps, seed=1)
    e_after = np.array([e.total_energy(w) for w in range(4)])
    e_start = np.array([0.0] * 4)
    st3 = e.cmc_state()
    e.set_occupancy_all(occ)
    e_start = np.array([e.total_energy(w) for w in range(4)])
    assert np.max(np.abs((e_after - e_start) - st3["energy"])) < 5e-9


@pytest.mark.parametrize("f,edge", [(4, 0), (6, 12), (10, 7), (16, 8), (16, 10), ((5, 9, 14), 6), ((12, 4, 7), 0)])
def test_trajectory_does_not_depend_on_the_launch_shape(coef_json, monkeypatch, f, edge):
    """Same seed => same occupancy, steps and accepted counts for every (lanes, speculative rounds), for the A + B table
    form against the difference tables, and for several domains per lane group; energies equal to the rounding of dE
    (the lanes of a trial sum their terms in another order)."""
    e = capi.Engine(f, id_order=capi.ORDER_GENERATE, n_walkers=2, device=0)
    e.load_coefficients(coef_json)
    occ = np.stack([synth.random_alloy(f, 0.06, 0.06, seed=31 + w, vacancy_site=None) for w in range(2)])
    ref = None
    for lanes, spec, env in ((8, 4, {}), (8, 1, {}), (8, 2, {}), (16, 2, {}), (16, 1, {}), (32, 1, {}), (8, 4, {"LMC_CMC_DOMAIN_TABLES": "0"}),
                             (8, 1, {"LMC_CMC_DOMAIN_PASSES": "3"}), (8, 1, {"LMC_CMC_DOMAIN_64REG": "1"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        st, final, de_total = _run(e, occ, [15000, 9000], domain_edge=edge, lanes=lanes, speculate=spec)
        for k in env:
            monkeypatch.delenv(k)
        assert np.max(np.abs(de_total - st["energy"])) < 5e-9
        assert e.cmc_domain_last_shape()["lanes"] == lanes and e.cmc_domain_last_shape()["speculate"] == spec
        if ref is None:
            ref = (st, final)
            continue
        assert np.array_equal(final, ref[1]), (lanes, spec, env)
        assert np.array_equal(st["steps"], ref[0]["steps"]) and np.array_equal(st["accepted"], ref[0]["accepted"]), (lanes, spec, env)
        assert np.max(np.abs(st["energy"] - ref[0]["energy"])) < 1e-9, (lanes, spec, env)


def test_domain_driver_simulated_annealing_and_vacancy(coef_json):
    f = 10
    e = capi.Engine(f, id_order=capi.ORDER_GENERATE, device=0)
    e.load_coefficients(coef_json)
    occ = synth.random_alloy(f, 0.05, 0.05, seed=2)[None, :]                       # one vacancy: a non-solvent site like any other
    max_steps = 400000
    st, final, de_total = _run(e, occ, [max_steps], reset=(700.0, max_steps))
    steps = int(st["steps"][0])
    assert steps >= max_steps
    base = 700.0 * np.exp(-3.0 * steps / max_steps)               # SimulatedAnnealing.cpp:134, lowered / raised by the window and reheat rules
    assert 0.3 * base < st["temperature"][0] <= base * 1.1 ** 5 * (1 + 1e-9), (st["temperature"][0], base)
    assert abs(de_total[0] - st["energy"][0]) < 5e-9 and st["energy"][0] < 0.0
    assert np.array_equal(np.bincount(final[0], minlength=4), np.bincount(occ[0], minlength=4))


def test_two_vacancies_in_range_raise_like_the_reference(coef_json):
    f = 8
    e = capi.Engine(f, id_order=capi.ORDER_GENERATE, device=0)
    e.load_coefficients(coef_json)
    occ = synth.random_alloy(f, 0.05, 0.05, seed=3, vacancy_site=None)
    occ[:] = np.where(np.arange(occ.size) % 5 == 0, 0, occ)      # vacancies everywhere: some trial meets two in one neighbourhood
    e.set_occupancy(occ)
    e.cmc_reset()
    with pytest.raises(capi.LmcOutOfRange):
        e.cmc_domain_run(50000, temperature=900.0, seed=1)


def test_domain_driver_full_size_lattices(coef_json):
    """BASELINE configs[1] (40^3) and configs[3] (100^3, simulated annealing) at full size: energy bookkeeping against the
    total-energy difference, composition conserved."""
    for f, trials, sa in ((40, 3000000, ()), (100, 40000000, (900.0, 4000000000))):
        e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, n_walkers=1, device=0)
        e.load_coefficients(coef_json)
        occ = synth.random_alloy(f, 0.02, 0.02, seed=1000, vacancy_site=None)[None, :]
        st, final, de_total = _run(e, occ, [trials], reset=sa)
        assert st["steps"][0] >= trials and st["accepted"][0] > 0.05 * trials
        assert abs(de_total[0] - st["energy"][0]) < 1e-7 * max(1.0, abs(st["energy"][0]))
        assert np.array_equal(np.bincount(final[0], minlength=4), np.bincount(occ[0], minlength=4))
        e.close()


def test_domain_parameters_are_checked(coef_json):
    e = capi.Engine(8, n_walkers=1, device=0)
    e.load_coefficients(coef_json)
    e.set_occupancy(synth.random_alloy(8, 0.05, 0.05, seed=3, vacancy_site=None))
    for bad in (dict(domain_edge=3), dict(domain_edge=60), dict(lanes=4), dict(lanes=32, speculate=2), dict(lanes=16, speculate=4)):
        with pytest.raises(capi.LmcInvalidArgument):
            e.cmc_domain_run(1000, **bad)

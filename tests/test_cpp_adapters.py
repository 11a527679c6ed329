"""The C++ host-side mirror of the reference's interfaces (include/lmc_b200_adapters.hpp): it must compile and link
against the C ABI on the CPU box, fail loudly without a device, and reproduce the Python/C-ABI numbers on the GPU."""
import os
import re
import subprocess

import numpy as np
import pytest

from latticemontecarlo_b200 import build as _build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    _build.build()
    exe = str(tmp_path_factory.mktemp("demo") / "adapter_demo")
    libdir = os.path.join(ROOT, "latticemontecarlo_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "adapter_demo.cpp"), "-L" + libdir, "-llmc_b200", "-Wl,-rpath," + libdir, "-o", exe],
                   check=True)
    return exe


def test_adapters_compile_and_fail_loudly_without_gpu(demo, coef_json):
    if capi.device_count() > 0:
        pytest.skip("a GPU is present; the loud-failure path is for the CPU box")
    res = subprocess.run([demo, coef_json], capture_output=True, text=True)
    assert res.returncode == 1 and "not available" in res.stderr


@pytest.mark.gpu
def test_adapter_demo_matches_c_abi(demo, coef_json):
    res = subprocess.run([demo, coef_json, "6"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    ea = np.array([float(m) for m in re.findall(r"Ea = ([-0-9.e+]+) eV", res.stdout)])
    de = np.array([float(m) for m in re.findall(r"dE = ([-+0-9.e]+) eV", res.stdout)])
    assert len(ea) == 12 and "std::out_of_range as in the reference" in res.stdout and "KMC: 1000 steps" in res.stdout
    assert "chain KMC: 500 steps" in res.stdout
    # CanonicalMcOmp adapter, global pairs and domain decomposition: the accumulated dE equals the change of the total energy
    m = re.search(r"^CMC: (\d+) trials, (\d+) accepted, energy drift ([-0-9.e+]+) eV", res.stdout, re.M)
    assert m and int(m.group(1)) >= 2000 and int(m.group(2)) > 0 and float(m.group(3)) < 1e-8
    m = re.search(r"^domain CMC: (\d+) trials, (\d+) accepted, energy drift ([-0-9.e+]+) eV", res.stdout, re.M)
    assert m and int(m.group(1)) >= 2000 and int(m.group(2)) > 0 and float(m.group(3)) < 1e-8
    # the same occupancy through the Python binding: the demo draws it with std::mt19937_64(42), re-created here
    f = 6
    e = capi.Engine(f, device=0)
    e.load_coefficients(coef_json)
    occ = _mt19937_64_alloy(4 * f ** 3)
    e.set_occupancy(occ)
    vac = len(occ) // 2 + 3
    ea2, de2 = e.eval_barriers(np.full(12, vac), e.neighbors(1, vac))
    assert np.max(np.abs(ea - ea2)) < 1e-11 and np.max(np.abs(de - de2)) < 1e-11


@pytest.mark.gpu
def test_adapter_alternative_predictors(demo, coef_json, tmp_path):
    """VacancyMigrationPredictorE0Lru / EnergyChangePredictorPair / EnergyChangePredictorSite adapters against the C ABI,
    sharing one Config with the quartic predictor (model hand-over on the engine)."""
    from latticemontecarlo_b200 import synth
    e0_json = str(tmp_path / "e0.json")
    synth.write_synthetic_json(e0_json, model="e0", k_mmm=6)
    res = subprocess.run([demo, coef_json, "6", e0_json], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    rows = np.array([[float(v) for v in m] for m in
                     re.findall(r"E0 model \d+ -> \d+  barrier ([-0-9.e+]+)  change ([-+0-9.e]+)  pair-predictor ([-+0-9.e]+)", res.stdout)])
    assert rows.shape == (12, 3) and "pair predictor: std::out_of_range" in res.stdout
    f = 6
    e = capi.Engine(f, device=0)
    occ = _mt19937_64_alloy(4 * f ** 3)
    e.set_occupancy(occ)
    vac = len(occ) // 2 + 3
    nb = e.neighbors(1, vac)
    e.load_coefficients(e0_json, model=capi.BARRIER_E0)
    ea, de = e.eval_barriers(np.full(12, vac), nb)
    assert np.max(np.abs(rows[:, 0] - ea)) < 1e-11 and np.max(np.abs(rows[:, 1] - de)) < 1e-11
    # the engine holds one coefficient file at a time: the pair / site predictors see the E0 file's Base.theta here
    assert np.max(np.abs(rows[:, 2] - e.eval_pair_de(np.full(12, vac), nb))) < 1e-11
    m = re.search(r"site \d+ -> Mg: ([-+0-9.e]+)", res.stdout)
    assert abs(float(m.group(1)) - e.eval_site_de([nb[0]], [2])[0]) < 1e-11
    e.load_coefficients(coef_json)
    q = re.search(r"quartic after E0: ([-0-9.e+]+) ([-+0-9.e]+)", res.stdout)
    ea_q, de_q = e.eval_barriers([vac], [nb[0]])
    assert abs(float(q.group(1)) - ea_q[0]) < 1e-11 and abs(float(q.group(2)) - de_q[0]) < 1e-11


def _mt19937_64_alloy(n):
    """std::mt19937_64(42) + uniform_real_distribution<double>(0,1): one 64-bit draw x, u = double(x) / 2^64
    (libstdc++ generate_canonical)."""
    nn, mm = 312, 156
    mt = [0] * nn
    mt[0] = 42
    for i in range(1, nn):
        mt[i] = (6364136223846793005 * (mt[i - 1] ^ (mt[i - 1] >> 62)) + i) & 0xFFFFFFFFFFFFFFFF
    idx = nn
    occ = np.ones(n, dtype=np.uint8)
    for k in range(n):
        if idx >= nn:
            for i in range(nn):
                x = (mt[i] & 0xFFFFFFFF80000000) | (mt[(i + 1) % nn] & 0x7FFFFFFF)
                mt[i] = mt[(i + mm) % nn] ^ (x >> 1) ^ (0xB5026F5AA96619E9 if x & 1 else 0)
            idx = 0
        y = mt[idx]; idx += 1
        y ^= (y >> 29) & 0x5555555555555555
        y ^= (y << 17) & 0x71D67FFFEDA60000
        y ^= (y << 37) & 0xFFF7EEE000000000
        y ^= y >> 43
        u = float(y) / 18446744073709551616.0
        occ[k] = 2 if u < 0.02 else (3 if u < 0.04 else 1)
    occ[n // 2 + 3] = 0
    return occ


@pytest.fixture(scope="module")
def energy_demo(tmp_path_factory):
    _build.build()
    exe = str(tmp_path_factory.mktemp("demo2") / "adapter_energy_demo")
    libdir = os.path.join(ROOT, "latticemontecarlo_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "adapter_energy_demo.cpp"), "-L" + libdir, "-llmc_b200", "-Wl,-rpath," + libdir, "-o", exe],
                   check=True)
    return exe


def test_energy_demo_compiles_and_fails_loudly_without_gpu(energy_demo, coef_json, tmp_path):
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    (tmp_path / "occ.bin").write_bytes(bytes([1] * 256))
    res = subprocess.run([energy_demo, coef_json, str(tmp_path / "occ.bin"), "4", "0"], capture_output=True, text=True)
    assert res.returncode == 1 and "not available" in res.stderr


@pytest.mark.gpu
def test_atom_id_entry_points_and_energy_predictor_surface(energy_demo, golden, tmp_path):
    """GetBarrierAndDiffFromAtomIdPair, GetDeFromAtomIdPair / Site, Config read accessors, EnergyPredictor::GetEnergyOfCluster /
    GetEncode / GetChemicalPotential through the C++ adapters, against the reference's own values
    (tests/golden/golden_energy_v1.npz, make_golden_energy.py) and the C ABI."""
    from tests import helpers as H
    g = np.load(os.path.join(ROOT, "tests", "golden", "golden_energy_v1.npz"), allow_pickle=False)
    js = H.golden_json(golden, tmp_path)
    f, reassign = (int(v) for v in g["B_params"])
    occ = g["B_occ"]
    (tmp_path / "occ.bin").write_bytes(occ.tobytes())
    res = subprocess.run([energy_demo, js, str(tmp_path / "occ.bin"), str(f), str(reassign)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = res.stdout
    val = lambda pat: float(re.search(pat, out).group(1))
    assert abs(val(r"E = ([-0-9.e+]+)") - g["B_energy"][0]) < 1e-9
    assert abs(val(r"encode n = 95 sum = ([-0-9.e+]+)") - g["B_encode"].sum()) < 1e-9
    mu = {int(a): float(b) for a, b in re.findall(r"mu\[(\d+)\] = ([-0-9.e+]+)", out)}
    assert set(mu) == set(int(v) for v in g["mu_elements"])
    for el, v in zip(g["mu_elements"], g["mu_values"]):
        assert abs(mu[int(el)] - v) < 1e-9
    assert "(same atom: 1)" in out and "clone independent: 1" in out
    assert abs(val(r"dE check ([-0-9.e+]+)")) < 1e-9
    m = re.search(r"swap dE by atom ids ([-+0-9.e]+) by lattice ids ([-+0-9.e]+)", out)
    assert m.group(1) == m.group(2)
    # the same calls through the C ABI (atom id == lattice id before the jump in GenerateFCC order)
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED if reassign else capi.ORDER_GENERATE, device=0)
    e.load_coefficients(js)
    e.set_occupancy(occ)
    vac, n_vac = e.find_element(0)
    assert n_vac == 1 and re.search(r"vacancy lattice %d atom %d element 0" % (vac, vac), out)
    assert abs(val(r"E_cluster = ([-0-9.e+]+)") - e.energy_of_cluster([3, 17, 40, vac])) < 1e-11
    j = int(e.neighbors(1, vac)[4])
    ea, de = e.eval_barriers([vac], [j])
    m = re.search(r"jump by atom ids: Ea = ([-0-9.e+]+) dE = ([-+0-9.e]+)", out)
    assert abs(float(m.group(1)) - ea[0]) < 1e-11 and abs(float(m.group(2)) - de[0]) < 1e-11
    e.lattice_jump(vac, j)
    # after the jump the vacancy ATOM sits on lattice site j, the moved atom on the old vacancy site
    assert abs(val(r"E_cluster after jump = ([-0-9.e+]+)") - e.energy_of_cluster([3, 17, 40, j])) < 1e-11
    assert abs(val(r"site dE by atom id ([-+0-9.e]+)") - e.eval_site_de([vac], [3])[0]) < 1e-11

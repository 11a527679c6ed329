"""`lmc_b200.exe -p kmc_param.txt` against what the reference's `lmc.exe -p` wrote for the same start.cfg, parameter
file and random stream (tests/golden/cli_v1, generated from oracle/_ref by tests/golden/make_golden_cli.py)."""
import gzip
import os
import shutil
import subprocess

import numpy as np
import pytest

from latticemontecarlo_b200 import build as _build, capi, synth
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "cli_v1")


@pytest.fixture(scope="module")
def exe():
    _build.build()
    return _build.build_cli()


def _rows(text):
    lines = text.strip().split("\n")
    return lines[0], [line.replace("\t", " ").split() for line in lines[1:]]


def test_cli_reports_missing_inputs(exe, tmp_path):
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 1 and "No input parameter filename." in res.stdout           # main.cpp:25-31
    (tmp_path / "p.txt").write_text("simulation_method KineticMcFirstOmp\nconfig_filename nope.cfg\nelement_set Al Mg Zn\n")
    res = subprocess.run([exe, "-p", "p.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 1 and "Cannot open nope.cfg" in res.stderr                  # Config.cpp:556-558
    (tmp_path / "q.txt").write_text("# comment\nsimulation_method Nonsense\nunknown_key 1\n")
    res = subprocess.run([exe, "-p", "q.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert "No such method: Nonsense" in res.stdout                                       # Home.cpp:123


@pytest.mark.gpu
def test_kmc_cli_reproduces_reference_log_and_dumps(exe, golden, tmp_path):
    for name in ("start.cfg", "kmc_param.txt", "uniforms.txt", "time_temperature.dat"):
        shutil.copy(os.path.join(GOLD, name), tmp_path / name)
    shutil.copy(H.golden_json(golden, tmp_path), tmp_path / "coefficients.json")
    res = subprocess.run([exe, "-p", "kmc_param.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 0, res.stderr
    head, mine = _rows((tmp_path / "kmc_log.txt").read_text())
    head_ref, ref = _rows(open(os.path.join(GOLD, "kmc_log.txt")).read())
    assert head == head_ref and len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert a[0] == b[0] and a[6] == b[6]                                   # step number, selected atom id: exact
        va, vb = np.array([float(x) for x in a]), np.array([float(x) for x in b])
        assert np.allclose(va[[1, 2]], vb[[1, 2]], rtol=1e-9, atol=0)         # time, temperature
        assert np.max(np.abs(va[3:6] - vb[3:6])) < 1e-9                        # energy, Ea, dE  (eV)
        assert np.max(np.abs(va[7:10] - vb[7:10])) < 1e-9                      # unwrapped vacancy position (A)
    # text formatting of the log follows the reference's stream state (default float on the first row, fixed afterwards)
    first, second = (tmp_path / "kmc_log.txt").read_text().split("\n")[1:3]
    assert first.split("\t")[1] == "0" and "." in second.split("\t")[2] and len(second.split("\t")[2].split(".")[1]) == 16
    # configuration dumps: identical text (atom identities, positions, periodic image counters)
    for name in ("0", "25", "end"):
        got = gzip.open(tmp_path / (name + ".cfg.gz"), "rt").read()
        assert got == open(os.path.join(GOLD, name + ".cfg.txt")).read(), name


@pytest.mark.gpu
def test_chain_kmc_cli_reproduces_reference_log_and_dumps(exe, golden, tmp_path):
    """simulation_method KineticMcChainOmpi (what script/kmc_param.txt selects), solute_disp on: the reference ran as 12
    ranks, the engine as one block of 12 half-warps; same log rows (incl. the solute centre-of-mass columns) and dumps."""
    gold = os.path.join(ROOT, "tests", "golden", "cli_chain_v1")
    for name in ("kmc_param.txt", "uniforms.txt"):
        shutil.copy(os.path.join(gold, name), tmp_path / name)
    for name in ("start.cfg", "time_temperature.dat"):
        shutil.copy(os.path.join(GOLD, name), tmp_path / name)
    shutil.copy(H.golden_json(golden, tmp_path), tmp_path / "coefficients.json")
    res = subprocess.run([exe, "-p", "kmc_param.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 0, res.stderr
    head, mine = _rows((tmp_path / "kmc_log.txt").read_text())
    head_ref, ref = _rows(open(os.path.join(gold, "kmc_log.txt")).read())
    assert head == head_ref and len(mine) == len(ref)
    for a, b in zip(mine, ref):
        assert a[0] == b[0] and a[6] == b[6] and len(a) == len(b) == 13
        va, vb = np.array([float(x) for x in a]), np.array([float(x) for x in b])
        assert np.allclose(va[[1, 2]], vb[[1, 2]], rtol=1e-9, atol=0)         # second-order time, temperature
        assert np.max(np.abs(va[3:6] - vb[3:6])) < 1e-9                        # energy, Ea, dE  (eV)
        assert np.max(np.abs(va[7:13] - vb[7:13])) < 1e-9                      # vacancy position, solute centre of mass (A)
    for name in ("0", "25", "end"):
        got = gzip.open(tmp_path / (name + ".cfg.gz"), "rt").read()
        assert got == open(os.path.join(gold, name + ".cfg.txt")).read(), name


@pytest.mark.gpu
def test_kmc_cli_from_map_files(exe, golden, tmp_path):
    """map_filename input (Config::ReadMap): lattice.txt / element.txt / map.txt as written by the reference
    (tests/golden/cli_map_v1; GenerateFCC id order, kept as is).  The reference itself cannot run from this input (see
    make_golden_cli.from_map), so the CLI's log is checked against the C ABI driven in the same id order with the same
    uniforms; the cfg path's parity with the reference is covered by the tests above."""
    gold = os.path.join(ROOT, "tests", "golden", "cli_map_v1")
    for name in ("lattice.txt", "element.txt", "map.txt", "kmc_param.txt", "uniforms.txt"):
        shutil.copy(os.path.join(gold, name), tmp_path / name)
    shutil.copy(os.path.join(GOLD, "time_temperature.dat"), tmp_path / "time_temperature.dat")
    js = H.golden_json(golden, tmp_path)
    shutil.copy(js, tmp_path / "coefficients.json")
    res = subprocess.run([exe, "-p", "kmc_param.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 0, res.stderr
    _, rows = _rows((tmp_path / "kmc_log.txt").read_text())
    occ = np.load(os.path.join(gold, "occupancy_by_lattice_id.npy"))
    u = np.loadtxt(os.path.join(gold, "uniforms.txt"))
    tt = np.loadtxt(os.path.join(GOLD, "time_temperature.dat"))
    e = capi.Engine(4, id_order=capi.ORDER_GENERATE, device=0)
    e.load_coefficients(js)
    e.set_occupancy(occ)
    e.kmc_reset()
    n = len(u)
    tr = e.kmc_run(n, temperature=500.0, time_temperature=tt, rate_corrector=True, replay_u1=u[:, 0], replay_u2=u[:, 1], trace=True)
    assert len(rows) == n
    time = np.concatenate([[0.0], np.cumsum(tr["dt"][0])])
    energy = np.concatenate([[0.0], np.cumsum(tr["dE"][0])])
    for s, r in enumerate(rows):
        v = [float(x) for x in r]
        assert int(v[0]) == s
        assert np.isclose(v[1], time[s], rtol=1e-9, atol=0) and abs(v[3] - energy[s]) < 1e-9
        assert abs(v[4] - tr["Ea"][0][s]) < 1e-9 and abs(v[5] - tr["dE"][0][s]) < 1e-9
    # bad inputs fail like the reference's ReadMap (Config.cpp:819-879)
    (tmp_path / "map.txt").write_text("0\n0\n")
    res = subprocess.run([exe, "-p", "kmc_param.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 1 and "Duplicate lattice id in map file: 0" in res.stderr
    os.remove(tmp_path / "lattice.txt")
    res = subprocess.run([exe, "-p", "kmc_param.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 1 and "Cannot open lattice.txt" in res.stderr


@pytest.mark.gpu
def test_cmc_and_sa_cli_run_and_log(exe, golden, coef_json, tmp_path):
    """CanonicalMcOmp / SimulatedAnnealing through the CLI: log format, monotone step counter, energy bookkeeping
    against the total energy of the dumped configurations (re-read through the engine)."""
    from oracle import lmc_oracle as O
    f = 6
    occ = synth.random_alloy(f, 0.06, 0.06, seed=31, vacancy_site=None)
    # start.cfg in the reference's format, written by this repository's own writer through a throw-away KMC-less path:
    # build it from the oracle's positions
    cfg = O.Config.generate_fcc(f, occ)
    with open(tmp_path / "start.cfg", "w") as fh:
        fh.write("Number of particles = %d\nA = 1.0 Angstrom (basic length-scale)\n" % cfg.num_sites)
        for i in range(3):
            for j in range(3):
                fh.write("H0(%d,%d) = %.16g A\n" % (i + 1, j + 1, cfg.basis[i][j]))
        fh.write(".NO_VELOCITY.\nentry_count = 6\nauxiliary[0] = ix\nauxiliary[1] = iy\nauxiliary[2] = iz\n")
        for k in range(cfg.num_sites):
            name = synth.ELEMENT_NAMES[int(occ[k])]
            fh.write("%.16g\n%s\n%.16f %.16f %.16f 0 0 0\n" % (synth.ELEMENT_MASS[name], name, *cfg.rel[k]))
    shutil.copy(coef_json, tmp_path / "c.json")
    (tmp_path / "cmc.txt").write_text("simulation_method CanonicalMcOmp\njson_coefficients_filename c.json\nconfig_filename start.cfg\n"
                                      "log_dump_steps 500\nconfig_dump_steps 100000\nmaximum_steps 3000\n"
                                      "thermodynamic_averaging_steps 100\ntemperature 800\nelement_set Al Mg Zn\nseed 5\n")
    res = subprocess.run([exe, "-p", "cmc.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 0, res.stderr
    head, rows = _rows((tmp_path / "cmc_log.txt").read_text())
    assert head.split("\t") == ["steps", "temperature", "energy", "average_energy", "absolute_energy"]
    steps = [int(r[0]) for r in rows]
    assert steps[0] == 0 and steps == sorted(steps) and steps[-1] == 3000
    # absolute_energy - energy is the constant initial total energy; the final dump has that absolute energy
    offs = [float(r[4]) - float(r[2]) for r in rows]
    assert max(offs) - min(offs) < 1e-8
    e = capi.Engine(f, id_order=capi.ORDER_REASSIGNED, device=0)
    e.load_coefficients(coef_json)

    def occupancy_of(path):
        lines = gzip.open(path, "rt").read().split("\n")
        start = next(i for i, l in enumerate(lines) if l.startswith("auxiliary[2]")) + 1
        out = np.zeros(e.num_sites, np.uint8)
        for a in range(e.num_sites):
            name = lines[start + 3 * a + 1]
            x, y, z = (float(v) for v in lines[start + 3 * a + 2].split()[:3])
            X, Y, Z = (int(round(v * 2 * f)) for v in (x, y, z))
            out[X * 2 * f * f + Y * f + Z // 2] = synth.ELEMENT_CODES[name]
        return out

    final = occupancy_of(tmp_path / "end.cfg.gz")
    assert np.array_equal(np.sort(final), np.sort(occ))
    e.set_occupancy(final)
    assert abs(e.total_energy() - float(rows[-1][4])) < 1e-6
    # the same run with the domain-decomposed driver (extension keys): whole sweeps, same bookkeeping
    os.rename(tmp_path / "cmc_log.txt", tmp_path / "cmc_log_global.txt")
    (tmp_path / "cmc_dom.txt").write_text((tmp_path / "cmc.txt").read_text() + "domain_edge 6\nrounds_per_sweep 32\n")
    res = subprocess.run([exe, "-p", "cmc_dom.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 0, res.stderr
    head, rows = _rows((tmp_path / "cmc_log.txt").read_text())
    steps = [int(r[0]) for r in rows]
    assert steps[0] == 0 and steps == sorted(steps) and steps[-1] == 3000
    offs = [float(r[4]) - float(r[2]) for r in rows]
    assert max(offs) - min(offs) < 1e-8
    final = occupancy_of(tmp_path / "end.cfg.gz")
    assert np.array_equal(np.sort(final), np.sort(occ)) and not np.array_equal(final, occ)
    e.set_occupancy(final)
    assert abs(e.total_energy() - float(rows[-1][4])) < 1e-6
    (tmp_path / "sa.txt").write_text("simulation_method SimulatedAnnealing\njson_coefficients_filename c.json\nfactor 6\nsolvent_element Al\n"
                                     "solute_element_set Mg Zn\nsolute_number_set 12 15\nlog_dump_steps 1000\nconfig_dump_steps 100000\n"
                                     "maximum_steps 6000\ninitial_temperature 700\nseed 9\n")
    res = subprocess.run([exe, "-p", "sa.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 0 and "initial_energy = " in res.stdout, res.stderr
    head, rows = _rows((tmp_path / "sa_log.txt").read_text())
    assert head.split("\t") == ["steps", "temperature", "energy", "lowest_energy", "absolute_energy"]
    temps = [float(r[1]) for r in rows]
    assert temps[0] == 700.0 and temps[-1] < 0.2 * temps[0]                  # T0 exp(-3) at the end (SimulatedAnnealing.cpp:134)
    final = occupancy_of(tmp_path / "end.cfg.gz")
    assert (final == 2).sum() == 12 and (final == 3).sum() == 15
    assert os.path.exists(tmp_path / "lowest_energy.cfg.gz")
    # simulated annealing with the domain-decomposed driver: same log format, schedule evaluated per sweep
    (tmp_path / "sa_dom.txt").write_text((tmp_path / "sa.txt").read_text() + "domain_edge 6\nrounds_per_sweep 16\n")
    res = subprocess.run([exe, "-p", "sa_dom.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 0 and "initial_energy = " in res.stdout, res.stderr
    head, rows = _rows((tmp_path / "sa_log.txt").read_text())
    temps = [float(r[1]) for r in rows]
    assert temps[0] == 700.0 and temps[-1] < 0.2 * temps[0] and int(rows[-1][0]) == 6000
    assert float(rows[-1][3]) <= float(rows[0][2]) + 1e-9                        # lowest_energy never above the start
    final = occupancy_of(tmp_path / "end.cfg.gz")
    assert (final == 2).sum() == 12 and (final == 3).sum() == 15


@pytest.mark.gpu
def test_cmc_cli_replay_reproduces_reference_log_average_energy_and_dumps(exe, golden, tmp_path):
    """simulation_method CanonicalMcSerial with the reference's trial stream (replay_trials_filename): cmc_log.txt incl. the
    `average_energy` column -- mc::ThermodynamicAveraging over a 50-step sliding window, fed every step
    (mc/src/ThermodynamicAveraging.cpp:5-39, CanonicalMcSerial.cpp:40-51) -- and the .cfg dumps (atom identities through
    Config::LatticeJump) against what the reference wrote (tests/golden/cli_cmc_v1, make_golden_cli.cmc)."""
    gold = os.path.join(ROOT, "tests", "golden", "cli_cmc_v1")
    for name in ("start.cfg", "cmc_param.txt", "trials.txt"):
        shutil.copy(os.path.join(gold, name), tmp_path / name)
    shutil.copy(H.golden_json(golden, tmp_path), tmp_path / "coefficients.json")
    res = subprocess.run([exe, "-p", "cmc_param.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 0, res.stderr
    head, mine = _rows((tmp_path / "cmc_log.txt").read_text())
    head_ref, ref = _rows(open(os.path.join(gold, "cmc_log.txt")).read())
    assert head == head_ref and len(mine) == len(ref) and len(ref) > 30
    for a, b in zip(mine, ref):
        assert a[0] == b[0] and float(a[1]) == float(b[1])                     # steps, temperature
        va, vb = np.array([float(x) for x in a[2:]]), np.array([float(x) for x in b[2:]])
        assert np.max(np.abs(va - vb)) < 1e-9, (a, b)                          # energy, average_energy, absolute_energy (eV)
    assert any(abs(float(r[2]) - float(r[3])) > 1e-3 for r in ref[5:])        # the window average is not just the energy
    for name in ("0", "100", "300", "end"):
        got = gzip.open(tmp_path / (name + ".cfg.gz"), "rt").read()
        assert got == open(os.path.join(gold, name + ".cfg.txt")).read(), name


@pytest.mark.gpu
def test_kmc_cli_restart_continues_the_run(exe, golden, tmp_path):
    """Restart (mc/src/McAbstract.cpp:24-34, script/restart.py:51-105): a run resumed from the 30-step dump with
    restart_steps / restart_energy / restart_time taken from the log must continue exactly where the uninterrupted run
    went -- same T(t) (the clock resumes at restart_time), same random stream (the Philox counter is the step number),
    appended log, identical end.cfg.gz."""
    full, part = tmp_path / "full", tmp_path / "part"
    for d in (full, part):
        d.mkdir()
        for name in ("start.cfg", "time_temperature.dat"):
            shutil.copy(os.path.join(GOLD, name), d / name)
        shutil.copy(H.golden_json(golden, tmp_path), d / "coefficients.json")
    base = ("simulation_method KineticMcFirstOmp\njson_coefficients_filename coefficients.json\ntime_temperature_filename time_temperature.dat\n"
            "log_dump_steps 1\nconfig_dump_steps 30\nthermodynamic_averaging_steps 0\ntemperature 500\nelement_set Al Mg Zn\n"
            "rate_corrector true\nearly_stop false\nsolute_disp false\nseed 1234\n")
    (full / "p.txt").write_text(base + "config_filename start.cfg\nmaximum_steps 60\nrestart_steps 0\nrestart_energy 0\nrestart_time 0\n")
    assert subprocess.run([exe, "-p", "p.txt"], capture_output=True, text=True, cwd=full).returncode == 0
    (part / "p.txt").write_text(base + "config_filename start.cfg\nmaximum_steps 30\nrestart_steps 0\nrestart_energy 0\nrestart_time 0\n")
    assert subprocess.run([exe, "-p", "p.txt"], capture_output=True, text=True, cwd=part).returncode == 0
    _, rows = _rows((part / "kmc_log.txt").read_text())
    last = rows[-1]
    assert last[0] == "30"
    with gzip.open(part / "30.cfg.gz", "rt") as fh:
        (part / "restart.cfg").write_text(fh.read())
    (part / "p2.txt").write_text(base + "config_filename restart.cfg\nmaximum_steps 60\nrestart_steps 30\nrestart_energy %s\nrestart_time %s\n" % (last[3], last[1]))
    res = subprocess.run([exe, "-p", "p2.txt"], capture_output=True, text=True, cwd=part)
    assert res.returncode == 0, res.stderr
    _, a = _rows((full / "kmc_log.txt").read_text())
    _, b = _rows((part / "kmc_log.txt").read_text())
    assert [r[0] for r in a] == [r[0] for r in b] == [str(s) for s in range(61)]          # the restarted run appends rows 31..60
    for ra, rb in zip(a, b):
        va, vb = np.array([float(x) for x in ra]), np.array([float(x) for x in rb])
        # (the log prints times in fixed notation with 16 decimals: ~12 significant digits at 3e-5 s)
        assert ra[6] == rb[6] and np.allclose(va[[1, 2]], vb[[1, 2]], rtol=1e-9, atol=0), (ra, rb)     # atom, time, temperature
        assert np.max(np.abs(va[3:6] - vb[3:6])) < 1e-9 and np.max(np.abs(va[7:10] - vb[7:10])) < 1e-9
    assert float(a[45][2]) != float(a[0][2])                                   # the temperature ramp is live in the resumed half
    assert gzip.open(full / "end.cfg.gz", "rt").read() == gzip.open(part / "end.cfg.gz", "rt").read()


def test_cli_rejects_cells_outside_the_engine_geometry(exe, tmp_path):
    """The integer FCC geometry needs an orthogonal cubic-axes cell whose lattice constant keeps the reference's distance
    cutoffs (3.5 / 4.8 / 5.3 A) on the first three FCC shells: anything else is an error, not a silent assumption."""
    text = open(os.path.join(GOLD, "start.cfg")).read()
    base = "simulation_method KineticMcFirstOmp\njson_coefficients_filename none.json\nconfig_filename %s\nelement_set Al Mg Zn\nmaximum_steps 1\n"
    (tmp_path / "sheared.cfg").write_text(text.replace("H0(1,2) = 0 A", "H0(1,2) = 1.5 A"))
    (tmp_path / "p1.txt").write_text(base % "sheared.cfg")
    res = subprocess.run([exe, "-p", "p1.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 1 and "must be orthogonal" in res.stderr, res.stderr
    import re
    stretched = re.sub(r"H0\((\d),\1\) = ([0-9.]+) A", lambda m: "H0(%s,%s) = %.6f A" % (m.group(1), m.group(1), float(m.group(2)) * 1.1), text)
    (tmp_path / "stretched.cfg").write_text(stretched)
    (tmp_path / "p2.txt").write_text(base % "stretched.cfg")
    res = subprocess.run([exe, "-p", "p2.txt"], capture_output=True, text=True, cwd=tmp_path)
    assert res.returncode == 1 and "lattice constant" in res.stderr, res.stderr

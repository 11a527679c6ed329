"""Generate tests/golden/cli_v1/ from the reference itself (oracle/_ref): what `lmc.exe -p kmc_param.txt` with
simulation_method KineticMcFirstOmp writes for a small seeded run -- kmc_log.txt, 0.cfg.gz, end.cfg.gz -- together with
the start.cfg it read and the (u1, u2) stream its std::mt19937_64 delivered.  The coefficient file is the K=4/5 set of
golden_v1.npz.   python tests/golden/make_golden_cli.py"""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from latticemontecarlo_b200 import synth  # noqa: E402
from oracle import ref_lib as R  # noqa: E402
from tests import helpers as H  # noqa: E402

STEPS, SEED, F = 60, 21, 4


def main():
    assert R.build()
    out = os.path.join(ROOT, "tests", "golden", "cli_v1")
    os.makedirs(out, exist_ok=True)
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    work = tempfile.mkdtemp()
    js = H.golden_json(golden, work)
    tt = os.path.join(work, "time_temperature.dat")
    synth.write_time_temperature(tt, points=((0.0, 450.0), (2e-8, 500.0), (1e-6, 650.0)))
    occ = synth.random_alloy(F, 0.03, 0.03, seed=104)
    R.RefConfig.fcc(F, occ, reassign=False).write(os.path.join(work, "start.cfg"))
    cfg = R.RefConfig.read(os.path.join(work, "start.cfg"), reassign=True)
    trace = R.kmc_first_omp(cfg, js, temperature=500.0, maximum_steps=STEPS, seed=SEED, threads=1, tt_file=tt, rate_corrector=True)
    R.kmc_first_omp_with_logs(cfg, js, work, temperature=500.0, maximum_steps=STEPS, log_dump_steps=5, config_dump_steps=25,
                              seed=SEED, tt_file=tt, rate_corrector=True)
    np.savetxt(os.path.join(out, "uniforms.txt"), np.stack([trace["u1"], trace["u2"]], axis=1), fmt="%.17g")
    for name in ("start.cfg", "kmc_log.txt", "0.cfg.gz", "25.cfg.gz", "end.cfg.gz", "time_temperature.dat"):
        dst = name.replace(".cfg.gz", ".cfg.txt")        # the shim's gzip is a pass-through: these are plain text
        shutil.copy(os.path.join(work, name), os.path.join(out, dst))
    with open(os.path.join(out, "kmc_param.txt"), "w") as f:
        f.write("simulation_method KineticMcFirstOmp\njson_coefficients_filename coefficients.json\n"
                "time_temperature_filename time_temperature.dat\nconfig_filename start.cfg\nlog_dump_steps 5\n"
                "config_dump_steps 25\nmaximum_steps %d\nthermodynamic_averaging_steps 0\ntemperature 500\n"
                "element_set Al Mg Zn\nrestart_steps 0\nrestart_energy 0\nrestart_time 0\nrate_corrector true\n"
                "early_stop false\nsolute_disp false\n# extension of lmc_b200 (ignored by the reference):\n"
                "replay_uniforms_filename uniforms.txt\n" % STEPS)
    print(open(os.path.join(out, "kmc_log.txt")).read()[:600])
    print(sorted(os.listdir(out)))
    chain(golden, js, tt)
    from_map(js, tt)
    cmc(js)


def cmc(js):
    """mc::CanonicalMcSerial as shipped (cmc_log.txt with the thermodynamic average over a 50-step window, 0 / 100 / 200 /
    300.cfg.gz, end.cfg.gz) on a vacancy-free 5x5x5 cell, together with the trial stream (site_a, site_b, u) its
    std::mt19937_64 delivered -> tests/golden/cli_cmc_v1/."""
    out = os.path.join(ROOT, "tests", "golden", "cli_cmc_v1")
    os.makedirs(out, exist_ok=True)
    work = tempfile.mkdtemp()
    steps, seed, f = 300, 33, 5
    occ = synth.random_alloy(f, 0.08, 0.08, seed=105, vacancy_site=None)
    R.RefConfig.fcc(f, occ, reassign=False).write(os.path.join(work, "start.cfg"))
    cfg = R.RefConfig.read(os.path.join(work, "start.cfg"), reassign=True)
    trace = R.cmc_serial(cfg, js, temperature=800.0, maximum_steps=steps, seed=seed)
    R.cmc_serial_with_logs(cfg, js, work, temperature=800.0, maximum_steps=steps, log_dump_steps=10, config_dump_steps=100,
                           thermodynamic_averaging_steps=50, seed=seed)
    with open(os.path.join(out, "trials.txt"), "w") as fh:
        for a, b, u in zip(trace["a"], trace["b"], trace["u"]):
            fh.write("%d %d %.17g\n" % (a, b, u))
    for name in ("start.cfg", "cmc_log.txt", "0.cfg.gz", "100.cfg.gz", "300.cfg.gz", "end.cfg.gz"):
        shutil.copy(os.path.join(work, name), os.path.join(out, name.replace(".cfg.gz", ".cfg.txt")))
    with open(os.path.join(out, "cmc_param.txt"), "w") as fh:
        fh.write("simulation_method CanonicalMcSerial\njson_coefficients_filename coefficients.json\nconfig_filename start.cfg\n"
                 "log_dump_steps 10\nconfig_dump_steps 100\nmaximum_steps %d\nthermodynamic_averaging_steps 50\ntemperature 800\n"
                 "element_set Al Mg Zn\nrestart_steps 0\nrestart_energy 0\n# extension of lmc_b200 (ignored by the reference):\n"
                 "replay_trials_filename trials.txt\n" % steps)
    print(open(os.path.join(out, "cmc_log.txt")).read()[:500])
    print(sorted(os.listdir(out)))


def from_map(js, tt):
    """map_filename input (api/src/Home.cpp:133-138 -> Config::ReadMap): lattice.txt / element.txt / map.txt written by the
    reference's own Config::WriteLattice / WriteElement / WriteMap from the GenerateFCC-ordered start configuration (ids are
    NOT reassigned on this path) -> tests/golden/cli_map_v1/.  Only the file formats can be pinned: a reference run started
    through ReadMap crashes in Config::LatticeJump / GetUnwrappedCartesianPositionOfLattice because ReadMap never allocates
    map_shift_list_ (cfg/src/Config.cpp:814-885 vs :283,:448).  The test therefore checks the CLI against this repository's
    C ABI in the same id order."""
    out = os.path.join(ROOT, "tests", "golden", "cli_map_v1")
    os.makedirs(out, exist_ok=True)
    start = R.RefConfig.read(os.path.join(ROOT, "tests", "golden", "cli_v1", "start.cfg"), reassign=False)
    start.write_map_files(os.path.join(out, "lattice.txt"), os.path.join(out, "element.txt"), os.path.join(out, "map.txt"))
    back = R.RefConfig.read_map(os.path.join(out, "lattice.txt"), os.path.join(out, "element.txt"), os.path.join(out, "map.txt"))
    assert np.array_equal(back.occupancy(), start.occupancy())
    np.save(os.path.join(out, "occupancy_by_lattice_id.npy"), start.occupancy())
    rng = np.random.default_rng(77)
    np.savetxt(os.path.join(out, "uniforms.txt"), rng.random((STEPS + 1, 2)), fmt="%.17g")
    with open(os.path.join(out, "kmc_param.txt"), "w") as f:
        f.write("simulation_method KineticMcFirstOmp\njson_coefficients_filename coefficients.json\n"
                "time_temperature_filename time_temperature.dat\nmap_filename map.txt\nlog_dump_steps 1\n"
                "config_dump_steps 1000\nmaximum_steps %d\nthermodynamic_averaging_steps 0\ntemperature 500\n"
                "element_set Al Mg Zn\nrestart_steps 0\nrestart_energy 0\nrestart_time 0\nrate_corrector true\n"
                "early_stop false\nsolute_disp false\nreplay_uniforms_filename uniforms.txt\n" % STEPS)
    print(sorted(os.listdir(out)))


def chain(golden, js, tt):
    """The same start.cfg through mc::KineticMcChainOmpi (12 thread-ranks), the method script/kmc_param.txt selects, with
    the solute centre-of-mass columns on -> tests/golden/cli_chain_v1/."""
    out = os.path.join(ROOT, "tests", "golden", "cli_chain_v1")
    os.makedirs(out, exist_ok=True)
    work = tempfile.mkdtemp()
    shutil.copy(os.path.join(ROOT, "tests", "golden", "cli_v1", "start.cfg"), os.path.join(work, "start.cfg"))
    cfg = R.RefConfig.read(os.path.join(work, "start.cfg"), reassign=True)
    trace = R.kmc_chain_ompi(cfg, js, temperature=500.0, maximum_steps=STEPS, seed=SEED + 1, tt_file=tt, rate_corrector=True)
    R.kmc_chain_ompi_with_logs(cfg, js, work, temperature=500.0, maximum_steps=STEPS, log_dump_steps=5, config_dump_steps=25,
                               seed=SEED + 1, tt_file=tt, rate_corrector=True, solute_disp=True)
    np.savetxt(os.path.join(out, "uniforms.txt"), trace["u2"], fmt="%.17g")
    for name in ("kmc_log.txt", "0.cfg.gz", "25.cfg.gz", "end.cfg.gz"):
        shutil.copy(os.path.join(work, name), os.path.join(out, name.replace(".cfg.gz", ".cfg.txt")))
    with open(os.path.join(out, "kmc_param.txt"), "w") as f:
        f.write("simulation_method KineticMcChainOmpi\njson_coefficients_filename coefficients.json\n"
                "time_temperature_filename time_temperature.dat\nconfig_filename start.cfg\nlog_dump_steps 5\n"
                "config_dump_steps 25\nmaximum_steps %d\nthermodynamic_averaging_steps 0\ntemperature 500\n"
                "element_set Al Mg Zn\nrestart_steps 0\nrestart_energy 0\nrestart_time 0\nrate_corrector true\n"
                "early_stop false\nsolute_disp true\n# extension of lmc_b200 (ignored by the reference):\n"
                "replay_uniforms_filename uniforms.txt\n" % STEPS)
    print(open(os.path.join(out, "kmc_log.txt")).read()[:600])
    print(sorted(os.listdir(out)))


if __name__ == "__main__":
    main()

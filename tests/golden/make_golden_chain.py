"""Generate tests/golden/golden_chain_v1.npz: traces of the reference's second-order KMC driver
(mc::KineticMcChainOmpi, unmodified sources compiled by oracle/Makefile, its 12 MPI ranks run as 12 threads over
oracle/shims/mpi.h) on the two small cases of golden_v1.npz (same coefficients, same start configurations).
Run from the repo root where /root/reference exists:   python tests/golden/make_golden_chain.py
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from latticemontecarlo_b200 import synth  # noqa: E402
from oracle import ref_lib as R  # noqa: E402
import helpers as H  # noqa: E402

ELEMENTS = ("Al", "Mg", "Zn")
KEYS = ("u2", "from", "to", "slot", "dt", "time", "energy", "Ea", "dE", "temperature", "total_rate", "final_occ")


def main():
    assert R.build(), "needs /root/reference to compile oracle/_ref"
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"))
    tmp = tempfile.mkdtemp()
    js = os.path.join(tmp, "coef.json")
    with open(js, "w") as f:
        json.dump(H.golden_coefficients(golden), f)
    tt = os.path.join(tmp, "tt.dat")
    synth.write_time_temperature(tt)
    out = {}
    for tag in ("A", "B"):
        f, reassign = (int(v) for v in golden[tag + "_factor"])
        cfg = R.RefConfig.fcc(f, golden[tag + "_occ_generate_order"], reassign=bool(reassign))
        assert np.array_equal(cfg.occupancy(), golden[tag + "_occ"])
        for name, kw in (("chain", dict(temperature=500.0, seed=21)),
                         ("chain_tt", dict(temperature=500.0, seed=22, tt_file=tt, rate_corrector=True))):
            tr = R.kmc_chain_ompi(cfg, js, ELEMENTS, maximum_steps=120, **kw)
            for k in KEYS:
                out["%s_%s_%s" % (tag, name, k)] = tr[k]
    path = os.path.join(ROOT, "tests", "golden", "golden_chain_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f kB" % (os.path.getsize(path) / 1e3), len(out), "arrays")


if __name__ == "__main__":
    main()

"""Golden trace of the reference's mc::CanonicalMcOmp (mc/src/CanonicalMcOmp.cpp:40-92): batches of `threads` mutually
non-interfering trials built by its greedy serial pass (unavailable_position_ rule), dE evaluated on the batch-start
configuration, events accepted in order.  Produced by oracle/_ref (ref_cmc_omp_traced) with the coefficient file stored in
golden_v1.npz.  Output: tests/golden/golden_omp_v1.npz.       python tests/golden/make_golden_omp.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from latticemontecarlo_b200 import synth  # noqa: E402
from oracle import ref_lib as R  # noqa: E402
from tests import helpers as H  # noqa: E402


def main():
    if not R.build():
        raise SystemExit("oracle/_ref not available")
    golden = np.load(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), allow_pickle=False)
    out = {}
    with tempfile.TemporaryDirectory() as d:
        js = H.golden_json(golden, d)
        for tag, f, reassign, threads, temperature, seed in (("P", 6, False, 6, 800.0, 5), ("Q", 5, True, 3, 1500.0, 8)):
            occ = synth.random_alloy(f, 0.10, 0.10, seed=400 + f, vacancy_site=None)
            cfg = R.RefConfig.fcc(f, occ, reassign=reassign)
            tr = R.cmc_omp_traced(cfg, js, temperature=temperature, maximum_steps=600, seed=seed, threads=threads)
            out[tag + "_params"] = np.array([f, int(reassign), threads, temperature, seed], dtype=np.float64)
            out[tag + "_occ"] = cfg.occupancy()
            for k in ("a", "b", "u", "dE", "energy_before", "batch", "final_occ"):
                out["%s_%s" % (tag, k)] = tr[k]
            out[tag + "_final_energy"] = np.array([tr["final_energy"]])
            print(tag, "steps", tr["steps"], "batches", int(tr["batch"][-1]) + 1, "accepted", int(np.count_nonzero(np.diff(tr["energy_before"]))))
    path = os.path.join(ROOT, "tests", "golden", "golden_omp_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.1f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    main()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref/liblmc_ref.so (the compiled reference)")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref). Built on demand where /root/reference exists, else skipped."""
    from oracle import ref_lib
    if not ref_lib.build():
        pytest.skip("oracle/_ref/liblmc_ref.so unavailable (no /root/reference here)")
    return ref_lib


@pytest.fixture(scope="session")
def coef_json(tmp_path_factory):
    from latticemontecarlo_b200 import synth
    path = tmp_path_factory.mktemp("coef") / "quartic_coefficients.json"
    synth.write_synthetic_json(path)
    return str(path)


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
    return np.load(path, allow_pickle=False)


@pytest.fixture(scope="session")
def golden_chain():
    """Traces of the reference's second-order KMC (tests/golden/make_golden_chain.py)."""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "golden_chain_v1.npz"), allow_pickle=False)

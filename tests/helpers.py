"""Shared helpers for the test-suite (oracle side). Only tests/ may import oracle/."""
import json
import os

import numpy as np

from oracle import lmc_oracle as O

ELEMENTS = ("Al", "Mg", "Zn")
CODES = [1, 2, 3]


def golden_coefficients(golden):
    co = {}
    for key in golden.files:
        if key.startswith("coef__"):
            _, top, name = key.split("__")
            val = golden[key]
            co.setdefault(top, {})[name] = float(val) if val.ndim == 0 else val.tolist()
    return co


def golden_json(golden, tmpdir):
    path = os.path.join(str(tmpdir), "golden_coefficients.json")
    with open(path, "w") as f:
        json.dump(golden_coefficients(golden), f)
    return path


def oracle_config(golden, tag, occ=None):
    """Oracle config of golden case `tag` ('A': f=4 reassigned order, 'B': f=5 generate order)."""
    f, reassign = (int(v) for v in golden[tag + "_factor"])
    cfg = O.Config.generate_fcc(f)
    if reassign:
        cfg.reassign_lattice_vector()
    cfg.occ = np.array(golden[tag + "_occ"] if occ is None else occ, dtype=np.uint8)
    return cfg


def unflatten_mapping(flat):
    pos, groups = 1, []
    for _ in range(int(flat[0])):
        c, l = int(flat[pos]), int(flat[pos + 1])
        pos += 2
        groups.append([tuple(int(v) for v in row) for row in np.asarray(flat[pos:pos + c * l]).reshape(c, l)])
        pos += c * l
    return groups

"""Synthetic inputs for tests and bench.py: the reference ships no .cfg, no JSON
coefficients and no time-temperature table, so everything is generated here, deterministically.

Element codes are the reference's ElementName enum values (lmc/cfg/include/Element.hpp:7).
"""
from __future__ import annotations

import json
import math

import numpy as np

X, AL, MG, ZN, CU, SN = 0, 1, 2, 3, 4, 5
ELEMENT_CODES = {"X": X, "Al": AL, "Mg": MG, "Zn": ZN, "Cu": CU, "Sn": SN}
ELEMENT_NAMES = {v: k for k, v in ELEMENT_CODES.items()}
ELEMENT_MASS = {"X": 0.00, "Al": 26.98, "Mg": 24.31, "Zn": 65.38, "Cu": 63.55, "Sn": 118.71}
LATTICE_CONSTANT = 4.046  # lmc/cfg/include/Constants.hpp:6

ORDER_GENERATE = 0    # cfg::GenerateFCC order  (cfg/src/Config.cpp:1073-1090): id = ((k*fy + j)*fx + i)*4 + b
ORDER_REASSIGNED = 1  # Config::ReassignLatticeVector order (cfg/src/Config.cpp:466-552): sorted by x, then y, then z


def encode_lengths(n_elements: int):
    """(n_cluster_types, len_mmm, len_mm2) for an element set of n_elements species (vacancy excluded)."""
    n = n_elements
    n_types = (n + 1) + 3 * (n * (n + 1) // 2 + n) + 4 * (math.comb(n + 2, 3) + n * (n + 1) // 2)
    len_mmm = 11 * n + 66 * n * n + 14 * (n * (n + 1) // 2)
    len_mm2 = 20 * n + 137 * n * n + 18 * (n * (n + 1) // 2)
    return n_types, len_mmm, len_mm2


def fcc_half_coords(factors, order=ORDER_GENERATE):
    """Integer half-lattice-constant coordinates (N x 3, int32) of every lattice id for the given id order."""
    if np.isscalar(factors):
        factors = (int(factors),) * 3
    fx, fy, fz = (int(v) for v in factors)
    k, j, i, b = np.meshgrid(np.arange(fz), np.arange(fy), np.arange(fx), np.arange(4), indexing="ij")
    basis = np.array([[0, 0, 0], [1, 1, 0], [1, 0, 1], [0, 1, 1]], dtype=np.int32)
    xyz = np.stack([2 * i, 2 * j, 2 * k], axis=-1).astype(np.int32) + basis[b]
    xyz = xyz.reshape(-1, 3)
    if order == ORDER_GENERATE:
        return xyz
    key = np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))
    return xyz[key]


def generate_to_reassigned_permutation(factors):
    """perm[new_id] = generate-order id of the site that ReassignLatticeVector puts at new_id."""
    xyz = fcc_half_coords(factors, ORDER_GENERATE)
    return np.lexsort((xyz[:, 2], xyz[:, 1], xyz[:, 0]))


def random_alloy(factors, p_mg=0.02, p_zn=0.02, seed=42, vacancy_site="default"):
    """i.i.d. random Al-Mg-Zn occupancy by lattice id (uint8 enum codes), one vacancy at N/2+3 by default."""
    if np.isscalar(factors):
        factors = (int(factors),) * 3
    n = 4 * int(factors[0]) * int(factors[1]) * int(factors[2])
    u = np.random.default_rng(seed).random(n)
    occ = np.full(n, AL, dtype=np.uint8)
    occ[u < p_mg + p_zn] = ZN
    occ[u < p_mg] = MG
    if vacancy_site == "default":
        vacancy_site = n // 2 + 3
    if vacancy_site is not None:
        occ[int(vacancy_site)] = X
    return occ


def synthetic_coefficients(seed=20240611, elements=("Al", "Mg", "Zn"), k_mmm=24, k_mm2=32):
    """The SURVEY.md §8(d) recipe; key names as parsed at pred/src/VacancyMigrationPredictorQuartic.cpp:44-62."""
    rng = np.random.default_rng(seed)
    n_types, len_mmm, len_mm2 = encode_lengths(len(elements))
    out = {"Base": {"theta": rng.normal(0.0, 50.0, n_types).tolist()}}
    for e in elements:
        out[e] = {
            "mu_x_mmm": rng.uniform(0.0, 0.5, len_mmm).tolist(),
            "sigma_x_mmm": rng.uniform(0.5, 1.5, len_mmm).tolist(),
            "mu_x_mm2": rng.uniform(0.0, 1.0, len_mm2).tolist(),
            "sigma_x_mm2": rng.uniform(0.5, 1.5, len_mm2).tolist(),
            "U_mmm": rng.normal(0.0, 1.0 / math.sqrt(len_mmm), (k_mmm, len_mmm)).tolist(),
            "U_mm2": rng.normal(0.0, 1.0 / math.sqrt(len_mm2), (k_mm2, len_mm2)).tolist(),
            "theta_D": rng.normal(0.0, 0.05, k_mmm).tolist(),
            "theta_Ks": rng.normal(0.0, 0.2, k_mm2).tolist(),
            "mu_D": math.log(2.86),
            "sigma_D": 0.02,
            "mu_Ks": math.log(4.7),
            "sigma_Ks": 0.15,
        }
    return out


def synthetic_coefficients_e0(seed=20240612, elements=("Al", "Mg", "Zn"), k_mmm=24, mu_e0=math.log(0.55), sigma_e0=0.2):
    """Synthetic file for the E0 model; key names as parsed at pred/src/VacancyMigrationPredictorE0.cpp:24-37."""
    rng = np.random.default_rng(seed)
    n_types, len_mmm, _ = encode_lengths(len(elements))
    out = {"Base": {"theta": rng.normal(0.0, 50.0, n_types).tolist()}}
    for e in elements:
        out[e] = {
            "mu_x_mmm": rng.uniform(0.0, 0.5, len_mmm).tolist(),
            "sigma_x_mmm": rng.uniform(0.5, 1.5, len_mmm).tolist(),
            "U_mmm": rng.normal(0.0, 1.0 / math.sqrt(len_mmm), (k_mmm, len_mmm)).tolist(),
            "theta_e0": rng.normal(0.0, 0.3, k_mmm).tolist(),
            "mu_e0": float(mu_e0),
            "sigma_e0": float(sigma_e0),
        }
    return out


def write_synthetic_json(path, model="quartic", **kw):
    coeffs = synthetic_coefficients_e0(**kw) if model == "e0" else synthetic_coefficients(**kw)
    with open(path, "w") as f:
        json.dump(coeffs, f)
    return coeffs


def write_time_temperature(path, points=((0.0, 300.0), (1e-3, 500.0), (1e-1, 700.0))):
    """First data line must start with the character '0' (pred/src/TimeTemperatureInterpolator.cpp:19)."""
    with open(path, "w") as f:
        for t, temp in points:
            f.write(("0" if t == 0 else repr(float(t))) + " " + repr(float(temp)) + "\n")

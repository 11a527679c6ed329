"""ctypes binding of oracle/_ref/liblmc_ref.so -- TEST INFRASTRUCTURE ONLY.

The library is the UNMODIFIED reference (zhucongx/LatticeMonteCarlo, lmc/{cfg,pred,mc}) compiled by
oracle/Makefile behind the C-ABI harness oracle/ref_harness.cpp.  It is the ground truth that pins the
numpy oracle (oracle/lmc_oracle.py) and generates tests/golden/*; bench.py uses it for the
`cpu_baseline` / `--impl reference` legs.  Nothing under latticemontecarlo_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "liblmc_ref.so")
REFERENCE_ROOT = "/root/reference"

# ElementName enum of the reference (lmc/cfg/include/Element.hpp:7)
ELEMENT_CODES = {"X": 0, "Al": 1, "Mg": 2, "Zn": 3, "Cu": 4, "Sn": 5}


def build(force: bool = False) -> bool:
    """Compile oracle/_ref/liblmc_ref.so if the reference sources are present. Returns availability."""
    have_ref = os.path.isdir(os.path.join(REFERENCE_ROOT, "lmc"))
    stale = have_ref and os.path.exists(_SO) and os.path.getmtime(_SO) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("ref_harness.cpp", "Makefile"))      # the harness grew an entry point
    if os.path.exists(_SO) and not force and not stale:
        return True
    if not have_ref:
        return os.path.exists(_SO)
    subprocess.run(["make", "-C", _HERE, "-j8", "all"], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(_SO)


def available() -> bool:
    return os.path.exists(_SO)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise RuntimeError("oracle/_ref/liblmc_ref.so missing: run `make -C oracle` where /root/reference exists")
        _lib = C.CDLL(_SO)
        _lib.ref_last_error.restype = C.c_char_p
        for name in ("ref_config_create", "ref_config_read", "ref_config_read_map", "ref_config_clone", "ref_quartic_create",
                     "ref_pairsite_create", "ref_energy_create", "ref_e0_create", "ref_pair_create", "ref_site_create"):
            getattr(_lib, name).restype = C.c_void_p
        for name in ("ref_config_num_sites", "ref_mapping", "ref_config_vacancy"):
            getattr(_lib, name).restype = C.c_int64
        for name in ("ref_kmc_first_omp", "ref_cmc_serial", "ref_cmc_omp", "ref_cmc_omp_traced", "ref_sa", "ref_rate_correction",
                     "ref_kmc_first_omp_with_logs", "ref_cmc_serial_with_logs", "ref_kmc_chain_ompi", "ref_kmc_chain_ompi_with_logs"):
            getattr(_lib, name).restype = C.c_double
    return _lib


def _err():
    return lib().ref_last_error().decode()


def _p(a, ctype=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _codes(elements):
    arr = np.array([ELEMENT_CODES[e] if isinstance(e, str) else int(e) for e in elements], dtype=np.int32)
    return arr, _p(arr), C.c_int(len(arr))


class RefConfig:
    """cfg::Config of the reference."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("reference config creation failed: " + _err())
        self.h = C.c_void_p(handle)

    @classmethod
    def fcc(cls, factors, occ=None, reassign=False):
        """GenerateFCC(factors, Al) + per-lattice-id elements `occ` (GenerateFCC order). With reassign=True the
        config goes through WriteConfig/ReadConfig/ReassignLatticeVector like a run started from a .cfg file."""
        if np.isscalar(factors):
            factors = (factors,) * 3
        occ_arr = None if occ is None else np.ascontiguousarray(occ, dtype=np.uint8)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "start.cfg").encode()
            h = lib().ref_config_create(int(factors[0]), int(factors[1]), int(factors[2]), _p(occ_arr),
                                        int(bool(reassign)), path)
        return cls(h)

    @classmethod
    def read(cls, path, reassign=True):
        return cls(lib().ref_config_read(str(path).encode(), int(bool(reassign))))

    @classmethod
    def read_map(cls, lattice_path, element_path, map_path):
        """Config::ReadMap: lattice ids as written in lattice.txt (no reassignment)."""
        return cls(lib().ref_config_read_map(str(lattice_path).encode(), str(element_path).encode(), str(map_path).encode()))

    def write_map_files(self, lattice_path, element_path, map_path):
        if lib().ref_config_write_map_files(self.h, str(lattice_path).encode(), str(element_path).encode(), str(map_path).encode()) != 0:
            raise RuntimeError(_err())

    def clone(self):
        return RefConfig(lib().ref_config_clone(self.h))

    def write(self, path):
        if lib().ref_config_write(self.h, str(path).encode()) != 0:
            raise RuntimeError(_err())

    def __del__(self):
        try:
            if self.h:
                lib().ref_config_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def num_sites(self):
        return int(lib().ref_config_num_sites(self.h))

    def occupancy(self):
        out = np.empty(self.num_sites, dtype=np.uint8)
        lib().ref_config_get_occupancy(self.h, _p(out))
        return out

    def basis(self):
        out = np.empty(9, dtype=np.float64)
        lib().ref_config_get_basis(self.h, _p(out))
        return out.reshape(3, 3)

    def positions(self):
        out = np.empty((self.num_sites, 3), dtype=np.float64)
        lib().ref_config_get_positions(self.h, _p(out))
        return out

    def neighbors(self, shell):
        k = {1: 12, 2: 6, 3: 24}[shell]
        out = np.empty((self.num_sites, k), dtype=np.int64)
        lib().ref_config_get_neighbors(self.h, int(shell), _p(out))
        return out

    def maps(self):
        n = self.num_sites
        l2a = np.empty(n, dtype=np.int64)
        a2l = np.empty(n, dtype=np.int64)
        shift = np.empty((n, 3), dtype=np.int32)
        lib().ref_config_get_maps(self.h, _p(l2a), _p(a2l), _p(shift))
        return l2a, a2l, shift

    def lattice_jump(self, a, b):
        lib().ref_config_lattice_jump(self.h, C.c_int64(a), C.c_int64(b))

    def set_element(self, lattice_id, code):
        lib().ref_config_set_element(self.h, C.c_int64(lattice_id), int(code))

    def vacancy(self):
        return int(lib().ref_config_vacancy(self.h))

    # ---- pred:: free functions
    def mapping(self, which):
        """which: 'state_pair' | 'mmm' | 'mm2' | 'state_site' -> list of groups, each a list of tuples (-1 = SIZE_MAX)."""
        w = {"state_pair": 0, "mmm": 1, "mm2": 2, "state_site": 3}[which]
        n = lib().ref_mapping(self.h, w, None, C.c_int64(0))
        if n < 0:
            raise RuntimeError(_err())
        flat = np.empty(n, dtype=np.int64)
        lib().ref_mapping(self.h, w, _p(flat), C.c_int64(n))
        return unflatten_mapping(flat)

    def pair_lists(self, i, j):
        s = np.empty(60, dtype=np.int64)
        m = np.empty(58, dtype=np.int64)
        m2 = np.empty(58, dtype=np.int64)
        if lib().ref_pair_lists(self.h, C.c_int64(i), C.c_int64(j), _p(s), _p(m), _p(m2)) != 0:
            raise RuntimeError(_err())
        return s, m, m2

    def site_list(self, i):
        s = np.empty(43, dtype=np.int64)
        if lib().ref_site_list(self.h, C.c_int64(i), _p(s)) != 0:
            raise RuntimeError(_err())
        return s


def unflatten_mapping(flat):
    pos = 0
    groups = []
    g = int(flat[pos]); pos += 1
    for _ in range(g):
        c = int(flat[pos]); l = int(flat[pos + 1]); pos += 2
        arr = flat[pos:pos + c * l].reshape(c, l); pos += c * l
        groups.append([tuple(int(v) for v in row) for row in arr])
    return groups


def cluster_types(elements):
    """Cluster types in ClusterIndexer order: list of (label, (codes...))."""
    _, p, n = _codes(elements)
    cnt = lib().ref_cluster_types(p, n, None, 0)
    out = np.empty((cnt, 5), dtype=np.int32)
    lib().ref_cluster_types(p, n, _p(out), cnt)
    return [(int(r[0]), tuple(int(v) for v in r[2:2 + r[1]])) for r in out]


class RefQuartic:
    """pred::VacancyMigrationPredictorQuartic (lru_size=0) or ...QuarticLru (lru_size>0)."""

    def __init__(self, json_path, config: RefConfig, elements=("Al", "Mg", "Zn"), lru_size=0):
        self._codes = _codes(elements)
        self.h = C.c_void_p(lib().ref_quartic_create(str(json_path).encode(), config.h, self._codes[1], self._codes[2],
                                                     C.c_int64(lru_size)))
        if not self.h:
            raise RuntimeError("reference predictor creation failed: " + _err())

    def __del__(self):
        try:
            if self.h:
                lib().ref_quartic_free(self.h)
                self.h = None
        except Exception:
            pass

    def eval(self, config: RefConfig, i, j, threads=1):
        i = np.ascontiguousarray(i, dtype=np.int64)
        j = np.ascontiguousarray(j, dtype=np.int64)
        ea = np.empty(len(i), dtype=np.float64)
        de = np.empty(len(i), dtype=np.float64)
        if lib().ref_quartic_eval(self.h, config.h, C.c_int64(len(i)), _p(i), _p(j), _p(ea), _p(de), int(threads)) != 0:
            raise RuntimeError(_err())
        return ea, de

    def parts(self, config: RefConfig, i, j, n_types=95, len_mmm=711, len_mm2=1401):
        de = C.c_double(); d = C.c_double(); ks = C.c_double(); nt = C.c_int64()
        sc = np.zeros(512, dtype=np.int32); ec = np.zeros(512, dtype=np.int32)
        em = np.zeros(4096, dtype=np.float64); ef = np.zeros(4096, dtype=np.float64); eb = np.zeros(4096, dtype=np.float64)
        s = np.empty(60, dtype=np.int64); m = np.empty(58, dtype=np.int64); m2 = np.empty(58, dtype=np.int64)
        rc = lib().ref_quartic_parts(self.h, config.h, C.c_int64(i), C.c_int64(j), C.byref(de), C.byref(d), C.byref(ks),
                                     _p(sc), _p(ec), C.byref(nt), _p(em), _p(ef), _p(eb), _p(s), _p(m), _p(m2))
        if rc != 0:
            raise RuntimeError(_err())
        return dict(dE=de.value, D=d.value, Ks=ks.value, start_counts=sc[:nt.value].copy(), end_counts=ec[:nt.value].copy(),
                    enc_mmm=em[:len_mmm].copy(), enc_mm2_f=ef[:len_mm2].copy(), enc_mm2_b=eb[:len_mm2].copy(),
                    state=s, mmm=m, mm2=m2)


class RefPairSite:
    """pred::EnergyChangePredictorPairSite."""

    def __init__(self, json_path, config: RefConfig, elements=("Al", "Mg", "Zn")):
        self._codes = _codes(elements)
        self.h = C.c_void_p(lib().ref_pairsite_create(str(json_path).encode(), config.h, self._codes[1], self._codes[2]))
        if not self.h:
            raise RuntimeError("reference predictor creation failed: " + _err())

    def __del__(self):
        try:
            if self.h:
                lib().ref_pairsite_free(self.h)
                self.h = None
        except Exception:
            pass

    def de_pair(self, config: RefConfig, a, b, threads=1):
        a = np.ascontiguousarray(a, dtype=np.int64)
        b = np.ascontiguousarray(b, dtype=np.int64)
        out = np.empty(len(a), dtype=np.float64)
        if lib().ref_pairsite_de_pair(self.h, config.h, C.c_int64(len(a)), _p(a), _p(b), _p(out), int(threads)) != 0:
            raise RuntimeError(_err())
        return out

    def de_site(self, config: RefConfig, site, new_code, threads=1):
        site = np.ascontiguousarray(site, dtype=np.int64)
        new_code = np.ascontiguousarray(new_code, dtype=np.uint8)
        out = np.empty(len(site), dtype=np.float64)
        if lib().ref_pairsite_de_site(self.h, config.h, C.c_int64(len(site)), _p(site), _p(new_code), _p(out), int(threads)) != 0:
            raise RuntimeError(_err())
        return out

    def site_counts(self, config: RefConfig, site, new_code, n_types=95):
        de = C.c_double()
        sc = np.zeros(512, dtype=np.int32); ec = np.zeros(512, dtype=np.int32)
        if lib().ref_pairsite_site_counts(self.h, config.h, C.c_int64(site), int(new_code), C.byref(de), _p(sc), _p(ec)) != 0:
            raise RuntimeError(_err())
        return de.value, sc[:n_types].copy(), ec[:n_types].copy()



class RefE0:
    """pred::VacancyMigrationPredictorE0 (lru_size=0) or ...E0Lru (lru_size>0)."""

    def __init__(self, json_path, config: RefConfig, elements=("Al", "Mg", "Zn"), lru_size=0):
        self._codes = _codes(elements)
        self.lru = lru_size > 0
        self.h = C.c_void_p(lib().ref_e0_create(str(json_path).encode(), config.h, self._codes[1], self._codes[2], C.c_int64(lru_size)))
        if not self.h:
            raise RuntimeError("reference predictor creation failed: " + _err())

    def __del__(self):
        try:
            if self.h:
                lib().ref_e0_free(self.h)
                self.h = None
        except Exception:
            pass

    def eval(self, config: RefConfig, i, j):
        """(Ea, dE, e0); e0 is None for the LRU predictor (GetE0 is protected there)."""
        i = np.ascontiguousarray(i, dtype=np.int64)
        j = np.ascontiguousarray(j, dtype=np.int64)
        ea, de = np.empty(len(i)), np.empty(len(i))
        e0 = None if self.lru else np.empty(len(i))
        if lib().ref_e0_eval(self.h, config.h, C.c_int64(len(i)), _p(i), _p(j), _p(ea), _p(de), _p(e0)) != 0:
            raise RuntimeError(_err())
        return ea, de, e0


class RefPair:
    """pred::EnergyChangePredictorPair (first-neighbour pairs only: anything else throws std::out_of_range)."""

    def __init__(self, json_path, config: RefConfig, elements=("Al", "Mg", "Zn")):
        self._codes = _codes(elements)
        self.h = C.c_void_p(lib().ref_pair_create(str(json_path).encode(), config.h, self._codes[1], self._codes[2]))
        if not self.h:
            raise RuntimeError("reference predictor creation failed: " + _err())

    def __del__(self):
        try:
            if self.h:
                lib().ref_pair_free(self.h)
                self.h = None
        except Exception:
            pass

    def de_pair(self, config: RefConfig, a, b):
        a = np.ascontiguousarray(a, dtype=np.int64)
        b = np.ascontiguousarray(b, dtype=np.int64)
        out = np.empty(len(a), dtype=np.float64)
        if lib().ref_pair_de(self.h, config.h, C.c_int64(len(a)), _p(a), _p(b), _p(out)) != 0:
            raise RuntimeError(_err())
        return out


class RefSite:
    """pred::EnergyChangePredictorSite."""

    def __init__(self, json_path, config: RefConfig, elements=("Al", "Mg", "Zn")):
        self._codes = _codes(elements)
        self.h = C.c_void_p(lib().ref_site_create(str(json_path).encode(), config.h, self._codes[1], self._codes[2]))
        if not self.h:
            raise RuntimeError("reference predictor creation failed: " + _err())

    def __del__(self):
        try:
            if self.h:
                lib().ref_site_free(self.h)
                self.h = None
        except Exception:
            pass

    def de_site(self, config: RefConfig, site, new_code):
        site = np.ascontiguousarray(site, dtype=np.int64)
        new_code = np.ascontiguousarray(new_code, dtype=np.uint8)
        out = np.empty(len(site), dtype=np.float64)
        if lib().ref_site_de(self.h, config.h, C.c_int64(len(site)), _p(site), _p(new_code), _p(out)) != 0:
            raise RuntimeError(_err())
        return out

class RefEnergy:
    """pred::EnergyPredictor."""

    def __init__(self, json_path, elements=("Al", "Mg", "Zn")):
        self._codes = _codes(elements)
        self.h = C.c_void_p(lib().ref_energy_create(str(json_path).encode(), self._codes[1], self._codes[2]))
        if not self.h:
            raise RuntimeError("reference predictor creation failed: " + _err())

    def __del__(self):
        try:
            if self.h:
                lib().ref_energy_free(self.h)
                self.h = None
        except Exception:
            pass

    def energy(self, config: RefConfig, n_types=95):
        e = C.c_double()
        enc = np.zeros(512, dtype=np.float64)
        if lib().ref_energy_get(self.h, config.h, C.byref(e), _p(enc), 512) != 0:
            raise RuntimeError(_err())
        return e.value, enc[:n_types].copy()


    def energy_of_cluster(self, config: RefConfig, atom_ids, n_types=95):
        ids = np.ascontiguousarray(atom_ids, dtype=np.int64)
        e = C.c_double()
        enc = np.zeros(512, dtype=np.float64)
        if lib().ref_energy_of_cluster(self.h, config.h, _p(ids), C.c_int64(len(ids)), C.byref(e), _p(enc), 512) != 0:
            raise RuntimeError(_err())
        return e.value, enc[:n_types].copy()

    def chemical_potential(self, solvent="Al"):
        codes = np.zeros(16, dtype=np.int32); mu = np.zeros(16, dtype=np.float64)
        k = lib().ref_chemical_potential(self.h, ELEMENT_CODES[solvent], _p(codes), _p(mu), 16)
        if k < 0:
            raise RuntimeError(_err())
        return {int(c): float(v) for c, v in zip(codes[:k], mu[:k])}


def rate_correction(c_vac, c_solute, temperature):
    return float(lib().ref_rate_correction(C.c_double(c_vac), C.c_double(c_solute), C.c_double(temperature)))


def tt_interpolate(path, times):
    times = np.ascontiguousarray(times, dtype=np.float64)
    out = np.empty_like(times)
    if lib().ref_tt_interpolate(str(path).encode(), C.c_int64(len(times)), _p(times), _p(out)) != 0:
        raise RuntimeError(_err())
    return out


def kmc_first_omp(config: RefConfig, json_path, elements=("Al", "Mg", "Zn"), temperature=500.0, maximum_steps=100,
                  seed=1, threads=1, tt_file=None, rate_corrector=False, trace=True):
    """mc::KineticMcFirstOmp::Simulate() with a seeded generator. Returns dict (trace arrays have maximum_steps+1 rows)."""
    _, p, n = _codes(elements)
    cap = int(maximum_steps) + 1 if trace else 0
    f = lambda: np.zeros(cap, dtype=np.float64)
    g = lambda: np.zeros(cap, dtype=np.int64)
    t = dict(u1=f(), u2=f(), dt=f(), time=f(), energy=f(), Ea=f(), dE=f(), temperature=f(), total_rate=f())
    t.update({"from": g(), "to": g(), "slot": g()})
    occ = np.empty(config.num_sites, dtype=np.uint8)
    summary = np.zeros(4, dtype=np.float64)
    with tempfile.TemporaryDirectory() as d:
        sec = lib().ref_kmc_first_omp(config.h, str(json_path).encode(), p, n,
                                      str(tt_file).encode() if tt_file else None, int(bool(rate_corrector)),
                                      C.c_double(temperature), C.c_uint64(int(maximum_steps)), C.c_uint64(int(seed)),
                                      int(threads), d.encode(), C.c_int64(cap), _p(t["u1"]), _p(t["u2"]), _p(t["from"]),
                                      _p(t["to"]), _p(t["slot"]), _p(t["dt"]), _p(t["time"]), _p(t["energy"]),
                                      _p(t["Ea"]), _p(t["dE"]), _p(t["temperature"]), _p(t["total_rate"]), _p(occ),
                                      _p(summary))
    if sec < 0:
        raise RuntimeError(_err())
    t.update(seconds=sec, final_occ=occ, final_time=summary[0], final_energy=summary[1], absolute_energy=summary[2],
             steps=int(summary[3]))
    return t


def kmc_chain_ompi(config: RefConfig, json_path, elements=("Al", "Mg", "Zn"), temperature=500.0, maximum_steps=100,
                   seed=1, tt_file=None, rate_corrector=False, trace=True):
    """mc::KineticMcChainOmpi::Simulate() (second-order KMC) with its 12 MPI ranks run as 12 threads of this process and a
    seeded rank-0 generator.  Trace as kmc_first_omp; u2 is the one uniform a step consumes (u1 is unused)."""
    _, p, n = _codes(elements)
    cap = int(maximum_steps) + 1 if trace else 0
    f = lambda: np.zeros(cap, dtype=np.float64)
    g = lambda: np.zeros(cap, dtype=np.int64)
    t = dict(u1=f(), u2=f(), dt=f(), time=f(), energy=f(), Ea=f(), dE=f(), temperature=f(), total_rate=f())
    t.update({"from": g(), "to": g(), "slot": g()})
    occ = np.empty(config.num_sites, dtype=np.uint8)
    summary = np.zeros(4, dtype=np.float64)
    with tempfile.TemporaryDirectory() as d:
        sec = lib().ref_kmc_chain_ompi(config.h, str(json_path).encode(), p, n,
                                       str(tt_file).encode() if tt_file else None, int(bool(rate_corrector)),
                                       C.c_double(temperature), C.c_uint64(int(maximum_steps)), C.c_uint64(int(seed)),
                                       d.encode(), C.c_int64(cap), _p(t["u1"]), _p(t["u2"]), _p(t["from"]),
                                       _p(t["to"]), _p(t["slot"]), _p(t["dt"]), _p(t["time"]), _p(t["energy"]),
                                       _p(t["Ea"]), _p(t["dE"]), _p(t["temperature"]), _p(t["total_rate"]), _p(occ),
                                       _p(summary))
    if sec < 0:
        raise RuntimeError(_err())
    t.update(seconds=sec, final_occ=occ, final_time=summary[0], final_energy=summary[1], absolute_energy=summary[2],
             steps=int(summary[3]))
    return t


def cmc_serial(config: RefConfig, json_path, elements=("Al", "Mg", "Zn"), temperature=800.0, maximum_steps=100, seed=1,
               trace=True):
    _, p, n = _codes(elements)
    cap = int(maximum_steps) + 1 if trace else 0
    a = np.zeros(cap, dtype=np.int64); b = np.zeros(cap, dtype=np.int64)
    de = np.zeros(cap, dtype=np.float64); eb = np.zeros(cap, dtype=np.float64); u = np.zeros(cap, dtype=np.float64)
    occ = np.empty(config.num_sites, dtype=np.uint8)
    fe = C.c_double()
    with tempfile.TemporaryDirectory() as d:
        sec = lib().ref_cmc_serial(config.h, str(json_path).encode(), p, n, C.c_double(temperature),
                                   C.c_uint64(int(maximum_steps)), C.c_uint64(int(seed)), d.encode(), C.c_int64(cap),
                                   _p(a), _p(b), _p(de), _p(eb), _p(u), _p(occ), C.byref(fe))
    if sec < 0:
        raise RuntimeError(_err())
    return dict(seconds=sec, a=a, b=b, dE=de, energy_before=eb, u=u, final_occ=occ, final_energy=fe.value)


def cmc_omp(config: RefConfig, json_path, elements=("Al", "Mg", "Zn"), temperature=800.0, maximum_steps=100, seed=1,
            threads=0):
    _, p, n = _codes(elements)
    occ = np.empty(config.num_sites, dtype=np.uint8)
    fe = C.c_double(); steps = C.c_uint64()
    with tempfile.TemporaryDirectory() as d:
        sec = lib().ref_cmc_omp(config.h, str(json_path).encode(), p, n, C.c_double(temperature),
                                C.c_uint64(int(maximum_steps)), C.c_uint64(int(seed)), int(threads), d.encode(),
                                _p(occ), C.byref(fe), C.byref(steps))
    if sec < 0:
        raise RuntimeError(_err())
    return dict(seconds=sec, final_occ=occ, final_energy=fe.value, steps=int(steps.value))


def cmc_omp_traced(config: RefConfig, json_path, elements=("Al", "Mg", "Zn"), temperature=800.0, maximum_steps=100, seed=1, threads=4):
    """mc::CanonicalMcOmp with `threads` OMP threads (= batch size) and one trace record per event."""
    _, p, n = _codes(elements)
    cap = int(maximum_steps) + 2 * int(threads) + 2            # the last batch may run past maximum_steps
    a = np.zeros(cap, dtype=np.int64); b = np.zeros(cap, dtype=np.int64); batch = np.zeros(cap, dtype=np.int64)
    de = np.zeros(cap, dtype=np.float64); eb = np.zeros(cap, dtype=np.float64); u = np.zeros(cap, dtype=np.float64)
    occ = np.empty(config.num_sites, dtype=np.uint8)
    fe = C.c_double(); steps = C.c_uint64()
    with tempfile.TemporaryDirectory() as d:
        sec = lib().ref_cmc_omp_traced(config.h, str(json_path).encode(), p, n, C.c_double(temperature), C.c_uint64(int(maximum_steps)),
                                       C.c_uint64(int(seed)), int(threads), d.encode(), C.c_int64(cap), _p(a), _p(b), _p(de), _p(eb), _p(u),
                                       _p(batch), _p(occ), C.byref(fe), C.byref(steps))
    if sec < 0:
        raise RuntimeError(_err())
    k = int(steps.value)
    return dict(seconds=sec, a=a[:k], b=b[:k], dE=de[:k], energy_before=eb[:k], u=u[:k], batch=batch[:k], final_occ=occ, final_energy=fe.value, steps=k)


def simulated_annealing(factor, solvent, solute_counts: dict, occ, json_path, initial_temperature=700.0,
                        maximum_steps=100, seed=1, trace=True):
    names = list(solute_counts)
    codes = np.array([ELEMENT_CODES[e] for e in names], dtype=np.int32)
    # with an explicit occupancy the constructor's own random placement is discarded, so ask it for one solute
    # atom per species (keeps the element set, avoids its O(N*n_solute) ">= 4NN apart" generator)
    counts = np.array([solute_counts[e] if occ is None else 1 for e in names], dtype=np.int64)
    cap = int(maximum_steps) + 1 if trace else 0
    a = np.zeros(cap, dtype=np.int64); b = np.zeros(cap, dtype=np.int64)
    eb = np.zeros(cap, dtype=np.float64); tb = np.zeros(cap, dtype=np.float64); u = np.zeros(cap, dtype=np.float64)
    n_sites = 4 * factor ** 3
    occ_in = None if occ is None else np.ascontiguousarray(occ, dtype=np.uint8)
    occ_out = np.empty(n_sites, dtype=np.uint8)
    e0 = C.c_double(); fe = C.c_double(); ft = C.c_double()
    with tempfile.TemporaryDirectory() as d:
        sec = lib().ref_sa(int(factor), ELEMENT_CODES[solvent], _p(codes), _p(counts), len(names), _p(occ_in),
                           str(json_path).encode(), C.c_double(initial_temperature), C.c_uint64(int(maximum_steps)),
                           C.c_uint64(int(seed)), d.encode(), C.c_int64(cap), _p(a), _p(b), _p(eb), _p(tb), _p(u), _p(occ_out),
                           C.byref(e0), C.byref(fe), C.byref(ft))
    if sec < 0:
        raise RuntimeError(_err())
    return dict(seconds=sec, a=a, b=b, energy_before=eb, temperature_before=tb, u=u, final_occ=occ_out, energy0=e0.value,
                final_energy=fe.value, final_temperature=ft.value)


def kmc_first_omp_with_logs(config: RefConfig, json_path, workdir, elements=("Al", "Mg", "Zn"), temperature=500.0,
                            maximum_steps=100, log_dump_steps=10, config_dump_steps=1000, seed=1, tt_file=None,
                            rate_corrector=False):
    """mc::KineticMcFirstOmp::Simulate() as shipped (logs + dumps written into workdir), seeded generator."""
    _, p, n = _codes(elements)
    sec = lib().ref_kmc_first_omp_with_logs(config.h, str(json_path).encode(), p, n, str(tt_file).encode() if tt_file else None,
                                            int(bool(rate_corrector)), C.c_double(temperature), C.c_uint64(int(log_dump_steps)),
                                            C.c_uint64(int(config_dump_steps)), C.c_uint64(int(maximum_steps)),
                                            C.c_uint64(int(seed)), str(workdir).encode())
    if sec < 0:
        raise RuntimeError(_err())
    return sec


def kmc_chain_ompi_with_logs(config: RefConfig, json_path, workdir, elements=("Al", "Mg", "Zn"), temperature=500.0,
                             maximum_steps=100, log_dump_steps=10, config_dump_steps=1000, seed=1, tt_file=None,
                             rate_corrector=False, solute_disp=False):
    """mc::KineticMcChainOmpi::Simulate() as shipped (rank 0 writes logs + dumps into workdir), 12 thread-ranks, seeded."""
    _, p, n = _codes(elements)
    sec = lib().ref_kmc_chain_ompi_with_logs(config.h, str(json_path).encode(), p, n, str(tt_file).encode() if tt_file else None,
                                             int(bool(rate_corrector)), C.c_double(temperature), C.c_uint64(int(log_dump_steps)),
                                             C.c_uint64(int(config_dump_steps)), C.c_uint64(int(maximum_steps)),
                                             C.c_uint64(int(seed)), int(bool(solute_disp)), str(workdir).encode())
    if sec < 0:
        raise RuntimeError(_err())
    return sec


def cmc_serial_with_logs(config: RefConfig, json_path, workdir, elements=("Al", "Mg", "Zn"), temperature=800.0,
                         maximum_steps=100, log_dump_steps=10, config_dump_steps=1000, thermodynamic_averaging_steps=0, seed=1):
    _, p, n = _codes(elements)
    sec = lib().ref_cmc_serial_with_logs(config.h, str(json_path).encode(), p, n, C.c_double(temperature),
                                         C.c_uint64(int(log_dump_steps)), C.c_uint64(int(config_dump_steps)),
                                         C.c_uint64(int(maximum_steps)), C.c_uint64(int(thermodynamic_averaging_steps)),
                                         C.c_uint64(int(seed)), str(workdir).encode())
    if sec < 0:
        raise RuntimeError(_err())
    return sec


def uniform_real_stream(seed, n):
    out = np.empty(int(n), dtype=np.float64)
    lib().ref_rng_uniform_real(C.c_uint64(int(seed)), C.c_int64(int(n)), _p(out))
    return out

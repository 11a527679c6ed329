"""ctypes binding of the C ABI in include/lmc_b200.h (liblmc_b200.so).

This is plumbing for tests / bench.py / the Python host mirror; the product is the shared library.
There is no fallback: if the library is missing or no CUDA device is usable, compute calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblmc_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "lmc_b200.h")

LMC_OK, LMC_ERR_INVALID_ARGUMENT, LMC_ERR_OUT_OF_RANGE, LMC_ERR_RUNTIME, LMC_ERR_NO_DEVICE, LMC_ERR_CUDA = 0, -1, -2, -3, -4, -5
ORDER_GENERATE, ORDER_REASSIGNED = 0, 1
BARRIER_QUARTIC, BARRIER_E0 = 0, 1


class LmcError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("[lmc %d] %s" % (code, message))
        self.code = code


class LmcInvalidArgument(LmcError, ValueError):
    pass


class LmcOutOfRange(LmcError, IndexError):
    """std::out_of_range of the reference (non-neighbour pair, cluster without index)."""


_lib = None


def declared_symbols():
    """Every function name declared in include/lmc_b200.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lmc_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("LMC_B200_LIB", LIB_PATH)      # A/B runs of two builds (tools/ab.sh); the default is the in-tree library
        if not os.path.exists(path):
            raise RuntimeError("liblmc_b200.so not built: run `python -m latticemontecarlo_b200.build` "
                               "(there is no CPU fallback)")
        _lib = C.CDLL(path)
        _lib.lmc_last_error.restype = C.c_char_p
        _lib.lmc_engine_num_sites.restype = C.c_int64
        _lib.lmc_tables_mapping.restype = C.c_int64
        _lib.lmc_engine_destroy.restype = None
        _lib.lmc_engine_cuda_stream.restype = C.c_void_p
        _lib.lmc_engine_last_kernel_ms.restype = C.c_double
        _lib.lmc_engine_launch_count.restype = C.c_int64
        _lib.lmc_engine_get_tables.restype = C.c_int64
        _lib.lmc_engine_find_element.restype = C.c_int64
        _lib.lmc_engine_coefficients_path.restype = C.c_char_p
    return _lib


def _check(rc):
    if rc >= 0:
        return rc
    msg = lib().lmc_last_error().decode()
    cls = {LMC_ERR_INVALID_ARGUMENT: LmcInvalidArgument, LMC_ERR_OUT_OF_RANGE: LmcOutOfRange}.get(rc, LmcError)
    raise cls(rc, msg)


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def device_count():
    return int(lib().lmc_device_count())


def tables_mapping(which):
    """which: 'state_pair' | 'mmm' | 'mm2' | 'state_site' -> list of groups of position tuples (-1 = SIZE_MAX)."""
    w = {"state_pair": 0, "mmm": 1, "mm2": 2, "state_site": 3}[which]
    n = _check(lib().lmc_tables_mapping(w, None, C.c_int64(0)))
    flat = np.empty(n, dtype=np.int64)
    _check(lib().lmc_tables_mapping(w, _p(flat), C.c_int64(n)))
    pos, groups = 1, []
    for _ in range(int(flat[0])):
        c, l = int(flat[pos]), int(flat[pos + 1])
        pos += 2
        groups.append([tuple(int(v) for v in row) for row in flat[pos:pos + c * l].reshape(c, l)])
        pos += c * l
    return groups


def tables_cluster_types(element_set):
    es = np.ascontiguousarray(element_set, dtype=np.int32)
    n = _check(lib().lmc_tables_cluster_types(_p(es), len(es), None, 0))
    rows = np.empty((n, 5), dtype=np.int32)
    _check(lib().lmc_tables_cluster_types(_p(es), len(es), _p(rows), n))
    return [(int(r[0]), tuple(int(v) for v in r[2:2 + r[1]])) for r in rows]


def tables_group_sizes(which, n_elements):
    w = {"mmm": 1, "mm2": 2}[which]
    n = _check(lib().lmc_tables_group_sizes(w, int(n_elements), None, 0))
    sizes = np.zeros(n, dtype=np.int32)
    _check(lib().lmc_tables_group_sizes(w, int(n_elements), _p(sizes), n))
    return sizes


def tables_env_pairs(which):
    w = {"pair": 0, "site": 1}[which]
    n = _check(lib().lmc_tables_env_pairs(w, None, 0))
    out = np.empty((n, 2), dtype=np.int16)
    _check(lib().lmc_tables_env_pairs(w, _p(out), n))
    return out


class KmcParams(C.Structure):
    _fields_ = [("temperature", C.c_double), ("temperatures", C.c_void_p), ("n_time_temperature", C.c_int32),
                ("tt_time", C.c_void_p), ("tt_temperature", C.c_void_p), ("rate_corrector", C.c_int32), ("seed", C.c_uint64)]


class KmcTrace(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("from_", "to", "slot", "dt", "Ea", "dE", "total_rate", "temperature")]


class CmcParams(C.Structure):
    _fields_ = [("temperature", C.c_double), ("temperatures", C.c_void_p), ("seed", C.c_uint64), ("batch_size", C.c_int32)]


class CmcDomainParams(C.Structure):
    _fields_ = [("domain_edge", C.c_int32), ("rounds_per_sweep", C.c_int32), ("speculate", C.c_int32), ("lanes", C.c_int32),
                ("passes", C.c_double)]


class Engine:
    """One lmc_engine (one GPU, or host-only with device=-1)."""

    def __init__(self, factors, id_order=ORDER_REASSIGNED, element_set=(1, 2, 3), solvent=1, n_walkers=1, device=0):
        if np.isscalar(factors):
            factors = (int(factors),) * 3
        self.factors = tuple(int(f) for f in factors)
        f = (C.c_int32 * 3)(*self.factors)
        es = np.ascontiguousarray(element_set, dtype=np.int32)
        self.element_set = tuple(int(e) for e in es)
        h = C.c_void_p()
        _check(lib().lmc_engine_create(C.byref(h), f, int(id_order), _p(es), len(es), int(solvent), int(n_walkers), int(device)))
        self.h = h
        self.num_sites = int(lib().lmc_engine_num_sites(self.h))
        self.n_walkers = int(n_walkers)
        self.device = int(device)
        self.n_types = len(tables_cluster_types(self.element_set))

    def close(self):
        if getattr(self, "h", None):
            lib().lmc_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state
    def load_coefficients(self, json_path, model=BARRIER_QUARTIC):
        """model: BARRIER_QUARTIC (VacancyMigrationPredictorQuartic) or BARRIER_E0 (VacancyMigrationPredictorE0)."""
        if model == BARRIER_QUARTIC:
            _check(lib().lmc_engine_load_coefficients(self.h, str(json_path).encode()))
        else:
            _check(lib().lmc_engine_load_coefficients_model(self.h, str(json_path).encode(), C.c_int32(int(model))))

    def set_occupancy(self, occ, walker=0):
        occ = np.ascontiguousarray(occ, dtype=np.uint8)
        _check(lib().lmc_engine_set_occupancy(self.h, int(walker), _p(occ), C.c_int64(occ.size)))

    def get_occupancy(self, walker=0):
        out = np.empty(self.num_sites, dtype=np.uint8)
        _check(lib().lmc_engine_get_occupancy(self.h, int(walker), _p(out), C.c_int64(out.size)))
        return out

    def set_occupancy_all(self, occ):
        occ = np.ascontiguousarray(occ, dtype=np.uint8)
        _check(lib().lmc_engine_set_occupancy_all(self.h, _p(occ), C.c_int64(occ.size)))

    def get_occupancy_all(self):
        out = np.empty((self.n_walkers, self.num_sites), dtype=np.uint8)
        _check(lib().lmc_engine_get_occupancy_all(self.h, _p(out), C.c_int64(out.size)))
        return out

    def lattice_jump(self, a, b, walker=0):
        _check(lib().lmc_engine_lattice_jump(self.h, int(walker), C.c_int64(int(a)), C.c_int64(int(b))))

    def kmc_folded_tables(self):
        """The (dE, log E0) tables the KMC kernels walk, on their binary grid, and the grid bits per component."""
        lib().lmc_engine_get_tables.restype = C.c_int64
        raw = []
        for which in (7, 8, 9):
            n = _check(lib().lmc_engine_get_tables(self.h, which, None, C.c_int64(0)))
            buf = np.empty(n, dtype=np.float64)
            _check(lib().lmc_engine_get_tables(self.h, which, _p(buf), C.c_int64(n)))
            raw.append(buf)
        bits = (C.c_int32 * 2)()
        _check(lib().lmc_engine_kmc_table_grid_bits(self.h, bits))
        n = len(self.element_set)
        return dict(C=raw[0].reshape(n, 2), A=raw[1].reshape(n, 58, n, 2), B=raw[2].reshape(n, -1, n, n, 2), bits=(int(bits[0]), int(bits[1])))

    def get_tables(self):
        """Host copies of the contracted coefficient tables, reshaped (see lmc_engine_get_tables)."""
        lib().lmc_engine_get_tables.restype = C.c_int64
        raw = []
        for which in range(6):
            n = _check(lib().lmc_engine_get_tables(self.h, which, None, C.c_int64(0)))
            buf = np.empty(n, dtype=np.float64)
            _check(lib().lmc_engine_get_tables(self.h, which, _p(buf), C.c_int64(n)))
            raw.append(buf)
        n = len(self.element_set)
        m = n + 1
        return dict(pair_C=raw[0].reshape(n, 3), pair_A=raw[1].reshape(n, 58, n, 3), pair_B=raw[2].reshape(n, -1, n, n, 3),
                    site_C=raw[3].reshape(m), site_A=raw[4].reshape(m, 42, m), site_B=raw[5].reshape(m, -1, m, m))

    # ---- hot path
    def eval_barriers(self, site_i, site_j, walker=None, want_parts=False):
        i, j = _i64(site_i), _i64(site_j)
        n = len(i)
        w = None if walker is None else np.ascontiguousarray(walker, dtype=np.int32)
        ea = np.empty(n); de = np.empty(n)
        d = np.empty(n) if want_parts else None
        ks = np.empty(n) if want_parts else None
        _check(lib().lmc_eval_barriers(self.h, C.c_int64(n), _p(w), _p(i), _p(j), _p(ea), _p(de), _p(d), _p(ks)))
        return (ea, de, d, ks) if want_parts else (ea, de)

    def eval_vacancy_events(self, vacancy_site, walker=None):
        """All 12 jumps of each listed vacancy (KineticMcFirstOmp::BuildEventList order): (neighbour ids, Ea, dE), each [n, 12]."""
        v = _i64(vacancy_site)
        n = len(v)
        w = None if walker is None else np.ascontiguousarray(walker, dtype=np.int32)
        nb = np.empty((n, 12), dtype=np.int64); ea = np.empty((n, 12)); de = np.empty((n, 12))
        _check(lib().lmc_eval_vacancy_events(self.h, C.c_int64(n), _p(w), _p(v), _p(nb), _p(ea), _p(de)))
        return nb, ea, de

    def eval_vacancy_events_dev(self, n, walker_ptr, vacancy_ptr, neighbour_ptr, ea_ptr, de_ptr):
        _check(lib().lmc_eval_vacancy_events_dev(self.h, C.c_int64(int(n)), C.c_void_p(walker_ptr or None), C.c_void_p(vacancy_ptr),
                                                 C.c_void_p(neighbour_ptr), C.c_void_p(ea_ptr), C.c_void_p(de_ptr)))

    def eval_swap_de(self, site_a, site_b, walker=None):
        a, b = _i64(site_a), _i64(site_b)
        w = None if walker is None else np.ascontiguousarray(walker, dtype=np.int32)
        out = np.empty(len(a))
        _check(lib().lmc_eval_swap_de(self.h, C.c_int64(len(a)), _p(w), _p(a), _p(b), _p(out)))
        return out

    def eval_pair_de(self, site_a, site_b, walker=None):
        """EnergyChangePredictorPair semantics: first-neighbour pairs only (LmcOutOfRange otherwise)."""
        a, b = _i64(site_a), _i64(site_b)
        w = None if walker is None else np.ascontiguousarray(walker, dtype=np.int32)
        out = np.empty(len(a), dtype=np.float64)
        _check(lib().lmc_eval_pair_de(self.h, C.c_int64(len(a)), _p(w), _p(a), _p(b), _p(out)))
        return out

    def barrier_model(self):
        return int(lib().lmc_engine_barrier_model(self.h))

    def eval_site_de(self, site, new_element, walker=None):
        s = _i64(site)
        e = np.ascontiguousarray(new_element, dtype=np.uint8)
        w = None if walker is None else np.ascontiguousarray(walker, dtype=np.int32)
        out = np.empty(len(s))
        _check(lib().lmc_eval_site_de(self.h, C.c_int64(len(s)), _p(w), _p(s), _p(e), _p(out)))
        return out

    def total_energy(self, walker=0, want_counts=False):
        e = C.c_double()
        counts = np.zeros(self.n_types, dtype=np.int64) if want_counts else None
        _check(lib().lmc_total_energy(self.h, int(walker), C.byref(e), _p(counts), self.n_types if want_counts else 0))
        return (e.value, counts) if want_counts else e.value

    # ---- measurement
    def get_elements(self, lattice_ids, walker=0):
        ids = _i64(lattice_ids)
        out = np.empty(len(ids), np.uint8)
        _check(lib().lmc_engine_get_elements(self.h, int(walker), C.c_int64(len(ids)), _p(ids), _p(out)))
        return out

    def find_element(self, element, walker=0):
        """(lowest lattice id holding `element` or -1, number of such sites)."""
        cnt = C.c_int64()
        first = lib().lmc_engine_find_element(self.h, int(walker), int(element), C.byref(cnt))
        if first < -1:
            _check(int(first) + 1000)
        return int(first), int(cnt.value)

    def energy_of_cluster(self, lattice_ids, walker=0, want_counts=False):
        """EnergyPredictor::GetEnergyOfCluster on lattice ids (lmc_energy_of_cluster)."""
        ids = _i64(lattice_ids)
        e = C.c_double()
        counts = np.zeros(self.n_types, np.int64) if want_counts else None
        _check(lib().lmc_energy_of_cluster(self.h, int(walker), _p(ids) if len(ids) else None, C.c_int64(len(ids)), C.byref(e), _p(counts),
                                           self.n_types if want_counts else 0))
        return (e.value, counts) if want_counts else e.value

    def energy_encode(self, lattice_ids=None, walker=0):
        """EnergyPredictor::GetEncode (lattice_ids None) / GetEncodeOfCluster."""
        out = np.empty(self.n_types, np.float64)
        if lattice_ids is None:
            _check(lib().lmc_energy_encode(self.h, int(walker), None, C.c_int64(-1), _p(out), self.n_types))
        else:
            ids = _i64(lattice_ids)
            _check(lib().lmc_energy_encode(self.h, int(walker), _p(ids) if len(ids) else None, C.c_int64(len(ids)), _p(out), self.n_types))
        return out

    def chemical_potential(self, solvent=1):
        """EnergyPredictor::GetChemicalPotential(solvent): {element enum: mu} incl. the vacancy X (0)."""
        n = _check(lib().lmc_chemical_potential(self.h, int(solvent), None, None, 0))
        el = np.zeros(n, np.int32); mu = np.zeros(n, np.float64)
        _check(lib().lmc_chemical_potential(self.h, int(solvent), _p(el), _p(mu), n))
        return {int(a): float(b) for a, b in zip(el, mu)}

    def cuda_stream(self):
        return int(lib().lmc_engine_cuda_stream(self.h) or 0)

    def synchronize(self):
        _check(lib().lmc_engine_synchronize(self.h))

    def last_kernel_ms(self):
        return float(lib().lmc_engine_last_kernel_ms(self.h))

    def launch_count(self):
        return int(lib().lmc_engine_launch_count(self.h))

    def kmc_last_launch_lanes(self):
        """0: half-warp-per-walker kernel; 8 / 16 / 32: block-per-walker latency kernel with that many lanes per jump."""
        return int(lib().lmc_kmc_last_launch_lanes(self.h))

    def kmc_last_launch_handoff(self):
        """True if the last first-order launch ran the half-warp kernel and handed its tail to the latency kernel."""
        return bool(lib().lmc_kmc_last_launch_handoff(self.h))

    def kmc_last_launch_resident_occupancy(self):
        """True if the last first-order launch was a latency-kernel launch with the walkers' occupancy in shared memory."""
        return bool(lib().lmc_kmc_last_launch_resident_occupancy(self.h))

    # ---- device-resident batches (inputs already in HBM): pointers are raw device addresses (e.g. tensor.data_ptr())
    def eval_barriers_dev(self, n, walker_ptr, i_ptr, j_ptr, ea_ptr, de_ptr):
        _check(lib().lmc_eval_barriers_dev(self.h, C.c_int64(int(n)), C.c_void_p(walker_ptr or None), C.c_void_p(i_ptr), C.c_void_p(j_ptr),
                                           C.c_void_p(ea_ptr), C.c_void_p(de_ptr), None, None))

    def eval_swap_de_dev(self, n, walker_ptr, a_ptr, b_ptr, de_ptr):
        _check(lib().lmc_eval_swap_de_dev(self.h, C.c_int64(int(n)), C.c_void_p(walker_ptr or None), C.c_void_p(a_ptr), C.c_void_p(b_ptr),
                                          C.c_void_p(de_ptr)))

    # ---- KMC driver (mc::KineticMcFirstOmp semantics over all walkers)
    def kmc_reset(self):
        _check(lib().lmc_kmc_reset(self.h))

    def kmc_run(self, n_steps, temperature=500.0, temperatures=None, time_temperature=None, rate_corrector=False, seed=0,
                replay_u1=None, replay_u2=None, trace=False, second_order=False):
        """Advance every walker by n_steps (first-order KMC; second_order=True: mc::KineticMcChainOmpi, one uniform per step
        passed as replay_u2).  Returns the trace dict (arrays [n_walkers, n_steps]) if trace else None."""
        keep = []
        prm = KmcParams()
        prm.temperature = float(temperature)
        if temperatures is not None:
            t = np.ascontiguousarray(temperatures, dtype=np.float64); keep.append(t)
            assert t.size == self.n_walkers
            prm.temperatures = t.ctypes.data
        if time_temperature is not None:
            tt = np.ascontiguousarray(time_temperature, dtype=np.float64)
            tt = tt[np.argsort(tt[:, 0], kind="stable")]
            tt_t = np.ascontiguousarray(tt[:, 0]); tt_v = np.ascontiguousarray(tt[:, 1]); keep += [tt_t, tt_v]
            prm.n_time_temperature = len(tt_t); prm.tt_time = tt_t.ctypes.data; prm.tt_temperature = tt_v.ctypes.data
        prm.rate_corrector = int(bool(rate_corrector))
        prm.seed = int(seed)
        u1 = u2 = None
        if replay_u1 is not None:
            u1 = np.ascontiguousarray(replay_u1, dtype=np.float64).reshape(self.n_walkers, n_steps)
        if replay_u2 is not None:
            u2 = np.ascontiguousarray(replay_u2, dtype=np.float64).reshape(self.n_walkers, n_steps)
        tr, out = None, None
        if trace:
            shape = (self.n_walkers, int(n_steps))
            out = {"from": np.zeros(shape, np.int64), "to": np.zeros(shape, np.int64), "slot": np.zeros(shape, np.int32)}
            for k in ("dt", "Ea", "dE", "total_rate", "temperature"):
                out[k] = np.zeros(shape, np.float64)
            tr = KmcTrace(out["from"].ctypes.data, out["to"].ctypes.data, out["slot"].ctypes.data, out["dt"].ctypes.data,
                          out["Ea"].ctypes.data, out["dE"].ctypes.data, out["total_rate"].ctypes.data, out["temperature"].ctypes.data)
        if second_order:
            _check(lib().lmc_kmc_chain_run(self.h, C.byref(prm), C.c_int64(int(n_steps)), _p(u2), C.byref(tr) if tr else None))
        else:
            _check(lib().lmc_kmc_run(self.h, C.byref(prm), C.c_int64(int(n_steps)), _p(u1), _p(u2), C.byref(tr) if tr else None))
        return out

    def kmc_chain_run(self, n_steps, replay_u=None, **kw):
        """mc::KineticMcChainOmpi::Simulate over all walkers (lmc_kmc_chain_run)."""
        return self.kmc_run(n_steps, replay_u2=replay_u, second_order=True, **kw)

    def kmc_set_state(self, time=None, energy=None, steps=None):
        """Restart: overwrite the per-walker clocks after kmc_reset (lmc_kmc_set_state)."""
        t = None if time is None else np.ascontiguousarray(np.broadcast_to(time, self.n_walkers), dtype=np.float64)
        e = None if energy is None else np.ascontiguousarray(np.broadcast_to(energy, self.n_walkers), dtype=np.float64)
        s = None if steps is None else np.ascontiguousarray(np.broadcast_to(steps, self.n_walkers), dtype=np.int64)
        _check(lib().lmc_kmc_set_state(self.h, _p(t), _p(e), _p(s)))

    def kmc_state(self):
        n = self.n_walkers
        out = dict(time=np.empty(n), energy=np.empty(n), steps=np.empty(n, np.int64), vacancy=np.empty(n, np.int64),
                   temperature=np.empty(n))
        _check(lib().lmc_kmc_get_state(self.h, _p(out["time"]), _p(out["energy"]), _p(out["steps"]), _p(out["vacancy"]),
                                       _p(out["temperature"])))
        return out

    # ---- CMC / SA driver (mc::CanonicalMcOmp / SimulatedAnnealing semantics over all replicas)
    def cmc_reset(self, sa_initial_temperature=0.0, sa_maximum_steps=0):
        _check(lib().lmc_cmc_reset(self.h, C.c_double(sa_initial_temperature), C.c_uint64(int(sa_maximum_steps))))

    def _cmc_params(self, temperature, temperatures, seed, batch_size, keep):
        prm = CmcParams()
        prm.temperature = float(temperature)
        if temperatures is not None:
            t = np.ascontiguousarray(temperatures, dtype=np.float64); keep.append(t)
            prm.temperatures = t.ctypes.data
        prm.seed = int(seed)
        prm.batch_size = int(batch_size)
        return prm

    def cmc_run(self, n_trials, temperature=800.0, temperatures=None, seed=0, batch_size=0):
        keep = []
        prm = self._cmc_params(temperature, temperatures, seed, batch_size, keep)
        _check(lib().lmc_cmc_run(self.h, C.byref(prm), C.c_int64(int(n_trials))))

    def cmc_grid_run(self, n_trials, temperature=800.0, seed=0, batch_size=0):
        """ONE lattice on the whole GPU (and, after cmc_attach_peers, on several GPUs): lmc_cmc_grid_run."""
        keep = []
        prm = self._cmc_params(temperature, None, seed, batch_size, keep)
        _check(lib().lmc_cmc_grid_run(self.h, C.byref(prm), C.c_int64(int(n_trials))))

    def cmc_domain_run(self, n_trials, temperature=800.0, temperatures=None, seed=0, domain_edge=0, rounds_per_sweep=0,
                       speculate=0, lanes=0, passes=0.0):
        """Domain-decomposed ("sublattice") CMC / SA driver: lmc_cmc_domain_run (one lattice or many replicas; after
        cmc_domain_attach_peers one lattice over several GPUs)."""
        keep = []
        prm = self._cmc_params(temperature, temperatures, seed, 0, keep)
        dom = CmcDomainParams(int(domain_edge), int(rounds_per_sweep), int(speculate), int(lanes), float(passes))
        _check(lib().lmc_cmc_domain_run(self.h, C.byref(prm), C.byref(dom), C.c_int64(int(n_trials))))

    def cmc_domain_last_shape(self):
        out = (C.c_int32 * 7)()
        _check(lib().lmc_cmc_domain_last_shape(self.h, out))
        return dict(zip(("domain_edge", "domains", "lanes", "threads", "blocks", "rounds_per_sweep", "speculate"), (int(v) for v in out)))

    def cmc_domain_handles(self):
        """192 bytes: CUDA IPC handles of the two occupancy buffers and the line buffer (to be all-gathered over the ranks)."""
        buf = C.create_string_buffer(192)
        _check(lib().lmc_cmc_domain_handles(self.h, buf))
        return buf.raw

    def cmc_domain_attach_peers(self, rank, world, handles):
        """handles: list of `world` 192-byte blobs in rank order (entry `rank` is ignored)."""
        blob = b"".join(bytes(h) for h in handles)
        assert len(blob) == 192 * world
        _check(lib().lmc_cmc_domain_attach_peers(self.h, int(rank), int(world), blob))

    def cmc_exchange_handle(self):
        """64-byte CUDA IPC handle of this engine's exchange buffer (to be all-gathered over the ranks)."""
        buf = C.create_string_buffer(64)
        _check(lib().lmc_cmc_exchange_handle(self.h, buf))
        return buf.raw

    def cmc_attach_peers(self, rank, world, handles, grid_ctas=0):
        """handles: list of `world` 64-byte handles in rank order (entry `rank` is ignored)."""
        blob = b"".join(bytes(h) for h in handles)
        assert len(blob) == 64 * world
        _check(lib().lmc_cmc_attach_peers(self.h, int(rank), int(world), blob, int(grid_ctas)))

    def cmc_replay(self, site_a, site_b, u, temperature=800.0, walker=0, batch_size=0):
        a, b = _i64(site_a), _i64(site_b)
        u = np.ascontiguousarray(u, dtype=np.float64)
        n = len(a)
        out = dict(dE=np.empty(n), energy_before=np.empty(n), temperature_before=np.empty(n), accepted=np.empty(n, np.uint8))
        keep = []
        prm = self._cmc_params(temperature, None, 0, batch_size, keep)
        _check(lib().lmc_cmc_replay(self.h, int(walker), C.byref(prm), C.c_int64(n), _p(a), _p(b), _p(u), _p(out["dE"]),
                                    _p(out["energy_before"]), _p(out["temperature_before"]), _p(out["accepted"])))
        out["accepted"] = out["accepted"].astype(bool)
        return out

    def cmc_state(self):
        n = self.n_walkers
        out = dict(energy=np.empty(n), steps=np.empty(n, np.int64), accepted=np.empty(n, np.int64), temperature=np.empty(n))
        _check(lib().lmc_cmc_get_state(self.h, _p(out["energy"]), _p(out["steps"]), _p(out["accepted"]), _p(out["temperature"])))
        return out

    # ---- debug taps (device)
    def debug_pair(self, i, j, walker=0):
        n_e = len(self.element_set)
        len_mmm = len(tables_group_sizes("mmm", n_e)); len_mm2 = len(tables_group_sizes("mm2", n_e))
        out = dict(state=np.empty(60, np.int64), mmm=np.empty(58, np.int64), mm2=np.empty(58, np.int64),
                   mm2_backward=np.empty(58, np.int64), start_counts=np.empty(self.n_types, np.int32),
                   end_counts=np.empty(self.n_types, np.int32), enc_mmm=np.empty(len_mmm, np.int32),
                   enc_mm2_f=np.empty(len_mm2, np.int32), enc_mm2_b=np.empty(len_mm2, np.int32))
        _check(lib().lmc_debug_pair(self.h, int(walker), C.c_int64(int(i)), C.c_int64(int(j)), _p(out["state"]), _p(out["mmm"]),
                                    _p(out["mm2"]), _p(out["mm2_backward"]), _p(out["start_counts"]), _p(out["end_counts"]),
                                    _p(out["enc_mmm"]), _p(out["enc_mm2_f"]), _p(out["enc_mm2_b"])))
        return out

    def debug_site(self, site, new_element, walker=0):
        out = dict(state=np.empty(43, np.int64), start_counts=np.empty(self.n_types, np.int32),
                   end_counts=np.empty(self.n_types, np.int32))
        _check(lib().lmc_debug_site(self.h, int(walker), C.c_int64(int(site)), int(new_element), _p(out["state"]),
                                    _p(out["start_counts"]), _p(out["end_counts"])))
        return out

    # ---- host geometry
    def neighbors(self, shell, site):
        out = np.empty({1: 12, 2: 6, 3: 24}[shell], dtype=np.int64)
        _check(lib().lmc_engine_neighbors(self.h, int(shell), C.c_int64(int(site)), _p(out)))
        return out

    def kmc_event_order(self, site):
        """The 12 first neighbours of `site` in the KMC kernels' event order (class table); equals neighbors(1, site)."""
        out = np.empty(12, dtype=np.int64)
        _check(lib().lmc_engine_kmc_event_order(self.h, C.c_int64(int(site)), _p(out)))
        return out

    def site_coords(self, site):
        out = (C.c_int32 * 3)()
        _check(lib().lmc_engine_site_coords(self.h, C.c_int64(int(site)), out))
        return tuple(out)

    def pair_lists(self, i, j):
        s = np.empty(60, np.int64); m = np.empty(58, np.int64); m2 = np.empty(58, np.int64); mb = np.empty(58, np.int64)
        _check(lib().lmc_engine_pair_lists(self.h, C.c_int64(int(i)), C.c_int64(int(j)), _p(s), _p(m), _p(m2), _p(mb)))
        return s, m, m2, mb

    def site_list(self, site):
        s = np.empty(43, np.int64)
        _check(lib().lmc_engine_site_list(self.h, C.c_int64(int(site)), _p(s)))
        return s

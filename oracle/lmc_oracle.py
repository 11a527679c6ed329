"""oracle/lmc_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

A numpy restatement of the reference's algorithm for the hot path named in BASELINE.json's north_star:
local-environment energy-change / migration-barrier evaluation and the KMC / CMC / SA event loops that
consume it (zhucongx/LatticeMonteCarlo, lmc/{cfg,pred,mc}).  Every function cites the reference
file:line it follows.  It deliberately follows the reference's *floating point geometry* (relative
positions, epsilon comparators, rotation matrices) so that it is independent of the product's
integer-offset tables.

Parity status: PINNED.  The reference has no tests or golden vectors of its own, so the
oracle is pinned against outputs of the reference itself compiled here (oracle/_ref/liblmc_ref.so, built
from /root/reference by oracle/Makefile) -- see tests/test_oracle_vs_reference.py -- and against the
committed fixtures tests/golden/*.npz generated from that library by tests/golden/make_golden.py.

Element codes are the reference's ElementName enum values (lmc/cfg/include/Element.hpp:7): X=0, Al=1, ...
Pure-Python loops are used only in one-off setup (mappings for one reference pair/site); everything that
scales with the number of events is vectorised numpy.
"""
from __future__ import annotations

import functools
import itertools
import json
import math

import numpy as np

K_EPSILON = 1e-4                      # lmc/cfg/include/VectorMatrix.hpp:65
LATTICE_CONSTANT = 4.046              # lmc/cfg/include/Constants.hpp:6
CUTOFFS = (3.5, 4.8, 5.3)             # Constants.hpp:8-10
K_BOLTZMANN = 8.617333262145e-5       # Constants.hpp:31
K_PREFACTOR = 1e13                    # Constants.hpp:32
ELEMENT_NAMES = {0: "X", 1: "Al", 2: "Mg", 3: "Zn", 4: "Cu", 5: "Sn",
                 6: "pAl", 7: "pMg", 8: "pZn", 9: "pCu", 10: "pSn"}   # Element.hpp:7,51-66
ELEMENT_CODES = {v: k for k, v in ELEMENT_NAMES.items()}
# normalisers: pred/src/VacancyMigrationPredictorQuartic.cpp:14 == EnergyChangePredictorPairSite.cpp:11
DE_CLUSTER_COUNTER = (256, 1536, 768, 3072, 2048, 3072, 6144, 6144, 6144, 6144, 2048)
# pred/src/EnergyPredictor.cpp:8
E_CLUSTER_COUNTER = (256, 3072, 1536, 6144, 12288, 6144, 12288, 6144, 12288, 12288, 12288)


# ----------------------------------------------------------------------------------------------- cfg
def _eps_rank(values):
    """Rank values along the last axis, merging values closer than K_EPSILON (chain merge).  Equivalent to the
    reference's `diff < -eps / diff > eps` comparisons as long as equal-in-exact-arithmetic values agree to << eps
    and distinct ones differ by >> eps (asserted by callers where it matters)."""
    values = np.asarray(values, dtype=np.float64)
    order = np.argsort(values, axis=-1, kind="stable")
    srt = np.take_along_axis(values, order, axis=-1)
    step = np.concatenate([np.zeros(srt.shape[:-1] + (1,), dtype=np.int64),
                           (np.diff(srt, axis=-1) > K_EPSILON).astype(np.int64)], axis=-1)
    rank_sorted = np.cumsum(step, axis=-1)
    ranks = np.empty_like(rank_sorted)
    np.put_along_axis(ranks, order, rank_sorted, axis=-1)
    return ranks


def _lexsort_rows(keys):
    """keys: list of (E, n) int arrays, most significant first -> (E, n) permutation (stable)."""
    e, n = keys[0].shape
    radix = n + 1
    combined = np.zeros((e, n), dtype=np.int64)
    for k in keys:
        combined = combined * radix + k
    return np.argsort(combined, axis=-1, kind="stable")


class Config:
    """Restatement of cfg::Config (lmc/cfg/include/Config.h:14-112) limited to what the hot path reads:
    basis, relative positions, element per lattice id, and the three sorted adjacency lists."""

    def __init__(self, basis, rel, occ):
        self.basis = np.asarray(basis, dtype=np.float64)
        self.rel = np.asarray(rel, dtype=np.float64)
        self.occ = np.asarray(occ, dtype=np.uint8).copy()
        self.nn = None
        self.update_neighbors()

    @property
    def num_sites(self):
        return self.rel.shape[0]

    @classmethod
    def generate_fcc(cls, factors, occ=None):
        """cfg::GenerateFCC (cfg/src/Config.cpp:1060-1095): k, j, i loops (outer->inner), 4 basis sites."""
        if np.isscalar(factors):
            factors = (factors,) * 3
        fx, fy, fz = (int(v) for v in factors)
        basis = np.diag([LATTICE_CONSTANT * fx, LATTICE_CONSTANT * fy, LATTICE_CONSTANT * fz])
        k, j, i = np.meshgrid(np.arange(fz), np.arange(fy), np.arange(fx), indexing="ij")
        xr, yr, zr = i.reshape(-1, 1).astype(float), j.reshape(-1, 1).astype(float), k.reshape(-1, 1).astype(float)
        off = np.array([[0, 0, 0], [.5, .5, 0], [.5, 0, .5], [0, .5, .5]])
        rel = np.stack([(xr + off[:, 0]) / fx, (yr + off[:, 1]) / fy, (zr + off[:, 2]) / fz], axis=-1).reshape(-1, 3)
        if occ is None:
            occ = np.full(rel.shape[0], ELEMENT_CODES["Al"], dtype=np.uint8)
        return cls(basis, rel, occ)

    def reassign_lattice_vector(self):
        """Config::ReassignLatticeVector (Config.cpp:466-552): sort sites by relative position with the epsilon
        `operator<` of Vector_d (VectorMatrix.hpp:80-90), renumber, carry the occupancy along, re-sort adjacency."""
        keys = [_eps_rank(self.rel[None, :, d])[0][None, :] for d in range(3)]
        n = self.num_sites
        combined = (keys[0].astype(np.int64) * (n + 1) + keys[1]) * (n + 1) + keys[2]
        perm = np.argsort(combined[0], kind="stable")      # perm[new] = old
        self.rel = self.rel[perm]
        self.occ = self.occ[perm]
        self.update_neighbors()
        return perm

    def update_neighbors(self):
        """Config::UpdateNeighbors (Config.cpp:955-1045): bonded if |d|^2 < cutoff^2 under the minimum-image
        convention; three shells; each list sorted ascending.  (The reference's cell list is an acceleration
        structure; a periodic KD-tree gives the same sets.)"""
        from scipy.spatial import cKDTree
        lengths = np.array([np.linalg.norm(self.basis[d]) for d in range(3)])
        assert np.allclose(self.basis, np.diag(np.diag(self.basis))), "oracle supports orthorhombic cells"
        cart = (self.rel % 1.0) * lengths
        cart = np.where(cart >= lengths, cart - lengths, cart)
        tree = cKDTree(cart, boxsize=lengths)
        pairs = tree.query_pairs(CUTOFFS[2], output_type="ndarray")
        d = self.rel[pairs[:, 1]] - self.rel[pairs[:, 0]]
        d -= np.round(d)                                    # Lattice.hpp:64-76 minimum image
        r2 = np.sum((d * lengths) ** 2, axis=1)
        n = self.num_sites
        lists = []
        lo = 0.0
        for cutoff, count in zip(CUTOFFS, (12, 6, 24)):
            sel = (r2 >= lo) & (r2 < cutoff * cutoff)
            a = np.concatenate([pairs[sel, 0], pairs[sel, 1]])
            b = np.concatenate([pairs[sel, 1], pairs[sel, 0]])
            order = np.lexsort((b, a))
            a, b = a[order], b[order]
            assert np.all(np.bincount(a, minlength=n) == count), "unexpected shell population (cell too small?)"
            lists.append(b.reshape(n, count).astype(np.int64))
            lo = cutoff * cutoff
        self.nn = lists

    # Config::FindDistanceLabelBetweenLattice (Config.cpp:354-371)
    def distance_label(self, a, b):
        for label, lst in enumerate(self.nn, start=1):
            if b in lst[a]:
                return label
        return -1

    def distance_label_matrix(self, ids):
        """labels[p, q] for a list of lattice ids (0 on the diagonal / -1 if not within 3NN)."""
        ids = np.asarray(ids)
        out = np.full((len(ids), len(ids)), -1, dtype=np.int64)
        for label, lst in enumerate(self.nn, start=1):
            hit = (lst[ids][:, :, None] == ids[None, None, :]).any(axis=1)
            out[hit] = label
        return out

    def lattice_jump(self, a, b):
        """Config::LatticeJump (Config.cpp:431-456) as seen through GetElementAtLatticeId: the elements swap."""
        self.occ[a], self.occ[b] = self.occ[b], self.occ[a]

    def neighbors_set_of_site(self, i):          # Config.cpp:312-326
        return np.concatenate([[i], self.nn[0][i], self.nn[1][i], self.nn[2][i]])

    def neighbors_set_of_pair(self, i, j):       # Config.cpp:336-352
        return np.unique(np.concatenate([self.neighbors_set_of_site(i), self.neighbors_set_of_site(j)]))


# ----------------------------------------------------------------------------- ordered neighbourhoods
def _pair_center(cfg, i, j):
    """Config::GetLatticePairCenter (Config.cpp:228-245), vectorised over pairs."""
    first = cfg.rel[i].copy()
    second = cfg.rel[j]
    for _ in range(4):
        period = np.trunc((first - second) / 0.5)
        if not np.any(period):
            break
        first = first - period
    return 0.5 * (first + second)


def _rel_distance(cfg, a, b):
    """GetRelativeDistanceVectorLattice (Lattice.hpp:64-76): wrap into [-0.5, 0.5)."""
    d = cfg.rel[b] - cfg.rel[a]
    d = np.where(d >= 0.5, d - 1.0, d)
    d = np.where(d < -0.5, d + 1.0, d)
    return d


def _normalize(v):
    return v / np.sqrt(np.sum(v * v, axis=-1, keepdims=True))


def _pair_rotation(cfg, i, j):
    """Config::GetLatticePairRotationMatrix (Config.cpp:247-264), vectorised: rows = (pair direction,
    first 1NN of `first` -- in ascending lattice id order -- perpendicular to it, their cross product);
    returns the three row vectors (E,3) each; a position r maps to (r.d, r.v, r.c)."""
    d = _normalize(_rel_distance(cfg, i, j) @ cfg.basis)
    nbr = cfg.nn[0][i]                                                     # (E,12) ascending ids
    jump = _normalize(np.stack([_rel_distance(cfg, i, nbr[:, q]) @ cfg.basis for q in range(12)], axis=1))
    dots = np.abs(np.einsum("eqd,ed->eq", jump, d))
    first_perp = np.argmax(dots < K_EPSILON, axis=1)                       # first hit in id order
    v = jump[np.arange(len(i)), first_perp]
    c = np.cross(d, v)
    return d, v, c


def _centered_rotated(cfg, i, j, ids):
    """Move the pair centre to (.5,.5,.5), wrap, rotate about the centre, wrap (EnergyUtility.cpp:261-287 and
    Config.cpp:1047-1058 RotateLatticeVector).  ids: (E,n) lattice ids -> (E,n,3) relative positions."""
    move = 0.5 - _pair_center(cfg, i, j)                                   # (E,3)
    rel = cfg.rel[ids] + move[:, None, :]
    rel -= np.floor(rel)
    d, v, c = _pair_rotation(cfg, i, j)
    rot = np.stack([np.einsum("end,ed->en", rel, d), np.einsum("end,ed->en", rel, v),
                    np.einsum("end,ed->en", rel, c)], axis=-1)
    half = np.full(3, 0.5)
    shift = 0.5 - np.stack([d @ half, v @ half, c @ half], axis=-1)        # 0.5 - (0.5,0.5,0.5)*R
    rot += shift[:, None, :]
    rot -= np.floor(rot)
    return rot


def pair_neighbor_ids(cfg, i, j):
    """(E,60) union of the 1-3NN shells of both sites incl. the sites themselves (Config.cpp:336-352)."""
    i = np.atleast_1d(i); j = np.atleast_1d(j)
    both = np.concatenate([i[:, None], cfg.nn[0][i], cfg.nn[1][i], cfg.nn[2][i],
                           j[:, None], cfg.nn[0][j], cfg.nn[1][j], cfg.nn[2][j]], axis=1)
    both.sort(axis=1)
    keep = np.concatenate([np.ones((len(i), 1), bool), both[:, 1:] != both[:, :-1]], axis=1)
    assert np.all(keep.sum(axis=1) == 60), "pair neighbourhood must hold 60 sites (Constants.hpp:20)"
    return both[keep].reshape(len(i), 60)


def sorted_lists_of_pairs(cfg, i, j):
    """State / MMM / MM2 ordered id lists for many jump pairs at once:
    GetSortedLatticeVectorStateOfPair (EnergyUtility.cpp:261-287, PositionCompareState LatticeCluster.hpp:62-75),
    GetSymmetricallySortedLatticeVectorMMM/MM2 (EnergyUtility.cpp:45-104, PositionCompareMMM/MM2
    LatticeCluster.hpp:105-154).  Returns (E,60), (E,58), (E,58)."""
    i = np.atleast_1d(np.asarray(i, dtype=np.int64)); j = np.atleast_1d(np.asarray(j, dtype=np.int64))
    ids = pair_neighbor_ids(cfg, i, j)
    rot = _centered_rotated(cfg, i, j, ids)
    kx, ky, kz = (_eps_rank(rot[:, :, d]) for d in range(3))
    state = np.take_along_axis(ids, _lexsort_rows([kx, ky, kz]), axis=1)
    # mmm / mm2 exclude the jump pair itself
    not_pair = (ids != i[:, None]) & (ids != j[:, None])
    ids58 = ids[not_pair].reshape(len(i), 58)
    rot58 = rot[not_pair].reshape(len(i), 58, 3)
    norm = _eps_rank(np.sum((rot58 - 0.5) ** 2, axis=-1))
    absx = _eps_rank(np.abs(rot58[:, :, 0] - 0.5))
    kx, ky, kz = (_eps_rank(rot58[:, :, d]) for d in range(3))
    mmm = np.take_along_axis(ids58, _lexsort_rows([norm, absx, kx, ky, kz]), axis=1)
    mm2 = np.take_along_axis(ids58, _lexsort_rows([norm, kx, ky, kz]), axis=1)
    return state, mmm, mm2


def _sorted_positions_of_pair(cfg, i, j, which):
    """Single pair: ordered ids and their rotated positions (needed to build the mappings)."""
    ii = np.array([i]); jj = np.array([j])
    ids = pair_neighbor_ids(cfg, ii, jj)
    rot = _centered_rotated(cfg, ii, jj, ids)[0]
    ids = ids[0]
    if which != "state":
        keep = (ids != i) & (ids != j)
        ids, rot = ids[keep], rot[keep]
    cmp = {"state": _position_compare_state, "mmm": _position_compare_mmm, "mm2": _position_compare_mm2}[which]
    order = sorted(range(len(ids)), key=functools.cmp_to_key(lambda a, b: _as_cmp(cmp, rot[a], rot[b])))
    return ids[order], rot[order]


def sorted_list_of_sites(cfg, sites):
    """GetSortedLatticeVectorStateOfSite (EnergyUtility.cpp:288-313): centre the site, wrap, sort (x,y,z). (E,43)."""
    sites = np.atleast_1d(np.asarray(sites, dtype=np.int64))
    ids = np.concatenate([sites[:, None], cfg.nn[0][sites], cfg.nn[1][sites], cfg.nn[2][sites]], axis=1)
    rel = cfg.rel[ids] + (0.5 - cfg.rel[sites])[:, None, :]
    rel -= np.floor(rel)
    kx, ky, kz = (_eps_rank(rel[:, :, d]) for d in range(3))
    return np.take_along_axis(ids, _lexsort_rows([kx, ky, kz]), axis=1)


# scalar comparators, used for the (small) mapping construction exactly as written in the reference
def _as_cmp(less, a, b):
    if less(a, b):
        return -1
    if less(b, a):
        return 1
    return 0


def _position_compare_state(l, r):            # LatticeCluster.hpp:62-75
    for d in range(3):
        diff = l[d] - r[d]
        if diff < -K_EPSILON:
            return True
        if diff > K_EPSILON:
            return False
    return False


def _inner_centered(p):
    return (p[0] - 0.5) ** 2 + (p[1] - 0.5) ** 2 + (p[2] - 0.5) ** 2


def _group_compare_mmm(l, r):                 # LatticeCluster.hpp:77-90
    diff = _inner_centered(l) - _inner_centered(r)
    if diff < -K_EPSILON:
        return True
    if diff > K_EPSILON:
        return False
    return abs(l[0] - 0.5) - abs(r[0] - 0.5) < -K_EPSILON


def _group_compare_mm2(l, r):                 # LatticeCluster.hpp:91-103
    diff = _inner_centered(l) - _inner_centered(r)
    if diff < -K_EPSILON:
        return True
    if diff > K_EPSILON:
        return False
    return l[0] - r[0] < -K_EPSILON


def _position_compare_mmm(l, r):              # LatticeCluster.hpp:105-132
    diffs = [_inner_centered(l) - _inner_centered(r), abs(l[0] - 0.5) - abs(r[0] - 0.5),
             l[0] - r[0], l[1] - r[1], l[2] - r[2]]
    for diff in diffs:
        if diff < -K_EPSILON:
            return True
        if diff > K_EPSILON:
            return False
    return False


def _position_compare_mm2(l, r):              # LatticeCluster.hpp:133-154
    diffs = [_inner_centered(l) - _inner_centered(r), l[0] - r[0], l[1] - r[1], l[2] - r[2]]
    for diff in diffs:
        if diff < -K_EPSILON:
            return True
        if diff > K_EPSILON:
            return False
    return False


# ---------------------------------------------------------------------------------- cluster mappings
def _triplet_label(l01, l12, l20):
    """GetLabel for three sites (EnergyUtility.cpp:345-379)."""
    b = tuple(sorted((l01, l12, l20)))
    return {(1, 1, 1): 4, (1, 1, 2): 5, (1, 1, 3): 6, (1, 2, 3): 7, (1, 3, 3): 8, (2, 3, 3): 9, (3, 3, 3): 10}.get(b, -1)


def _state_mapping(cfg, ids, centre_ids):
    """GetClusterParametersMappingStatePair / ...StateSite (EnergyUtility.cpp:393-581): clusters (as positions in
    the ordered list, index1 > index2 > index3) that touch a centre site: labels 0 singlet, 1-3 pairs, 4-7 triplets."""
    n = len(ids)
    lab = cfg.distance_label_matrix(ids)
    centre = np.isin(ids, centre_ids)
    mapping = [[] for _ in range(8)]
    for p1 in range(n):
        if centre[p1]:
            mapping[0].append((p1,))
        for p2 in range(p1):
            if centre[p1] or centre[p2]:
                if lab[p1, p2] in (1, 2, 3):
                    mapping[lab[p1, p2]].append((p1, p2))
                else:
                    continue                                  # `default: continue` (EnergyUtility.cpp:434)
            for p3 in range(p2):
                if centre[p1] or centre[p2] or centre[p3]:
                    t = _triplet_label(lab[p1, p2], lab[p2, p3], lab[p3, p1])
                    if 4 <= t <= 7:
                        mapping[t].append((p1, p2, p3))
    return mapping


def reference_pair(cfg):
    return 0, int(cfg.nn[0][0][0])            # {0, first 1NN of site 0} (EnergyUtility.cpp:171,217,395)


def mapping_state_pair(cfg):
    i, j = reference_pair(cfg)
    ids, _ = _sorted_positions_of_pair(cfg, i, j, "state")
    return _state_mapping(cfg, ids, [i, j])


def mapping_state_site(cfg):
    ids = sorted_list_of_sites(cfg, [0])[0]
    return _state_mapping(cfg, ids, [0])


def _average_mapping(cfg, which):
    """GetAverageClusterParametersMappingMMM / MM2 (EnergyUtility.cpp:169-259) with the grouping helper
    (:106-167): singlets and 1NN/2NN/3NN pairs among the 58 ordered sites, each family sorted with
    IsClusterSmallerSymmetrically* (LatticeCluster.hpp:212-235) and cut into runs of equivalent clusters.
    Symmetric clusters (FindSymmetryLabel, LatticeCluster.hpp:171-180,199-208) carry a leading -1 (SIZE_MAX)."""
    i, j = reference_pair(cfg)
    ids, pos = _sorted_positions_of_pair(cfg, i, j, which)
    group_less = _group_compare_mmm if which == "mmm" else _group_compare_mm2
    lab = cfg.distance_label_matrix(ids)

    def cluster_less(a, b):                   # clusters = tuples of list positions, already position-sorted
        for pa, pb in zip(a, b):
            if group_less(pos[pa], pos[pb]):
                return True
            if group_less(pos[pb], pos[pa]):
                return False
        return False

    def symmetric(c):
        return any(not group_less(pos[c[x]], pos[c[y]]) and not group_less(pos[c[y]], pos[c[x]])
                   for x in range(len(c)) for y in range(x))

    families = [[(p,) for p in range(len(ids))], [], [], []]
    for p1 in range(len(ids)):
        for p2 in range(p1):
            if lab[p1, p2] in (1, 2, 3):
                # LatticeClusterMMM/MM2::Sort orders the members by PositionCompare*; the list is already in that
                # order, so the sorted cluster is (smaller position, larger position)
                families[lab[p1, p2]].append((p2, p1))
    mapping = []
    for family in families:
        family = sorted(family, key=functools.cmp_to_key(lambda a, b: _as_cmp(cluster_less, a, b)))
        start = 0
        while start < len(family):
            end = start + 1
            while end < len(family) and not cluster_less(family[start], family[end]):
                end += 1
            mapping.append([((-1,) + c) if symmetric(c) else c for c in family[start:end]])
            start = end
    return mapping


def mapping_mmm(cfg):
    return _average_mapping(cfg, "mmm")


def mapping_mm2(cfg):
    return _average_mapping(cfg, "mm2")


def canonical_mapping(mapping):
    """Order-insensitive form (clusters sorted inside each group; group order kept)."""
    return [sorted(tuple(int(v) for v in c) for c in g) for g in mapping]


def mapping_checksum(mapping):
    """FNV-1a-64 checksum defined in SURVEY.md Appendix A.7."""
    h = 0xcbf29ce484222325
    def feed(byte):
        nonlocal h
        h ^= byte
        h = (h * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    for g in canonical_mapping(mapping):
        feed(0xFE)
        for c in g:
            feed(0xFD)
            for v in c:
                feed(255 if v < 0 else v)
    return "%016x" % h


# ------------------------------------------------------------------------------------ cluster types
def element_set_sorted(codes):
    """std::set<Element> order = by GetString() (Element.hpp:44-46)."""
    return sorted(set(int(c) for c in codes), key=lambda c: ELEMENT_NAMES[c])


def cluster_types(element_codes):
    """InitializeClusterHashMap (EnergyUtility.cpp:314-343) ordered like std::map<ElementCluster,...>
    (ElementCluster.hpp:40-50): by size, then label, then element strings.  Returns list of (label, codes)."""
    es = element_set_sorted(list(element_codes) + [0])
    name = ELEMENT_NAMES
    types = set()
    for e1 in es:
        types.add((0, (e1,)))
        for e2 in es:
            if e2 == 0:
                continue
            if e1 == 0 and name[e2][0] == "p":
                continue
            for label in (1, 2, 3):
                types.add((label, tuple(sorted((e1, e2), key=lambda c: name[c]))))
            for e3 in es:
                if e3 == 0 or name[e3][0] == "p":
                    continue
                for label in (4, 5, 6, 7):
                    types.add((label, tuple(sorted((e1, e2, e3), key=lambda c: name[c]))))
    return sorted(types, key=lambda t: (len(t[1]), t[0], [name[c] for c in t[1]]))


class ClusterIndexer:
    """pred::ClusterIndexer (EnergyUtility.cpp:798-819) as dense lookup tables over element codes."""

    def __init__(self, element_codes, counter):
        self.types = cluster_types(element_codes)
        self.size = len(self.types)
        self.total_bonds = np.array([counter[t[0]] for t in self.types], dtype=np.float64)
        m = max(max(t[1]) for t in self.types) + 1
        self.lut1 = np.full(m, -1, dtype=np.int64)
        self.lut2 = np.full((8, m, m), -1, dtype=np.int64)
        self.lut3 = np.full((8, m, m, m), -1, dtype=np.int64)
        for idx, (label, el) in enumerate(self.types):
            for perm in set(itertools.permutations(el)):
                if len(perm) == 1:
                    self.lut1[perm[0]] = idx
                elif len(perm) == 2:
                    self.lut2[label, perm[0], perm[1]] = idx
                else:
                    self.lut3[label, perm[0], perm[1], perm[2]] = idx

    def index(self, label, elems):
        """elems: (..., arity) codes -> type index (raises like std::out_of_range if unknown)."""
        elems = np.asarray(elems, dtype=np.int64)
        arity = elems.shape[-1]
        if arity == 1:
            out = self.lut1[elems[..., 0]]
        elif arity == 2:
            out = self.lut2[label, elems[..., 0], elems[..., 1]]
        else:
            out = self.lut3[label, elems[..., 0], elems[..., 1], elems[..., 2]]
        if np.any(out < 0):
            raise IndexError("Cluster not found in ClusterIndexer")
        return out


def _count_types(indexer, mapping, elem_of_pos):
    """Per-event histogram over cluster types.  elem_of_pos: (E, n_positions) element codes."""
    e = elem_of_pos.shape[0]
    counts = np.zeros((e, indexer.size), dtype=np.int32)
    rows = np.arange(e)[:, None]
    for label, clusters in enumerate(mapping):
        if not clusters:
            continue
        pos = np.asarray(clusters, dtype=np.int64)                       # (C, arity)
        idx = indexer.index(label, elem_of_pos[:, pos])                  # (E, C)
        np.add.at(counts, (np.broadcast_to(rows, idx.shape), idx), 1)
    return counts


# --------------------------------------------------------------------------------- one-hot encoding
def one_hot_encode(mapping, elem_of_pos, element_codes):
    """GetOneHotParametersFromMap (EnergyUtility.cpp:743-796) with GetOneHotEncodeHashmap (:7-43).
    elem_of_pos: (E, 58) codes in mmm / mm2 list order -> (E, L) averaged one-hots; also returns the integer
    counts (E, L) and the per-slot group sizes (L,)."""
    es = element_set_sorted(element_codes)
    n = len(es)
    rank = np.full(max(ELEMENT_NAMES) + 1, -1, dtype=np.int64)
    for r, c in enumerate(es):
        rank[c] = r
    tri = np.zeros((n, n), dtype=np.int64)
    ct = 0
    for a in range(n):
        for b in range(a, n):
            tri[a, b] = tri[b, a] = ct
            ct += 1
    e = elem_of_pos.shape[0]
    cols, sizes = [], []
    r_all = rank[elem_of_pos]
    if np.any(r_all < 0):
        raise KeyError("element without one-hot code (e.g. a second vacancy) in the 58-site list")
    rows = np.arange(e)[:, None]
    for group in mapping:
        first = group[0]
        if first[0] < 0:
            length = n * (n + 1) // 2
        else:
            length = n ** len(first)
        cnt = np.zeros((e, length), dtype=np.int32)
        arr = np.asarray(group, dtype=np.int64)
        if arr[0, 0] < 0:
            t = tri[r_all[:, arr[:, 1]], r_all[:, arr[:, 2]]]
        elif arr.shape[1] == 1:
            t = r_all[:, arr[:, 0]]
        else:
            t = r_all[:, arr[:, 0]] * n + r_all[:, arr[:, 1]]
        np.add.at(cnt, (np.broadcast_to(rows, t.shape), t), 1)
        cols.append(cnt)
        sizes.append(np.full(length, len(group), dtype=np.float64))
    counts = np.concatenate(cols, axis=1)
    sizes = np.concatenate(sizes)
    return counts / sizes, counts, sizes


# --------------------------------------------------------------------------------------- predictors
def load_coefficients(json_path_or_dict):
    if isinstance(json_path_or_dict, dict):
        return json_path_or_dict
    with open(json_path_or_dict) as f:
        return json.load(f)


class VacancyMigrationPredictorQuartic:
    """pred::VacancyMigrationPredictorQuartic (pred/src/VacancyMigrationPredictorQuartic.cpp)."""

    def __init__(self, coefficients, reference_config, element_codes):
        self.elements = element_set_sorted(element_codes)
        co = load_coefficients(coefficients)
        self.mapping_mmm = mapping_mmm(reference_config)            # :21
        self.mapping_mm2 = mapping_mm2(reference_config)            # :22
        self.mapping_state = mapping_state_pair(reference_config)   # :23
        self.indexer = ClusterIndexer(self.elements, DE_CLUSTER_COUNTER)   # :26-35
        self.base_theta = np.asarray(co["Base"]["theta"], dtype=np.float64)
        self.params = {}
        for name, p in co.items():                                   # :44-62
            if name == "Base":
                continue
            self.params[ELEMENT_CODES.get(name, 0)] = {k: np.asarray(v, dtype=np.float64) for k, v in p.items()}

    # ---- GetDe (:112-166)
    def de_counts(self, cfg, i, j):
        i = np.atleast_1d(np.asarray(i, dtype=np.int64)); j = np.atleast_1d(np.asarray(j, dtype=np.int64))
        state, _, _ = sorted_lists_of_pairs(cfg, i, j)
        start = cfg.occ[state].astype(np.int64)
        mig = cfg.occ[j].astype(np.int64)
        end = start.copy()
        end[state == i[:, None]] = mig                               # first site receives the migrating element
        end[state == j[:, None]] = 0                                 # second site becomes the vacancy
        return _count_types(self.indexer, self.mapping_state, start), _count_types(self.indexer, self.mapping_state, end)

    def get_de(self, cfg, i, j):
        sc, ec = self.de_counts(cfg, i, j)
        enc = (ec.astype(np.float64) - sc.astype(np.float64)) / self.indexer.total_bonds
        return _seq_dot(self.base_theta, enc)

    # ---- GetD (:216-246), GetKs (:167-215)
    def encodes(self, cfg, i, j):
        i = np.atleast_1d(np.asarray(i, dtype=np.int64)); j = np.atleast_1d(np.asarray(j, dtype=np.int64))
        _, mmm, mm2_f = sorted_lists_of_pairs(cfg, i, j)
        _, _, mm2_b = sorted_lists_of_pairs(cfg, j, i)
        enc_mmm = one_hot_encode(self.mapping_mmm, cfg.occ[mmm], self.elements)
        enc_f = one_hot_encode(self.mapping_mm2, cfg.occ[mm2_f], self.elements)
        enc_b = one_hot_encode(self.mapping_mm2, cfg.occ[mm2_b], self.elements)
        return enc_mmm, enc_f, enc_b

    def _log_model(self, x, mig, mu_x, sigma_x, u, theta, mu_y, sigma_y):
        out = np.empty(x.shape[0], dtype=np.float64)
        for code in np.unique(mig):
            p = self.params[int(code)]
            sel = mig == code
            z = (x[sel] - p[mu_x]) / p[sigma_x]
            proj = _seq_matvec(p[u], z)                               # (E', K)
            out[sel] = _seq_dot(p[theta], proj) * float(p[sigma_y]) + float(p[mu_y])
        return out

    def get_d_ks(self, cfg, i, j):
        i = np.atleast_1d(np.asarray(i, dtype=np.int64)); j = np.atleast_1d(np.asarray(j, dtype=np.int64))
        (x_mmm, _, _), (x_f, _, _), (x_b, _, _) = self.encodes(cfg, i, j)
        mig = cfg.occ[j]
        log_d = self._log_model(x_mmm, mig, "mu_x_mmm", "sigma_x_mmm", "U_mmm", "theta_D", "mu_D", "sigma_D")
        log_ks = self._log_model(x_f + x_b, mig, "mu_x_mm2", "sigma_x_mm2", "U_mm2", "theta_Ks", "mu_Ks", "sigma_Ks")
        return np.exp(log_d), np.exp(log_ks)

    # ---- GetBarrierAndDiffFromLatticeIdPair (:247-276)
    def barrier_and_diff(self, cfg, i, j):
        de = self.get_de(cfg, i, j)
        d, ks = self.get_d_ks(cfg, i, j)
        return quartic_barrier(de, d, ks), de


def quartic_barrier(de, d, ks):
    """Closed form of pred/src/VacancyMigrationPredictorQuartic.cpp:266-275."""
    b = 4 * de / (d * d * d)
    a = ks / (4 * d * d)
    c = (9 * b * b - 16 * a * a * d * d) / (32 * a)
    delta = np.sqrt(np.abs(9 * b * b - 32 * a * c))
    return (3 * b + delta) * (3 * b + delta) * (3 * b * b - 16 * a * c + b * delta) / np.power(a, 3) / 2048


def _seq_dot(theta, x):
    """Sequential-order dot product like the shim Eigen (row by row), x: (E, L) -> (E,)."""
    return x @ theta


def _seq_matvec(u, x):
    return x @ u.T


class EnergyChangePredictorPairSite:
    """pred::EnergyChangePredictorPairSite (pred/src/EnergyChangePredictorPairSite.cpp)."""

    def __init__(self, coefficients, reference_config, element_codes):
        self.elements = element_set_sorted(element_codes)
        co = load_coefficients(coefficients)
        self.site_mapping = mapping_state_site(reference_config)    # :17
        self.indexer = ClusterIndexer(self.elements, DE_CLUSTER_COUNTER)
        self.base_theta = np.asarray(co["Base"]["theta"], dtype=np.float64)

    def _helper(self, sc, ec):                                       # GetDeHelper :83-94
        enc = (ec.astype(np.float64) - sc.astype(np.float64)) / self.indexer.total_bonds
        return _seq_dot(self.base_theta, enc)

    def site_counts(self, cfg, sites, new_codes):                    # GetDeFromLatticeIdSite :154-192
        sites = np.atleast_1d(np.asarray(sites, dtype=np.int64))
        new_codes = np.atleast_1d(np.asarray(new_codes, dtype=np.int64))
        lists = sorted_list_of_sites(cfg, sites)
        start = cfg.occ[lists].astype(np.int64)
        end = start.copy()
        end[lists == sites[:, None]] = new_codes
        return _count_types(self.indexer, self.site_mapping, start), _count_types(self.indexer, self.site_mapping, end)

    def de_site(self, cfg, sites, new_codes):
        sites = np.atleast_1d(np.asarray(sites, dtype=np.int64))
        new_codes = np.atleast_1d(np.asarray(new_codes, dtype=np.int64))
        sc, ec = self.site_counts(cfg, sites, new_codes)
        out = self._helper(sc, ec)
        out[cfg.occ[sites] == new_codes] = 0.0                       # :157-160
        return out

    def de_pair(self, cfg, a, b):                                    # GetDeFromLatticeIdPair :70-81
        a = np.atleast_1d(np.asarray(a, dtype=np.int64)); b = np.atleast_1d(np.asarray(b, dtype=np.int64))
        out = np.zeros(len(a), dtype=np.float64)
        ea, eb = cfg.occ[a], cfg.occ[b]
        differ = ea != eb
        nbrs = np.concatenate([a[:, None], cfg.nn[0][a], cfg.nn[1][a], cfg.nn[2][a]], axis=1)
        coupled = (nbrs == b[:, None]).any(axis=1) & differ
        plain = differ & ~coupled
        if plain.any():                                              # WithoutCoupling :136-146
            out[plain] = self.de_site(cfg, a[plain], eb[plain]) + self.de_site(cfg, b[plain], ea[plain])
        for k in np.nonzero(coupled)[0]:                             # WithCoupling :96-134
            out[k] = self._de_pair_with_coupling(cfg, int(a[k]), int(b[k]))
        return out

    def _de_pair_with_coupling(self, cfg, a, b):
        """GetClusterParametersMappingStatePairOf (EnergyUtility.cpp:582-663): every singlet / 1-3NN pair /
        label 4-7 triplet inside the pair's neighbourhood that touches a or b, counted before and after the swap."""
        ids = cfg.neighbors_set_of_pair(a, b)
        mapping = _state_mapping(cfg, ids, [a, b])
        start = cfg.occ[ids].astype(np.int64)[None, :]
        end = start.copy()
        end[0, ids == a] = cfg.occ[b]
        end[0, ids == b] = cfg.occ[a]
        sc = _count_types(self.indexer, mapping, start)
        ec = _count_types(self.indexer, mapping, end)
        return float(self._helper(sc, ec)[0])


class VacancyMigrationPredictorE0(VacancyMigrationPredictorQuartic):
    """pred::VacancyMigrationPredictorE0 (pred/src/VacancyMigrationPredictorE0.cpp): GetDe (:77-125) is the quartic
    predictor's cluster-count difference with the same normalisers; GetE0 (:127-151) is the mmm one-hot model with
    theta_e0 / mu_e0 / sigma_e0; Ea = max(0, e0 + dE / 2) (:153-159)."""

    def get_e0(self, cfg, i, j):
        i = np.atleast_1d(np.asarray(i, dtype=np.int64)); j = np.atleast_1d(np.asarray(j, dtype=np.int64))
        _, mmm, _ = sorted_lists_of_pairs(cfg, i, j)
        x_mmm, _, _ = one_hot_encode(self.mapping_mmm, cfg.occ[mmm], self.elements)
        return np.exp(self._log_model(x_mmm, cfg.occ[j], "mu_x_mmm", "sigma_x_mmm", "U_mmm", "theta_e0", "mu_e0", "sigma_e0"))

    def barrier_and_diff(self, cfg, i, j):
        de = self.get_de(cfg, i, j)
        return np.maximum(0.0, self.get_e0(cfg, i, j) + de / 2), de


class EnergyChangePredictorPair:
    """pred::EnergyChangePredictorPair (pred/src/EnergyChangePredictorPair.cpp:69-122): swap dE of a FIRST-NEIGHBOUR pair
    from the 473 state clusters of the pair (first <- element of second, second <- element of first); equal elements
    give 0 (:71-75); a pair that is not a first-neighbour pair has no cached list (`.at` throws, :83)."""

    def __init__(self, coefficients, reference_config, element_codes):
        self.elements = element_set_sorted(element_codes)
        co = load_coefficients(coefficients)
        self.mapping_state = mapping_state_pair(reference_config)   # :18
        self.indexer = ClusterIndexer(self.elements, DE_CLUSTER_COUNTER)
        self.base_theta = np.asarray(co["Base"]["theta"], dtype=np.float64)

    def de_pair(self, cfg, a, b):
        a = np.atleast_1d(np.asarray(a, dtype=np.int64)); b = np.atleast_1d(np.asarray(b, dtype=np.int64))
        if not (cfg.nn[0][a] == b[:, None]).any(axis=1).all():
            raise IndexError("EnergyChangePredictorPair: not a first-neighbour pair")
        lo, hi = np.minimum(a, b), np.maximum(a, b)                   # the list of the id-sorted pair is used (:80-83)
        state, _, _ = sorted_lists_of_pairs(cfg, lo, hi)
        start = cfg.occ[state].astype(np.int64)
        end = start.copy()
        end[state == a[:, None]] = cfg.occ[b]
        end[state == b[:, None]] = cfg.occ[a]
        sc = _count_types(self.indexer, self.mapping_state, start)
        ec = _count_types(self.indexer, self.mapping_state, end)
        out = _seq_dot(self.base_theta, (ec.astype(np.float64) - sc.astype(np.float64)) / self.indexer.total_bonds)
        out[cfg.occ[a] == cfg.occ[b]] = 0.0
        return out


class EnergyChangePredictorSite(EnergyChangePredictorPairSite):
    """pred::EnergyChangePredictorSite (pred/src/EnergyChangePredictorSite.cpp:56-98): the single-site half of the
    PairSite predictor (per-site cluster lists instead of one translated mapping; same clusters, same counts)."""

    def de_pair(self, cfg, a, b):
        raise AttributeError("EnergyChangePredictorSite has no pair interface")


class EnergyPredictor:
    """pred::EnergyPredictor (pred/src/EnergyPredictor.cpp:40-96,173-177): ordered-tuple counting."""

    def __init__(self, coefficients, element_codes):
        self.elements = element_set_sorted(element_codes)
        co = load_coefficients(coefficients)
        self.indexer = ClusterIndexer(self.elements, E_CLUSTER_COUNTER)
        self.base_theta = np.asarray(co["Base"]["theta"], dtype=np.float64)

    def counts(self, cfg):
        ix = self.indexer
        occ = cfg.occ.astype(np.int64)
        n = cfg.num_sites
        nn1, nn2, nn3 = cfg.nn
        counts = np.zeros(ix.size, dtype=np.int64)
        e1 = occ
        np.add.at(counts, ix.index(0, e1[:, None]), 1)
        for shell, lst in ((1, nn1), (2, nn2), (3, nn3)):
            np.add.at(counts, ix.index(shell, np.stack([np.repeat(e1, lst.shape[1]), occ[lst].ravel()], axis=-1)), 1)
        # triplets: site1 -> 1NN site2 -> (1NN | 2NN) site3, classified by the shell of site3 around site1
        label_of = np.full((n, 0), 0)
        for q in range(12):
            s2 = nn1[:, q]
            for lst3, options in ((nn1, ((nn1, 4), (nn2, 5), (nn3, 6))), (nn2, ((nn3, 7),))):
                s3 = lst3[s2]                                            # (N, k)
                lab = np.zeros(s3.shape, dtype=np.int64)
                for shell_list, label in options:
                    hit = (s3[:, :, None] == shell_list[:, None, :]).any(axis=2) & (lab == 0)
                    lab[hit] = label
                for label in {o[1] for o in options}:
                    sel = lab == label
                    if sel.any():
                        rows = np.nonzero(sel)
                        el = np.stack([e1[rows[0]], occ[s2[rows[0]]], occ[s3[sel]]], axis=-1)
                        np.add.at(counts, ix.index(label, el), 1)
        return counts

    def encode(self, cfg):
        return self.counts(cfg).astype(np.float64) / self.indexer.total_bonds

    def energy(self, cfg):
        return float(self.encode(cfg) @ self.base_theta)


# ------------------------------------------------------------------------------------------ helpers
def rate_correction_factor(c_vac, c_solute, temperature):
    """pred::RateCorrector (pred/include/RateCorrector.hpp:17-24)."""
    correct = 1.64 * math.exp(-(0.66 / K_BOLTZMANN / temperature - 0.7))
    return c_vac / correct / (1 - 13 * c_solute)


class TimeTemperatureInterpolator:
    """pred::TimeTemperatureInterpolator (pred/src/TimeTemperatureInterpolator.cpp:10-65)."""

    def __init__(self, path=None, points=None):
        pts = []
        if path:
            with open(path) as f:
                lines = f.read().split("\n")
            k = 0
            while k < len(lines) and not lines[k].startswith("0"):      # skip until a line starting with '0' (:19)
                k += 1
            for line in lines[k:]:
                parts = line.split()
                if len(parts) < 2:
                    break
                pts.append((float(parts[0]), float(parts[1])))
        if points:
            pts = list(points)
        self.points = sorted(pts)

    def temperature(self, time):
        xs = [p[0] for p in self.points]
        import bisect
        k = bisect.bisect_left(xs, time)
        if k == len(xs):
            return self.points[-1][1]
        if k == 0 and time <= xs[0]:
            return self.points[0][1]
        (x0, y0), (x1, y1) = self.points[k - 1], self.points[k]
        return y0 + ((time - x0) / (x1 - x0)) * (y1 - y0)


# ------------------------------------------------------------------------------------------ drivers
def kmc_first(cfg, predictor, temperature, u1, u2, tt=None, rate_corrector=False, solvent_code=1):
    """mc::KineticMcFirstOmp::Simulate with host-supplied uniforms (replay): per step, in the order of
    KineticMcFirstAbstract::OneStepSimulation (mc/src/KineticMcAbstract.cpp:140-182):
    T(t) -> 12 events (KineticMcFirstOmp.cpp:52-78) -> dt = -ln(u1)/sum/1e13*corr (:79-82) -> select first slot with
    cumulative p >= u2 (KineticMcAbstract.cpp:106-123) -> time, energy, jump.  cfg is modified in place."""
    vac = int(np.nonzero(cfg.occ == 0)[0][0])
    c_vac = float(np.mean(cfg.occ == 0))
    c_sol = float(np.mean((cfg.occ != ELEMENT_CODES["Al"]) & (cfg.occ != 0)))   # KineticMcAbstract.cpp:35
    time, energy = 0.0, 0.0
    beta = 1.0 / K_BOLTZMANN / temperature
    trace = {k: [] for k in ("from", "to", "slot", "dt", "time", "energy", "Ea", "dE", "temperature", "total_rate")}
    for s in range(len(u1)):
        if tt is not None:
            temperature = tt.temperature(time)
            beta = 1.0 / K_BOLTZMANN / temperature
        nbrs = cfg.nn[0][vac]
        ea, de = predictor.barrier_and_diff(cfg, np.full(12, vac), nbrs)
        rates = np.exp(-ea * beta)                                       # JumpEvent.cpp:13
        total = 0.0
        for r in rates:
            total += r
        cumulative, acc = [], 0.0
        for r in rates:
            acc += r / total
            cumulative.append(acc)
        corr = rate_correction_factor(c_vac, c_sol, temperature) if rate_corrector else 1.0
        dt = -math.log(u1[s]) / total / K_PREFACTOR * corr
        slot = next((q for q, c in enumerate(cumulative) if not c < u2[s]), 11)
        to = int(nbrs[slot])
        time += dt
        energy += float(de[slot])
        cfg.lattice_jump(vac, to)
        for key, val in (("from", vac), ("to", to), ("slot", slot), ("dt", dt), ("time", time), ("energy", energy),
                         ("Ea", float(ea[slot])), ("dE", float(de[slot])), ("temperature", temperature),
                         ("total_rate", total)):
            trace[key].append(val)
        vac = to
    return {k: np.asarray(v) for k, v in trace.items()}


def kmc_chain(cfg, predictor, temperature, u, tt=None, rate_corrector=False):
    """mc::KineticMcChainOmpi::Simulate (second-order KMC) with host-supplied uniforms (one per step).  Per step
    (KineticMcAbstract.cpp:140-182,260-263; KineticMcChainOmpi.cpp:56-152), k = vacancy site, rank r handles the r-th first
    neighbour i (ascending lattice ids):
      BuildEventList: jump k->i, evaluate the 12 events i->l in that state, total_rate_i = sum exp(-Ea_il beta); the event
        k->i is the REVERSE of i->k (barrier Ea_ik - dE_ik, change -dE_ik; JumpEvent.cpp:55-58); jump back;
        total_rate_k = sum_r exp(-Ea_ki beta), p_ki = rate/total.
      CalculateTime: p_ik = backward rate of (k->i) / total_rate_i; the eight sums of MpiData in rank order; t_2 and the
        second-order probabilities (the branch depends on whether i is the site the vacancy came from).
      SelectEvent: first slot with cumulative p >= u (last if none).  previous_j starts as first neighbour 0 of the initial
        vacancy site (KineticMcAbstract.cpp:255).  cfg is modified in place."""
    vac = int(np.nonzero(cfg.occ == 0)[0][0])
    c_vac = float(np.mean(cfg.occ == 0))
    c_sol = float(np.mean((cfg.occ != ELEMENT_CODES["Al"]) & (cfg.occ != 0)))
    previous_j = int(cfg.nn[0][vac][0])
    time, energy = 0.0, 0.0
    beta = 1.0 / K_BOLTZMANN / temperature
    trace = {k: [] for k in ("from", "to", "slot", "dt", "time", "energy", "Ea", "dE", "temperature", "total_rate")}
    for s in range(len(u)):
        if tt is not None:
            temperature = tt.temperature(time)
            beta = 1.0 / K_BOLTZMANN / temperature
        k = vac
        nbrs = [int(v) for v in cfg.nn[0][k]]
        barrier_ki, de_ki, fwd, bwd, total_i = [], [], [], [], []
        for i in nbrs:
            cfg.lattice_jump(k, i)
            ls = cfg.nn[0][i]
            ea, de = predictor.barrier_and_diff(cfg, np.full(12, i), ls)
            cfg.lattice_jump(i, k)
            tot = 0.0
            for q in range(12):
                tot += math.exp(-float(ea[q]) * beta)
            q = [int(v) for v in ls].index(k)
            b_ki, d_ki = float(ea[q]) - float(de[q]), -float(de[q])            # GetReverseJumpEvent
            barrier_ki.append(b_ki); de_ki.append(d_ki)
            fwd.append(math.exp(-b_ki * beta))                                 # forward_rate_
            bwd.append(math.exp((d_ki - b_ki) * beta))                         # GetBackwardRate
            total_i.append(tot)
        total_k = 0.0
        for r in fwd:
            total_k += r
        t_1 = 1.0 / total_k / K_PREFACTOR
        sums = [0.0] * 8
        per_rank = []
        for r in range(12):
            p_ki = fwd[r] / total_k
            p_ik = bwd[r] / total_i[r]
            bb, b = p_ki * p_ik, p_ki * (1 - p_ik)
            prev = nbrs[r] == previous_j
            t_i = 1.0 / total_i[r] / K_PREFACTOR
            contrib = (bb, b, 0.0 if prev else bb, 0.0 if prev else b, b if prev else 0.0, p_ki if prev else 0.0,
                       (t_1 + t_i) * bb, 0.0 if prev else (t_1 + t_i) * bb)
            for q in range(8):
                sums[q] = contrib[q] if r == 0 else sums[q] + contrib[q]
            per_rank.append((b, prev))
        beta_bar_k, beta_k, gamma_bar_k_j, gamma_k_j, beta_k_j, alpha_k_j, ts_num, ts_j_num = sums
        ts = ts_num / beta_bar_k
        ts_j = ts_j_num / gamma_bar_k_j
        inv = 1.0 / (1.0 - alpha_k_j)
        t_2 = inv * (gamma_k_j * t_1 + gamma_bar_k_j * (ts_j + t_1 + beta_bar_k / beta_k * ts))
        cumulative, acc = [], 0.0
        for b, prev in per_rank:
            acc += inv * (gamma_bar_k_j / beta_k) * beta_k_j if prev else inv * (1 + gamma_bar_k_j / beta_k) * b
            cumulative.append(acc)
        corr = rate_correction_factor(c_vac, c_sol, temperature) if rate_corrector else 1.0
        dt = t_2 * corr
        slot = next((q for q, cp in enumerate(cumulative) if not cp < u[s]), 11)
        to = nbrs[slot]
        time += dt
        energy += de_ki[slot]
        cfg.lattice_jump(k, to)
        for key, val in (("from", k), ("to", to), ("slot", slot), ("dt", dt), ("time", time), ("energy", energy),
                         ("Ea", barrier_ki[slot]), ("dE", de_ki[slot]), ("temperature", temperature), ("total_rate", total_k)):
            trace[key].append(val)
        previous_j = k
        vac = to
    return {k: np.asarray(v) for k, v in trace.items()}


def metropolis_trials(cfg, predictor, a, b, u, temperature=None, sa_schedule=None):
    """CanonicalMcSerial::Simulate (mc/src/CanonicalMcSerial.cpp:40-51) / SimulatedAnnealing::Simulate
    (mc/src/SimulatedAnnealing.cpp:168-185) replayed on host-supplied trial pairs and uniforms: accept if dE < 0,
    else if u < exp(-dE*beta) (CanonicalMcAbstract.cpp:86-101).  With sa_schedule (a SaSchedule) the temperature
    follows SimulatedAnnealing::UpdateTemperature (:99-139)."""
    energy = 0.0
    out = {k: [] for k in ("dE", "accepted", "energy_before", "temperature_before")}
    for k in range(len(a)):
        t_now = sa_schedule.temperature if sa_schedule else temperature
        beta = 1.0 / K_BOLTZMANN / max(t_now, 1e-12) if sa_schedule else 1.0 / K_BOLTZMANN / t_now
        de = float(predictor.de_pair(cfg, [a[k]], [b[k]])[0])
        out["energy_before"].append(energy)
        out["temperature_before"].append(t_now)
        accepted = de < 0 or u[k] < math.exp(-de * beta)
        if accepted:
            cfg.lattice_jump(int(a[k]), int(b[k]))
            energy += de
        if sa_schedule:
            sa_schedule.update(accepted, energy, k)
        out["dE"].append(de)
        out["accepted"].append(accepted)
    return {k: np.asarray(v) for k, v in out.items()}


class SaSchedule:
    """SimulatedAnnealing::UpdateTemperature (mc/src/SimulatedAnnealing.cpp:99-139; constants
    mc/include/SimulatedAnnealing.h:42-68)."""

    def __init__(self, initial_temperature, maximum_steps, initial_energy=0.0):
        self.temperature = float(initial_temperature)
        self.max_steps = int(maximum_steps)
        self.reheat_trigger = max(1, int(self.max_steps * 0.05))
        self.reheat_cooldown = max(1, int(self.max_steps * 0.10))
        self.window = max(1, int(self.max_steps * 0.001))
        self.trials = 0
        self.accepts = 0
        self.recent_best = initial_energy
        self.last_improvement = 0
        self.last_reheat = 0
        self.reheats = 0

    def update(self, accepted, energy, step):
        self.trials += 1
        if accepted:
            self.accepts += 1
            if energy < self.recent_best - K_EPSILON:
                self.recent_best = energy
                self.last_improvement = step
        if self.trials >= self.window:
            if self.accepts / self.trials > 0.50:
                self.temperature *= 0.99
            self.trials = 0
            self.accepts = 0
        acc_est = self.accepts / self.trials if self.trials > 0 else 1.0
        if (self.reheats < 5 and step - self.last_improvement >= self.reheat_trigger
                and step - self.last_reheat >= self.reheat_cooldown and acc_est < 0.05):
            self.temperature *= 1.10
            self.last_improvement = step
            self.last_reheat = step
            self.recent_best = energy
            self.reheats += 1
        self.temperature *= math.exp(-3.0 / max(1, self.max_steps))

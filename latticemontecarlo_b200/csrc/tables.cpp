// tables.cpp -- see tables.h.  All geometry is exact integer arithmetic in half-lattice-constant units.
#include "tables.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <tuple>

namespace lmc {
namespace {

int bond_label(Int3 a, Int3 b) {
  // Config::FindDistanceLabelBetweenLattice (cfg/src/Config.cpp:354-371) in integer form: squared distance in
  // half-units is 2 / 4 / 6 for the first / second / third shell (cut-offs 3.5 / 4.8 / 5.3 A, Constants.hpp:8-10)
  switch (norm2(a - b)) {
    case 2: return 1;
    case 4: return 2;
    case 6: return 3;
    default: return -1;
  }
}

int triplet_label(int l01, int l12, int l20) {
  // GetLabel (pred/src/EnergyUtility.cpp:345-379)
  int b[3] = {l01, l12, l20};
  std::sort(b, b + 3);
  if (b[0] == 1 && b[1] == 1 && b[2] == 1) return 4;
  if (b[0] == 1 && b[1] == 1 && b[2] == 2) return 5;
  if (b[0] == 1 && b[1] == 1 && b[2] == 3) return 6;
  if (b[0] == 1 && b[1] == 2 && b[2] == 3) return 7;
  if (b[0] == 1 && b[1] == 3 && b[2] == 3) return 8;
  if (b[0] == 2 && b[1] == 3 && b[2] == 3) return 9;
  if (b[0] == 3 && b[1] == 3 && b[2] == 3) return 10;
  return -1;
}

// GetClusterParametersMappingStatePair / StateSite (pred/src/EnergyUtility.cpp:393-581): same loop nest,
// `centre` marks the positions of the jump pair (or the single site).
template <size_t N>
std::vector<Cluster> build_state_mapping(const std::array<Int3, N> &sites, const std::vector<int> &centre_pos) {
  auto is_centre = [&](int p) { return std::find(centre_pos.begin(), centre_pos.end(), p) != centre_pos.end(); };
  std::vector<std::vector<Cluster>> by_label(8);
  const int n = static_cast<int>(N);
  for (int p1 = 0; p1 < n; ++p1) {
    if (is_centre(p1)) by_label[0].push_back(Cluster{0, 1, {static_cast<int16_t>(p1), -1, -1}, -1, false});
    for (int p2 = 0; p2 < p1; ++p2) {
      const int l12 = bond_label(sites[p1], sites[p2]);
      if (is_centre(p1) || is_centre(p2)) {
        if (l12 >= 1 && l12 <= 3)
          by_label[l12].push_back(Cluster{static_cast<int8_t>(l12), 2, {static_cast<int16_t>(p1), static_cast<int16_t>(p2), -1}, -1, false});
        else
          continue;  // `default: continue` of the reference: no triplets through a non-bonded centre pair
      }
      for (int p3 = 0; p3 < p2; ++p3) {
        if (!(is_centre(p1) || is_centre(p2) || is_centre(p3))) continue;
        const int t = triplet_label(l12, bond_label(sites[p2], sites[p3]), bond_label(sites[p3], sites[p1]));
        if (t >= 4 && t <= 7)
          by_label[t].push_back(Cluster{static_cast<int8_t>(t), 3,
                                        {static_cast<int16_t>(p1), static_cast<int16_t>(p2), static_cast<int16_t>(p3)}, -1, false});
      }
    }
  }
  std::vector<Cluster> out;
  for (auto &v : by_label) out.insert(out.end(), v.begin(), v.end());
  return out;
}

using Key2 = std::pair<int, int>;

// GetAverageClusterParametersMappingMMM / MM2 (pred/src/EnergyUtility.cpp:169-259) with the grouping helper
// (:106-167).  `gkey[a]` is the integer form of GroupCompareMMM/MM2 (cfg/include/LatticeCluster.hpp:77-103).
void build_average_mapping(const std::vector<Int3> &list_sites, const std::vector<Key2> &gkey, std::vector<Cluster> &out,
                           int &n_groups, std::vector<int8_t> &g_arity, std::vector<bool> &g_sym,
                           std::vector<int32_t> &g_size) {
  const int n = static_cast<int>(list_sites.size());
  struct Item { std::vector<Key2> key; Cluster c; };
  std::vector<std::vector<Item>> family(4);
  for (int a = 0; a < n; ++a) {
    family[0].push_back({{gkey[a]}, Cluster{0, 1, {static_cast<int16_t>(a), -1, -1}, -1, false}});
    for (int b = a + 1; b < n; ++b) {
      const int l = bond_label(list_sites[a], list_sites[b]);
      if (l < 1) continue;
      // members sorted by PositionCompare* == ascending list position (the list itself is sorted that way)
      family[l].push_back({{gkey[a], gkey[b]},
                           Cluster{static_cast<int8_t>(l), 2, {static_cast<int16_t>(a), static_cast<int16_t>(b), -1}, -1, gkey[a] == gkey[b]}});
    }
  }
  n_groups = 0;
  out.clear();
  for (auto &fam : family) {
    std::stable_sort(fam.begin(), fam.end(), [](const Item &l, const Item &r) { return l.key < r.key; });
    for (size_t s = 0; s < fam.size();) {
      size_t e = s;
      while (e < fam.size() && fam[e].key == fam[s].key) ++e;
      for (size_t q = s; q < e; ++q) {
        Cluster c = fam[q].c;
        c.group = static_cast<int16_t>(n_groups);
        out.push_back(c);
      }
      g_arity.push_back(fam[s].c.arity);
      g_sym.push_back(fam[s].c.symmetric);
      g_size.push_back(static_cast<int32_t>(e - s));
      ++n_groups;
      s = e;
    }
  }
}

Geometry build_geometry() {
  Geometry g{};
  // --- shells
  {
    int n1 = 0, n2 = 0, n3 = 0;
    for (int x = -2; x <= 2; ++x)
      for (int y = -2; y <= 2; ++y)
        for (int z = -2; z <= 2; ++z) {
          const int r2 = x * x + y * y + z * z;
          if (((x + y + z) & 1) != 0) continue;
          if (r2 == 2) g.nn1[n1++] = {x, y, z};
          if (r2 == 4) g.nn2[n2++] = {x, y, z};
          if (r2 == 6) g.nn3[n3++] = {x, y, z};
        }
    if (n1 != 12 || n2 != 6 || n3 != 24) throw std::logic_error("shell enumeration failed");
  }
  std::vector<Int3> shell_all{{0, 0, 0}};
  shell_all.insert(shell_all.end(), g.nn1.begin(), g.nn1.end());
  shell_all.insert(shell_all.end(), g.nn2.begin(), g.nn2.end());
  shell_all.insert(shell_all.end(), g.nn3.begin(), g.nn3.end());

  // --- jump frames: x along d, y along a first neighbour perpendicular to d (Config.cpp:247-264)
  for (int k = 0; k < 12; ++k) {
    int found = 0;
    for (const auto &v : g.nn1)
      if (dot(v, g.nn1[k]) == 0 && found++ == 0) g.frame_p[k] = v;
    if (found != 2) throw std::logic_error("expected exactly two perpendicular first neighbours");
  }
  auto pair_sites = [&](Int3 d) {
    std::vector<Int3> s;
    for (const auto &o : shell_all) {
      for (const Int3 r : {o, d + o})
        if (std::find(s.begin(), s.end(), r) == s.end()) s.push_back(r);
    }
    return s;
  };
  // --- canonical keys from direction 0, frame flag 0
  {
    const Int3 d = g.nn1[0], p = g.frame_p[0], c = cross(d, p);
    auto sites = pair_sites(d);
    if (sites.size() != kPairSites) throw std::logic_error("pair neighbourhood must hold 60 sites");
    std::vector<Int3> keys;
    for (const auto &r : sites) {
      const Int3 q{2 * r.x - d.x, 2 * r.y - d.y, 2 * r.z - d.z};
      keys.push_back({dot(q, d), dot(q, p), dot(q, c)});
    }
    // PositionCompareState (LatticeCluster.hpp:62-75): lexicographic (x, y, z)
    std::sort(keys.begin(), keys.end(), [](Int3 a, Int3 b) { return std::tie(a.x, a.y, a.z) < std::tie(b.x, b.y, b.z); });
    std::copy(keys.begin(), keys.end(), g.pair_keys.begin());
  }
  // --- offsets of the 60 canonical positions for every direction / frame flag
  for (int k = 0; k < 12; ++k)
    for (int s = 0; s < 2; ++s) {
      const Int3 d = g.nn1[k];
      const Int3 p = s == 0 ? g.frame_p[k] : Int3{-g.frame_p[k].x, -g.frame_p[k].y, -g.frame_p[k].z};
      const Int3 c = cross(d, p);
      auto expect = pair_sites(d);
      for (int t = 0; t < kPairSites; ++t) {
        const Int3 key = g.pair_keys[t];
        // q = (alpha/2) d + (beta/2) p + (gamma/4) c  (|d|^2 = |p|^2 = 2, |c|^2 = 4);  r = (q + d) / 2
        const Int3 q4{2 * key.x * d.x + 2 * key.y * p.x + key.z * c.x, 2 * key.x * d.y + 2 * key.y * p.y + key.z * c.y,
                      2 * key.x * d.z + 2 * key.y * p.z + key.z * c.z};
        if (q4.x % 4 || q4.y % 4 || q4.z % 4) throw std::logic_error("non-integer quarter coordinate");
        const Int3 q{q4.x / 4, q4.y / 4, q4.z / 4};
        if ((q.x + d.x) % 2 || (q.y + d.y) % 2 || (q.z + d.z) % 2) throw std::logic_error("non-integer site offset");
        const Int3 r{(q.x + d.x) / 2, (q.y + d.y) / 2, (q.z + d.z) / 2};
        if (std::find(expect.begin(), expect.end(), r) == expect.end()) throw std::logic_error("offset outside the neighbourhood");
        g.pair_offsets[k][s][t] = r;
      }
    }
  // --- positions of the jump pair, env indexing
  g.pair_first_pos = g.pair_second_pos = -1;
  for (int t = 0; t < kPairSites; ++t) {
    if (g.pair_offsets[0][0][t] == Int3{0, 0, 0}) g.pair_first_pos = t;
    if (g.pair_offsets[0][0][t] == g.nn1[0]) g.pair_second_pos = t;
  }
  {
    int e = 0;
    for (int t = 0; t < kPairSites; ++t)
      g.env_of_state[t] = t == g.pair_first_pos ? -1 : (t == g.pair_second_pos ? -2 : e++);
  }
  std::vector<int> state_pos_of_env;
  for (int t = 0; t < kPairSites; ++t)
    if (g.env_of_state[t] >= 0) state_pos_of_env.push_back(t);
  // --- symmetric orders (PositionCompareMMM / MM2, LatticeCluster.hpp:105-154).  In quarter units the centred
  // position is q = alpha d/2 + beta p/2 + gamma c/4, so |q|^2 = (2 alpha^2 + 2 beta^2 + gamma^2) / 4
  auto norm_key = [&](int env) {
    const Int3 k = g.pair_keys[state_pos_of_env[env]];
    return 2 * k.x * k.x + 2 * k.y * k.y + k.z * k.z;
  };
  auto key_of = [&](int env) { return g.pair_keys[state_pos_of_env[env]]; };
  {
    std::vector<int> order(kPairEnv);
    for (int e = 0; e < kPairEnv; ++e) order[e] = e;
    auto by_mmm = order, by_mm2 = order;
    std::sort(by_mmm.begin(), by_mmm.end(), [&](int a, int b) {
      const Int3 ka = key_of(a), kb = key_of(b);
      return std::make_tuple(norm_key(a), std::abs(ka.x), ka.x, ka.y, ka.z) < std::make_tuple(norm_key(b), std::abs(kb.x), kb.x, kb.y, kb.z);
    });
    std::sort(by_mm2.begin(), by_mm2.end(), [&](int a, int b) {
      const Int3 ka = key_of(a), kb = key_of(b);
      return std::make_tuple(norm_key(a), ka.x, ka.y, ka.z) < std::make_tuple(norm_key(b), kb.x, kb.y, kb.z);
    });
    std::copy(by_mmm.begin(), by_mmm.end(), g.env_of_mmm.begin());
    std::copy(by_mm2.begin(), by_mm2.end(), g.env_of_mm2.begin());
    // backward list (j -> i): same centre, x axis reversed; y axis equal (variant 1) or opposite (variant 0) to the
    // forward frame's, z = x cross y accordingly.
    for (int v = 0; v < 2; ++v) {
      auto by_b = order;
      auto bkey = [&](int e) {
        const Int3 k = key_of(e);
        return v == 1 ? Int3{-k.x, k.y, -k.z} : Int3{-k.x, -k.y, k.z};
      };
      std::sort(by_b.begin(), by_b.end(), [&](int a, int b) {
        const Int3 ka = bkey(a), kb = bkey(b);
        return std::make_tuple(norm_key(a), ka.x, ka.y, ka.z) < std::make_tuple(norm_key(b), kb.x, kb.y, kb.z);
      });
      std::copy(by_b.begin(), by_b.end(), g.env_of_mm2_backward[v].begin());
    }
  }
  // --- site neighbourhood: GetSortedLatticeVectorStateOfSite (EnergyUtility.cpp:288-313)
  {
    auto s = shell_all;
    std::sort(s.begin(), s.end(), [](Int3 a, Int3 b) { return std::tie(a.x, a.y, a.z) < std::tie(b.x, b.y, b.z); });
    std::copy(s.begin(), s.end(), g.site_offsets.begin());
    g.site_centre_pos = -1;
    for (int t = 0; t < kSiteSites; ++t)
      if (g.site_offsets[t] == Int3{0, 0, 0}) g.site_centre_pos = t;
  }
  // --- cluster mappings
  g.state_pair = build_state_mapping(g.pair_offsets[0][0], {g.pair_first_pos, g.pair_second_pos});
  g.state_site = build_state_mapping(g.site_offsets, {g.site_centre_pos});
  for (int which = 0; which < 2; ++which) {
    const auto &env_of = which == 0 ? g.env_of_mmm : g.env_of_mm2;
    std::vector<Int3> sites;
    std::vector<Key2> gkey;
    for (int a = 0; a < kPairEnv; ++a) {
      const int env = env_of[a];
      sites.push_back(g.pair_offsets[0][0][state_pos_of_env[env]]);
      const Int3 k = key_of(env);
      gkey.push_back({norm_key(env), which == 0 ? std::abs(k.x) : k.x});
    }
    if (which == 0)
      build_average_mapping(sites, gkey, g.mmm, g.n_groups_mmm, g.group_arity_mmm, g.group_sym_mmm, g.group_size_mmm);
    else
      build_average_mapping(sites, gkey, g.mm2, g.n_groups_mm2, g.group_arity_mm2, g.group_sym_mm2, g.group_size_mm2);
  }
  // --- env pair adjacency of the jump neighbourhood
  {
    for (int t = 0; t < kPairEnv; ++t) {
      g.env_pair_base[t] = static_cast<uint16_t>(g.env_pairs.size());
      g.env_pair_mask_hi[t] = 0;
      for (int u = t + 1; u < kPairEnv; ++u) {
        if (bond_label(g.pair_offsets[0][0][state_pos_of_env[t]], g.pair_offsets[0][0][state_pos_of_env[u]]) >= 1) {
          g.env_pairs.push_back({static_cast<int16_t>(t), static_cast<int16_t>(u)});
          g.env_pair_mask_hi[t] |= 1ULL << u;
        }
      }
    }
  }
  // --- site env: pairs (t<u) that form a label 4-7 triplet with the centre
  {
    std::vector<Int3> env;
    for (int t = 0; t < kSiteSites; ++t)
      if (t != g.site_centre_pos) env.push_back(g.site_offsets[t]);
    const Int3 origin{0, 0, 0};
    for (int t = 0; t < kSiteEnv; ++t) {
      g.site_env_shell[t] = static_cast<int8_t>(bond_label(origin, env[t]));
      g.site_pair_base[t] = static_cast<uint16_t>(g.site_env_pairs.size());
      g.site_pair_mask_hi[t] = 0;
      for (int u = t + 1; u < kSiteEnv; ++u) {
        const int lab = triplet_label(bond_label(origin, env[t]), bond_label(env[t], env[u]), bond_label(env[u], origin));
        if (lab >= 4 && lab <= 7) {
          g.site_env_pairs.push_back({static_cast<int16_t>(t), static_cast<int16_t>(u)});
          g.site_pair_mask_hi[t] |= 1ULL << u;
        }
      }
    }
  }
  // --- total-energy walk (EnergyPredictor.cpp:47-84): site1 -> 1NN site2 -> {1NN, 2NN} site3
  for (const auto &v1 : g.nn1) {
    for (const auto &v2 : g.nn1) {
      const int l = bond_label(Int3{0, 0, 0}, v1 + v2);
      if (l >= 1) g.energy_triplets.push_back({v1, v1 + v2, static_cast<int8_t>(3 + l)});   // labels 4, 5, 6
    }
    for (const auto &v2 : g.nn2)
      if (bond_label(Int3{0, 0, 0}, v1 + v2) == 3) g.energy_triplets.push_back({v1, v1 + v2, 7});
  }
  return g;
}

}  // namespace

const Geometry &geometry() {
  static const Geometry g = build_geometry();
  return g;
}

// ------------------------------------------------------------------------------------------------ species
namespace {
const char *kElementNames[] = {"X", "Al", "Mg", "Zn", "Cu", "Sn", "pAl", "pMg", "pZn", "pCu", "pSn"};  // Element.hpp:7
constexpr int kNumElementNames = 11;
}  // namespace

const char *element_name(int enum_code) {
  return (enum_code >= 0 && enum_code < kNumElementNames) ? kElementNames[enum_code] : "X";
}
int element_enum_from_name(const std::string &name) {
  for (int i = 0; i < kNumElementNames; ++i)
    if (name == kElementNames[i]) return i;
  return 0;  // unknown strings become X (Element.hpp:25-26)
}

Species make_species(const int32_t *enum_codes, int n, int solvent_enum) {
  Species sp;
  std::vector<int> codes;
  for (int i = 0; i < n; ++i) {
    if (enum_codes[i] <= 0 || enum_codes[i] >= kNumElementNames) throw std::invalid_argument("bad element code in element set");
    if (std::find(codes.begin(), codes.end(), enum_codes[i]) == codes.end()) codes.push_back(enum_codes[i]);
  }
  if (codes.empty() || static_cast<int>(codes.size()) > kMaxSpecies) throw std::invalid_argument("element set size must be 1..7");
  // std::set<Element> order: by GetString() (Element.hpp:44-46)
  std::sort(codes.begin(), codes.end(), [](int a, int b) { return std::string(kElementNames[a]) < std::string(kElementNames[b]); });
  sp.n = static_cast<int>(codes.size());
  sp.enum_of_code = codes;
  sp.enum_of_code.push_back(0);
  sp.code_of_enum.fill(-1);
  for (int c = 0; c <= sp.n; ++c) {
    sp.code_of_enum[sp.enum_of_code[c]] = static_cast<int8_t>(c);
    sp.names.push_back(kElementNames[sp.enum_of_code[c]]);
  }
  sp.solvent = 0;
  if (solvent_enum > 0) {
    if (sp.code_of_enum[solvent_enum] < 0) throw std::invalid_argument("solvent element is not in the element set");
    sp.solvent = sp.code_of_enum[solvent_enum];
  }
  return sp;
}

std::vector<ClusterType> cluster_types(const Species &sp) {
  // InitializeClusterHashMap (EnergyUtility.cpp:314-343) over element_set + X, then std::map order
  // (ElementCluster.hpp:40-50): by size, label, element strings.
  const int m = sp.n + 1;
  auto name = [&](int c) { return sp.names[c]; };
  auto is_x = [&](int c) { return c == sp.n; };
  auto is_pseudo = [&](int c) { return name(c)[0] == 'p'; };
  std::vector<ClusterType> types;
  auto add = [&](int label, std::vector<int> codes) {
    std::sort(codes.begin(), codes.end(), [&](int a, int b) { return name(a) < name(b); });
    ClusterType t{static_cast<int8_t>(label), static_cast<int8_t>(codes.size()), {0, 0, 0}};
    for (size_t i = 0; i < codes.size(); ++i) t.code[i] = static_cast<int8_t>(codes[i]);
    for (const auto &o : types)
      if (o.label == t.label && o.arity == t.arity && o.code[0] == t.code[0] && o.code[1] == t.code[1] && o.code[2] == t.code[2]) return;
    types.push_back(t);
  };
  for (int e1 = 0; e1 < m; ++e1) {
    add(0, {e1});
    for (int e2 = 0; e2 < m; ++e2) {
      if (is_x(e2)) continue;
      if (is_x(e1) && is_pseudo(e2)) continue;
      for (int label = 1; label <= 3; ++label) add(label, {e1, e2});
      for (int e3 = 0; e3 < m; ++e3) {
        if (is_x(e3) || is_pseudo(e3)) continue;
        for (int label = 4; label < 8; ++label) add(label, {e1, e2, e3});
      }
    }
  }
  std::sort(types.begin(), types.end(), [&](const ClusterType &a, const ClusterType &b) {
    if (a.arity != b.arity) return a.arity < b.arity;
    if (a.label != b.label) return a.label < b.label;
    for (int i = 0; i < a.arity; ++i)
      if (name(a.code[i]) != name(b.code[i])) return name(a.code[i]) < name(b.code[i]);
    return false;
  });
  return types;
}

TypeLut make_type_lut(const Species &sp, const std::vector<ClusterType> &types) {
  TypeLut lut;
  lut.m = sp.n + 1;
  const int m = lut.m;
  lut.lut.assign(static_cast<size_t>(8) * m * m * m, -1);
  for (size_t idx = 0; idx < types.size(); ++idx) {
    const auto &t = types[idx];
    int perm[3] = {0, 1, 2};
    std::vector<int> c(t.code, t.code + t.arity);
    std::sort(perm, perm + t.arity);
    do {
      int cc[3] = {0, 0, 0};
      for (int i = 0; i < t.arity; ++i) cc[i] = c[perm[i]];
      lut.lut[((t.label * m + cc[0]) * m + cc[1]) * m + cc[2]] = static_cast<int16_t>(idx);
    } while (std::next_permutation(perm, perm + t.arity));
  }
  return lut;
}

std::vector<GroupInfo> group_layout(const Geometry &g, bool mm2, int n, int *encode_length) {
  // GetOneHotParametersFromMap (EnergyUtility.cpp:743-796): per group n, n^2 or n(n+1)/2 slots
  const int ng = mm2 ? g.n_groups_mm2 : g.n_groups_mmm;
  const auto &arity = mm2 ? g.group_arity_mm2 : g.group_arity_mmm;
  const auto &sym = mm2 ? g.group_sym_mm2 : g.group_sym_mmm;
  const auto &size = mm2 ? g.group_size_mm2 : g.group_size_mmm;
  std::vector<GroupInfo> out(ng);
  int off = 0;
  for (int q = 0; q < ng; ++q) {
    const int len = sym[q] ? n * (n + 1) / 2 : (arity[q] == 1 ? n : n * n);
    out[q] = GroupInfo{size[q], off, len, arity[q], static_cast<bool>(sym[q])};
    off += len;
  }
  if (encode_length) *encode_length = off;
  return out;
}

// ------------------------------------------------------------------------------------------------ JSON
namespace {
// Minimal JSON reader for the coefficient file: objects, arrays, numbers, strings, true/false/null.
struct JsonValue {
  enum Kind { kNull, kNumber, kString, kArray, kObject } kind{kNull};
  double number{0};
  std::string str;
  std::vector<JsonValue> array;
  std::vector<std::pair<std::string, JsonValue>> object;
  const JsonValue *find(const std::string &key) const {
    for (const auto &kv : object)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  const JsonValue &at(const std::string &key) const {
    const JsonValue *v = find(key);
    if (!v) throw std::out_of_range("key '" + key + "' not found");   // nlohmann::json::at throws out_of_range
    return *v;
  }
};

class JsonParser {
 public:
  explicit JsonParser(const std::string &text) : s_(text) {}
  JsonValue parse() {
    JsonValue v = value();
    skip();
    if (pos_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string &s_;
  size_t pos_{0};
  [[noreturn]] void fail(const char *what) const {
    throw std::runtime_error(std::string("JSON parse error: ") + what + " at byte " + std::to_string(pos_));
  }
  void skip() {
    while (pos_ < s_.size() && (s_[pos_] == ' ' || s_[pos_] == '\n' || s_[pos_] == '\t' || s_[pos_] == '\r')) ++pos_;
  }
  JsonValue value() {
    skip();
    if (pos_ >= s_.size()) fail("unexpected end");
    const char c = s_[pos_];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') {
      JsonValue v;
      v.kind = JsonValue::kString;
      v.str = string();
      return v;
    }
    if (s_.compare(pos_, 4, "true") == 0) { pos_ += 4; JsonValue v; v.kind = JsonValue::kNumber; v.number = 1; return v; }
    if (s_.compare(pos_, 5, "false") == 0) { pos_ += 5; JsonValue v; v.kind = JsonValue::kNumber; v.number = 0; return v; }
    if (s_.compare(pos_, 4, "null") == 0) { pos_ += 4; return JsonValue{}; }
    return number();
  }
  JsonValue number() {
    const char *begin = s_.c_str() + pos_;
    char *end = nullptr;
    const double d = std::strtod(begin, &end);
    if (end == begin) fail("bad number");
    pos_ += static_cast<size_t>(end - begin);
    JsonValue v;
    v.kind = JsonValue::kNumber;
    v.number = d;
    return v;
  }
  std::string string() {
    ++pos_;
    std::string out;
    while (pos_ < s_.size() && s_[pos_] != '"') {
      if (s_[pos_] == '\\' && pos_ + 1 < s_.size()) ++pos_;
      out.push_back(s_[pos_++]);
    }
    if (pos_ >= s_.size()) fail("unterminated string");
    ++pos_;
    return out;
  }
  JsonValue array() {
    JsonValue v;
    v.kind = JsonValue::kArray;
    ++pos_;
    skip();
    if (pos_ < s_.size() && s_[pos_] == ']') { ++pos_; return v; }
    for (;;) {
      v.array.push_back(value());
      skip();
      if (pos_ >= s_.size()) fail("unterminated array");
      if (s_[pos_] == ',') { ++pos_; continue; }
      if (s_[pos_] == ']') { ++pos_; return v; }
      fail("expected , or ]");
    }
  }
  JsonValue object() {
    JsonValue v;
    v.kind = JsonValue::kObject;
    ++pos_;
    skip();
    if (pos_ < s_.size() && s_[pos_] == '}') { ++pos_; return v; }
    for (;;) {
      skip();
      if (pos_ >= s_.size() || s_[pos_] != '"') fail("expected key");
      std::string key = string();
      skip();
      if (pos_ >= s_.size() || s_[pos_] != ':') fail("expected :");
      ++pos_;
      v.object.emplace_back(std::move(key), value());
      skip();
      if (pos_ >= s_.size()) fail("unterminated object");
      if (s_[pos_] == ',') { ++pos_; continue; }
      if (s_[pos_] == '}') { ++pos_; return v; }
      fail("expected , or }");
    }
  }
};

std::vector<double> to_vector(const JsonValue &v) {
  if (v.kind != JsonValue::kArray) throw std::runtime_error("JSON: expected an array of numbers");
  std::vector<double> out;
  out.reserve(v.array.size());
  for (const auto &e : v.array) {
    if (e.kind != JsonValue::kNumber) throw std::runtime_error("JSON: expected a number");
    out.push_back(e.number);
  }
  return out;
}
std::vector<double> to_matrix(const JsonValue &v, int *rows) {
  if (v.kind != JsonValue::kArray) throw std::runtime_error("JSON: expected an array of arrays");
  std::vector<double> out;
  *rows = static_cast<int>(v.array.size());
  size_t cols = 0;
  for (const auto &row : v.array) {
    auto r = to_vector(row);
    if (cols == 0) cols = r.size();
    if (r.size() != cols) throw std::runtime_error("JSON: ragged matrix");
    out.insert(out.end(), r.begin(), r.end());
  }
  return out;
}
double to_number(const JsonValue &v) {
  if (v.kind != JsonValue::kNumber) throw std::runtime_error("JSON: expected a number");
  return v.number;
}
}  // namespace

Coefficients parse_coefficients_json(const std::string &path) {
  std::ifstream ifs(path);
  if (!ifs) throw std::runtime_error("Cannot open " + path);   // same message as the reference predictors
  std::stringstream ss;
  ss << ifs.rdbuf();
  const std::string text = ss.str();
  const JsonValue root = JsonParser(text).parse();
  if (root.kind != JsonValue::kObject) throw std::runtime_error("JSON: top level must be an object");
  Coefficients co;
  for (const auto &kv : root.object) {
    if (kv.first == "Base") {
      co.base_theta = to_vector(kv.second.at("theta"));
      continue;
    }
    // every other top-level key is an element block.  Quartic files (VacancyMigrationPredictorQuartic.cpp:44-62) carry
    // the mmm and mm2 blocks with theta_D / theta_Ks, E0 files (VacancyMigrationPredictorE0.cpp:30-37) the mmm block with
    // theta_e0 / mu_e0 / sigma_e0; a block may carry both.  Blocks with neither are ignored here.
    if (!kv.second.find("mu_x_mmm") || !kv.second.find("U_mmm")) continue;
    const bool quartic = kv.second.find("U_mm2") != nullptr, e0 = kv.second.find("theta_e0") != nullptr;
    if (!quartic && !e0) continue;
    ElementCoefficients ec;
    ec.mu_x_mmm = to_vector(kv.second.at("mu_x_mmm"));
    ec.sigma_x_mmm = to_vector(kv.second.at("sigma_x_mmm"));
    ec.U_mmm = to_matrix(kv.second.at("U_mmm"), &ec.k_mmm);
    if (quartic) {
      ec.mu_x_mm2 = to_vector(kv.second.at("mu_x_mm2"));
      ec.sigma_x_mm2 = to_vector(kv.second.at("sigma_x_mm2"));
      ec.U_mm2 = to_matrix(kv.second.at("U_mm2"), &ec.k_mm2);
      ec.theta_D = to_vector(kv.second.at("theta_D"));
      ec.theta_Ks = to_vector(kv.second.at("theta_Ks"));
      ec.mu_D = to_number(kv.second.at("mu_D"));
      ec.mu_Ks = to_number(kv.second.at("mu_Ks"));
      ec.sigma_D = to_number(kv.second.at("sigma_D"));
      ec.sigma_Ks = to_number(kv.second.at("sigma_Ks"));
      ec.present = true;
    }
    if (e0) {
      ec.theta_e0 = to_vector(kv.second.at("theta_e0"));
      ec.mu_e0 = to_number(kv.second.at("mu_e0"));
      ec.sigma_e0 = to_number(kv.second.at("sigma_e0"));
      ec.present_e0 = true;
    }
    co.element[element_enum_from_name(kv.first)] = std::move(ec);
  }
  return co;
}

// ------------------------------------------------------------------------------------------------ contraction
namespace {
// normalisers by label: pred/src/VacancyMigrationPredictorQuartic.cpp:14 == EnergyChangePredictorPairSite.cpp:11
const double kDeClusterCounter[11] = {256, 1536, 768, 3072, 2048, 3072, 6144, 6144, 6144, 6144, 2048};
// pred/src/EnergyPredictor.cpp:8
const double kEnergyClusterCounter[11] = {256, 3072, 1536, 6144, 12288, 6144, 12288, 6144, 12288, 12288, 12288};
const double kNaN = std::numeric_limits<double>::quiet_NaN();

int tri_index(int a, int b, int n) {   // upper-triangular row-major (GetOneHotEncodeHashmap, EnergyUtility.cpp:31-40)
  if (a > b) std::swap(a, b);
  return a * n - a * (a - 1) / 2 + (b - a);
}

// theta[idx] / normaliser[label] by (label, codes), NaN when the reference has no such type (would throw)
struct ThetaLookup {
  const TypeLut &lut;
  const std::vector<ClusterType> &types;
  const std::vector<double> &theta;
  const double *counter;
  double operator()(int label, int c1, int c2 = 0, int c3 = 0) const {
    const int idx = lut.index(label, c1, c2, c3);
    if (idx < 0 || idx >= static_cast<int>(theta.size())) return kNaN;
    return theta[idx] / counter[label];
  }
};

void check_theta(const Species &sp, const Coefficients &co, size_t n_types) {
  if (co.base_theta.size() != n_types)
    throw std::invalid_argument("Base.theta has " + std::to_string(co.base_theta.size()) + " entries, the element set needs " +
                                std::to_string(n_types) + " (n_species=" + std::to_string(sp.n) + ")");
}
}  // namespace

PairTables build_pair_tables(const Species &sp, const Coefficients &co, int model) {
  const Geometry &g = geometry();
  const int n = sp.n;
  const auto types = cluster_types(sp);
  check_theta(sp, co, types.size());
  const TypeLut lut = make_type_lut(sp, types);
  const ThetaLookup th{lut, types, co.base_theta, kDeClusterCounter};
  const int np = static_cast<int>(g.env_pairs.size());
  // pair index lookup
  std::vector<int> pair_index(kPairEnv * kPairEnv, -1);
  for (int p = 0; p < np; ++p) pair_index[g.env_pairs[p][0] * kPairEnv + g.env_pairs[p][1]] = p;

  using ld = long double;
  const size_t a_sz = static_cast<size_t>(n) * kPairEnv * n * 3, b_sz = static_cast<size_t>(n) * np * n * n * 3;
  std::vector<ld> C0(static_cast<size_t>(n) * 3, 0), A0(a_sz, 0), B0(b_sz, 0);
  auto a_at = [&](int m, int t, int e, int q) -> ld & { return A0[((static_cast<size_t>(m) * kPairEnv + t) * n + e) * 3 + q]; };
  auto b_at = [&](int m, int p, int a, int b, int q) -> ld & {
    return B0[(((static_cast<size_t>(m) * np + p) * n + a) * n + b) * 3 + q];
  };
  PairTables out;
  out.n = n;
  out.model = model;
  if (model != kModelQuartic && model != kModelE0) throw std::invalid_argument("unknown barrier model");
  int len_mmm = 0, len_mm2 = 0;
  const auto groups_mmm = group_layout(g, false, n, &len_mmm);
  const auto groups_mm2 = group_layout(g, true, n, &len_mm2);

  for (int m = 0; m < n; ++m) {
    // ---- dE (VacancyMigrationPredictorQuartic::GetDe, :112-166): first site X -> m, second site m -> X
    const int X = n;
    for (const auto &c : g.state_pair) {
      int env[3], n_env = 0, role[3];
      for (int i = 0; i < c.arity; ++i) {
        role[i] = g.env_of_state[c.pos[i]];
        if (role[i] >= 0) env[n_env++] = i;
      }
      auto contribution = [&](const int *env_codes) -> ld {
        int start[3] = {0, 0, 0}, end[3] = {0, 0, 0};
        int k = 0;
        for (int i = 0; i < c.arity; ++i) {
          if (role[i] == -1) { start[i] = X; end[i] = m; }
          else if (role[i] == -2) { start[i] = m; end[i] = X; }
          else { start[i] = end[i] = env_codes[k++]; }
        }
        return static_cast<ld>(th(c.label, end[0], end[1], end[2])) - static_cast<ld>(th(c.label, start[0], start[1], start[2]));
      };
      if (n_env == 0) {
        C0[m * 3 + 0] += contribution(nullptr);
      } else if (n_env == 1) {
        const int t = role[env[0]];
        for (int e = 0; e < n; ++e) a_at(m, t, e, 0) += contribution(&e);
      } else {
        int t = role[env[0]], u = role[env[1]];
        const bool swapped = t > u;
        if (swapped) std::swap(t, u);
        const int p = pair_index[t * kPairEnv + u];
        if (p < 0) throw std::logic_error("state triplet over a non-adjacent env pair");
        for (int a = 0; a < n; ++a)
          for (int b = 0; b < n; ++b) {
            const int codes[2] = {swapped ? b : a, swapped ? a : b};   // codes[] follows the cluster's member order
            b_at(m, p, a, b, 0) += contribution(codes);
          }
      }
    }
    // ---- logD / logKs (GetD :216-246, GetKs :167-215)
    const auto it = co.element.find(sp.enum_of_code[m]);
    if (it == co.element.end() || !(model == kModelE0 ? it->second.present_e0 : it->second.present)) {
      C0[m * 3 + 1] = C0[m * 3 + 2] = kNaN;
      continue;
    }
    out.has_barrier = true;
    const ElementCoefficients &ec = it->second;
    auto weights = [&](const std::vector<double> &U, int K, const std::vector<double> &theta, const std::vector<double> &sigma_x,
                       const std::vector<double> &mu_x, double sigma_y, double mu_y, int len, std::vector<ld> &w, ld &c0,
                       const char *what) {
      if (static_cast<int>(sigma_x.size()) != len || static_cast<int>(mu_x.size()) != len ||
          static_cast<int>(U.size()) != K * len || static_cast<int>(theta.size()) != K)
        throw std::invalid_argument(std::string("coefficient block '") + what + "' has the wrong shape for this element set");
      w.assign(len, 0);
      c0 = mu_y;
      for (int i = 0; i < len; ++i) {
        ld s = 0;
        for (int r = 0; r < K; ++r) s += static_cast<ld>(theta[r]) * static_cast<ld>(U[static_cast<size_t>(r) * len + i]);
        w[i] = static_cast<ld>(sigma_y) * s / static_cast<ld>(sigma_x[i]);
        c0 -= w[i] * static_cast<ld>(mu_x[i]);
      }
    };
    if (model == kModelE0) {
      // log e0 = mu_e0 + sigma_e0 theta_e0^T U_mmm ((x - mu_x) / sigma_x)   (VacancyMigrationPredictorE0.cpp:128-151); it is
      // stored as quantity 2 with quantity 1 == 0, so that the folded pair (dE, q2 + 2 q1) is (dE, log e0)
      std::vector<ld> w_e0;
      ld c_e0;
      weights(ec.U_mmm, ec.k_mmm, ec.theta_e0, ec.sigma_x_mmm, ec.mu_x_mmm, ec.sigma_e0, ec.mu_e0, len_mmm, w_e0, c_e0, "mmm (E0)");
      C0[m * 3 + 2] += c_e0;
      for (const auto &c : g.mmm) {
        const GroupInfo &gi = groups_mmm[c.group];
        const ld inv = static_cast<ld>(1) / static_cast<ld>(gi.size);
        if (c.arity == 1) {
          const int t = g.env_of_mmm[c.pos[0]];
          for (int e = 0; e < n; ++e) a_at(m, t, e, 2) += w_e0[gi.offset + e] * inv;
        } else {
          int t = g.env_of_mmm[c.pos[0]], u = g.env_of_mmm[c.pos[1]];
          const bool swapped = t > u;
          if (swapped) std::swap(t, u);
          const int p = pair_index[t * kPairEnv + u];
          if (p < 0) throw std::logic_error("mmm pair over a non-adjacent env pair");
          for (int e1 = 0; e1 < n; ++e1)
            for (int e2 = 0; e2 < n; ++e2) {
              const int slot = gi.symmetric ? tri_index(e1, e2, n) : e1 * n + e2;
              b_at(m, p, swapped ? e2 : e1, swapped ? e1 : e2, 2) += w_e0[gi.offset + slot] * inv;
            }
        }
      }
      continue;
    }
    std::vector<ld> w_d, w_ks;
    ld c_d, c_ks;
    weights(ec.U_mmm, ec.k_mmm, ec.theta_D, ec.sigma_x_mmm, ec.mu_x_mmm, ec.sigma_D, ec.mu_D, len_mmm, w_d, c_d, "mmm");
    weights(ec.U_mm2, ec.k_mm2, ec.theta_Ks, ec.sigma_x_mm2, ec.mu_x_mm2, ec.sigma_Ks, ec.mu_Ks, len_mm2, w_ks, c_ks, "mm2");
    C0[m * 3 + 1] += c_d;
    C0[m * 3 + 2] += c_ks;
    auto scatter = [&](const std::vector<Cluster> &clusters, const std::vector<GroupInfo> &groups, const std::array<int, kPairEnv> &env_of,
                       const std::vector<ld> &w, int q) {
      for (const auto &c : clusters) {
        const GroupInfo &gi = groups[c.group];
        const ld inv = static_cast<ld>(1) / static_cast<ld>(gi.size);
        if (c.arity == 1) {
          const int t = env_of[c.pos[0]];
          for (int e = 0; e < n; ++e) a_at(m, t, e, q) += w[gi.offset + e] * inv;
        } else {
          int t = env_of[c.pos[0]], u = env_of[c.pos[1]];   // (first, second) member in list order
          const bool swapped = t > u;
          if (swapped) std::swap(t, u);
          const int p = pair_index[t * kPairEnv + u];
          if (p < 0) throw std::logic_error("mmm/mm2 pair over a non-adjacent env pair");
          for (int e1 = 0; e1 < n; ++e1)       // e1: element of the first member, e2: of the second
            for (int e2 = 0; e2 < n; ++e2) {
              const int slot = gi.symmetric ? tri_index(e1, e2, n) : e1 * n + e2;
              b_at(m, p, swapped ? e2 : e1, swapped ? e1 : e2, q) += w[gi.offset + slot] * inv;
            }
        }
      }
    };
    scatter(g.mmm, groups_mmm, g.env_of_mmm, w_d, 1);
    scatter(g.mm2, groups_mm2, g.env_of_mm2, w_ks, 2);
    scatter(g.mm2, groups_mm2, g.env_of_mm2_backward[1], w_ks, 2);   // x = enc_forward + enc_backward (:204-205)
  }
  // ---- delta form relative to the solvent species
  const int s0 = sp.solvent;
  out.C.assign(static_cast<size_t>(n) * 3, 0);
  out.A.assign(a_sz, 0);
  out.B.assign(b_sz, 0);
  for (int m = 0; m < n; ++m)
    for (int q = 0; q < 3; ++q) {
      ld c = C0[m * 3 + q];
      for (int t = 0; t < kPairEnv; ++t) c += a_at(m, t, s0, q);
      for (int p = 0; p < np; ++p) c += b_at(m, p, s0, s0, q);
      out.C[m * 3 + q] = static_cast<double>(c);
      std::vector<ld> a_delta(static_cast<size_t>(kPairEnv) * n, 0);
      for (int t = 0; t < kPairEnv; ++t)
        for (int e = 0; e < n; ++e) a_delta[t * n + e] = a_at(m, t, e, q) - a_at(m, t, s0, q);
      for (int p = 0; p < np; ++p) {
        const int t = g.env_pairs[p][0], u = g.env_pairs[p][1];
        for (int e = 0; e < n; ++e) {
          a_delta[t * n + e] += b_at(m, p, e, s0, q) - b_at(m, p, s0, s0, q);
          a_delta[u * n + e] += b_at(m, p, s0, e, q) - b_at(m, p, s0, s0, q);
        }
        for (int a = 0; a < n; ++a)
          for (int b = 0; b < n; ++b)
            out.B[(((static_cast<size_t>(m) * np + p) * n + a) * n + b) * 3 + q] =
                static_cast<double>(b_at(m, p, a, b, q) - b_at(m, p, a, s0, q) - b_at(m, p, s0, b, q) + b_at(m, p, s0, s0, q));
      }
      for (int t = 0; t < kPairEnv; ++t)
        for (int e = 0; e < n; ++e)
          out.A[((static_cast<size_t>(m) * kPairEnv + t) * n + e) * 3 + q] = static_cast<double>(a_delta[t * n + e]);
    }
  return out;
}

SiteTables build_site_tables(const Species &sp, const Coefficients &co) {
  const Geometry &g = geometry();
  const int m = sp.n + 1;
  const auto types = cluster_types(sp);
  check_theta(sp, co, types.size());
  const TypeLut lut = make_type_lut(sp, types);
  const ThetaLookup th{lut, types, co.base_theta, kDeClusterCounter};
  const int np = static_cast<int>(g.site_env_pairs.size());
  std::vector<Int3> env;
  for (int t = 0; t < kSiteSites; ++t)
    if (t != g.site_centre_pos) env.push_back(g.site_offsets[t]);
  using ld = long double;
  const int s0 = sp.solvent;
  SiteTables out;
  out.m = m;
  out.C.assign(m, 0);
  out.A.assign(static_cast<size_t>(m) * kSiteEnv * m, 0);
  out.B.assign(static_cast<size_t>(m) * np * m * m, 0);
  for (int x = 0; x < m; ++x) {
    ld c = th(0, x);
    std::vector<ld> a_delta(static_cast<size_t>(kSiteEnv) * m, 0);
    for (int t = 0; t < kSiteEnv; ++t) {
      const int shell = g.site_env_shell[t];
      c += static_cast<ld>(th(shell, x, s0));
      for (int e = 0; e < m; ++e) a_delta[t * m + e] = static_cast<ld>(th(shell, x, e)) - static_cast<ld>(th(shell, x, s0));
    }
    for (int p = 0; p < np; ++p) {
      const int t = g.site_env_pairs[p][0], u = g.site_env_pairs[p][1];
      const int label = triplet_label(bond_label(Int3{0, 0, 0}, env[t]), bond_label(env[t], env[u]), bond_label(env[u], Int3{0, 0, 0}));
      const ld b00 = th(label, x, s0, s0);
      c += b00;
      for (int e = 0; e < m; ++e) {
        a_delta[t * m + e] += static_cast<ld>(th(label, x, e, s0)) - b00;
        a_delta[u * m + e] += static_cast<ld>(th(label, x, s0, e)) - b00;
      }
      for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b)
          out.B[((static_cast<size_t>(x) * np + p) * m + a) * m + b] = static_cast<double>(
              static_cast<ld>(th(label, x, a, b)) - static_cast<ld>(th(label, x, a, s0)) - static_cast<ld>(th(label, x, s0, b)) + b00);
    }
    out.C[x] = static_cast<double>(c);
    for (int t = 0; t < kSiteEnv; ++t)
      for (int e = 0; e < m; ++e) out.A[(static_cast<size_t>(x) * kSiteEnv + t) * m + e] = static_cast<double>(a_delta[t * m + e]);
  }
  return out;
}

double energy_cluster_counter(int label) { return kEnergyClusterCounter[label]; }

EnergyTables build_energy_tables(const Species &sp, const Coefficients &co) {
  const int m = sp.n + 1;
  const auto types = cluster_types(sp);
  check_theta(sp, co, types.size());
  const TypeLut lut = make_type_lut(sp, types);
  const ThetaLookup th{lut, types, co.base_theta, kEnergyClusterCounter};
  EnergyTables out;
  out.m = m;
  out.single.resize(m);
  out.pair.resize(static_cast<size_t>(3) * m * m);
  out.triplet.resize(static_cast<size_t>(4) * m * m * m);
  for (int a = 0; a < m; ++a) {
    out.single[a] = th(0, a);
    for (int b = 0; b < m; ++b) {
      for (int s = 1; s <= 3; ++s) out.pair[((s - 1) * m + a) * m + b] = th(s, a, b);
      for (int c = 0; c < m; ++c)
        for (int l = 4; l <= 7; ++l) out.triplet[(((l - 4) * m + a) * m + b) * m + c] = th(l, a, b, c);
    }
  }
  return out;
}

}  // namespace lmc

// tables.h -- host-side construction of every constant table the kernels use.
//
// This is the B200-first replacement of the reference's per-site / per-pair precomputed std::vector tables
// (pred/src/VacancyMigrationPredictorQuartic.cpp:64-100: 12*N*(60+58+58) size_t;
//  pred/src/EnergyChangePredictorPairSite.cpp:41-57: N*43 size_t + N unordered_sets): because the lattice is
// translation invariant, the symmetry-ordered neighbourhoods are *constant offset patterns* and the cluster
// mappings are constant index tables, all derived here with exact integer geometry, and the JSON coefficients are
// contracted into per-position / per-pair lookup tables (SURVEY.md §7 "coefficient contraction").
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "lattice.h"

namespace lmc {

constexpr int kNumDirections = 12;  // first-neighbour jump directions
constexpr int kPairSites = 60;      // constants::kNumThirdNearestSetSizeOfPair (cfg/include/Constants.hpp:20)
constexpr int kPairEnv = 58;        // the 60 minus the jump pair
constexpr int kSiteSites = 43;      // constants::kNumThirdNearestSetSizeOfSite (Constants.hpp:21)
constexpr int kSiteEnv = 42;
constexpr int kNumQuantities = 3;   // dE, logD, logKs
constexpr int kMaxSpecies = 7;      // chemical species without the vacancy (compact codes 0..n-1, vacancy = n)

struct Int3 {
  int x, y, z;
};
inline Int3 operator+(Int3 a, Int3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Int3 operator-(Int3 a, Int3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline bool operator==(Int3 a, Int3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline int dot(Int3 a, Int3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Int3 cross(Int3 a, Int3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline int norm2(Int3 a) { return dot(a, a); }

// A cluster of the reference's mappings, as positions in an ordered neighbourhood list.
struct Cluster {
  int8_t label;        // state mappings: 0 singlet, 1-3 pairs, 4-7 triplets.  mmm/mm2: 0 singlet, 1-3 pair shell
  int8_t arity;        // 1..3
  int16_t pos[3];      // list positions (state lists: index1 > index2 > index3 as in the reference)
  int16_t group;       // mmm/mm2 only: group number in mapping order
  bool symmetric;      // mmm/mm2 only: carries the SIZE_MAX marker in the reference
};

struct GroupInfo {
  int32_t size;        // number of clusters
  int32_t offset;      // first slot in the encode vector
  int32_t length;      // n, n*n or n(n+1)/2
  int8_t arity;
  bool symmetric;
};

// Geometry that depends on nothing but the FCC structure.
struct Geometry {
  std::array<Int3, 12> nn1;
  std::array<Int3, 6> nn2;
  std::array<Int3, 24> nn3;
  // frame of each jump direction: d = nn1[k], p = canonical perpendicular 1NN, c = d x p
  std::array<Int3, 12> frame_p;
  // canonical rotated keys (alpha,beta,gamma) of the 60 pair-neighbourhood sites in state order
  std::array<Int3, kPairSites> pair_keys;
  int pair_first_pos, pair_second_pos;     // positions of the jump pair in the state list (21 and 38)
  std::array<int, kPairEnv> env_of_mmm;    // mmm list position  -> env index (state order without the pair)
  std::array<int, kPairEnv> env_of_mm2;    // mm2 list position  -> env index
  // backward mm2 list (pair j->i) in terms of forward env indices; [0]: backward frame uses -p_f, [1]: +p_f
  std::array<std::array<int, kPairEnv>, 2> env_of_mm2_backward;
  std::array<int, kPairSites> env_of_state;  // state position -> env index, -1 / -2 for first / second
  // lattice offsets (relative to the first site of the pair) of the 60 state-ordered sites, per direction and
  // per frame flag s (0: y axis = +frame_p[k], 1: y axis = -frame_p[k])
  std::array<std::array<std::array<Int3, kPairSites>, 2>, 12> pair_offsets;
  std::array<Int3, kSiteSites> site_offsets;  // state order of a site neighbourhood (centre at position 21)
  int site_centre_pos;

  std::vector<Cluster> state_pair;   // 473 clusters (positions in the 60-list)
  std::vector<Cluster> state_site;   // 247 clusters (positions in the 43-list)
  std::vector<Cluster> mmm, mm2;     // 614 clusters each (positions in the 58-lists), grouped
  int n_groups_mmm{0}, n_groups_mm2{0};
  std::vector<int8_t> group_arity_mmm, group_arity_mm2;
  std::vector<bool> group_sym_mmm, group_sym_mm2;
  std::vector<int32_t> group_size_mmm, group_size_mm2;

  // all (t<u) env pairs of the jump neighbourhood within 3NN of each other, in (t,u) lexicographic order
  std::vector<std::array<int16_t, 2>> env_pairs;             // 556
  std::array<uint64_t, kPairEnv> env_pair_mask_hi;           // bit u set iff u>t and (t,u) is a pair
  std::array<uint16_t, kPairEnv> env_pair_base;              // index of the first pair of t in env_pairs
  // site neighbourhood: env index = state position without the centre; pairs (t<u) forming a triplet with it
  std::vector<std::array<int16_t, 2>> site_env_pairs;        // 204
  std::array<uint64_t, kSiteEnv> site_pair_mask_hi;
  std::array<uint16_t, kSiteEnv> site_pair_base;
  std::array<int8_t, kSiteEnv> site_env_shell;               // 1..3: shell of env site around the centre
  // total-energy walk (pred/src/EnergyPredictor.cpp:40-96): ordered (offset2, offset3, label) tuples
  struct Walk { Int3 o2, o3; int8_t label; };
  std::vector<Walk> energy_triplets;                          // 144
};

const Geometry &geometry();   // built once, thread-safe

// Chemistry: element set, cluster types, coefficient tables.
struct Species {
  int n{0};                                  // species without vacancy
  std::vector<int> enum_of_code;             // compact code -> reference ElementName value; code n = vacancy (0)
  std::array<int8_t, 16> code_of_enum;       // ElementName value -> compact code, -1 if absent
  std::vector<std::string> names;            // by compact code (code n = "X")
  int solvent{0};                            // compact code used as the expansion origin of the delta tables
};
Species make_species(const int32_t *enum_codes, int n, int solvent_enum);
const char *element_name(int enum_code);
int element_enum_from_name(const std::string &name);

struct ClusterType {
  int8_t label, arity;
  int8_t code[3];                            // compact codes, sorted by element name (vacancy = n)
};
// Cluster types in pred::ClusterIndexer order (pred/src/EnergyUtility.cpp:314-343,798-811).
std::vector<ClusterType> cluster_types(const Species &sp);
// dense lookup: index = type_lut[label][c1][c2][c3] (unused trailing codes = 0), -1 if no such type
struct TypeLut {
  int m;                                     // n+1 codes
  std::vector<int16_t> lut;                  // 8*m*m*m
  int index(int label, int c1, int c2 = 0, int c3 = 0) const { return lut[((label * m + c1) * m + c2) * m + c3]; }
};
TypeLut make_type_lut(const Species &sp, const std::vector<ClusterType> &types);

std::vector<GroupInfo> group_layout(const Geometry &g, bool mm2, int n_species, int *encode_length);

// Raw coefficients as parsed from the reference's JSON format
// (pred/src/VacancyMigrationPredictorQuartic.cpp:38-63).
struct ElementCoefficients {
  std::vector<double> mu_x_mmm, sigma_x_mmm, mu_x_mm2, sigma_x_mm2, theta_D, theta_Ks;
  std::vector<double> U_mmm, U_mm2;          // row-major K x L
  int k_mmm{0}, k_mm2{0};
  double mu_D{0}, sigma_D{1}, mu_Ks{0}, sigma_Ks{1};
  bool present{false};                       // carries the quartic keys (mmm + mm2 blocks, theta_D / theta_Ks, ...)
  // E0 model (pred/src/VacancyMigrationPredictorE0.cpp:30-37): mmm block + theta_e0 / mu_e0 / sigma_e0
  std::vector<double> theta_e0;
  double mu_e0{0}, sigma_e0{1};
  bool present_e0{false};
};
// which closed form turns (dE, environment) into a barrier
constexpr int kModelQuartic = 0;             // VacancyMigrationPredictorQuartic: Ea(dE, D, Ks)
constexpr int kModelE0 = 1;                  // VacancyMigrationPredictorE0: Ea = max(0, e0 + dE / 2)
struct Coefficients {
  std::vector<double> base_theta;
  std::map<int, ElementCoefficients> element;   // key: ElementName enum value
};
Coefficients parse_coefficients_json(const std::string &path);

// Contracted lookup tables (see DESIGN.md "coefficient contraction").  Every quantity Q in {dE, logD, logKs} of a
// jump with migrating species m is   Q = C[m] + sum_t A[m][t][e_t] + sum_{(t,u)} B[m][(t,u)][e_t][e_u]
// over the 58 environment sites; the tables are stored in "delta" form relative to the solvent species s0, so
// terms with e_t == s0 vanish identically and may be skipped.
struct PairTables {
  int n{0};
  std::vector<double> C;   // [m][3]
  std::vector<double> A;   // [m][t][e][3]
  std::vector<double> B;   // [m][pair][a][b][3]
  bool has_barrier{false}; // false if the JSON has no per-element blocks (dE only)
  int model{kModelQuartic}; // kModelE0: quantity 1 is 0 and quantity 2 is log e0
};
// Site tables:  H(x, env) = Cs[x] + sum_t As[x][t][e_t] + sum_{(t,u)} Bs[x][(t,u)][e_t][e_u],  codes incl. vacancy;
// dE(site: old->new) = H(new) - H(old)   (pred/src/EnergyChangePredictorPairSite.cpp:154-192)
struct SiteTables {
  int m{0};                // n+1 codes
  std::vector<double> C;   // [x]
  std::vector<double> A;   // [x][t][e]
  std::vector<double> B;   // [x][pair][a][b]
};
// Total-energy tables (pred/src/EnergyPredictor.cpp:8,40-96,173-177): theta[idx]/normaliser by type
struct EnergyTables {
  int m{0};
  std::vector<double> single;   // [c]
  std::vector<double> pair;     // [shell-1][c1][c2]
  std::vector<double> triplet;  // [label-4][c1][c2][c3]
};

PairTables build_pair_tables(const Species &sp, const Coefficients &co, int model = kModelQuartic);
SiteTables build_site_tables(const Species &sp, const Coefficients &co);
EnergyTables build_energy_tables(const Species &sp, const Coefficients &co);
double energy_cluster_counter(int label);   // normaliser of EnergyPredictor::GetEncode by cluster label (pred/src/EnergyPredictor.cpp:8)

}  // namespace lmc

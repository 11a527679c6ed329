// lmc_cli.cpp -- `lmc_b200.exe -p <param file>`: the reference's command-line surface on top of the C ABI.
//
// Mirrors (file:line under /root/reference/lmc/):
//   api::Parameter::ReadParam            api/src/Parameter.cpp:23-130   key/value file, '#' comments, unknown keys ignored
//   api::Run dispatch on simulation_method api/src/Home.cpp:97-125       KineticMcFirstOmp/Mpi, KineticMcChainOmpi, CanonicalMcSerial/Omp, SimulatedAnnealing
//   Config::ReadConfig / WriteConfig      cfg/src/Config.cpp:554-710     .cfg (and .cfg.gz through zlib)
//   KineticMcFirstAbstract::Dump          mc/src/KineticMcAbstract.cpp:59-104   kmc_log.txt, N.cfg.gz, end.cfg.gz
//   CanonicalMcAbstract::Dump             mc/src/CanonicalMcAbstract.cpp:53-84  cmc_log.txt
//   SimulatedAnnealing::Dump              mc/src/SimulatedAnnealing.cpp:81-97   sa_log.txt
//   ThermodynamicAveraging                mc/src/ThermodynamicAveraging.cpp:5-39
// The event loop itself runs on the GPU (lmc_kmc_run / lmc_cmc_run); this file only parses, logs and dumps.
// Ansys / Reformat are not part of the accelerated path.
#include <zlib.h>

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/lmc_b200.h"

namespace {

using Vec3 = std::array<double, 3>;
using Mat3 = std::array<Vec3, 3>;
constexpr double kLatticeConstant = 4.046;     // cfg/include/Constants.hpp:6
constexpr double kBoltzmann = 8.617333262145e-5;
constexpr double kEpsilon = 1e-4;              // cfg/include/VectorMatrix.hpp:65

const char *kNames[] = {"X", "Al", "Mg", "Zn", "Cu", "Sn", "pAl", "pMg", "pZn", "pCu", "pSn"};
const double kMass[] = {0.00, 26.98, 24.31, 65.38, 63.55, 118.71, 26.98, 24.31, 65.38, 63.55, 118.71};   // Element.hpp:67-84
int element_from_string(const std::string &s) {
  for (int i = 0; i < 11; ++i)
    if (s == kNames[i]) return i;
  return 0;   // unknown strings are X (Element.hpp:25-26)
}

void check(int rc) {
  if (rc >= 0) return;
  const std::string msg = lmc_last_error();
  if (rc == LMC_ERR_OUT_OF_RANGE) throw std::out_of_range(msg);
  if (rc == LMC_ERR_INVALID_ARGUMENT) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);
}

// ------------------------------------------------------------------------------------------------ Parameter
struct Parameter {
  std::string parameters_filename, method, config_filename, map_filename, json_coefficients_filename, time_temperature_filename;
  unsigned long long log_dump_steps{0}, config_dump_steps{0}, maximum_steps{0}, thermodynamic_averaging_steps{0}, restart_steps{0};
  double temperature{0}, initial_temperature{0}, restart_energy{0}, restart_time{0};
  std::vector<std::string> element_set, solute_element_set;
  std::vector<size_t> solute_number_set;
  bool rate_corrector{false}, early_stop{false}, solute_disp{false};
  size_t factor{0};
  std::string solvent_element;
  // extensions of this engine (ignored by the reference, which skips unknown keys)
  unsigned long long seed{0};
  bool seed_given{false};
  std::string replay_uniforms_filename;   // validation: "u1 u2" per KMC step instead of the device RNG
  std::string replay_trials_filename;     // validation: "site_a site_b u" per CMC trial (the reference's serial stream)
  int domain_edge{0};                     // extension: > 0 runs CMC / SA with the domain-decomposed driver (lmc_cmc_domain_run)
  int rounds_per_sweep{0};
  int device{0};

  static std::vector<std::string> split(const std::string &s) {
    std::vector<std::string> out;
    std::istringstream iss(s);
    for (std::string tok; iss >> tok;) out.push_back(tok);
    return out;
  }
  void parse_args(int argc, char **argv) {
    for (int i = 0; i < argc; ++i)
      if ((!std::strcmp(argv[i], "--p") || !std::strcmp(argv[i], "-p")) && i + 1 < argc) parameters_filename = argv[++i];
  }
  void read(const std::string &filename) {
    std::ifstream ifs(filename);
    if (!ifs.is_open()) throw std::runtime_error("Cannot open " + filename);
    for (std::string line; std::getline(ifs, line);) {
      if (line.empty() || line[0] == '#') continue;
      const auto segs = split(line);
      if (segs.size() < 2) continue;
      const std::string &k = segs[0], &v = segs[1];
      auto flag = [&] { return v == "true"; };
      if (k == "simulation_method") method = v;
      else if (k == "config_filename") config_filename = v;
      else if (k == "map_filename") map_filename = v;
      else if (k == "json_coefficients_filename") json_coefficients_filename = v;
      else if (k == "time_temperature_filename") time_temperature_filename = v;
      else if (k == "log_dump_steps") log_dump_steps = std::stoull(v);
      else if (k == "config_dump_steps") config_dump_steps = std::stoull(v);
      else if (k == "maximum_steps") maximum_steps = std::stoull(v);
      else if (k == "thermodynamic_averaging_steps") thermodynamic_averaging_steps = std::stoull(v);
      else if (k == "temperature") temperature = std::stod(v);
      else if (k == "initial_temperature") initial_temperature = std::stod(v);
      else if (k == "element_set") element_set.assign(segs.begin() + 1, segs.end());
      else if (k == "restart_steps") restart_steps = std::stoull(v);
      else if (k == "restart_energy") restart_energy = std::stod(v);
      else if (k == "restart_time") restart_time = std::stod(v);
      else if (k == "rate_corrector") rate_corrector = flag();
      else if (k == "early_stop") early_stop = flag();
      else if (k == "solute_disp") solute_disp = flag();
      else if (k == "factor") factor = std::stoul(v);
      else if (k == "solvent_element") solvent_element = v;
      else if (k == "solute_element_set") solute_element_set.assign(segs.begin() + 1, segs.end());
      else if (k == "solute_number_set") {
        solute_number_set.clear();
        for (size_t i = 1; i < segs.size(); ++i) solute_number_set.push_back(std::stoul(segs[i]));
      } else if (k == "seed") { seed = std::stoull(v); seed_given = true; }
      else if (k == "replay_uniforms_filename") replay_uniforms_filename = v;
      else if (k == "replay_trials_filename") replay_trials_filename = v;
      else if (k == "domain_edge") domain_edge = std::stoi(v);
      else if (k == "rounds_per_sweep") rounds_per_sweep = std::stoi(v);
      else if (k == "device") device = std::stoi(v);
    }
  }
};

// ------------------------------------------------------------------------------------------------ files (.cfg / .cfg.gz)
bool ends_with(const std::string &s, const std::string &suffix) {
  return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}
std::string read_text_file(const std::string &filename) {
  gzFile f = gzopen(filename.c_str(), "rb");   // transparently reads plain files too
  if (!f) throw std::runtime_error("Cannot open " + filename);
  std::string out;
  char buf[1 << 16];
  for (int n; (n = gzread(f, buf, sizeof buf)) > 0;) out.append(buf, static_cast<size_t>(n));
  gzclose(f);
  return out;
}
void write_text_file(const std::string &filename, const std::string &text) {
  if (ends_with(filename, ".gz")) {
    gzFile f = gzopen(filename.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot open " + filename);
    gzwrite(f, text.data(), static_cast<unsigned>(text.size()));
    gzclose(f);
  } else {
    std::ofstream ofs(filename, std::ios::binary);
    if (!ofs) throw std::runtime_error("Cannot open " + filename);
    ofs << text;
  }
}

// Host mirror of cfg::Config for FCC supercells: everything the logs and dumps need.
struct HostConfig {
  Mat3 basis{};
  int32_t factors[3]{};
  std::vector<uint8_t> element_of_atom;            // atom_vector_[a].element_
  std::vector<Vec3> rel_of_lattice;                // lattice_vector_[l].relative_position_
  std::vector<std::array<int, 3>> map_shift;       // map_shift_list_[a]
  std::vector<int64_t> atom_to_lattice, lattice_to_atom;
  lmc_id_order id_order{LMC_ID_ORDER_REASSIGNED};  // how lattice ids map to sites (a .cfg is reassigned; a map keeps lattice.txt's order)
  size_t n() const { return element_of_atom.size(); }
  int64_t lattice_id(int X, int Y, int Z) const {
    if (id_order == LMC_ID_ORDER_GENERATE) {        // cfg::GenerateFCC order (Config.cpp:1073-1090): ((k fy + j) fx + i) 4 + basis site
      const int b = (X & 1) ? ((Y & 1) ? 1 : 2) : ((Y & 1) ? 3 : 0);
      return ((static_cast<int64_t>(Z >> 1) * factors[1] + (Y >> 1)) * factors[0] + (X >> 1)) * 4 + b;
    }
    // ReassignLatticeVector order
    return static_cast<int64_t>(X) * (2LL * factors[1] * factors[2]) + static_cast<int64_t>(Y) * factors[2] + (Z >> 1);
  }
  uint8_t element_at_lattice(int64_t l) const { return element_of_atom[static_cast<size_t>(lattice_to_atom[static_cast<size_t>(l)])]; }
  std::vector<uint8_t> occupancy() const {
    std::vector<uint8_t> occ(n());
    for (size_t l = 0; l < n(); ++l) occ[l] = element_at_lattice(static_cast<int64_t>(l));
    return occ;
  }
  // Config::LatticeJump (cfg/src/Config.cpp:431-456)
  void lattice_jump(int64_t lhs, int64_t rhs) {
    const int64_t a_lhs = lattice_to_atom[static_cast<size_t>(lhs)], a_rhs = lattice_to_atom[static_cast<size_t>(rhs)];
    for (int d = 0; d < 3; ++d) {
      const double shift = rel_of_lattice[static_cast<size_t>(rhs)][d] - rel_of_lattice[static_cast<size_t>(lhs)][d];
      if (std::abs(shift) > 0.5 + 1e-8) {
        const int change = static_cast<int>(std::floor(shift + 0.5));
        map_shift[static_cast<size_t>(a_lhs)][d] -= change;
        map_shift[static_cast<size_t>(a_rhs)][d] += change;
      }
    }
    atom_to_lattice[static_cast<size_t>(a_lhs)] = rhs;
    atom_to_lattice[static_cast<size_t>(a_rhs)] = lhs;
    lattice_to_atom[static_cast<size_t>(lhs)] = a_rhs;
    lattice_to_atom[static_cast<size_t>(rhs)] = a_lhs;
  }
  Vec3 times_basis(const Vec3 &v) const {
    return {v[0] * basis[0][0] + v[1] * basis[1][0] + v[2] * basis[2][0], v[0] * basis[0][1] + v[1] * basis[1][1] + v[2] * basis[2][1],
            v[0] * basis[0][2] + v[1] * basis[1][2] + v[2] * basis[2][2]};
  }
  // Config::GetUnwrappedCartesianPositionOfLattice (Config.cpp:280-294)
  Vec3 unwrapped_position_of_lattice(int64_t l) const {
    const auto &r = rel_of_lattice[static_cast<size_t>(l)];
    const auto &s = map_shift[static_cast<size_t>(lattice_to_atom[static_cast<size_t>(l)])];
    return times_basis({r[0] + s[0], r[1] + s[1], r[2] + s[2]});
  }
  // Config::ReadConfig (Config.cpp:554-661) followed by ReassignLatticeVector (:466-552) as in api/src/Home.cpp:165-167
  static HostConfig read(const std::string &filename) {
    std::istringstream fis(read_text_file(filename));
    auto after_equals = [&](auto &value) {
      fis.ignore(std::numeric_limits<std::streamsize>::max(), '=');
      fis >> value;
    };
    HostConfig c;
    size_t num_atoms = 0;
    after_equals(num_atoms);
    fis.ignore(std::numeric_limits<std::streamsize>::max(), '=');   // "A = 1.0 Angstrom"
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) after_equals(c.basis[i][j]);
    fis.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    bool has_no_velocity = false;
    if (fis.peek() == '.') {
      has_no_velocity = true;
      fis.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    }
    size_t entry_count = 0;
    after_equals(entry_count);
    fis.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    while (fis.good()) {   // skip "auxiliary[...]" header lines: the data section starts with a number
      while (std::isspace(fis.peek()) && fis.peek() != '\n') fis.get();
      const int next = fis.peek();
      if (fis.eof()) break;
      if (std::isdigit(next) || next == '.' || next == '-') break;
      fis.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    }
    std::vector<Vec3> rel(num_atoms);
    c.element_of_atom.resize(num_atoms);
    c.map_shift.assign(num_atoms, {0, 0, 0});
    for (size_t a = 0; a < num_atoms; ++a) {
      double mass;
      std::string type;
      fis >> mass >> type >> rel[a][0] >> rel[a][1] >> rel[a][2];
      if (entry_count >= 6 && has_no_velocity) fis >> c.map_shift[a][0] >> c.map_shift[a][1] >> c.map_shift[a][2];
      if (!fis) throw std::runtime_error("Unexpected end of " + filename);
      c.element_of_atom[a] = static_cast<uint8_t>(element_from_string(type));
      fis.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    }
    // FCC supercell detection: factors from the basis lengths and the site count, sites on the half-unit grid
    c.detect_factors(num_atoms, filename);
    c.atom_to_lattice.assign(num_atoms, -1);
    c.lattice_to_atom.assign(num_atoms, -1);
    c.rel_of_lattice.resize(num_atoms);
    for (size_t a = 0; a < num_atoms; ++a) {
      int xyz[3];
      for (int d = 0; d < 3; ++d) {
        const double g = rel[a][d] * 2.0 * c.factors[d];
        const long r = std::lround(g);
        if (std::abs(g - static_cast<double>(r)) > 1e-3) throw std::runtime_error(filename + ": atom off the FCC lattice");
        xyz[d] = static_cast<int>(((r % (2 * c.factors[d])) + 2 * c.factors[d]) % (2 * c.factors[d]));
      }
      if ((xyz[0] + xyz[1] + xyz[2]) & 1) throw std::runtime_error(filename + ": atom on the wrong FCC sublattice");
      const int64_t l = c.lattice_id(xyz[0], xyz[1], xyz[2]);
      if (c.lattice_to_atom[static_cast<size_t>(l)] >= 0) throw std::runtime_error(filename + ": two atoms on one lattice site");
      c.lattice_to_atom[static_cast<size_t>(l)] = static_cast<int64_t>(a);
      c.atom_to_lattice[a] = l;
      c.rel_of_lattice[static_cast<size_t>(l)] = rel[a];
    }
    return c;
  }
  // Config::ReadMap (Config.cpp:814-885): lattice.txt (positions + neighbour lists, lattice-id order as written),
  // element.txt (element of every atom id), map file (lattice id of every atom id).  No reassignment (Home.cpp:133-138):
  // the ids of lattice.txt are kept, so the file must be in one of the two orders the engine knows (what
  // Config::WriteLattice produces from a generated or a reassigned configuration); the neighbour columns are checked.
  static HostConfig read_map(const std::string &lattice_filename, const std::string &element_filename, const std::string &map_filename) {
    std::ifstream ifs_lattice(lattice_filename);
    if (!ifs_lattice) throw std::runtime_error("Cannot open " + lattice_filename);
    HostConfig c;
    size_t num_atoms = 0;
    ifs_lattice >> num_atoms;
    ifs_lattice.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) ifs_lattice >> c.basis[i][j];
    c.rel_of_lattice.resize(num_atoms);
    std::vector<std::array<int64_t, 42>> neighbours(num_atoms);
    for (size_t l = 0; l < num_atoms; ++l) {
      ifs_lattice >> c.rel_of_lattice[l][0] >> c.rel_of_lattice[l][1] >> c.rel_of_lattice[l][2];
      ifs_lattice.ignore(std::numeric_limits<std::streamsize>::max(), '#');
      for (auto &v : neighbours[l]) ifs_lattice >> v;
      if (!ifs_lattice) throw std::runtime_error("Unexpected end of " + lattice_filename);
      ifs_lattice.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    }
    c.detect_factors(num_atoms, lattice_filename);
    // which id order?  compare every site's id under both orders with its line number
    std::vector<std::array<int, 3>> xyz(num_atoms);
    for (size_t l = 0; l < num_atoms; ++l) xyz[l] = c.grid_coordinates(c.rel_of_lattice[l], lattice_filename);
    bool found = false;
    for (lmc_id_order order : {LMC_ID_ORDER_REASSIGNED, LMC_ID_ORDER_GENERATE}) {
      c.id_order = order;
      bool ok = true;
      for (size_t l = 0; l < num_atoms && ok; ++l) ok = c.lattice_id(xyz[l][0], xyz[l][1], xyz[l][2]) == static_cast<int64_t>(l);
      if (ok) { found = true; break; }
    }
    if (!found) throw std::runtime_error(lattice_filename + ": lattice ids are neither in GenerateFCC nor in ReassignLatticeVector order");
    // the first-neighbour column must be what this order implies (ascending ids of the 12 nearest sites)
    for (size_t l = 0; l < num_atoms; l += std::max<size_t>(1, num_atoms / 64)) {
      std::vector<int64_t> want;
      static const int nn[12][3] = {{1, 1, 0}, {1, -1, 0}, {-1, 1, 0}, {-1, -1, 0}, {1, 0, 1}, {1, 0, -1}, {-1, 0, 1}, {-1, 0, -1},
                                    {0, 1, 1}, {0, 1, -1}, {0, -1, 1}, {0, -1, -1}};
      for (const auto &d : nn) {
        int q[3];
        for (int k = 0; k < 3; ++k) q[k] = ((xyz[l][k] + d[k]) % (2 * c.factors[k]) + 2 * c.factors[k]) % (2 * c.factors[k]);
        want.push_back(c.lattice_id(q[0], q[1], q[2]));
      }
      std::sort(want.begin(), want.end());
      if (!std::equal(want.begin(), want.end(), neighbours[l].begin()))
        throw std::runtime_error(lattice_filename + ": neighbour list of lattice id " + std::to_string(l) + " does not match the FCC supercell");
    }
    std::ifstream ifs_element(element_filename);
    if (!ifs_element) throw std::runtime_error("Cannot open " + element_filename);
    c.element_of_atom.resize(num_atoms);
    c.map_shift.assign(num_atoms, {0, 0, 0});
    for (size_t a = 0; a < num_atoms; ++a) {
      std::string type;
      ifs_element >> type;
      if (!ifs_element) throw std::runtime_error("Unexpected end of " + element_filename);
      c.element_of_atom[a] = static_cast<uint8_t>(element_from_string(type));
      ifs_element.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    }
    std::ifstream ifs_map(map_filename);
    if (!ifs_map) throw std::runtime_error("Cannot open " + map_filename);
    c.atom_to_lattice.assign(num_atoms, -1);
    c.lattice_to_atom.assign(num_atoms, -1);
    for (size_t a = 0; a < num_atoms; ++a) {
      long long l = -1;
      ifs_map >> l;
      if (!ifs_map) throw std::runtime_error("Unexpected end of " + map_filename);
      if (l < 0 || static_cast<size_t>(l) >= num_atoms) throw std::runtime_error("Lattice id out of bounds in map file: " + std::to_string(l));
      if (c.lattice_to_atom[static_cast<size_t>(l)] >= 0) throw std::runtime_error("Duplicate lattice id in map file: " + std::to_string(l));
      c.lattice_to_atom[static_cast<size_t>(l)] = static_cast<int64_t>(a);
      c.atom_to_lattice[a] = l;
      ifs_map.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
    }
    return c;
  }
  // FCC supercell detection: factors from the basis lengths and the site count
  void detect_factors(size_t num_atoms, const std::string &filename) {
    double len[3], volume_cells = static_cast<double>(num_atoms) / 4.0, prod = 1.0;
    for (int d = 0; d < 3; ++d) {
      len[d] = std::sqrt(basis[d][0] * basis[d][0] + basis[d][1] * basis[d][1] + basis[d][2] * basis[d][2]);
      prod *= len[d];
    }
    const double scale = std::cbrt(volume_cells / prod);
    size_t check_sites = 4;
    for (int d = 0; d < 3; ++d) {
      factors[d] = static_cast<int32_t>(std::lround(len[d] * scale));
      check_sites *= static_cast<size_t>(factors[d]);
    }
    if (check_sites != num_atoms) throw std::runtime_error(filename + ": not an FCC supercell (site count does not match the basis)");
    // What the engine's integer geometry assumes -- checked, not assumed: an orthogonal cell whose axes are the cubic axes
    // (the reference would accept any basis and find whatever neighbours its distance cutoffs give), and a lattice
    // constant for which the reference's cutoffs 3.5 / 4.8 / 5.3 A (cfg/include/Constants.hpp:6-10) select exactly the first,
    // second and third FCC shells: r1 = a / sqrt 2, r2 = a, r3 = a sqrt(3/2), r4 = a sqrt 2  =>  3.92 A < a <= 4.32 A.
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        if (i != j && std::abs(basis[i][j]) > 1e-6 * len[i])
          throw std::runtime_error(filename + ": the cell must be orthogonal with its axes along the cubic axes (off-diagonal basis entries found)");
    for (int d = 0; d < 3; ++d) {
      const double a = len[d] / factors[d];
      if (!(a * std::sqrt(1.5) <= 5.3 && a * std::sqrt(2.0) > 5.3 && a > 4.8 / std::sqrt(1.5)))
        throw std::runtime_error(filename + ": lattice constant " + std::to_string(a) + " A along axis " + std::to_string(d) +
                                 " is outside the range (3.92, 4.32] A in which the reference's neighbour cutoffs give the FCC 1-3NN shells");
      if (factors[d] < 4) throw std::runtime_error(filename + ": the supercell needs at least 4 conventional cells per axis");
    }
  }
  // half-unit grid coordinates of a relative position
  std::array<int, 3> grid_coordinates(const Vec3 &rel, const std::string &filename) const {
    std::array<int, 3> xyz{};
    for (int d = 0; d < 3; ++d) {
      const double g = rel[d] * 2.0 * factors[d];
      const long r = std::lround(g);
      if (std::abs(g - static_cast<double>(r)) > 1e-3) throw std::runtime_error(filename + ": site off the FCC lattice");
      xyz[static_cast<size_t>(d)] = static_cast<int>(((r % (2 * factors[d])) + 2 * factors[d]) % (2 * factors[d]));
    }
    if ((xyz[0] + xyz[1] + xyz[2]) & 1) throw std::runtime_error(filename + ": site on the wrong FCC sublattice");
    return xyz;
  }
  // cfg::GenerateFCC (Config.cpp:1060-1095) relabelled like ReassignLatticeVector would: atom id = lattice id
  static HostConfig generate_fcc(size_t f, int element) {
    HostConfig c;
    for (int d = 0; d < 3; ++d) {
      c.factors[d] = static_cast<int32_t>(f);
      c.basis[d] = {0, 0, 0};
      c.basis[d][d] = kLatticeConstant * static_cast<double>(f);
    }
    const size_t n = 4 * f * f * f;
    c.element_of_atom.assign(n, static_cast<uint8_t>(element));
    c.map_shift.assign(n, {0, 0, 0});
    c.atom_to_lattice.resize(n);
    c.lattice_to_atom.resize(n);
    c.rel_of_lattice.resize(n);
    for (int X = 0; X < 2 * static_cast<int>(f); ++X)
      for (int Y = 0; Y < 2 * static_cast<int>(f); ++Y)
        for (int Z = (X + Y) & 1; Z < 2 * static_cast<int>(f); Z += 2) {
          const int64_t l = c.lattice_id(X, Y, Z);
          c.rel_of_lattice[static_cast<size_t>(l)] = {X / (2.0 * f), Y / (2.0 * f), Z / (2.0 * f)};
          c.atom_to_lattice[static_cast<size_t>(l)] = l;
          c.lattice_to_atom[static_cast<size_t>(l)] = l;
        }
    return c;
  }
  // Config::WriteExtendedConfig (Config.cpp:667-710): atom-id order, current lattice position, image counters.
  // (the first mass is written before std::fixed is in effect, exactly like the reference's stream state)
  void write(const std::string &filename) const {
    std::ostringstream fos;
    fos.precision(16);
    fos << "Number of particles = " << n() << '\n';
    fos << "A = 1.0 Angstrom (basic length-scale)\n";
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) fos << "H0(" << i + 1 << "," << j + 1 << ") = " << basis[i][j] << " A\n";
    fos << ".NO_VELOCITY.\n";
    fos << "entry_count = 6\n";
    fos << "auxiliary[0] = ix\nauxiliary[1] = iy\nauxiliary[2] = iz\n";
    for (size_t a = 0; a < n(); ++a) {
      const auto &r = rel_of_lattice[static_cast<size_t>(atom_to_lattice[a])];
      fos << kMass[element_of_atom[a]] << '\n' << kNames[element_of_atom[a]] << '\n';
      fos << std::fixed << r[0] << ' ' << r[1] << ' ' << r[2] << ' ' << map_shift[a][0] << ' ' << map_shift[a][1] << ' ' << map_shift[a][2] << '\n';
    }
    write_text_file(filename, fos.str());
  }
};

// mc::ThermodynamicAveraging (mc/src/ThermodynamicAveraging.cpp:5-39)
class ThermodynamicAveraging {
 public:
  explicit ThermodynamicAveraging(size_t size) : size_(size) {}
  void AddEnergy(double value) {
    if (size_ == 0) return;
    if (energy_list_.size() == size_) { sum_ -= energy_list_.front(); energy_list_.pop_front(); }
    energy_list_.push_back(value);
    sum_ += value;
  }
  double GetThermodynamicAverage(double beta) const {
    if (size_ == 0 || energy_list_.empty()) return 0;
    const double average = sum_ / static_cast<double>(energy_list_.size());
    double partition = 0.0, weighted = 0.0;
    for (double energy : energy_list_) {
      energy -= average;
      const double e = std::exp(-energy * beta);
      weighted += energy * e;
      partition += e;
    }
    return weighted / partition + average;
  }
 private:
  size_t size_;
  std::deque<double> energy_list_;
  double sum_{0};
};

// the log cadence of KineticMcFirstAbstract::Dump / CanonicalMcAbstract::Dump
bool log_this_step(unsigned long long steps, unsigned long long log_dump_steps_param) {
  unsigned long long every;
  if (steps > 10 * log_dump_steps_param) {
    every = log_dump_steps_param;
  } else {
    every = static_cast<unsigned long long>(std::pow(10, static_cast<unsigned long long>(std::log10(static_cast<double>(steps + 1)) - 1)));
    every = std::max(every, 1ULL);
    every = std::min(every, log_dump_steps_param);
  }
  return every == 0 ? false : steps % every == 0;
}

std::vector<std::pair<double, double>> read_time_temperature(const std::string &filename) {
  // pred::TimeTemperatureInterpolator (pred/src/TimeTemperatureInterpolator.cpp:10-27): skip lines until one starts with '0'
  std::vector<std::pair<double, double>> points;
  if (filename.empty()) return points;
  std::ifstream ifs(filename);
  if (!ifs.is_open()) throw std::runtime_error("Cannot open " + filename);
  while (ifs.good() && ifs.peek() != '0') ifs.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
  double t, temp;
  while (ifs >> t >> temp) {
    points.emplace_back(t, temp);
    ifs.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
  }
  std::sort(points.begin(), points.end());
  return points;
}

struct EngineHandle {
  lmc_engine *e{nullptr};
  ~EngineHandle() { lmc_engine_destroy(e); }
};

std::vector<int32_t> element_codes(const std::vector<std::string> &names) {
  std::vector<int32_t> out;
  for (const auto &s : names) out.push_back(element_from_string(s));
  return out;
}
int solvent_of(const HostConfig &c) {   // Config::GetSolventElement (Config.cpp:414-425): most frequent element
  std::map<std::string, size_t> count;
  for (auto e : c.element_of_atom) count[kNames[e]]++;
  std::string best = "X";
  size_t most = 0;
  for (const auto &kv : count)
    if (kv.second > most) { most = kv.second; best = kv.first; }
  return element_from_string(best);
}
unsigned long long make_seed(const Parameter &p) {
  return p.seed_given ? p.seed : static_cast<unsigned long long>(std::chrono::system_clock::now().time_since_epoch().count());
}

// ------------------------------------------------------------------------------------------------ KineticMcFirstOmp
void run_kmc(const Parameter &p, bool second_order) {
  // api/src/Home.cpp:133-138: a map file replaces the .cfg (lattice.txt / element.txt are fixed names in the working directory)
  HostConfig config = p.map_filename.empty() ? HostConfig::read(p.config_filename) : HostConfig::read_map("lattice.txt", "element.txt", p.map_filename);
  std::cout << "Finish config reading. Start KMC." << std::endl;
  const auto elements = element_codes(p.element_set);
  const int solvent = solvent_of(config);
  EngineHandle eng;
  check(lmc_engine_create(&eng.e, config.factors, config.id_order, elements.data(), static_cast<int32_t>(elements.size()),
                          std::find(elements.begin(), elements.end(), solvent) != elements.end() ? solvent : 0, 1, p.device));
  check(lmc_engine_load_coefficients(eng.e, p.json_coefficients_filename.c_str()));
  const auto occ = config.occupancy();
  check(lmc_engine_set_occupancy(eng.e, 0, occ.data(), static_cast<int64_t>(occ.size())));
  check(lmc_kmc_reset(eng.e));
  if (p.restart_steps > 0) {
    // McAbstract.cpp:24-30: a restarted run resumes the clocks; on the device the time feeds T(t) / the rate corrector and
    // the step number is the Philox counter (the resumed run continues the random stream instead of repeating it)
    const double t0 = p.restart_time, e0 = p.restart_energy;
    const int64_t s0 = static_cast<int64_t>(p.restart_steps);
    check(lmc_kmc_set_state(eng.e, &t0, &e0, &s0));
  }
  double absolute_energy = 0;
  check(lmc_total_energy(eng.e, 0, &absolute_energy, nullptr, 0));   // McAbstract.cpp:26
  const auto tt = read_time_temperature(p.time_temperature_filename);
  std::vector<double> tt_t, tt_v;
  for (const auto &pt : tt) { tt_t.push_back(pt.first); tt_v.push_back(pt.second); }
  std::vector<double> ru1, ru2;
  if (!p.replay_uniforms_filename.empty()) {
    std::ifstream ifs(p.replay_uniforms_filename);
    if (!ifs) throw std::runtime_error("Cannot open " + p.replay_uniforms_filename);
    if (second_order) for (double a; ifs >> a;) ru2.push_back(a);               // one selecting uniform per step
    else for (double a, b; ifs >> a >> b;) { ru1.push_back(a); ru2.push_back(b); }
  }
  const bool restarted = p.restart_steps > 0;
  bool skip_first_dump = restarted;
  std::ofstream log("kmc_log.txt", restarted ? std::ofstream::app : std::ofstream::out);
  log.precision(16);
  unsigned long long steps = p.restart_steps;
  double time = p.restart_time, energy = p.restart_energy, temperature = p.temperature;
  int64_t vacancy = -1;
  for (size_t l = 0; l < config.n(); ++l)
    if (config.element_at_lattice(static_cast<int64_t>(l)) == 0) { vacancy = static_cast<int64_t>(l); break; }
  if (vacancy < 0) throw std::runtime_error("vacancy not found");
  double total_solute_mass = 0;
  for (auto e : config.element_of_atom)
    if (e != solvent) total_solute_mass += kMass[e];
  Vec3 solute_com{0, 0, 0};

  const unsigned long long total_steps = p.maximum_steps >= steps ? p.maximum_steps - steps + 1 : 0;   // while (steps_ <= maximum_steps_)
  const unsigned long long chunk_max = 1ULL << 16;
  const unsigned long long seed = make_seed(p);
  std::vector<int64_t> from, to;
  std::vector<int32_t> slot;
  std::vector<double> dt, Ea, dE, temp_trace;
  bool escaped = false;
  for (unsigned long long done = 0; done < total_steps && !escaped;) {
    const unsigned long long chunk = std::min(chunk_max, total_steps - done);
    from.resize(chunk); to.resize(chunk); slot.resize(chunk); dt.resize(chunk); Ea.resize(chunk); dE.resize(chunk); temp_trace.resize(chunk);
    lmc_kmc_params prm{};
    prm.temperature = p.temperature;
    prm.n_time_temperature = static_cast<int32_t>(tt_t.size());
    prm.tt_time = tt_t.data();
    prm.tt_temperature = tt_v.data();
    prm.rate_corrector = p.rate_corrector ? 1 : 0;
    prm.seed = seed;
    lmc_kmc_trace tr{};
    tr.from = from.data(); tr.to = to.data(); tr.slot = slot.data(); tr.dt = dt.data(); tr.Ea = Ea.data(); tr.dE = dE.data();
    tr.temperature = temp_trace.data();
    const double *u1 = nullptr, *u2 = nullptr;
    if (!ru2.empty()) {
      if (ru2.size() < done + chunk) throw std::runtime_error("replay_uniforms_filename holds too few rows");
      u1 = second_order ? nullptr : ru1.data() + done;
      u2 = ru2.data() + done;
    }
    // KineticMcChainOmpi differs from KineticMcFirstOmp only in BuildEventList / CalculateTime (mc/src/KineticMcChainOmpi.cpp:56-152);
    // Dump, IsEscaped and the state update are those of KineticMcFirstAbstract
    if (second_order) check(lmc_kmc_chain_run(eng.e, &prm, static_cast<int64_t>(chunk), u2, &tr));
    else check(lmc_kmc_run(eng.e, &prm, static_cast<int64_t>(chunk), u1, u2, &tr));
    // host replay of the chunk: IsEscaped, Dump, then the state update -- the order of OneStepSimulation (:140-182)
    for (unsigned long long s = 0; s < chunk; ++s) {
      temperature = temp_trace[s];
      if (p.early_stop) {   // KineticMcFirstAbstract::IsEscaped (:192-223): all 42 neighbours of the vacancy are solvent
        int64_t nbr[42];
        check(lmc_engine_neighbors(eng.e, 1, vacancy, nbr));
        check(lmc_engine_neighbors(eng.e, 2, vacancy, nbr + 12));
        check(lmc_engine_neighbors(eng.e, 3, vacancy, nbr + 18));
        bool all_solvent = true;
        for (int64_t l : nbr) all_solvent = all_solvent && config.element_at_lattice(l) == solvent;
        if (all_solvent) {
          config.write("escaped.cfg.gz");
          std::cout << "t_exit: " << time << std::endl;
          std::cout << "steps: " << steps << std::endl;
          escaped = true;
        }
      }
      // Dump (:59-104).  IsEscaped has already set steps_ = maximum_steps_ + 1 when the vacancy escaped (:219), so the dump of the
      // escape step is taken with that step number
      const unsigned long long dump_steps = escaped ? p.maximum_steps + 1 : steps;
      if (skip_first_dump) {
        skip_first_dump = false;
      } else {
        if (dump_steps == 0) {
          log << "steps\ttime\ttemperature\tenergy\tEa\tdE\tselected\tvac1\tvac2\tvac3";
          if (p.solute_disp) log << "\tsolute_com1\tsolute_com2\tsolute_com3";
          log << std::endl;
        }
        if (p.config_dump_steps && dump_steps % p.config_dump_steps == 0) config.write(std::to_string(dump_steps) + ".cfg.gz");
        if (dump_steps == p.maximum_steps) config.write("end.cfg.gz");
        if (log_this_step(dump_steps, p.log_dump_steps)) {
          const Vec3 v = config.unwrapped_position_of_lattice(vacancy);
          log << dump_steps << '\t' << time << '\t' << temperature << '\t' << energy << '\t' << Ea[s] << '\t' << dE[s] << '\t'
              << config.lattice_to_atom[static_cast<size_t>(to[s])] << '\t';
          log << std::fixed << v[0] << ' ' << v[1] << ' ' << v[2];   // operator<<(Vector_d) switches the stream to fixed for good
          if (p.solute_disp) log << '\t' << solute_com[0] << ' ' << solute_com[1] << ' ' << solute_com[2];
          log << std::endl;
        }
      }
      if (escaped) break;   // the reference sets steps_ = maximum_steps_ + 1 after this Dump and still applies the event below
      time += dt[s];
      energy += dE[s];
      absolute_energy += dE[s];
      if (p.solute_disp) {
        const uint8_t jumping = config.element_at_lattice(to[s]);
        if (jumping != solvent) {
          Vec3 d{};
          for (int k = 0; k < 3; ++k) {
            d[k] = config.rel_of_lattice[static_cast<size_t>(to[s])][k] - config.rel_of_lattice[static_cast<size_t>(vacancy)][k];
            while (d[k] >= 0.5) d[k] -= 1;
            while (d[k] < -0.5) d[k] += 1;
          }
          const Vec3 disp = config.times_basis(d);
          for (int k = 0; k < 3; ++k) solute_com[k] -= disp[k] * (kMass[jumping] / total_solute_mass);
        }
      }
      config.lattice_jump(vacancy, to[s]);
      ++steps;
      vacancy = to[s];
    }
    done += chunk;
  }
  (void)absolute_energy;
}

// ------------------------------------------------------------------------------------------------ CanonicalMc* / SimulatedAnnealing
// Batched drivers: log rows are written when the step counter crosses a logging step (the batch that crosses it has
// already been applied, so the row carries the state at most one batch later than the reference's).
void write_relabelled(HostConfig &config, const std::vector<uint8_t> &occ, const std::string &filename) {
  // species are interchangeable between atoms of one kind: after a batched run atom identities are re-assigned so that
  // atom a sits on lattice site a's original position holder; the dump is a valid .cfg of the final occupancy.
  for (size_t l = 0; l < occ.size(); ++l) config.element_of_atom[static_cast<size_t>(config.lattice_to_atom[l])] = occ[l];
  config.write(filename);
}

void run_swap_driver(const Parameter &p, bool annealing) {
  HostConfig config;
  std::vector<int32_t> elements;
  int solvent;
  unsigned long long seed = make_seed(p);
  if (annealing) {
    // SimulatedAnnealing constructor (mc/src/SimulatedAnnealing.cpp:30-79): own supercell, solutes >= 4NN apart
    solvent = element_from_string(p.solvent_element);
    config = HostConfig::generate_fcc(p.factor, solvent);
    elements.push_back(solvent);
    for (const auto &s : p.solute_element_set) elements.push_back(element_from_string(s));
  } else {
    config = p.map_filename.empty() ? HostConfig::read(p.config_filename) : HostConfig::read_map("lattice.txt", "element.txt", p.map_filename);
    std::cout << "Finish config reading. Start CMC." << std::endl;
    elements = element_codes(p.element_set);
    solvent = solvent_of(config);
  }
  EngineHandle eng;
  check(lmc_engine_create(&eng.e, config.factors, config.id_order, elements.data(), static_cast<int32_t>(elements.size()),
                          std::find(elements.begin(), elements.end(), solvent) != elements.end() ? solvent : 0, 1, p.device));
  check(lmc_engine_load_coefficients(eng.e, p.json_coefficients_filename.c_str()));
  if (annealing) {
    // cfg::GenerateSoluteConfigFromExcitingPure (cfg/src/Config.cpp:1097-1137): each solute on a random site whose 1-3NN
    // shells hold no other solute
    std::mt19937_64 gen(seed);
    std::vector<char> blocked(config.n(), 0);
    for (size_t k = 0; k < p.solute_element_set.size() && k < p.solute_number_set.size(); ++k) {
      for (size_t it = 0; it < p.solute_number_set[k]; ++it) {
        int64_t pick = -1;
        for (int tries = 0; tries < 10000; ++tries) {
          const int64_t l = static_cast<int64_t>(std::uniform_int_distribution<size_t>(0, config.n() - 1)(gen));
          if (!blocked[static_cast<size_t>(l)]) { pick = l; break; }
        }
        if (pick < 0) { std::cerr << "Size is too small. Cannot generate correct config.\n"; break; }
        config.element_of_atom[static_cast<size_t>(config.lattice_to_atom[static_cast<size_t>(pick)])] =
            static_cast<uint8_t>(element_from_string(p.solute_element_set[k]));
        blocked[static_cast<size_t>(pick)] = 1;
        int64_t nbr[42];
        check(lmc_engine_neighbors(eng.e, 1, pick, nbr));
        check(lmc_engine_neighbors(eng.e, 2, pick, nbr + 12));
        check(lmc_engine_neighbors(eng.e, 3, pick, nbr + 18));
        for (int64_t l : nbr) blocked[static_cast<size_t>(l)] = 1;
      }
    }
  }
  auto occ = config.occupancy();
  check(lmc_engine_set_occupancy(eng.e, 0, occ.data(), static_cast<int64_t>(occ.size())));
  double absolute_energy0 = 0;
  check(lmc_total_energy(eng.e, 0, &absolute_energy0, nullptr, 0));
  double energy0 = p.restart_energy;
  if (annealing) {
    // energy_ = [E(config) - E(pure solvent)] - sum_e mu_e * count_e with mu from 15^3 reference cells
    // (SimulatedAnnealing.cpp:58-70, EnergyPredictor::GetChemicalPotential :196-214)
    int32_t mu_el[16];
    double mu[16];
    const int32_t n_mu = lmc_chemical_potential(eng.e, solvent, mu_el, mu, 16);
    check(n_mu);
    double solution = 0, e_pure = 0;
    for (size_t k = 0; k < p.solute_element_set.size() && k < p.solute_number_set.size(); ++k) {
      const int el = element_from_string(p.solute_element_set[k]);
      for (int32_t q = 0; q < n_mu; ++q)
        if (mu_el[q] == el) solution += mu[q] * static_cast<double>(p.solute_number_set[k]);
    }
    std::vector<uint8_t> all_solvent(config.n(), static_cast<uint8_t>(solvent));
    check(lmc_engine_set_occupancy(eng.e, 0, all_solvent.data(), static_cast<int64_t>(all_solvent.size())));
    check(lmc_total_energy(eng.e, 0, &e_pure, nullptr, 0));
    check(lmc_engine_set_occupancy(eng.e, 0, occ.data(), static_cast<int64_t>(occ.size())));
    energy0 = (absolute_energy0 - e_pure) - solution;
    std::cout << "initial_energy = " << energy0 << std::endl;
  }
  check(lmc_cmc_reset(eng.e, annealing ? p.initial_temperature : 0.0, annealing ? p.maximum_steps : 0));
  const bool restarted = !annealing && p.restart_steps > 0;
  std::ofstream log(annealing ? "sa_log.txt" : "cmc_log.txt", restarted ? std::ofstream::app : std::ofstream::out);
  log.precision(16);
  ThermodynamicAveraging averaging(annealing ? 0 : static_cast<size_t>(std::min<unsigned long long>(p.thermodynamic_averaging_steps, 1ULL << 24)));
  const unsigned long long start_steps = annealing ? 0 : p.restart_steps;
  double lowest_energy = energy0;
  lmc_cmc_params prm{};
  prm.temperature = annealing ? p.initial_temperature : p.temperature;
  prm.seed = seed;
  auto read_state = [&](double &energy, unsigned long long &steps, double &temperature) {
    double e = 0, t = 0;
    int64_t s = 0;
    check(lmc_cmc_get_state(eng.e, &e, &s, nullptr, &t));
    energy = energy0 + e;
    steps = start_steps + static_cast<unsigned long long>(s);
    temperature = t;
  };
  auto log_row = [&](unsigned long long steps, double temperature, double energy) {
    if (annealing) {
      if (steps == 0) log << "steps\ttemperature\tenergy\tlowest_energy\tabsolute_energy" << std::endl;
      log << steps << '\t' << temperature << '\t' << energy << '\t' << lowest_energy << '\t' << absolute_energy0 + (energy - energy0) << std::endl;
    } else {
      if (steps == 0) log << "steps\ttemperature\tenergy\taverage_energy\tabsolute_energy" << std::endl;
      log << steps << '\t' << temperature << '\t' << energy << '\t' << averaging.GetThermodynamicAverage(1.0 / kBoltzmann / temperature) << '\t'
          << absolute_energy0 + (energy - energy0) << std::endl;
    }
  };
  if (!annealing && !p.replay_trials_filename.empty()) {
    // Replay of the reference's serial trial stream (CanonicalMcSerial::Simulate, mc/src/CanonicalMcSerial.cpp:40-51): per
    // step AddEnergy(energy_), Dump(), then the trial -- so the sliding window of ThermodynamicAveraging holds consecutive
    // steps, the log rows and the .cfg dumps (atom identities through Config::LatticeJump) are the reference's own.
    std::vector<int64_t> ta, tb;
    std::vector<double> tu;
    {
      std::ifstream ifs(p.replay_trials_filename);
      if (!ifs) throw std::runtime_error("Cannot open " + p.replay_trials_filename);
      long long a, b;
      for (double u; ifs >> a >> b >> u;) { ta.push_back(a); tb.push_back(b); tu.push_back(u); }
    }
    const unsigned long long n_total = p.maximum_steps >= start_steps ? p.maximum_steps - start_steps + 1 : 0;
    if (ta.size() < n_total) throw std::runtime_error("replay_trials_filename holds too few rows");
    bool skip_first_dump = restarted;
    unsigned long long steps_r = start_steps;
    double absolute_energy = absolute_energy0;
    std::vector<double> dE(4096), e_before(4096), t_before(4096);
    std::vector<uint8_t> accepted(4096);
    for (unsigned long long done = 0; done < n_total;) {
      const size_t chunk = static_cast<size_t>(std::min<unsigned long long>(4096, n_total - done));
      check(lmc_cmc_replay(eng.e, 0, &prm, static_cast<int64_t>(chunk), ta.data() + done, tb.data() + done, tu.data() + done, dE.data(),
                           e_before.data(), t_before.data(), accepted.data()));
      for (size_t s = 0; s < chunk; ++s, ++steps_r) {
        const double energy_now = energy0 + e_before[s];
        averaging.AddEnergy(energy_now);
        if (skip_first_dump) skip_first_dump = false;
        else {                                                  // CanonicalMcAbstract::Dump (mc/src/CanonicalMcAbstract.cpp:53-84)
          if (steps_r == 0) log << "steps\ttemperature\tenergy\taverage_energy\tabsolute_energy" << std::endl;
          if (p.config_dump_steps && steps_r % p.config_dump_steps == 0) config.write(std::to_string(steps_r) + ".cfg.gz");
          if (steps_r == p.maximum_steps) config.write("end.cfg.gz");
          if (log_this_step(steps_r, p.log_dump_steps))
            log << steps_r << '\t' << p.temperature << '\t' << energy_now << '\t' << averaging.GetThermodynamicAverage(1.0 / kBoltzmann / p.temperature)
                << '\t' << absolute_energy << std::endl;
        }
        if (accepted[s]) {
          config.lattice_jump(ta[done + s], tb[done + s]);
          absolute_energy += dE[s];
        }
      }
      done += chunk;
    }
    return;
  }
  double energy, temperature;
  unsigned long long steps;
  read_state(energy, steps, temperature);
  if (!restarted) {
    averaging.AddEnergy(energy);
    log_row(steps, temperature, energy);
    if (!annealing && p.config_dump_steps) config.write(std::to_string(steps) + ".cfg.gz");
  }
  const unsigned long long last = p.maximum_steps;   // while (steps_ <= maximum_steps_)
  unsigned long long next_cfg = p.config_dump_steps ? (steps / p.config_dump_steps + 1) * p.config_dump_steps : ~0ULL;
  while (steps < last) {
    // run to the next step the reference would log (its cadence is log-spaced up to 10 * log_dump_steps)
    unsigned long long target = steps + 1;
    while (target <= last && !log_this_step(target, p.log_dump_steps) && !(annealing && target % std::max(1ULL, p.log_dump_steps) == 0)) ++target;
    if (p.domain_edge > 0) {
      // swap partners drawn inside randomly shifted box domains; advances in whole sweeps (domains x rounds_per_sweep trials),
      // which is then the granularity of the log rows
      lmc_cmc_domain_params dom{};
      dom.domain_edge = p.domain_edge;
      dom.rounds_per_sweep = p.rounds_per_sweep;
      check(lmc_cmc_domain_run(eng.e, &prm, &dom, static_cast<int64_t>(target - steps)));
    } else {
      check(lmc_cmc_run(eng.e, &prm, static_cast<int64_t>(target - steps)));
    }
    read_state(energy, steps, temperature);
    averaging.AddEnergy(energy);
    if (annealing && energy < lowest_energy - kEpsilon) {
      lowest_energy = energy;
      check(lmc_engine_get_occupancy(eng.e, 0, occ.data(), static_cast<int64_t>(occ.size())));
      write_relabelled(config, occ, "lowest_energy.cfg.gz");
    }
    log_row(std::min(steps, last), temperature, energy);
    if (!annealing && steps >= next_cfg) {
      check(lmc_engine_get_occupancy(eng.e, 0, occ.data(), static_cast<int64_t>(occ.size())));
      write_relabelled(config, occ, std::to_string(next_cfg) + ".cfg.gz");
      next_cfg += p.config_dump_steps;
    }
  }
  check(lmc_engine_get_occupancy(eng.e, 0, occ.data(), static_cast<int64_t>(occ.size())));
  write_relabelled(config, occ, "end.cfg.gz");
}

}  // namespace

int main(int argc, char **argv) {
  std::cout << "lmc_b200 (B200-native LatticeMC hot path), compiled on " << __DATE__ << " at " << __TIME__ << std::endl;
  if (argc <= 2) {
    std::cout << "No input parameter filename." << std::endl;
    return 1;
  }
  try {
    Parameter p;
    p.parse_args(argc, argv);
    p.read(p.parameters_filename);
    std::cout << "Parameters\nsimulation_method: " << p.method << std::endl;
    if (p.method == "KineticMcFirstOmp" || p.method == "KineticMcFirstMpi") {
      run_kmc(p, false);   // the MPI variant only splits the same 12 events over 12 ranks (mc/src/KineticMcFirstMpi.cpp:46-76)
    } else if (p.method == "KineticMcChainOmpi") {
      run_kmc(p, true);    // second-order KMC; the reference's 12 ranks are the 12 half-warps of one block
    } else if (p.method == "CanonicalMcSerial" || p.method == "CanonicalMcOmp") {
      run_swap_driver(p, false);
    } else if (p.method == "SimulatedAnnealing") {
      run_swap_driver(p, true);
    } else if (p.method == "Ansys" || p.method == "Reformat") {
      std::cout << "simulation_method " << p.method << " is outside the accelerated hot path of this engine" << std::endl;
      return 2;
    } else {
      std::cout << "No such method: " << p.method << std::endl;   // api/src/Home.cpp:123
    }
  } catch (const std::exception &e) {
    std::cerr << "error: " << e.what() << std::endl;
    return 1;
  }
  return 0;
}

// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin C-ABI wrapper that links the UNMODIFIED reference sources where they lie under
// /root/reference/lmc/{cfg,pred,mc} (compiled by oracle/Makefile against oracle/shims/) into
// oracle/_ref/liblmc_ref.so.  It exists so that tests, golden-vector generation and bench.py's
// cpu_baseline / `--impl reference` leg can call the reference's own predictors and drivers:
//   cfg::Config, pred::VacancyMigrationPredictorQuartic[Lru], pred::EnergyChangePredictorPairSite,
//   pred::EnergyPredictor, mc::KineticMcFirstOmp, mc::CanonicalMcSerial/Omp, mc::SimulatedAnnealing.
// Nothing in the product path (latticemontecarlo_b200/) may link or load this library.
//
// Everything here is new code written for this repository; no reference source is copied.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#include <omp.h>

#include "Config.h"
#include "EnergyUtility.h"
#include "EnergyPredictor.h"
#include "EnergyChangePredictorPairSite.h"
#include "VacancyMigrationPredictorQuartic.h"
#include "VacancyMigrationPredictorQuarticLru.h"
#include "VacancyMigrationPredictorE0.h"
#include "VacancyMigrationPredictorE0Lru.h"
#include "EnergyChangePredictorPair.h"
#include "EnergyChangePredictorSite.h"
#include "TimeTemperatureInterpolator.h"
#include "RateCorrector.hpp"
#include "KineticMcFirstOmp.h"
#include "KineticMcChainOmpi.h"
#include "CanonicalMcSerial.h"
// CanonicalMcOmp keeps its batch (event_vector_, BuildEventVector) private and offers no hook between the batch build and
// the per-event accept loop; the traced driver below re-runs CanonicalMcOmp::Simulate's loop with the class's own
// building blocks, which needs access to them.  Access specifiers do not change the layout, and every header
// CanonicalMcOmp.h pulls in is already included (guarded) above, so the redefinition touches that one class only.
#define private public
#include "CanonicalMcOmp.h"
#undef private
#include "SimulatedAnnealing.h"

namespace {

thread_local std::string g_last_error;

int fail(const std::exception &e) {
  g_last_error = e.what();
  return -1;
}

Element element_from_code(int code) { return Element(static_cast<ElementName>(code)); }
int code_from_element(Element e) { return static_cast<int>(static_cast<ElementName>(e)); }

std::set<Element> element_set_from_codes(const int *codes, int n) {
  std::set<Element> s;
  for (int i = 0; i < n; ++i) s.insert(element_from_code(codes[i]));
  return s;
}

// RAII chdir so the reference's hard-wired log/cfg file names land in a scratch directory.
struct ScopedChdir {
  std::string old_;
  explicit ScopedChdir(const char *dir) {
    char buf[4096];
    if (getcwd(buf, sizeof buf)) old_ = buf;
    if (dir && *dir && chdir(dir) != 0) throw std::runtime_error(std::string("cannot chdir to ") + dir);
  }
  ~ScopedChdir() {
    if (!old_.empty()) { int r = chdir(old_.c_str()); (void)r; }
  }
};

// silence std::cout chatter ("Using N threads.") from the reference constructors
struct ScopedQuietCout {
  std::streambuf *old_;
  std::ostringstream sink_;
  ScopedQuietCout() : old_(std::cout.rdbuf(sink_.rdbuf())) {}
  ~ScopedQuietCout() { std::cout.rdbuf(old_); }
};

int64_t flatten_mapping(const std::vector<std::vector<std::vector<size_t>>> &mapping, int64_t *out, int64_t cap) {
  // layout: G, then per group: C, L, then C*L entries (SIZE_MAX -> -1)
  std::vector<int64_t> flat;
  flat.push_back(static_cast<int64_t>(mapping.size()));
  for (const auto &group : mapping) {
    flat.push_back(static_cast<int64_t>(group.size()));
    flat.push_back(group.empty() ? 0 : static_cast<int64_t>(group[0].size()));
    for (const auto &cluster : group)
      for (auto v : cluster) flat.push_back(v == SIZE_MAX ? -1 : static_cast<int64_t>(v));
  }
  if (out) {
    const int64_t n = std::min<int64_t>(cap, static_cast<int64_t>(flat.size()));
    std::memcpy(out, flat.data(), static_cast<size_t>(n) * sizeof(int64_t));
  }
  return static_cast<int64_t>(flat.size());
}

// --- predictor subclass exposing the reference's protected pieces -------------------------------
class QuarticProbe : public pred::VacancyMigrationPredictorQuartic {
 public:
  using pred::VacancyMigrationPredictorQuartic::VacancyMigrationPredictorQuartic;
  using pred::VacancyMigrationPredictorQuartic::GetDe;
  using pred::VacancyMigrationPredictorQuartic::GetD;
  using pred::VacancyMigrationPredictorQuartic::GetKs;
  using pred::VacancyMigrationPredictorQuartic::GetPairFlatIndex;
  const std::vector<std::vector<std::vector<size_t>>> &mapping(int which) const {
    return which == 0 ? mapping_state_ : (which == 1 ? mapping_mmm_ : mapping_mm2_);
  }
  const std::vector<size_t> &list(int which, size_t flat) const {
    return which == 0 ? site_bond_cluster_state_flat_.at(flat)
                      : (which == 1 ? site_bond_cluster_mmm_flat_.at(flat) : site_bond_cluster_mm2_flat_.at(flat));
  }
  const std::set<Element> &element_set() const { return element_set_; }
  const std::unordered_map<std::string, std::vector<double>> &one_hot() const { return one_hot_encode_hash_map_; }
};

// --- traced drivers ----------------------------------------------------------------------------
struct KmcTrace {
  int64_t cap{0}, n{0};
  double *u1{nullptr}, *u2{nullptr}, *dt{nullptr}, *time{nullptr}, *energy{nullptr}, *Ea{nullptr}, *dE{nullptr},
      *temperature{nullptr}, *total_rate{nullptr};
  int64_t *from{nullptr}, *to{nullptr}, *slot{nullptr};
};

class TracedKmcFirstOmp : public mc::KineticMcFirstOmp {
 public:
  using mc::KineticMcFirstOmp::KineticMcFirstOmp;
  void Reseed(uint64_t seed) { generator_.seed(seed); }
  void SetTrace(KmcTrace *t) { trace_ = t; }
  const cfg::Config &config() const { return config_; }
  double absolute_energy() const { return absolute_energy_; }
  double energy() const { return energy_; }
  double time() const { return time_; }
  unsigned long long steps() const { return steps_; }

 protected:
  void Dump() const override {}  // no log / cfg I/O while tracing or timing
  void OneStepSimulation() override {
    if (!trace_ || trace_->n >= trace_->cap) {
      mc::KineticMcFirstAbstract::OneStepSimulation();
      return;
    }
    auto g = generator_;  // peek the two uniforms this step will consume (u1: time, u2: select)
    std::uniform_real_distribution<double> d(0.0, 1.0);
    const double u1 = d(g), u2 = d(g);
    const size_t from = vacancy_lattice_id_;
    const double t0 = time_;
    mc::KineticMcFirstAbstract::OneStepSimulation();
    const int64_t k = trace_->n++;
    const size_t to = vacancy_lattice_id_;
    if (trace_->u1) trace_->u1[k] = u1;
    if (trace_->u2) trace_->u2[k] = u2;
    if (trace_->from) trace_->from[k] = static_cast<int64_t>(from);
    if (trace_->to) trace_->to[k] = static_cast<int64_t>(to);
    if (trace_->slot) {
      const auto &nn = config_.GetFirstNeighborsAdjacencyList()[from];
      int64_t s = -1;
      for (size_t q = 0; q < nn.size(); ++q)
        if (nn[q] == to) s = static_cast<int64_t>(q);
      trace_->slot[k] = s;
    }
    if (trace_->dt) trace_->dt[k] = time_ - t0;
    if (trace_->time) trace_->time[k] = time_;
    if (trace_->energy) trace_->energy[k] = energy_;
    if (trace_->Ea) trace_->Ea[k] = event_k_i_.GetForwardBarrier();
    if (trace_->dE) trace_->dE[k] = event_k_i_.GetEnergyChange();
    if (trace_->temperature) trace_->temperature[k] = temperature_;
    if (trace_->total_rate) trace_->total_rate[k] = total_rate_k_;
  }

 private:
  KmcTrace *trace_{nullptr};
};

// mc::KineticMcChainOmpi needs exactly 12 MPI ranks (KineticMcChainOmpi.cpp:37-42); the harness runs them as 12 threads
// over the threads-as-ranks shim (shims/mpi.h).  Only rank 0's generator decides (SelectEvent broadcasts its choice,
// KineticMcAbstract.cpp:116-120) and consumes ONE uniform per step (the second-order time is not sampled).
class TracedKmcChainOmpi : public mc::KineticMcChainOmpi {
 public:
  using mc::KineticMcChainOmpi::KineticMcChainOmpi;
  void Reseed(uint64_t seed) { generator_.seed(seed); }
  void SetTrace(KmcTrace *t) { trace_ = t; }
  const cfg::Config &config() const { return config_; }
  double absolute_energy() const { return absolute_energy_; }
  double energy() const { return energy_; }
  double time() const { return time_; }
  unsigned long long steps() const { return steps_; }

 protected:
  void Dump() const override {}
  void OneStepSimulation() override {
    if (!trace_ || trace_->n >= trace_->cap) {
      mc::KineticMcChainAbstract::OneStepSimulation();
      return;
    }
    auto g = generator_;
    std::uniform_real_distribution<double> d(0.0, 1.0);
    const double u = d(g);
    const size_t from = vacancy_lattice_id_;
    const double t0 = time_;
    mc::KineticMcChainAbstract::OneStepSimulation();
    const int64_t k = trace_->n++;
    const size_t to = vacancy_lattice_id_;
    if (trace_->u1) trace_->u1[k] = 0.0;
    if (trace_->u2) trace_->u2[k] = u;
    if (trace_->from) trace_->from[k] = static_cast<int64_t>(from);
    if (trace_->to) trace_->to[k] = static_cast<int64_t>(to);
    if (trace_->slot) {
      const auto &nn = config_.GetFirstNeighborsAdjacencyList()[from];
      int64_t s = -1;
      for (size_t q = 0; q < nn.size(); ++q)
        if (nn[q] == to) s = static_cast<int64_t>(q);
      trace_->slot[k] = s;
    }
    if (trace_->dt) trace_->dt[k] = time_ - t0;
    if (trace_->time) trace_->time[k] = time_;
    if (trace_->energy) trace_->energy[k] = energy_;
    if (trace_->Ea) trace_->Ea[k] = event_k_i_.GetForwardBarrier();
    if (trace_->dE) trace_->dE[k] = event_k_i_.GetEnergyChange();
    if (trace_->temperature) trace_->temperature[k] = temperature_;
    if (trace_->total_rate) trace_->total_rate[k] = total_rate_k_;
  }

 private:
  KmcTrace *trace_{nullptr};
};

// the driver exactly as shipped (Dump() writes kmc_log.txt and N.cfg.gz), only with a reproducible seed
class SeededKmcFirstOmp : public mc::KineticMcFirstOmp {
 public:
  using mc::KineticMcFirstOmp::KineticMcFirstOmp;
  void Reseed(uint64_t seed) { generator_.seed(seed); }
};
class SeededKmcChainOmpi : public mc::KineticMcChainOmpi {
 public:
  using mc::KineticMcChainOmpi::KineticMcChainOmpi;
  void Reseed(uint64_t seed) { generator_.seed(seed); }
};
class SeededCmcSerial : public mc::CanonicalMcSerial {
 public:
  using mc::CanonicalMcSerial::CanonicalMcSerial;
  void Reseed(uint64_t seed) { generator_.seed(seed); }
};

struct SwapTrace {
  int64_t cap{0}, n{0};
  int64_t *a{nullptr}, *b{nullptr};
  double *energy_before{nullptr}, *temperature{nullptr}, *u{nullptr};
};

// Peek the pair GenerateLatticeIdJumpPair() is about to draw (same distribution type/params, copied engine).
// Also returns (in *u_next) the uniform real that SelectEvent would consume right after (only used if dE >= 0).
template <class Gen>
std::pair<size_t, size_t> PeekPair(Gen g, const cfg::Config &config, double *u_next) {
  std::uniform_int_distribution<size_t> sel(0, config.GetNumAtoms() - 1);
  size_t a, b;
  do {
    a = sel(g);
    b = sel(g);
  } while (config.GetElementAtLatticeId(a) == config.GetElementAtLatticeId(b));
  if (u_next) {
    std::uniform_real_distribution<double> d(0.0, 1.0);
    *u_next = d(g);
  }
  return {a, b};
}

class TracedCmcSerial : public mc::CanonicalMcSerial {
 public:
  using mc::CanonicalMcSerial::CanonicalMcSerial;
  void Reseed(uint64_t seed) { generator_.seed(seed); }
  void SetTrace(SwapTrace *t) { trace_ = t; }
  const cfg::Config &config() const { return config_; }
  double energy() const { return energy_; }

 protected:
  // Dump() runs once per trial *after* the pair was drawn; so the trace is taken by re-running the
  // serial loop here with the reference's own protected building blocks (same call order as
  // CanonicalMcSerial::Simulate).
 public:
  void SimulateTraced() {
    while (steps_ <= maximum_steps_) {
      auto pair = GenerateLatticeIdJumpPair();
      auto dE = energy_change_predictor_.GetDeFromLatticeIdPair(config_, pair);
      thermodynamic_averaging_.AddEnergy(energy_);
      if (trace_ && trace_->n < trace_->cap) {
        const int64_t k = trace_->n++;
        trace_->a[k] = static_cast<int64_t>(pair.first);
        trace_->b[k] = static_cast<int64_t>(pair.second);
        trace_->energy_before[k] = energy_;
        trace_->temperature[k] = dE;  // for CMC the slot carries dE (temperature is constant)
        if (trace_->u) {
          auto g = generator_;
          std::uniform_real_distribution<double> d(0.0, 1.0);
          trace_->u[k] = d(g);  // the real SelectEvent will consume iff dE >= 0
        }
      }
      SelectEvent(pair, dE);
      ++steps_;
    }
  }
  void Dump() const override {}

 private:
  SwapTrace *trace_{nullptr};
};

class QuietCmcOmp : public mc::CanonicalMcOmp {
 public:
  using mc::CanonicalMcOmp::CanonicalMcOmp;
  void Reseed(uint64_t seed) { generator_.seed(seed); }
  const cfg::Config &config() const { return config_; }
  double energy() const { return energy_; }
  unsigned long long steps() const { return steps_; }

 protected:
  void Dump() const override {}
};

// mc::CanonicalMcOmp::Simulate (mc/src/CanonicalMcOmp.cpp:80-92) with one trace record per event: the batch is built by the
// reference's own BuildEventVector (greedy serial pass with the unavailable_position_ rule, dE evaluated for the whole batch
// on the batch-start configuration), then the events are accepted / rejected in order.
class TracedCmcOmp : public mc::CanonicalMcOmp {
 public:
  using mc::CanonicalMcOmp::CanonicalMcOmp;
  void Reseed(uint64_t seed) { generator_.seed(seed); }
  void SetTrace(SwapTrace *t, int64_t *batch_of) { trace_ = t; batch_of_ = batch_of; }
  const cfg::Config &config() const { return config_; }
  double energy() const { return energy_; }
  unsigned long long steps() const { return steps_; }
  void SimulateTraced() {
    int64_t batch = 0;
    while (steps_ <= maximum_steps_) {
      BuildEventVector();
      for (auto [pair, dE] : event_vector_) {
        thermodynamic_averaging_.AddEnergy(energy_);
        if (trace_ && trace_->n < trace_->cap) {
          const int64_t k = trace_->n++;
          trace_->a[k] = static_cast<int64_t>(pair.first);
          trace_->b[k] = static_cast<int64_t>(pair.second);
          trace_->energy_before[k] = energy_;
          trace_->temperature[k] = dE;
          if (batch_of_) batch_of_[k] = batch;
          if (trace_->u) {
            auto g = generator_;
            std::uniform_real_distribution<double> d(0.0, 1.0);
            trace_->u[k] = d(g);  // SelectEvent consumes it iff dE >= 0
          }
        }
        SelectEvent(pair, dE);
        ++steps_;
      }
      ++batch;
    }
  }

 protected:
  void Dump() const override {}

 private:
  SwapTrace *trace_{nullptr};
  int64_t *batch_of_{nullptr};
};

class TracedSa : public mc::SimulatedAnnealing {
 public:
  using mc::SimulatedAnnealing::SimulatedAnnealing;
  void Reseed(uint64_t seed) { generator_.seed(seed); }
  void SetTrace(SwapTrace *t) { trace_ = t; }
  cfg::Config &mutable_config() { return config_; }
  const cfg::Config &config() const { return config_; }
  double energy() const { return energy_; }
  double temperature() const { return temperature_; }

 private:
  // SimulatedAnnealing::Dump is a private virtual: overriding it is legal and gives one hook per
  // trial, called right before the pair is drawn (mc/src/SimulatedAnnealing.cpp:168-185).
  void Dump() const override {
    if (!trace_ || trace_->n >= trace_->cap) return;
    const int64_t k = trace_->n++;
    double u = 0.0;
    const auto pair = PeekPair(generator_, config_, &u);
    if (trace_->u) trace_->u[k] = u;
    trace_->a[k] = static_cast<int64_t>(pair.first);
    trace_->b[k] = static_cast<int64_t>(pair.second);
    trace_->energy_before[k] = energy_;
    trace_->temperature[k] = temperature_;
  }
  SwapTrace *trace_{nullptr};
};

void copy_occupancy(const cfg::Config &config, uint8_t *out) {
  if (!out) return;
  for (size_t l = 0; l < config.GetNumAtoms(); ++l)
    out[l] = static_cast<uint8_t>(code_from_element(config.GetElementAtLatticeId(l)));
}

double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

const char *ref_last_error() { return g_last_error.c_str(); }
int ref_max_threads() { return omp_get_max_threads(); }
void ref_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

// ---------------------------------------------------------------- cfg::Config
// occ: element enum codes (ElementName) indexed by cfg::GenerateFCC lattice id (cfg/src/Config.cpp:1060-1095).
// reassign != 0 reproduces a run started from a .cfg file: WriteConfig -> ReadConfig -> ReassignLatticeVector
// (api/src/Home.cpp:165-167); scratch_path is where that file is written.
void *ref_config_create(int fx, int fy, int fz, const uint8_t *occ, int reassign, const char *scratch_path) {
  try {
    auto config = cfg::GenerateFCC({static_cast<size_t>(fx), static_cast<size_t>(fy), static_cast<size_t>(fz)},
                                   Element(ElementName::Al));
    if (occ)
      for (size_t l = 0; l < config.GetNumAtoms(); ++l) config.SetAtomElementTypeAtLattice(l, element_from_code(occ[l]));
    if (reassign) {
      config.WriteConfig(scratch_path);
      config = cfg::Config::ReadConfig(scratch_path);
      config.ReassignLatticeVector();
    }
    return new cfg::Config(std::move(config));
  } catch (const std::exception &e) { fail(e); return nullptr; }
}
void *ref_config_read(const char *path, int reassign) {
  try {
    auto config = cfg::Config::ReadConfig(path);
    if (reassign) config.ReassignLatticeVector();
    return new cfg::Config(std::move(config));
  } catch (const std::exception &e) { fail(e); return nullptr; }
}
// Config::ReadMap (cfg/src/Config.cpp:814-885): lattice.txt + element.txt + map file, lattice ids as written (no reassignment)
void *ref_config_read_map(const char *lattice, const char *element, const char *map) {
  try {
    return new cfg::Config(cfg::Config::ReadMap(lattice, element, map));
  } catch (const std::exception &e) { fail(e); return nullptr; }
}
// Config::WriteLattice / WriteElement / WriteMap (:887-922)
int ref_config_write_map_files(void *h, const char *lattice, const char *element, const char *map) {
  try {
    auto *config = static_cast<cfg::Config *>(h);
    config->WriteLattice(lattice);
    config->WriteElement(element);
    config->WriteMap(map);
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}
void *ref_config_clone(void *h) { return new cfg::Config(*static_cast<cfg::Config *>(h)); }
int ref_config_write(void *h, const char *path) {
  try { static_cast<cfg::Config *>(h)->WriteConfig(path); return 0; } catch (const std::exception &e) { return fail(e); }
}
void ref_config_free(void *h) { delete static_cast<cfg::Config *>(h); }
int64_t ref_config_num_sites(void *h) { return static_cast<int64_t>(static_cast<cfg::Config *>(h)->GetNumAtoms()); }
void ref_config_get_occupancy(void *h, uint8_t *out) { copy_occupancy(*static_cast<cfg::Config *>(h), out); }
void ref_config_get_basis(void *h, double *out9) {
  const auto &b = static_cast<cfg::Config *>(h)->GetBasis();
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) out9[3 * i + j] = b[i][j];
}
void ref_config_get_positions(void *h, double *out) {  // relative positions, N x 3, by lattice id
  const auto &lv = static_cast<cfg::Config *>(h)->GetLatticeVector();
  for (size_t l = 0; l < lv.size(); ++l)
    for (int d = 0; d < 3; ++d) out[3 * l + d] = lv[l].GetRelativePosition()[d];
}
void ref_config_get_neighbors(void *h, int shell, int64_t *out) {  // N x {12,6,24}
  const auto &c = *static_cast<cfg::Config *>(h);
  const auto &adj = shell == 1 ? c.GetFirstNeighborsAdjacencyList()
                               : (shell == 2 ? c.GetSecondNeighborsAdjacencyList() : c.GetThirdNeighborsAdjacencyList());
  size_t k = 0;
  for (const auto &row : adj) for (auto v : row) out[k++] = static_cast<int64_t>(v);
}
void ref_config_get_maps(void *h, int64_t *lattice_to_atom, int64_t *atom_to_lattice, int32_t *map_shift) {
  const auto &c = *static_cast<cfg::Config *>(h);
  for (size_t i = 0; i < c.GetNumAtoms(); ++i) {
    if (lattice_to_atom) lattice_to_atom[i] = static_cast<int64_t>(c.GetAtomIdFromLatticeId(i));
    if (atom_to_lattice) atom_to_lattice[i] = static_cast<int64_t>(c.GetLatticeIdFromAtomId(i));
    if (map_shift) for (int d = 0; d < 3; ++d) map_shift[3 * i + d] = c.GetMapShiftList()[i][static_cast<size_t>(d)];
  }
}
void ref_config_lattice_jump(void *h, int64_t a, int64_t b) {
  static_cast<cfg::Config *>(h)->LatticeJump({static_cast<size_t>(a), static_cast<size_t>(b)});
}
void ref_config_set_element(void *h, int64_t lattice_id, int code) {
  static_cast<cfg::Config *>(h)->SetAtomElementTypeAtLattice(static_cast<size_t>(lattice_id), element_from_code(code));
}
int64_t ref_config_vacancy(void *h) {
  try { return static_cast<int64_t>(static_cast<cfg::Config *>(h)->GetVacancyLatticeId()); }
  catch (const std::exception &e) { fail(e); return -1; }
}

// ---------------------------------------------------------------- pred:: free functions
// cluster types in ClusterIndexer order (pred/src/EnergyUtility.cpp:314-343,798-811): rows of 5 ints
// [label, size, e1, e2, e3] (enum codes, -1 padded). Returns the number of types.
int ref_cluster_types(const int *codes, int n, int32_t *out, int cap_rows) {
  auto set = element_set_from_codes(codes, n);
  set.emplace(ElementName::X);
  const auto hashmap = pred::InitializeClusterHashMap(set);
  const std::map<cfg::ElementCluster, int> ordered(hashmap.begin(), hashmap.end());
  int row = 0;
  for (const auto &kv : ordered) {
    if (out && row < cap_rows) {
      out[5 * row + 0] = kv.first.GetLabel();
      out[5 * row + 1] = static_cast<int32_t>(kv.first.GetSize());
      for (int q = 0; q < 3; ++q)
        out[5 * row + 2 + q] = q < static_cast<int>(kv.first.GetSize())
                                   ? code_from_element(kv.first.GetElementVector()[static_cast<size_t>(q)]) : -1;
    }
    ++row;
  }
  return row;
}
// which: 0 state-pair, 1 mmm, 2 mm2, 3 state-site
int64_t ref_mapping(void *config_h, int which, int64_t *out, int64_t cap) {
  try {
    const auto &c = *static_cast<cfg::Config *>(config_h);
    switch (which) {
      case 0: return flatten_mapping(pred::GetClusterParametersMappingStatePair(c), out, cap);
      case 1: return flatten_mapping(pred::GetAverageClusterParametersMappingMMM(c), out, cap);
      case 2: return flatten_mapping(pred::GetAverageClusterParametersMappingMM2(c), out, cap);
      default: return flatten_mapping(pred::GetClusterParametersMappingStateSite(c), out, cap);
    }
  } catch (const std::exception &e) { return fail(e); }
}
// ordered neighbourhood lists straight from the reference's sorters (pred/src/EnergyUtility.cpp:45-104,261-313)
int ref_pair_lists(void *config_h, int64_t i, int64_t j, int64_t *state60, int64_t *mmm58, int64_t *mm258) {
  try {
    const auto &c = *static_cast<cfg::Config *>(config_h);
    const std::pair<size_t, size_t> p{static_cast<size_t>(i), static_cast<size_t>(j)};
    if (state60) { size_t k = 0; for (const auto &l : pred::GetSortedLatticeVectorStateOfPair(c, p)) state60[k++] = static_cast<int64_t>(l.GetId()); }
    if (mmm58) { size_t k = 0; for (const auto &l : pred::GetSymmetricallySortedLatticeVectorMMM(c, p)) mmm58[k++] = static_cast<int64_t>(l.GetId()); }
    if (mm258) { size_t k = 0; for (const auto &l : pred::GetSymmetricallySortedLatticeVectorMM2(c, p)) mm258[k++] = static_cast<int64_t>(l.GetId()); }
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}
int ref_site_list(void *config_h, int64_t i, int64_t *state43) {
  try {
    const auto &c = *static_cast<cfg::Config *>(config_h);
    size_t k = 0;
    for (const auto &l : pred::GetSortedLatticeVectorStateOfSite(c, static_cast<size_t>(i))) state43[k++] = static_cast<int64_t>(l.GetId());
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}

// ---------------------------------------------------------------- VacancyMigrationPredictorQuartic
void *ref_quartic_create(const char *json, void *config_h, const int *codes, int n, int64_t lru_size) {
  try {
    ScopedQuietCout quiet;
    const auto &c = *static_cast<cfg::Config *>(config_h);
    if (lru_size > 0)
      return static_cast<pred::VacancyMigrationPredictorQuartic *>(new pred::VacancyMigrationPredictorQuarticLru(
          json, c, element_set_from_codes(codes, n), static_cast<size_t>(lru_size)));
    return static_cast<pred::VacancyMigrationPredictorQuartic *>(new QuarticProbe(json, c, element_set_from_codes(codes, n)));
  } catch (const std::exception &e) { fail(e); return nullptr; }
}
void ref_quartic_free(void *h) { delete static_cast<pred::VacancyMigrationPredictorQuartic *>(h); }
// Public virtual entry (works for both the plain and the LRU predictor); threads>1 mimics the OMP caller.
int ref_quartic_eval(void *h, void *config_h, int64_t n, const int64_t *i, const int64_t *j, double *Ea, double *dE,
                     int threads) {
  try {
    const auto *p = static_cast<pred::VacancyMigrationPredictorQuartic *>(h);
    const auto &c = *static_cast<cfg::Config *>(config_h);
    std::string err;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int64_t k = 0; k < n; ++k) {
      try {
        const auto r = p->GetBarrierAndDiffFromLatticeIdPair(c, {static_cast<size_t>(i[k]), static_cast<size_t>(j[k])});
        Ea[k] = r.first;
        dE[k] = r.second;
      } catch (const std::exception &e) {
#pragma omp critical
        err = e.what();
      }
    }
    if (!err.empty()) throw std::runtime_error(err);
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}
// Protected pieces (plain predictor only): dE, D, Ks, integer counts, raw one-hot encodes, cached id lists.
int ref_quartic_parts(void *h, void *config_h, int64_t i, int64_t j, double *dE, double *D, double *Ks,
                      int32_t *start_counts, int32_t *end_counts, int64_t *n_types, double *enc_mmm, double *enc_mm2_f,
                      double *enc_mm2_b, int64_t *state60, int64_t *mmm58, int64_t *mm258) {
  try {
    const auto *p = dynamic_cast<QuarticProbe *>(static_cast<pred::VacancyMigrationPredictorQuartic *>(h));
    if (!p) throw std::runtime_error("ref_quartic_parts needs the non-LRU predictor");
    const auto &c = *static_cast<cfg::Config *>(config_h);
    const std::pair<size_t, size_t> pr{static_cast<size_t>(i), static_cast<size_t>(j)};
    const std::pair<size_t, size_t> rp{pr.second, pr.first};
    if (D) *D = p->GetD(c, pr);
    if (Ks) *Ks = p->GetKs(c, pr);
    const double de = p->GetDe(c, pr);  // last: leaves the count buffers filled
    if (dE) *dE = de;
    const auto &sc = pred::GetThreadLocalStartCountsBuffer();
    const auto &ec = pred::GetThreadLocalEndCountsBuffer();
    if (n_types) *n_types = static_cast<int64_t>(sc.size());
    if (start_counts) for (size_t q = 0; q < sc.size(); ++q) start_counts[q] = sc[q];
    if (end_counts) for (size_t q = 0; q < ec.size(); ++q) end_counts[q] = ec[q];
    const auto flat = p->GetPairFlatIndex(pr);
    const auto flat_r = p->GetPairFlatIndex(rp);
    auto encode = [&](const std::vector<size_t> &ids, int which, double *out) {
      if (!out) return;
      std::vector<Element> ele;
      for (auto id : ids) ele.push_back(c.GetElementAtLatticeId(id));
      std::vector<double> enc;
      pred::GetOneHotParametersFromMap(ele, p->one_hot(), p->element_set().size(), p->mapping(which), enc);
      std::memcpy(out, enc.data(), enc.size() * sizeof(double));
    };
    encode(p->list(1, flat), 1, enc_mmm);
    encode(p->list(2, flat), 2, enc_mm2_f);
    encode(p->list(2, flat_r), 2, enc_mm2_b);
    auto copy_ids = [](const std::vector<size_t> &ids, int64_t *out) {
      if (out) for (size_t q = 0; q < ids.size(); ++q) out[q] = static_cast<int64_t>(ids[q]);
    };
    copy_ids(p->list(0, flat), state60);
    copy_ids(p->list(1, flat), mmm58);
    copy_ids(p->list(2, flat), mm258);
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}

// ---------------------------------------------------------------- EnergyChangePredictorPairSite
void *ref_pairsite_create(const char *json, void *config_h, const int *codes, int n) {
  try {
    return new pred::EnergyChangePredictorPairSite(json, *static_cast<cfg::Config *>(config_h), element_set_from_codes(codes, n));
  } catch (const std::exception &e) { fail(e); return nullptr; }
}
void ref_pairsite_free(void *h) { delete static_cast<pred::EnergyChangePredictorPairSite *>(h); }
int ref_pairsite_de_pair(void *h, void *config_h, int64_t n, const int64_t *a, const int64_t *b, double *out, int threads) {
  try {
    const auto *p = static_cast<pred::EnergyChangePredictorPairSite *>(h);
    const auto &c = *static_cast<cfg::Config *>(config_h);
    std::string err;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int64_t k = 0; k < n; ++k) {
      try {
        out[k] = p->GetDeFromLatticeIdPair(c, {static_cast<size_t>(a[k]), static_cast<size_t>(b[k])});
      } catch (const std::exception &e) {
#pragma omp critical
        err = e.what();
      }
    }
    if (!err.empty()) throw std::runtime_error(err);
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}
int ref_pairsite_de_site(void *h, void *config_h, int64_t n, const int64_t *site, const uint8_t *new_code, double *out,
                         int threads) {
  try {
    const auto *p = static_cast<pred::EnergyChangePredictorPairSite *>(h);
    const auto &c = *static_cast<cfg::Config *>(config_h);
    std::string err;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
    for (int64_t k = 0; k < n; ++k) {
      try {
        out[k] = p->GetDeFromLatticeIdSite(c, static_cast<size_t>(site[k]), element_from_code(new_code[k]));
      } catch (const std::exception &e) {
#pragma omp critical
        err = e.what();
      }
    }
    if (!err.empty()) throw std::runtime_error(err);
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}
// one site change with its integer start/end counts (thread-local buffers of the reference)
int ref_pairsite_site_counts(void *h, void *config_h, int64_t site, int new_code, double *dE, int32_t *start_counts,
                             int32_t *end_counts) {
  try {
    const auto *p = static_cast<pred::EnergyChangePredictorPairSite *>(h);
    const auto &c = *static_cast<cfg::Config *>(config_h);
    const double de = p->GetDeFromLatticeIdSite(c, static_cast<size_t>(site), element_from_code(new_code));
    if (dE) *dE = de;
    const auto &sc = pred::GetThreadLocalStartCountsBuffer();
    const auto &ec = pred::GetThreadLocalEndCountsBuffer();
    if (start_counts) for (size_t q = 0; q < sc.size(); ++q) start_counts[q] = sc[q];
    if (end_counts) for (size_t q = 0; q < ec.size(); ++q) end_counts[q] = ec[q];
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}

// ---------------------------------------------------------------- EnergyPredictor
void *ref_energy_create(const char *json, const int *codes, int n) {
  try { return new pred::EnergyPredictor(json, element_set_from_codes(codes, n)); }
  catch (const std::exception &e) { fail(e); return nullptr; }
}
void ref_energy_free(void *h) { delete static_cast<pred::EnergyPredictor *>(h); }
int ref_energy_get(void *h, void *config_h, double *energy, double *encode, int cap) {
  try {
    const auto *p = static_cast<pred::EnergyPredictor *>(h);
    const auto &c = *static_cast<cfg::Config *>(config_h);
    if (encode) {
      const auto enc = p->GetEncode(c);
      for (size_t q = 0; q < enc.size() && q < static_cast<size_t>(cap); ++q) encode[q] = enc[q];
    }
    if (energy) *energy = p->GetEnergy(c);
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}

// EnergyPredictor::GetEnergyOfCluster / GetEncodeOfCluster (pred/src/EnergyPredictor.cpp:97-184) on a list of ATOM ids
int ref_energy_of_cluster(void *h, void *config_h, const int64_t *atom_ids, int64_t n, double *energy, double *encode, int cap) {
  try {
    const auto *p = static_cast<pred::EnergyPredictor *>(h);
    const auto &c = *static_cast<cfg::Config *>(config_h);
    std::vector<size_t> ids(atom_ids, atom_ids + n);
    if (encode) {
      const auto enc = p->GetEncodeOfCluster(c, ids);
      for (size_t q = 0; q < enc.size() && q < static_cast<size_t>(cap); ++q) encode[q] = enc[q];
    }
    if (energy) *energy = p->GetEnergyOfCluster(c, ids);
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}
// EnergyPredictor::GetChemicalPotential (pred/src/EnergyPredictor.cpp:196-214): entries in std::map order
int ref_chemical_potential(void *h, int solvent_code, int *codes_out, double *mu_out, int cap) {
  try {
    const auto *p = static_cast<pred::EnergyPredictor *>(h);
    const auto mu = p->GetChemicalPotential(Element(static_cast<ElementName>(solvent_code)));
    int k = 0;
    for (const auto &[el, v] : mu) {
      if (k < cap) { codes_out[k] = static_cast<int>(static_cast<ElementName>(el)); mu_out[k] = v; }
      ++k;
    }
    return k;
  } catch (const std::exception &e) { fail(e); return -1; }
}

// ---------------------------------------------------------------- helpers (pred/include/RateCorrector.hpp, TimeTemperatureInterpolator)
double ref_rate_correction(double c_vac, double c_solute, double temperature) {
  return pred::RateCorrector(c_vac, c_solute).GetTimeCorrectionFactor(temperature);
}
int ref_tt_interpolate(const char *file, int64_t n, const double *time, double *temperature) {
  try {
    const pred::TimeTemperatureInterpolator tt{std::string(file)};
    for (int64_t k = 0; k < n; ++k) temperature[k] = tt.GetTemperature(time[k]);
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}

// ---------------------------------------------------------------- drivers
// mc::KineticMcFirstOmp (mc/src/KineticMcFirstOmp.cpp, mc/src/KineticMcAbstract.cpp:140-188) with the
// RNG reseeded and Dump() silenced. Runs maximum_steps+1 iterations exactly like Simulate().
// All trace pointers may be NULL. Returns wall seconds spent inside Simulate(), or <0 on error.
double ref_kmc_first_omp(void *config_h, const char *json, const int *codes, int ncodes, const char *tt_file,
                         int rate_corrector, double temperature, uint64_t maximum_steps, uint64_t seed, int threads,
                         const char *workdir, int64_t trace_cap, double *u1, double *u2, int64_t *from, int64_t *to,
                         int64_t *slot, double *dt, double *time, double *energy, double *Ea, double *dE,
                         double *temp_trace, double *total_rate, uint8_t *final_occ, double *summary4) {
  try {
    ScopedChdir cd(workdir);
    ScopedQuietCout quiet;
    if (threads > 0) omp_set_num_threads(threads);
    TracedKmcFirstOmp kmc(*static_cast<cfg::Config *>(config_h), 1ULL << 62, 1ULL << 62, maximum_steps, 0, 0, 0.0, 0.0,
                          temperature, element_set_from_codes(codes, ncodes), json, tt_file ? tt_file : "",
                          rate_corrector != 0, false, false);
    kmc.Reseed(seed);
    KmcTrace tr;
    tr.cap = trace_cap; tr.u1 = u1; tr.u2 = u2; tr.from = from; tr.to = to; tr.slot = slot; tr.dt = dt; tr.time = time;
    tr.energy = energy; tr.Ea = Ea; tr.dE = dE; tr.temperature = temp_trace; tr.total_rate = total_rate;
    kmc.SetTrace(trace_cap > 0 ? &tr : nullptr);
    const double t0 = now_s();
    kmc.Simulate();
    const double t1 = now_s();
    copy_occupancy(kmc.config(), final_occ);
    if (summary4) {
      summary4[0] = kmc.time(); summary4[1] = kmc.energy(); summary4[2] = kmc.absolute_energy();
      summary4[3] = static_cast<double>(kmc.steps());
    }
    return t1 - t0;
  } catch (const std::exception &e) { fail(e); return -1.0; }
}

// mc::KineticMcChainOmpi (mc/src/KineticMcChainOmpi.cpp:56-152, mc/src/KineticMcAbstract.cpp:140-188,260-263): the
// second-order ("chain") KMC, 12 ranks run as 12 threads.  Trace semantics as ref_kmc_first_omp; u1 is unused (0),
// u2 is the selecting uniform of rank 0, Ea/dE are those of the chosen k->i event, total_rate is total_rate_k_.
double ref_kmc_chain_ompi(void *config_h, const char *json, const int *codes, int ncodes, const char *tt_file,
                          int rate_corrector, double temperature, uint64_t maximum_steps, uint64_t seed,
                          const char *workdir, int64_t trace_cap, double *u1, double *u2, int64_t *from, int64_t *to,
                          int64_t *slot, double *dt, double *time, double *energy, double *Ea, double *dE,
                          double *temp_trace, double *total_rate, uint8_t *final_occ, double *summary4) {
  try {
    ScopedChdir cd(workdir);
    ScopedQuietCout quiet;
    constexpr int kRanks = 12;
    lmc_shim_mpi::World world(kRanks);
    const auto element_set = element_set_from_codes(codes, ncodes);
    const cfg::Config &start = *static_cast<cfg::Config *>(config_h);
    std::vector<std::string> errors(kRanks);
    double seconds = 0.0;
    KmcTrace tr;
    tr.cap = trace_cap; tr.u1 = u1; tr.u2 = u2; tr.from = from; tr.to = to; tr.slot = slot; tr.dt = dt; tr.time = time;
    tr.energy = energy; tr.Ea = Ea; tr.dE = dE; tr.temperature = temp_trace; tr.total_rate = total_rate;
    auto body = [&](int rank) {
      lmc_shim_mpi::attach(&world, rank);
      omp_set_num_threads(1);
      try {
        TracedKmcChainOmpi kmc(start, 1ULL << 62, 1ULL << 62, maximum_steps, 0, 0, 0.0, 0.0, temperature, element_set, json,
                               tt_file ? tt_file : "", rate_corrector != 0, false, false);
        kmc.Reseed(seed + static_cast<uint64_t>(rank));      // only rank 0's stream decides
        if (rank == 0 && trace_cap > 0) kmc.SetTrace(&tr);
        world.barrier();
        const double t0 = now_s();
        kmc.Simulate();
        world.barrier();
        if (rank == 0) {
          seconds = now_s() - t0;
          copy_occupancy(kmc.config(), final_occ);
          if (summary4) {
            summary4[0] = kmc.time(); summary4[1] = kmc.energy(); summary4[2] = kmc.absolute_energy();
            summary4[3] = static_cast<double>(kmc.steps());
          }
        }
      } catch (const std::exception &e) {
        errors[static_cast<size_t>(rank)] = e.what();
        std::cerr << "ref_kmc_chain_ompi rank " << rank << ": " << e.what() << std::endl;
        std::abort();                                         // the other ranks would wait forever in a collective
      }
      lmc_shim_mpi::detach();
    };
    std::vector<std::thread> ranks;
    for (int r = 0; r < kRanks; ++r) ranks.emplace_back(body, r);
    for (auto &t : ranks) t.join();
    return seconds;
  } catch (const std::exception &e) { fail(e); return -1.0; }
}

// mc::CanonicalMcSerial (mc/src/CanonicalMcSerial.cpp:40-51). Trace: pair, dE, energy before the trial.
double ref_cmc_serial(void *config_h, const char *json, const int *codes, int ncodes, double temperature,
                      uint64_t maximum_steps, uint64_t seed, const char *workdir, int64_t trace_cap, int64_t *a,
                      int64_t *b, double *dE, double *energy_before, double *u, uint8_t *final_occ,
                      double *final_energy) {
  try {
    ScopedChdir cd(workdir);
    ScopedQuietCout quiet;
    omp_set_num_threads(1);
    TracedCmcSerial cmc(*static_cast<cfg::Config *>(config_h), 1ULL << 62, 1ULL << 62, maximum_steps, 0, 0, 0.0,
                        temperature, element_set_from_codes(codes, ncodes), json);
    cmc.Reseed(seed);
    SwapTrace tr;
    tr.cap = trace_cap; tr.a = a; tr.b = b; tr.temperature = dE; tr.energy_before = energy_before; tr.u = u;
    cmc.SetTrace(trace_cap > 0 ? &tr : nullptr);
    const double t0 = now_s();
    if (trace_cap > 0) cmc.SimulateTraced(); else cmc.Simulate();
    const double t1 = now_s();
    copy_occupancy(cmc.config(), final_occ);
    if (final_energy) *final_energy = cmc.energy();
    return t1 - t0;
  } catch (const std::exception &e) { fail(e); return -1.0; }
}

// mc::CanonicalMcOmp (mc/src/CanonicalMcOmp.cpp:40-92): timing only (batch = OMP thread count).
double ref_cmc_omp(void *config_h, const char *json, const int *codes, int ncodes, double temperature,
                   uint64_t maximum_steps, uint64_t seed, int threads, const char *workdir, uint8_t *final_occ,
                   double *final_energy, uint64_t *steps_done) {
  try {
    ScopedChdir cd(workdir);
    ScopedQuietCout quiet;
    if (threads > 0) omp_set_num_threads(threads);
    QuietCmcOmp cmc(*static_cast<cfg::Config *>(config_h), 1ULL << 62, 1ULL << 62, maximum_steps, 0, 0, 0.0, temperature,
                    element_set_from_codes(codes, ncodes), json);
    cmc.Reseed(seed);
    const double t0 = now_s();
    cmc.Simulate();
    const double t1 = now_s();
    copy_occupancy(cmc.config(), final_occ);
    if (final_energy) *final_energy = cmc.energy();
    if (steps_done) *steps_done = cmc.steps();
    return t1 - t0;
  } catch (const std::exception &e) { fail(e); return -1.0; }
}

// mc::CanonicalMcOmp with a trace (pair, dE, energy before, the uniform SelectEvent would draw, batch number per event).
double ref_cmc_omp_traced(void *config_h, const char *json, const int *codes, int ncodes, double temperature,
                          uint64_t maximum_steps, uint64_t seed, int threads, const char *workdir, int64_t trace_cap, int64_t *a,
                          int64_t *b, double *dE, double *energy_before, double *u, int64_t *batch_of, uint8_t *final_occ,
                          double *final_energy, uint64_t *steps_done) {
  try {
    ScopedChdir cd(workdir);
    ScopedQuietCout quiet;
    if (threads > 0) omp_set_num_threads(threads);
    TracedCmcOmp cmc(*static_cast<cfg::Config *>(config_h), 1ULL << 62, 1ULL << 62, maximum_steps, 0, 0, 0.0, temperature,
                     element_set_from_codes(codes, ncodes), json);
    cmc.Reseed(seed);
    SwapTrace tr;
    tr.cap = trace_cap; tr.a = a; tr.b = b; tr.temperature = dE; tr.energy_before = energy_before; tr.u = u;
    cmc.SetTrace(&tr, batch_of);
    const double t0 = now_s();
    cmc.SimulateTraced();
    const double t1 = now_s();
    copy_occupancy(cmc.config(), final_occ);
    if (final_energy) *final_energy = cmc.energy();
    if (steps_done) *steps_done = cmc.steps();
    return t1 - t0;
  } catch (const std::exception &e) { fail(e); return -1.0; }
}

// mc::SimulatedAnnealing (mc/src/SimulatedAnnealing.cpp). The reference generates its own supercell with a
// clock-seeded generator; after construction the occupancy is overwritten with `occ` (GenerateFCC lattice-id
// order, same element counts) so that runs are reproducible. Energies in the trace are therefore relative to
// the constructor's initial value `energy0` (returned), which only shifts the log, not the decisions.
double ref_sa(int factor, int solvent_code, const int *solute_codes, const int64_t *solute_counts, int nsolute,
              const uint8_t *occ, const char *json, double initial_temperature, uint64_t maximum_steps, uint64_t seed,
              const char *workdir, int64_t trace_cap, int64_t *a, int64_t *b, double *energy_before,
              double *temperature_before, double *u, uint8_t *final_occ, double *energy0, double *final_energy,
              double *final_temperature) {
  try {
    ScopedChdir cd(workdir);
    ScopedQuietCout quiet;
    omp_set_num_threads(1);
    std::map<Element, size_t> counts;
    for (int s = 0; s < nsolute; ++s) counts[element_from_code(solute_codes[s])] = static_cast<size_t>(solute_counts[s]);
    const size_t f = static_cast<size_t>(factor);
    TracedSa sa({f, f, f}, element_from_code(solvent_code), counts, 1ULL << 62, 1ULL << 62, maximum_steps,
                initial_temperature, json);
    if (energy0) *energy0 = sa.energy();
    if (occ) {
      auto &c = sa.mutable_config();
      for (size_t l = 0; l < c.GetNumAtoms(); ++l) c.SetAtomElementTypeAtLattice(l, element_from_code(occ[l]));
    }
    sa.Reseed(seed);
    SwapTrace tr;
    tr.cap = trace_cap; tr.a = a; tr.b = b; tr.energy_before = energy_before; tr.temperature = temperature_before;
    tr.u = u;
    sa.SetTrace(trace_cap > 0 ? &tr : nullptr);
    const double t0 = now_s();
    sa.Simulate();
    const double t1 = now_s();
    copy_occupancy(sa.config(), final_occ);
    if (final_energy) *final_energy = sa.energy();
    if (final_temperature) *final_temperature = sa.temperature();
    return t1 - t0;
  } catch (const std::exception &e) { fail(e); return -1.0; }
}

// The drivers as shipped, logs and dumps included: after the call `workdir` holds kmc_log.txt / cmc_log.txt, 0.cfg.gz,
// end.cfg.gz ... exactly as `lmc.exe -p` would have written them (gzip is a pass-through in the shim, so they are text).
double ref_kmc_first_omp_with_logs(void *config_h, const char *json, const int *codes, int ncodes, const char *tt_file,
                                   int rate_corrector, double temperature, uint64_t log_dump_steps, uint64_t config_dump_steps,
                                   uint64_t maximum_steps, uint64_t seed, const char *workdir) {
  try {
    ScopedChdir cd(workdir);
    ScopedQuietCout quiet;
    omp_set_num_threads(1);
    SeededKmcFirstOmp kmc(*static_cast<cfg::Config *>(config_h), log_dump_steps, config_dump_steps, maximum_steps, 0, 0, 0.0, 0.0,
                          temperature, element_set_from_codes(codes, ncodes), json, tt_file ? tt_file : "", rate_corrector != 0,
                          false, false);
    kmc.Reseed(seed);
    const double t0 = now_s();
    kmc.Simulate();
    return now_s() - t0;
  } catch (const std::exception &e) { fail(e); return -1.0; }
}
// mc::KineticMcChainOmpi as shipped (rank 0 writes kmc_log.txt and the cfg dumps), 12 thread-ranks, seeded generators
double ref_kmc_chain_ompi_with_logs(void *config_h, const char *json, const int *codes, int ncodes, const char *tt_file,
                                    int rate_corrector, double temperature, uint64_t log_dump_steps, uint64_t config_dump_steps,
                                    uint64_t maximum_steps, uint64_t seed, int solute_disp, const char *workdir) {
  try {
    ScopedChdir cd(workdir);
    ScopedQuietCout quiet;
    constexpr int kRanks = 12;
    lmc_shim_mpi::World world(kRanks);
    const auto element_set = element_set_from_codes(codes, ncodes);
    const cfg::Config &start = *static_cast<cfg::Config *>(config_h);
    double seconds = 0.0;
    auto body = [&](int rank) {
      lmc_shim_mpi::attach(&world, rank);
      omp_set_num_threads(1);
      try {
        SeededKmcChainOmpi kmc(start, log_dump_steps, config_dump_steps, maximum_steps, 0, 0, 0.0, 0.0, temperature, element_set, json,
                               tt_file ? tt_file : "", rate_corrector != 0, false, solute_disp != 0);
        kmc.Reseed(seed + static_cast<uint64_t>(rank));
        world.barrier();
        const double t0 = now_s();
        kmc.Simulate();
        world.barrier();
        if (rank == 0) seconds = now_s() - t0;
      } catch (const std::exception &e) {
        std::cerr << "ref_kmc_chain_ompi_with_logs rank " << rank << ": " << e.what() << std::endl;
        std::abort();
      }
      lmc_shim_mpi::detach();
    };
    std::vector<std::thread> ranks;
    for (int r = 0; r < kRanks; ++r) ranks.emplace_back(body, r);
    for (auto &t : ranks) t.join();
    return seconds;
  } catch (const std::exception &e) { fail(e); return -1.0; }
}
double ref_cmc_serial_with_logs(void *config_h, const char *json, const int *codes, int ncodes, double temperature,
                                uint64_t log_dump_steps, uint64_t config_dump_steps, uint64_t maximum_steps,
                                uint64_t thermodynamic_averaging_steps, uint64_t seed, const char *workdir) {
  try {
    ScopedChdir cd(workdir);
    ScopedQuietCout quiet;
    omp_set_num_threads(1);
    SeededCmcSerial cmc(*static_cast<cfg::Config *>(config_h), log_dump_steps, config_dump_steps, maximum_steps,
                        thermodynamic_averaging_steps, 0, 0.0, temperature, element_set_from_codes(codes, ncodes), json);
    cmc.Reseed(seed);
    const double t0 = now_s();
    cmc.Simulate();
    return now_s() - t0;
  } catch (const std::exception &e) { fail(e); return -1.0; }
}

// libstdc++ draw helpers so tests can pre-generate exactly the stream the reference consumes
void ref_rng_uniform_real(uint64_t seed, int64_t n, double *out) {
  std::mt19937_64 g(seed);
  std::uniform_real_distribution<double> d(0.0, 1.0);
  for (int64_t k = 0; k < n; ++k) out[k] = d(g);
}

}  // extern "C"

// ---------------------------------------------------------------- VacancyMigrationPredictorE0[Lru], EnergyChangePredictorPair / Site
// (the alternative predictors of SURVEY 2b / 8(f)4; no live driver instantiates them)
namespace {
struct E0Probe : pred::VacancyMigrationPredictorE0 {   // opens the protected GetE0
  using pred::VacancyMigrationPredictorE0::VacancyMigrationPredictorE0;
  using pred::VacancyMigrationPredictorE0::GetE0;
};
}  // namespace
extern "C" {
void *ref_e0_create(const char *json, void *config_h, const int *codes, int n, int64_t lru_size) {
  try {
    ScopedQuietCout quiet;
    const auto &c = *static_cast<cfg::Config *>(config_h);
    if (lru_size > 0)
      return static_cast<pred::VacancyMigrationPredictorE0 *>(
          new pred::VacancyMigrationPredictorE0Lru(json, c, element_set_from_codes(codes, n), static_cast<size_t>(lru_size)));
    return static_cast<pred::VacancyMigrationPredictorE0 *>(new E0Probe(json, c, element_set_from_codes(codes, n)));
  } catch (const std::exception &e) { fail(e); return nullptr; }
}
void ref_e0_free(void *h) { delete static_cast<pred::VacancyMigrationPredictorE0 *>(h); }
// e0 (optional) needs the non-LRU predictor
int ref_e0_eval(void *h, void *config_h, int64_t n, const int64_t *i, const int64_t *j, double *Ea, double *dE, double *e0) {
  try {
    const auto *p = static_cast<pred::VacancyMigrationPredictorE0 *>(h);
    const auto *probe = dynamic_cast<const E0Probe *>(p);
    if (e0 && !probe) throw std::runtime_error("e0 output needs the non-LRU predictor");
    const auto &c = *static_cast<cfg::Config *>(config_h);
    for (int64_t k = 0; k < n; ++k) {
      const std::pair<size_t, size_t> pr{static_cast<size_t>(i[k]), static_cast<size_t>(j[k])};
      const auto r = p->GetBarrierAndDiffFromLatticeIdPair(c, pr);
      Ea[k] = r.first;
      dE[k] = r.second;
      if (e0) e0[k] = probe->GetE0(c, pr);
    }
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}
void *ref_pair_create(const char *json, void *config_h, const int *codes, int n) {
  try {
    return new pred::EnergyChangePredictorPair(json, *static_cast<cfg::Config *>(config_h), element_set_from_codes(codes, n));
  } catch (const std::exception &e) { fail(e); return nullptr; }
}
void ref_pair_free(void *h) { delete static_cast<pred::EnergyChangePredictorPair *>(h); }
int ref_pair_de(void *h, void *config_h, int64_t n, const int64_t *a, const int64_t *b, double *out) {
  try {
    const auto *p = static_cast<pred::EnergyChangePredictorPair *>(h);
    const auto &c = *static_cast<cfg::Config *>(config_h);
    for (int64_t k = 0; k < n; ++k) out[k] = p->GetDeFromLatticeIdPair(c, {static_cast<size_t>(a[k]), static_cast<size_t>(b[k])});
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}
void *ref_site_create(const char *json, void *config_h, const int *codes, int n) {
  try {
    return new pred::EnergyChangePredictorSite(json, *static_cast<cfg::Config *>(config_h), element_set_from_codes(codes, n));
  } catch (const std::exception &e) { fail(e); return nullptr; }
}
void ref_site_free(void *h) { delete static_cast<pred::EnergyChangePredictorSite *>(h); }
int ref_site_de(void *h, void *config_h, int64_t n, const int64_t *site, const uint8_t *new_code, double *out) {
  try {
    const auto *p = static_cast<pred::EnergyChangePredictorSite *>(h);
    const auto &c = *static_cast<cfg::Config *>(config_h);
    for (int64_t k = 0; k < n; ++k) out[k] = p->GetDeFromLatticeIdSite(c, static_cast<size_t>(site[k]), element_from_code(new_code[k]));
    return 0;
  } catch (const std::exception &e) { return fail(e); }
}
}  // extern "C"

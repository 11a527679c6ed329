// lmc_b200_adapters.hpp -- C++ host-side mirror of the reference's predictor / driver interfaces over the C ABI.
//
// The reference (zhucongx/LatticeMonteCarlo) reaches its hot path through three predictor classes that are
// const members of the mc:: drivers.  These adapters keep the same class names, constructor
// arguments, method names, argument meaning and exception types, but forward to liblmc_b200.so:
//
//   pred::VacancyMigrationPredictorQuartic[Lru]::GetBarrierAndDiffFromLatticeIdPair   pred/include/VacancyMigrationPredictorQuartic.h:25-27
//   pred::EnergyChangePredictorPairSite::GetDeFromLatticeIdPair / ...Site             pred/include/EnergyChangePredictorPairSite.h:20-25
//   pred::EnergyPredictor::GetEnergy                                                   pred/include/EnergyPredictor.h:19-25
//   pred::VacancyMigrationPredictorE0[Lru]::GetBarrierAndDiffFromLatticeIdPair        pred/include/VacancyMigrationPredictorE0.h:26-31
//   pred::EnergyChangePredictorPair::GetDeFromLatticeIdPair                            pred/include/EnergyChangePredictorPair.h:22-25
//   pred::EnergyChangePredictorSite::GetDeFromLatticeIdSite                            pred/include/EnergyChangePredictorSite.h:22-25
//   mc::KineticMcFirstOmp / CanonicalMcOmp / SimulatedAnnealing ::Simulate             mc/include/*.h
//
// What differs, by design: the configuration lives on the device (lmc_b200::Config owns an lmc_engine and the
// packed uint8 occupancy), batch entry points are added next to the scalar ones, and the LRU cache of
// VacancyMigrationPredictorQuarticLru is replaced by batch recomputation (the class is kept as an alias).
// Header-only; link with -llmc_b200.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "lmc_b200.h"

namespace lmc_b200 {

// ElementName values of the reference (lmc/cfg/include/Element.hpp:7)
enum class ElementName : int32_t { X = 0, Al, Mg, Zn, Cu, Sn, pAl, pMg, pZn, pCu, pSn };

inline void check(int rc) {
  if (rc >= 0) return;
  const std::string msg = lmc_last_error();
  switch (rc) {
    case LMC_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
    case LMC_ERR_OUT_OF_RANGE: throw std::out_of_range(msg);      // same type the reference throws (EnergyUtility.cpp:815-817)
    default: throw std::runtime_error(msg);                        // "Cannot open <file>", CUDA errors, no device
  }
}

namespace cfg {
// Device-resident counterpart of cfg::Config (lmc/cfg/include/Config.h:14-112) for FCC supercells.
//
// COPY SEMANTICS differ from the reference: copies of a Config share ONE engine (device occupancy) and one atom <-> lattice
// map, so a driver constructed from a Config (mc::KineticMcFirstOmp takes it by value, like the reference) advances the
// caller's object too.  Clone() makes the deep copy the reference's copy constructor would.
class Config {
 public:
  Config(const std::array<size_t, 3> &factors, lmc_id_order order, const std::set<ElementName> &element_set, ElementName solvent,
         int n_walkers = 1, int device = 0)
      : state_(std::make_shared<State>()) {
    const int32_t f[3] = {static_cast<int32_t>(factors[0]), static_cast<int32_t>(factors[1]), static_cast<int32_t>(factors[2])};
    for (auto e : element_set) state_->elements.push_back(static_cast<int32_t>(e));
    state_->factors = {f[0], f[1], f[2]};
    state_->order = order; state_->solvent = static_cast<int32_t>(solvent); state_->n_walkers = n_walkers; state_->device = device;
    lmc_engine *raw = nullptr;
    check(lmc_engine_create(&raw, f, order, state_->elements.data(), static_cast<int32_t>(state_->elements.size()), static_cast<int32_t>(solvent),
                            n_walkers, device));
    state_->engine.reset(raw, lmc_engine_destroy);
  }
  // deep copy: own engine with the same occupancy, coefficient file and atom <-> lattice maps
  [[nodiscard]] Config Clone() const {
    std::set<ElementName> es;
    for (auto e : state_->elements) es.insert(static_cast<ElementName>(e));
    Config out({static_cast<size_t>(state_->factors[0]), static_cast<size_t>(state_->factors[1]), static_cast<size_t>(state_->factors[2])},
               state_->order, es, static_cast<ElementName>(state_->solvent), state_->n_walkers, state_->device);
    for (int w = 0; w < state_->n_walkers; ++w) out.SetOccupancy(GetOccupancy(w), w);
    const std::string path = lmc_engine_coefficients_path(engine());
    if (!path.empty()) {
      const int model = lmc_engine_barrier_model(engine());
      check(lmc_engine_load_coefficients_model(out.engine(), path.c_str(), model == LMC_BARRIER_E0 ? LMC_BARRIER_E0 : LMC_BARRIER_QUARTIC));
    }
    out.state_->maps = state_->maps;
    return out;
  }
  [[nodiscard]] size_t GetNumAtoms() const { return static_cast<size_t>(lmc_engine_num_sites(engine())); }
  void SetOccupancy(const std::vector<uint8_t> &element_by_lattice_id, int walker = 0) {
    check(lmc_engine_set_occupancy(engine(), walker, element_by_lattice_id.data(), static_cast<int64_t>(element_by_lattice_id.size())));
  }
  [[nodiscard]] std::vector<uint8_t> GetOccupancy(int walker = 0) const {
    std::vector<uint8_t> out(GetNumAtoms());
    check(lmc_engine_get_occupancy(engine(), walker, out.data(), static_cast<int64_t>(out.size())));
    return out;
  }
  // Config::GetFirst/Second/ThirdNeighborsAdjacencyList()[lattice_id] (ascending ids)
  [[nodiscard]] std::vector<size_t> GetNeighbors(int shell, size_t lattice_id) const {
    int64_t buf[24];
    check(lmc_engine_neighbors(engine(), shell, static_cast<int64_t>(lattice_id), buf));
    const int n = shell == 1 ? 12 : (shell == 2 ? 6 : 24);
    return std::vector<size_t>(buf, buf + n);
  }
  // Config::GetElementAtLatticeId / GetElementAtAtomId (cfg/src/Config.cpp:128-135)
  [[nodiscard]] ElementName GetElementAtLatticeId(size_t lattice_id, int walker = 0) const {
    const int64_t id = static_cast<int64_t>(lattice_id);
    uint8_t e = 0;
    check(lmc_engine_get_elements(engine(), walker, 1, &id, &e));
    return static_cast<ElementName>(e);
  }
  [[nodiscard]] ElementName GetElementAtAtomId(size_t atom_id, int walker = 0) const { return GetElementAtLatticeId(GetLatticeIdFromAtomId(atom_id, walker), walker); }
  // Config::GetVacancyLatticeId / GetVacancyAtomId (cfg/src/Config.cpp:296-310): the first vacancy by lattice id
  [[nodiscard]] size_t GetVacancyLatticeId(int walker = 0) const {
    const int64_t id = lmc_engine_find_element(engine(), walker, 0, nullptr);
    if (id < -1) check(static_cast<int>(id + 1000));
    if (id < 0) throw std::runtime_error("vacancy not found");
    return static_cast<size_t>(id);
  }
  [[nodiscard]] size_t GetVacancyAtomId(int walker = 0) const { return GetAtomIdFromLatticeId(GetVacancyLatticeId(walker), walker); }
  // atom <-> lattice maps (cfg/src/Config.cpp:115-127): the identity until LatticeJump moves atoms
  [[nodiscard]] size_t GetLatticeIdFromAtomId(size_t atom_id, int walker = 0) const {
    const auto it = state_->maps.find(walker);
    return it == state_->maps.end() ? atom_id : it->second.atom_to_lattice.at(atom_id);
  }
  [[nodiscard]] size_t GetAtomIdFromLatticeId(size_t lattice_id, int walker = 0) const {
    const auto it = state_->maps.find(walker);
    return it == state_->maps.end() ? lattice_id : it->second.lattice_to_atom.at(lattice_id);
  }
  void LatticeJump(const std::pair<size_t, size_t> &lattice_id_jump_pair, int walker = 0) {   // Config.cpp:431-456
    check(lmc_engine_lattice_jump(engine(), walker, static_cast<int64_t>(lattice_id_jump_pair.first),
                                  static_cast<int64_t>(lattice_id_jump_pair.second)));
    auto &m = state_->maps[walker];
    if (m.atom_to_lattice.empty()) {
      m.atom_to_lattice.resize(GetNumAtoms());
      m.lattice_to_atom.resize(GetNumAtoms());
      for (size_t q = 0; q < m.atom_to_lattice.size(); ++q) m.atom_to_lattice[q] = m.lattice_to_atom[q] = q;
    }
    const size_t a = m.lattice_to_atom[lattice_id_jump_pair.first], b = m.lattice_to_atom[lattice_id_jump_pair.second];
    m.atom_to_lattice[a] = lattice_id_jump_pair.second; m.atom_to_lattice[b] = lattice_id_jump_pair.first;
    m.lattice_to_atom[lattice_id_jump_pair.first] = b; m.lattice_to_atom[lattice_id_jump_pair.second] = a;
  }
  void AtomJump(const std::pair<size_t, size_t> &atom_id_jump_pair, int walker = 0) {          // Config.cpp:458-462
    LatticeJump({GetLatticeIdFromAtomId(atom_id_jump_pair.first, walker), GetLatticeIdFromAtomId(atom_id_jump_pair.second, walker)}, walker);
  }
  [[nodiscard]] lmc_engine *engine() const { return state_->engine.get(); }

 private:
  struct Maps { std::vector<size_t> atom_to_lattice, lattice_to_atom; };
  struct State {
    std::shared_ptr<lmc_engine> engine;
    std::map<int, Maps> maps;                       // per walker, allocated by its first jump
    std::vector<int32_t> elements;
    std::array<int32_t, 3> factors{};
    lmc_id_order order{};
    int32_t solvent{0};
    int n_walkers{1}, device{0};
  };
  std::shared_ptr<State> state_;
};
}  // namespace cfg

namespace pred {
// One JSON file feeds all three predictors, exactly as in the reference (each of its constructors re-reads the file).
inline void LoadCoefficients(const cfg::Config &reference_config, const std::string &predictor_filename,
                             lmc_barrier_model model = LMC_BARRIER_QUARTIC) {
  check(lmc_engine_load_coefficients_model(reference_config.engine(), predictor_filename.c_str(), model));
}
// An engine holds the barrier tables of ONE model at a time.  A barrier predictor re-loads its file if another predictor
// object has put a different model (or none) on the shared engine since, so two predictor objects on one Config stay correct.
inline void EnsureBarrierModel(const cfg::Config &config, const std::string &predictor_filename, lmc_barrier_model model) {
  if (lmc_engine_barrier_model(config.engine()) != static_cast<int>(model) || predictor_filename != lmc_engine_coefficients_path(config.engine()))
    LoadCoefficients(config, predictor_filename, model);
}
// The dE / energy predictors read only "Base".theta: they re-load when another FILE has been put on the shared engine since
// (keeping whatever barrier model is there if the file carries its blocks).
inline void EnsureCoefficientFile(const cfg::Config &config, const std::string &predictor_filename) {
  if (predictor_filename == lmc_engine_coefficients_path(config.engine())) return;
  const int model = lmc_engine_barrier_model(config.engine());
  LoadCoefficients(config, predictor_filename, model == LMC_BARRIER_E0 ? LMC_BARRIER_E0 : LMC_BARRIER_QUARTIC);
}

class VacancyMigrationPredictorQuartic {
 public:
  VacancyMigrationPredictorQuartic(const std::string &predictor_filename, const cfg::Config &reference_config,
                                   const std::set<ElementName> & /*element_set: fixed by the Config*/)
      : VacancyMigrationPredictorQuartic(predictor_filename, reference_config, LMC_BARRIER_QUARTIC) {}
  virtual ~VacancyMigrationPredictorQuartic() = default;
  // {Ea, dE} of the vacancy at .first exchanging with the atom at .second
  [[nodiscard]] virtual std::pair<double, double> GetBarrierAndDiffFromLatticeIdPair(
      const cfg::Config &config, const std::pair<size_t, size_t> &lattice_id_jump_pair, int walker = 0) const {
    EnsureBarrierModel(config, filename_, model_);
    const int64_t i = static_cast<int64_t>(lattice_id_jump_pair.first), j = static_cast<int64_t>(lattice_id_jump_pair.second);
    const int32_t w = walker;
    double ea = 0, de = 0;
    check(lmc_eval_barriers(config.engine(), 1, &w, &i, &j, &ea, &de, nullptr, nullptr));
    return {ea, de};
  }
  // pred/include/VacancyMigrationPredictorQuartic.h:22-24
  [[nodiscard]] virtual std::pair<double, double> GetBarrierAndDiffFromAtomIdPair(const cfg::Config &config,
                                                                                   const std::pair<size_t, size_t> &atom_id_jump_pair,
                                                                                   int walker = 0) const {
    return GetBarrierAndDiffFromLatticeIdPair(config, {config.GetLatticeIdFromAtomId(atom_id_jump_pair.first, walker),
                                                       config.GetLatticeIdFromAtomId(atom_id_jump_pair.second, walker)}, walker);
  }
  // batch form: all candidate events of a step (or of many walkers) in one launch
  void GetBarrierAndDiffFromLatticeIdPairs(const cfg::Config &config, const std::vector<int32_t> &walker, const std::vector<int64_t> &first,
                                           const std::vector<int64_t> &second, std::vector<double> &Ea, std::vector<double> &dE) const {
    EnsureBarrierModel(config, filename_, model_);
    Ea.resize(first.size());
    dE.resize(first.size());
    check(lmc_eval_barriers(config.engine(), static_cast<int64_t>(first.size()), walker.empty() ? nullptr : walker.data(), first.data(),
                            second.data(), Ea.data(), dE.data(), nullptr, nullptr));
  }
  // the whole event list of a vacancy (KineticMcFirstOmp::BuildEventList, mc/src/KineticMcFirstOmp.cpp:52-68): neighbour ids
  // in adjacency order with their {Ea, dE}
  void GetEventListOfVacancy(const cfg::Config &config, size_t vacancy_lattice_id, std::array<size_t, 12> &neighbour_lattice_ids,
                             std::array<std::pair<double, double>, 12> &barrier_and_diff, int walker = 0) const {
    EnsureBarrierModel(config, filename_, model_);
    const int64_t v = static_cast<int64_t>(vacancy_lattice_id);
    const int32_t w = walker;
    int64_t nb[12];
    double ea[12], de[12];
    check(lmc_eval_vacancy_events(config.engine(), 1, &w, &v, nb, ea, de));
    for (int q = 0; q < 12; ++q) {
      neighbour_lattice_ids[static_cast<size_t>(q)] = static_cast<size_t>(nb[q]);
      barrier_and_diff[static_cast<size_t>(q)] = {ea[q], de[q]};
    }
  }

 protected:
  VacancyMigrationPredictorQuartic(const std::string &predictor_filename, const cfg::Config &reference_config, lmc_barrier_model model)
      : filename_(predictor_filename), model_(model) {
    LoadCoefficients(reference_config, predictor_filename, model);
    if (lmc_engine_barrier_model(reference_config.engine()) != static_cast<int>(model))   // the reference's json .at() would throw
      throw std::out_of_range("coefficient file " + predictor_filename + " has no element blocks for this barrier model");
  }

 private:
  std::string filename_;
  lmc_barrier_model model_;
};
// The LRU cache (pred/src/VacancyMigrationPredictorQuarticLru.cpp) is replaced by batch recomputation; cache_size is accepted and ignored.
class VacancyMigrationPredictorQuarticLru : public VacancyMigrationPredictorQuartic {
 public:
  VacancyMigrationPredictorQuarticLru(const std::string &predictor_filename, const cfg::Config &reference_config,
                                      const std::set<ElementName> &element_set, size_t /*cache_size*/)
      : VacancyMigrationPredictorQuartic(predictor_filename, reference_config, element_set) {}
};

// pred::VacancyMigrationPredictorE0 (pred/src/VacancyMigrationPredictorE0.cpp): same interface, coefficient keys
// mu_x_mmm / sigma_x_mmm / U_mmm / theta_e0 / mu_e0 / sigma_e0, Ea = max(0, e0 + dE / 2).  Batch and event-list forms are inherited.
class VacancyMigrationPredictorE0 : public VacancyMigrationPredictorQuartic {
 public:
  VacancyMigrationPredictorE0(const std::string &predictor_filename, const cfg::Config &reference_config, const std::set<ElementName> &)
      : VacancyMigrationPredictorQuartic(predictor_filename, reference_config, LMC_BARRIER_E0) {}
};
class VacancyMigrationPredictorE0Lru : public VacancyMigrationPredictorE0 {       // cache replaced by batch recomputation
 public:
  VacancyMigrationPredictorE0Lru(const std::string &predictor_filename, const cfg::Config &reference_config,
                                 const std::set<ElementName> &element_set, size_t /*cache_size*/)
      : VacancyMigrationPredictorE0(predictor_filename, reference_config, element_set) {}
};

class EnergyChangePredictorPairSite {
 public:
  EnergyChangePredictorPairSite(const std::string &predictor_filename, const cfg::Config &reference_config,
                                const std::set<ElementName> &)
      : filename_(predictor_filename) {
    EnsureCoefficientFile(reference_config, predictor_filename);
  }
  // pred/include/EnergyChangePredictorPairSite.h:20-25
  [[nodiscard]] double GetDeFromAtomIdPair(const cfg::Config &config, const std::pair<size_t, size_t> &atom_id_jump_pair, int walker = 0) const {
    return GetDeFromLatticeIdPair(config, {config.GetLatticeIdFromAtomId(atom_id_jump_pair.first, walker),
                                           config.GetLatticeIdFromAtomId(atom_id_jump_pair.second, walker)}, walker);
  }
  [[nodiscard]] double GetDeFromAtomIdSite(const cfg::Config &config, size_t atom_id, ElementName new_element, int walker = 0) const {
    return GetDeFromLatticeIdSite(config, config.GetLatticeIdFromAtomId(atom_id, walker), new_element, walker);
  }
  [[nodiscard]] double GetDeFromLatticeIdPair(const cfg::Config &config, const std::pair<size_t, size_t> &lattice_id_jump_pair,
                                              int walker = 0) const {
    EnsureCoefficientFile(config, filename_);
    const int64_t a = static_cast<int64_t>(lattice_id_jump_pair.first), b = static_cast<int64_t>(lattice_id_jump_pair.second);
    const int32_t w = walker;
    double de = 0;
    check(lmc_eval_swap_de(config.engine(), 1, &w, &a, &b, &de));
    return de;
  }
  [[nodiscard]] double GetDeFromLatticeIdSite(const cfg::Config &config, size_t lattice_id, ElementName new_element, int walker = 0) const {
    EnsureCoefficientFile(config, filename_);
    const int64_t s = static_cast<int64_t>(lattice_id);
    const uint8_t e = static_cast<uint8_t>(new_element);
    const int32_t w = walker;
    double de = 0;
    check(lmc_eval_site_de(config.engine(), 1, &w, &s, &e, &de));
    return de;
  }

 private:
  std::string filename_;
};

// pred::EnergyChangePredictorPair (pred/src/EnergyChangePredictorPair.cpp:69-122): exchange energy of a FIRST-NEIGHBOUR pair;
// equal species give 0, an unlike pair further apart throws std::out_of_range like the reference's unordered_map::at (:83).
// NB these dE predictors read only "Base".theta; they re-use whatever barrier model is on the engine.
class EnergyChangePredictorPair {
 public:
  EnergyChangePredictorPair(const std::string &predictor_filename, const cfg::Config &reference_config, const std::set<ElementName> &)
      : filename_(predictor_filename) {
    EnsureCoefficientFile(reference_config, predictor_filename);
  }
  [[nodiscard]] double GetDeFromAtomIdPair(const cfg::Config &config, const std::pair<size_t, size_t> &atom_id_jump_pair, int walker = 0) const {
    return GetDeFromLatticeIdPair(config, {config.GetLatticeIdFromAtomId(atom_id_jump_pair.first, walker),
                                           config.GetLatticeIdFromAtomId(atom_id_jump_pair.second, walker)}, walker);
  }
  [[nodiscard]] double GetDeFromLatticeIdPair(const cfg::Config &config, const std::pair<size_t, size_t> &lattice_id_jump_pair,
                                              int walker = 0) const {
    EnsureCoefficientFile(config, filename_);
    const int64_t a = static_cast<int64_t>(lattice_id_jump_pair.first), b = static_cast<int64_t>(lattice_id_jump_pair.second);
    const int32_t w = walker;
    double de = 0;
    check(lmc_eval_pair_de(config.engine(), 1, &w, &a, &b, &de));
    return de;
  }

 private:
  std::string filename_;
};
// pred::EnergyChangePredictorSite (pred/src/EnergyChangePredictorSite.cpp:56-98)
class EnergyChangePredictorSite {
 public:
  EnergyChangePredictorSite(const std::string &predictor_filename, const cfg::Config &reference_config, const std::set<ElementName> &)
      : filename_(predictor_filename) {
    EnsureCoefficientFile(reference_config, predictor_filename);
  }
  [[nodiscard]] double GetDeFromAtomIdSite(const cfg::Config &config, size_t atom_id, ElementName new_element, int walker = 0) const {
    return GetDeFromLatticeIdSite(config, config.GetLatticeIdFromAtomId(atom_id, walker), new_element, walker);
  }
  [[nodiscard]] double GetDeFromLatticeIdSite(const cfg::Config &config, size_t lattice_id, ElementName new_element, int walker = 0) const {
    EnsureCoefficientFile(config, filename_);
    const int64_t s = static_cast<int64_t>(lattice_id);
    const uint8_t e = static_cast<uint8_t>(new_element);
    const int32_t w = walker;
    double de = 0;
    check(lmc_eval_site_de(config.engine(), 1, &w, &s, &e, &de));
    return de;
  }

 private:
  std::string filename_;
};

class EnergyPredictor {      // pred/include/EnergyPredictor.h:14-30
 public:
  EnergyPredictor(const std::string &predictor_filename, const cfg::Config &reference_config) : filename_(predictor_filename) {
    EnsureCoefficientFile(reference_config, predictor_filename);
  }
  [[nodiscard]] double GetEnergy(const cfg::Config &config, int walker = 0) const {
    EnsureCoefficientFile(config, filename_);
    double e = 0;
    check(lmc_total_energy(config.engine(), walker, &e, nullptr, 0));
    return e;
  }
  // energy of the clusters inside { the listed atoms and their 1-3NN shells } (pred/src/EnergyPredictor.cpp:97-184)
  [[nodiscard]] double GetEnergyOfCluster(const cfg::Config &config, const std::vector<size_t> &atom_id_list, int walker = 0) const {
    EnsureCoefficientFile(config, filename_);
    const auto ids = LatticeIds(config, atom_id_list, walker);
    double e = 0;
    check(lmc_energy_of_cluster(config.engine(), walker, ids.data(), static_cast<int64_t>(ids.size()), &e, nullptr, 0));
    return e;
  }
  [[nodiscard]] std::vector<double> GetEncode(const cfg::Config &config, int walker = 0) const {
    EnsureCoefficientFile(config, filename_);
    std::vector<double> out(static_cast<size_t>(NumTypes(config)));
    check(lmc_energy_encode(config.engine(), walker, nullptr, -1, out.data(), static_cast<int32_t>(out.size())));
    return out;
  }
  [[nodiscard]] std::vector<double> GetEncodeOfCluster(const cfg::Config &config, const std::vector<size_t> &atom_id_list, int walker = 0) const {
    EnsureCoefficientFile(config, filename_);
    const auto ids = LatticeIds(config, atom_id_list, walker);
    std::vector<double> out(static_cast<size_t>(NumTypes(config)));
    check(lmc_energy_encode(config.engine(), walker, ids.data(), static_cast<int64_t>(ids.size()), out.data(), static_cast<int32_t>(out.size())));
    return out;
  }
  // pred/src/EnergyPredictor.cpp:196-214; needs a Config for the engine that holds the coefficient file
  [[nodiscard]] std::map<ElementName, double> GetChemicalPotential(const cfg::Config &config, ElementName solvent_element) const {
    EnsureCoefficientFile(config, filename_);
    int32_t el[16];
    double mu[16];
    const int32_t n = lmc_chemical_potential(config.engine(), static_cast<int32_t>(solvent_element), el, mu, 16);
    check(n);
    std::map<ElementName, double> out;
    for (int32_t q = 0; q < n; ++q) out[static_cast<ElementName>(el[q])] = mu[q];
    return out;
  }

 private:
  static std::vector<int64_t> LatticeIds(const cfg::Config &config, const std::vector<size_t> &atom_id_list, int walker) {
    std::vector<int64_t> ids;
    for (auto a : atom_id_list) ids.push_back(static_cast<int64_t>(config.GetLatticeIdFromAtomId(a, walker)));
    return ids;
  }
  static int32_t NumTypes(const cfg::Config &config) {     // Base.theta has one entry per cluster type of the element set (+ X)
    const int64_t n = lmc_engine_get_tables(config.engine(), 6, nullptr, 0);
    check(static_cast<int>(n < 0 ? n : 0));
    return static_cast<int32_t>(n);
  }
  std::string filename_;
};
}  // namespace pred

namespace mc {
// mc::KineticMcFirstOmp (mc/include/KineticMcFirstOmp.h) over all walkers of the Config.  Simulate() runs
// maximum_steps + 1 iterations like the reference's `while (steps_ <= maximum_steps_)` (KineticMcAbstract.cpp:184-188).
class KineticMcFirstOmp {
 public:
  KineticMcFirstOmp(cfg::Config config, unsigned long long maximum_steps, double temperature, const std::string &json_coefficients_filename,
                    const std::vector<std::pair<double, double>> &time_temperature = {}, bool is_rate_corrector = false,
                    uint64_t seed = 0)
      : config_(std::move(config)), maximum_steps_(maximum_steps), temperature_(temperature), tt_(time_temperature),
        rate_corrector_(is_rate_corrector), seed_(seed) {
    pred::LoadCoefficients(config_, json_coefficients_filename);
    check(lmc_kmc_reset(config_.engine()));
  }
  virtual ~KineticMcFirstOmp() = default;
  void Simulate() {
    std::vector<double> t, v;
    for (const auto &p : tt_) { t.push_back(p.first); v.push_back(p.second); }
    lmc_kmc_params prm{};
    prm.temperature = temperature_;
    prm.n_time_temperature = static_cast<int32_t>(t.size());
    prm.tt_time = t.data();
    prm.tt_temperature = v.data();
    prm.rate_corrector = rate_corrector_ ? 1 : 0;
    prm.seed = seed_;
    Run(prm, static_cast<int64_t>(maximum_steps_ + 1));
  }
  [[nodiscard]] const cfg::Config &GetConfig() const { return config_; }

 protected:
  virtual void Run(const lmc_kmc_params &prm, int64_t n_steps) {
    check(lmc_kmc_run(config_.engine(), &prm, n_steps, nullptr, nullptr, nullptr));
  }

 private:
  cfg::Config config_;
  unsigned long long maximum_steps_;
  double temperature_;
  std::vector<std::pair<double, double>> tt_;
  bool rate_corrector_;
  uint64_t seed_;
};

// mc::KineticMcChainOmpi (mc/include/KineticMcChainOmpi.h): second-order KMC; same constructor arguments, the 12 MPI ranks
// of the reference become the 12 half-warps of one thread block per walker.
class KineticMcChainOmpi : public KineticMcFirstOmp {
 public:
  using KineticMcFirstOmp::KineticMcFirstOmp;

 protected:
  void Run(const lmc_kmc_params &prm, int64_t n_steps) override {
    check(lmc_kmc_chain_run(GetConfig().engine(), &prm, n_steps, nullptr, nullptr));
  }
};

// mc::CanonicalMcOmp (mc/include/CanonicalMcOmp.h); with initial_temperature > 0 and sa_maximum_steps > 0 it is
// mc::SimulatedAnnealing's schedule (mc/include/SimulatedAnnealing.h:42-68).
class CanonicalMcOmp {
 public:
  CanonicalMcOmp(cfg::Config config, unsigned long long maximum_steps, double temperature, const std::string &json_coefficients_filename,
                 uint64_t seed = 0, bool simulated_annealing = false)
      : config_(std::move(config)), maximum_steps_(maximum_steps), temperature_(temperature), seed_(seed) {
    pred::LoadCoefficients(config_, json_coefficients_filename);
    check(lmc_cmc_reset(config_.engine(), simulated_annealing ? temperature : 0.0, simulated_annealing ? maximum_steps : 0));
  }
  // Extension (no reference counterpart): swap partners are drawn inside box domains of `domain_edge` half lattice constants
  // whose grid is shifted at random before every sweep of `rounds_per_sweep` trials per domain (0 = engine defaults: 6 / 216).
  // Same Metropolis rule and equilibrium distribution as the global-pair driver (tests/test_gpu_cmc_stat.py), 5-10x the
  // trial rate; trials advance in whole sweeps, so Simulate() may overshoot maximum_steps by less than one sweep.
  void SetDomainDecomposition(int domain_edge = 0, int rounds_per_sweep = 0) {
    domain_ = true;
    domain_params_ = lmc_cmc_domain_params{};
    domain_params_.domain_edge = domain_edge;
    domain_params_.rounds_per_sweep = rounds_per_sweep;
  }
  void Simulate() {
    lmc_cmc_params prm{};
    prm.temperature = temperature_;
    prm.seed = seed_;
    if (domain_) check(lmc_cmc_domain_run(config_.engine(), &prm, &domain_params_, static_cast<int64_t>(maximum_steps_ + 1)));
    else check(lmc_cmc_run(config_.engine(), &prm, static_cast<int64_t>(maximum_steps_ + 1)));
  }
  // energy relative to the start, trials done, trials accepted, current temperature (the SA schedule moves it)
  void GetState(double *energy, int64_t *steps, int64_t *accepted, double *temperature) const {
    check(lmc_cmc_get_state(config_.engine(), energy, steps, accepted, temperature));
  }
  [[nodiscard]] const cfg::Config &GetConfig() const { return config_; }

 private:
  cfg::Config config_;
  unsigned long long maximum_steps_;
  double temperature_;
  uint64_t seed_;
  bool domain_{false};
  lmc_cmc_domain_params domain_params_{};
};
}  // namespace mc

}  // namespace lmc_b200

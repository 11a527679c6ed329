// adapter_demo.cpp -- the reference's call pattern (KineticMcFirstOmp::BuildEventList, mc/src/KineticMcFirstOmp.cpp:52-68)
// written against the drop-in adapters: evaluate the 12 candidate jumps of the vacancy, then run the batched driver.
//   g++ -std=c++17 -Iinclude examples/adapter_demo.cpp -Llatticemontecarlo_b200 -llmc_b200 -Wl,-rpath,$PWD/latticemontecarlo_b200
//   ./a.out coefficients.json <factor> [e0_coefficients.json]
#include <cstdio>
#include <cstdlib>
#include <random>

#include "lmc_b200_adapters.hpp"

using namespace lmc_b200;

int main(int argc, char **argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s coefficients.json [factor]\n", argv[0]); return 2; }
  const size_t f = argc > 2 ? static_cast<size_t>(std::atoi(argv[2])) : 6;
  try {
    const std::set<ElementName> element_set{ElementName::Al, ElementName::Mg, ElementName::Zn};
    cfg::Config config({f, f, f}, LMC_ID_ORDER_REASSIGNED, element_set, ElementName::Al);
    std::vector<uint8_t> occ(config.GetNumAtoms(), static_cast<uint8_t>(ElementName::Al));
    std::mt19937_64 gen(42);
    std::uniform_real_distribution<double> u(0, 1);
    for (auto &e : occ) { const double r = u(gen); e = r < 0.02 ? 2 : (r < 0.04 ? 3 : 1); }
    const size_t vacancy = occ.size() / 2 + 3;
    occ[vacancy] = 0;
    config.SetOccupancy(occ);

    const pred::VacancyMigrationPredictorQuarticLru predictor(argv[1], config, element_set, 100000);
    double total_rate = 0.0;
    const double beta = 1.0 / 8.617333262145e-5 / 500.0;
    for (size_t j : config.GetNeighbors(1, vacancy)) {
      const auto [ea, de] = predictor.GetBarrierAndDiffFromLatticeIdPair(config, {vacancy, j});
      total_rate += std::exp(-ea * beta);
      std::printf("jump %zu -> %zu  Ea = %.12f eV  dE = %+.12f eV\n", vacancy, j, ea, de);
    }
    std::printf("total rate %.6e\n", total_rate);
    const pred::EnergyPredictor energy(argv[1], config);
    std::printf("E = %.9f eV\n", energy.GetEnergy(config));
    try {
      (void)predictor.GetBarrierAndDiffFromLatticeIdPair(config, {vacancy, vacancy + 7});   // not a first neighbour
    } catch (const std::out_of_range &e) {
      std::printf("std::out_of_range as in the reference: %s\n", e.what());
    }
    if (argc > 3) {
      // the alternative predictors (VacancyMigrationPredictorE0Lru, EnergyChangePredictorPair / Site) on the same Config;
      // the quartic predictor above stays valid: each barrier predictor puts its own model back on the engine when needed
      const pred::VacancyMigrationPredictorE0Lru e0_predictor(argv[3], config, element_set, 100000);
      const pred::EnergyChangePredictorPair pair_predictor(argv[3], config, element_set);
      const pred::EnergyChangePredictorSite site_predictor(argv[3], config, element_set);
      const auto neighbours = config.GetNeighbors(1, vacancy);
      for (size_t j : neighbours) {
        const auto [ea, de] = e0_predictor.GetBarrierAndDiffFromLatticeIdPair(config, {vacancy, j});
        std::printf("E0 model %zu -> %zu  barrier %.12f  change %+.12f  pair-predictor %+.12f\n", vacancy, j, ea, de,
                    pair_predictor.GetDeFromLatticeIdPair(config, {vacancy, j}));
      }
      std::printf("site %zu -> Mg: %+.12f\n", neighbours[0], site_predictor.GetDeFromLatticeIdSite(config, neighbours[0], ElementName::Mg));
      try {
        (void)pair_predictor.GetDeFromLatticeIdPair(config, {vacancy, config.GetNeighbors(2, vacancy)[0]});
      } catch (const std::out_of_range &) {
        std::printf("pair predictor: std::out_of_range for a second-neighbour pair\n");
      }
      const auto [ea_q, de_q] = predictor.GetBarrierAndDiffFromLatticeIdPair(config, {vacancy, neighbours[0]});   // quartic again
      std::printf("quartic after E0: %.12f %+.12f\n", ea_q, de_q);
    }
    mc::KineticMcFirstOmp kmc(config, 999, 500.0, argv[1]);
    kmc.Simulate();
    double t = 0, e = 0;
    int64_t steps = 0;
    check(lmc_kmc_get_state(kmc.GetConfig().engine(), &t, &e, &steps, nullptr, nullptr));
    std::printf("KMC: %lld steps, time %.6e s, energy %+.9f eV\n", static_cast<long long>(steps), t, e);
    mc::KineticMcChainOmpi chain(config, 499, 500.0, argv[1]);      // second-order KMC continues from the same state
    chain.Simulate();
    check(lmc_kmc_get_state(chain.GetConfig().engine(), &t, &e, &steps, nullptr, nullptr));
    std::printf("chain KMC: %lld steps, time %.6e s, energy %+.9f eV\n", static_cast<long long>(steps), t, e);
    {
      // canonical MC on a vacancy-free deep copy: global pairs (the reference's proposal), then the domain-decomposed driver
      cfg::Config cmc_config = config.Clone();
      occ[vacancy] = static_cast<uint8_t>(ElementName::Al);
      cmc_config.SetOccupancy(occ);
      const double e_start = energy.GetEnergy(cmc_config);
      mc::CanonicalMcOmp cmc(cmc_config, 1999, 800.0, argv[1], 7);
      cmc.Simulate();
      double de_sum = 0, temperature = 0;
      int64_t trials = 0, accepted = 0;
      cmc.GetState(&de_sum, &trials, &accepted, &temperature);
      std::printf("CMC: %lld trials, %lld accepted, energy drift %.3e eV\n", static_cast<long long>(trials), static_cast<long long>(accepted),
                  std::abs(energy.GetEnergy(cmc.GetConfig()) - (e_start + de_sum)));
      mc::CanonicalMcOmp domain_cmc(cmc_config, 1999, 800.0, argv[1], 7);     // continues on the same engine state
      domain_cmc.SetDomainDecomposition(6, 64);
      const double e_mid = energy.GetEnergy(cmc_config);
      domain_cmc.Simulate();
      domain_cmc.GetState(&de_sum, &trials, &accepted, &temperature);
      std::printf("domain CMC: %lld trials, %lld accepted, energy drift %.3e eV\n", static_cast<long long>(trials),
                  static_cast<long long>(accepted), std::abs(energy.GetEnergy(domain_cmc.GetConfig()) - (e_mid + de_sum)));
    }
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}

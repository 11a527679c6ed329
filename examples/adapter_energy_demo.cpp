// adapter_energy_demo.cpp -- the parts of the reference's interfaces that go through ATOM ids and the whole EnergyPredictor
// surface (pred/include/EnergyPredictor.h:19-25, EnergyChangePredictorPairSite.h:20-25, VacancyMigrationPredictorQuartic.h:22-24)
// written against the drop-in adapters.
//   ./a.out coefficients.json occupancy.bin <factor> <order: 0 generate | 1 reassigned>
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "lmc_b200_adapters.hpp"

using namespace lmc_b200;

int main(int argc, char **argv) {
  if (argc < 5) { std::fprintf(stderr, "usage: %s coefficients.json occupancy.bin factor order\n", argv[0]); return 2; }
  const size_t f = static_cast<size_t>(std::atoi(argv[3]));
  try {
    const std::set<ElementName> element_set{ElementName::Al, ElementName::Mg, ElementName::Zn};
    cfg::Config config({f, f, f}, std::atoi(argv[4]) ? LMC_ID_ORDER_REASSIGNED : LMC_ID_ORDER_GENERATE, element_set, ElementName::Al);
    std::vector<uint8_t> occ(config.GetNumAtoms());
    std::ifstream ifs(argv[2], std::ios::binary);
    if (!ifs.read(reinterpret_cast<char *>(occ.data()), static_cast<std::streamsize>(occ.size()))) throw std::runtime_error("Cannot open occupancy file");
    config.SetOccupancy(occ);
    const pred::EnergyPredictor energy(argv[1], config);
    const pred::EnergyChangePredictorPairSite pair_site(argv[1], config, element_set);
    const pred::VacancyMigrationPredictorQuartic quartic(argv[1], config, element_set);

    const size_t vac = config.GetVacancyLatticeId();
    std::printf("vacancy lattice %zu atom %zu element %d\n", vac, config.GetVacancyAtomId(), static_cast<int>(config.GetElementAtLatticeId(vac)));
    std::printf("E = %.12f\n", energy.GetEnergy(config));
    const std::vector<size_t> atoms{3, 17, 40, config.GetVacancyAtomId()};
    std::printf("E_cluster = %.12f\n", energy.GetEnergyOfCluster(config, atoms));
    const auto enc = energy.GetEncode(config);
    double sum = 0;
    for (double v : enc) sum += v;
    std::printf("encode n = %zu sum = %.12f\n", enc.size(), sum);
    for (const auto &[el, mu] : energy.GetChemicalPotential(config, ElementName::Al)) std::printf("mu[%d] = %.12f\n", static_cast<int>(el), mu);

    // move the vacancy: atom ids stay with the atoms, lattice ids with the sites (Config::LatticeJump, Config.cpp:431-456)
    const size_t j = config.GetNeighbors(1, vac)[4];
    const size_t moving_atom = config.GetAtomIdFromLatticeId(j), vac_atom = config.GetVacancyAtomId();
    const auto [ea, de] = quartic.GetBarrierAndDiffFromAtomIdPair(config, {vac_atom, moving_atom});
    std::printf("jump by atom ids: Ea = %.12f dE = %+.12f\n", ea, de);
    const double e_before = energy.GetEnergy(config);
    config.LatticeJump({vac, j});
    std::printf("after jump: vacancy lattice %zu atom %zu (same atom: %d), moved atom now at lattice %zu, dE check %.3e\n", config.GetVacancyLatticeId(),
                config.GetVacancyAtomId(), config.GetVacancyAtomId() == vac_atom, config.GetLatticeIdFromAtomId(moving_atom),
                (energy.GetEnergy(config) - e_before) - de);
    std::printf("swap dE by atom ids %+.12f by lattice ids %+.12f\n", pair_site.GetDeFromAtomIdPair(config, {3, 17}),
                pair_site.GetDeFromLatticeIdPair(config, {config.GetLatticeIdFromAtomId(3), config.GetLatticeIdFromAtomId(17)}));
    std::printf("site dE by atom id %+.12f\n", pair_site.GetDeFromAtomIdSite(config, moving_atom, ElementName::Zn));
    std::printf("E_cluster after jump = %.12f\n", energy.GetEnergyOfCluster(config, atoms));

    // deep copy: the clone evolves on its own
    cfg::Config clone = config.Clone();
    clone.LatticeJump({config.GetVacancyLatticeId(), config.GetNeighbors(1, config.GetVacancyLatticeId())[0]});
    std::printf("clone independent: %d (original vacancy %zu, clone vacancy %zu)\n", clone.GetVacancyLatticeId() != config.GetVacancyLatticeId(),
                config.GetVacancyLatticeId(), clone.GetVacancyLatticeId());
  } catch (const std::exception &e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}

// device_common.h -- constants and small device helpers shared by every kernel translation unit.
#pragma once
#include <cstdint>

namespace lmc {

// structural constants of the ordered neighbourhoods (verified against tables.cpp at engine creation)
constexpr int kFirstPos = 21, kSecondPos = 38, kCentrePos = 21;
constexpr int kEnvN = 58, kSiteEnvN = 42;

enum EventError : int { kErrNotNeighbour = 1, kErrNotVacancy = 2, kErrExtraVacancy = 4, kErrBadSite = 8 };

constexpr double kBoltzmannEv = 8.617333262145e-5;   // cfg/include/Constants.hpp:31
constexpr double kSaEpsilon = 1e-4;   // kEpsilon (cfg/include/VectorMatrix.hpp:65) used by SimulatedAnnealing.cpp:85,104

// batch energy totals are accumulated in 2^-44 eV fixed point: integer adds commute, so the total does not depend on the
// order in which thread blocks (or GPUs) contribute -- every rank gets the bit-identical energy (resolution 5.7e-14 eV)
constexpr double kEnergyFixedScale = 17592186044416.0;   // 2^44

#if defined(__CUDACC__)
// 53-bit uniforms like libstdc++'s generate_canonical<double,53> on a 64-bit engine: floor(x / 2^11) * 2^-53
__device__ __forceinline__ double uniform53(uint32_t lo, uint32_t hi) {
  const uint64_t x = (static_cast<uint64_t>(hi) << 32) | lo;
  return static_cast<double>(x >> 11) * (1.0 / 9007199254740992.0);
}
#endif

}  // namespace lmc

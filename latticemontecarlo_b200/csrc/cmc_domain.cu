// cmc_domain.cu -- translation unit of the domain-decomposed CMC / SA driver: the instantiations of the sweep kernel
// (cmc_domain_kernels.cuh) and the two small kernels around it.  Kept apart from engine.cu so that the two dozen
// instantiations compile in parallel with the rest of the library.
#ifdef LMC_DOM_PROFILE
#include <cstdio>
#endif
#include "cmc_domain_kernels.cuh"

namespace lmc {

const void *cmc_domain_kernel_for(int lanes, int k_tab, int speculate, bool small_block) {
  // instantiations: 8 / 16 / 32 lanes per trial (fewer lanes were slower everywhere: 22 or 43 loads per lane), blocks of
  // <= 512 threads (128 registers) or <= 1024 (64 registers)
#define LMC_DOM_CASE(LANES, TAB, SPEC, LL)                                                                      \
  if (lanes == LANES && (k_tab ? 1 : 0) == TAB && speculate == SPEC)                                            \
    return small_block ? reinterpret_cast<const void *>(cmc_domain_kernel<LL, TAB, 512, SPEC>)                  \
                       : reinterpret_cast<const void *>(cmc_domain_kernel<LL, TAB, kDomMaxThreads, SPEC>);
#define LMC_DOM_CASE_SMALL(LANES, TAB, SPEC, LL)                                                                \
  if (lanes == LANES && (k_tab ? 1 : 0) == TAB && speculate == SPEC)                                            \
    return small_block ? reinterpret_cast<const void *>(cmc_domain_kernel<LL, TAB, 512, SPEC>) : nullptr;
  LMC_DOM_CASE(8, 0, 1, 4) LMC_DOM_CASE(8, 1, 1, 4)
  LMC_DOM_CASE(8, 0, 2, 4) LMC_DOM_CASE(8, 1, 2, 4)
  LMC_DOM_CASE(8, 0, 4, 4) LMC_DOM_CASE(8, 1, 4, 4)
  LMC_DOM_CASE(16, 0, 1, 8) LMC_DOM_CASE(16, 1, 1, 8)
  LMC_DOM_CASE_SMALL(16, 0, 2, 8) LMC_DOM_CASE_SMALL(16, 1, 2, 8)
  LMC_DOM_CASE(32, 0, 1, 16) LMC_DOM_CASE(32, 1, 1, 16)
#undef LMC_DOM_CASE
#undef LMC_DOM_CASE_SMALL
  return nullptr;
}

void cmc_domain_state_init(int n_walkers, const CmcState &st, const double *temperatures, DomState *dst, cudaStream_t stream) {
  dom_state_init_kernel<<<static_cast<unsigned>((n_walkers + 127) / 128), 128, 0, stream>>>(n_walkers, st, temperatures, dst);
}

void cmc_domain_refresh_halo(const LatticeDesc &lat, uint8_t *padded, int n_walkers, cudaStream_t stream) {
  const unsigned blocks = static_cast<unsigned>((lat.padded_size + 255) / 256);
  for (int w0 = 0; w0 < n_walkers; w0 += 32768) {
    const unsigned ny = static_cast<unsigned>(n_walkers - w0 < 32768 ? n_walkers - w0 : 32768);
    dom_refresh_halo_kernel<<<dim3(blocks, ny), 256, 0, stream>>>(lat, padded + static_cast<int64_t>(w0) * lat.padded_size);
  }
}

}  // namespace lmc

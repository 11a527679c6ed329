// cmc_domain.cu -- translation unit of the domain-decomposed CMC / SA driver: the instantiations of the sweep kernel
// (cmc_domain_kernels.cuh) and the two small kernels around it.  Kept apart from engine.cu so that the two dozen
// instantiations compile in parallel with the rest of the library.
#include "cmc_domain_kernels.cuh"

namespace lmc {

const void *cmc_domain_kernel_for(int lanes, int k_tab, int speculate, bool small_block) {
#define LMC_DOM_PICK(LL, TT) (small_block ? reinterpret_cast<const void *>(cmc_domain_kernel<LL, TT, 512, 1>) \
                                          : reinterpret_cast<const void *>(cmc_domain_kernel<LL, TT, kDomMaxThreads, 1>))
  switch ((lanes * 2 + (k_tab ? 1 : 0)) * 8 + speculate) {
    case (4 + 0) * 8 + 1: return LMC_DOM_PICK(1, 0);
    case (4 + 1) * 8 + 1: return LMC_DOM_PICK(1, 1);
    case (8 + 0) * 8 + 1: return LMC_DOM_PICK(2, 0);
    case (8 + 1) * 8 + 1: return LMC_DOM_PICK(2, 1);
    case (16 + 0) * 8 + 1: return LMC_DOM_PICK(4, 0);
    case (16 + 1) * 8 + 1: return LMC_DOM_PICK(4, 1);
    case (32 + 0) * 8 + 1: return LMC_DOM_PICK(8, 0);
    case (32 + 1) * 8 + 1: return LMC_DOM_PICK(8, 1);
    case (64 + 0) * 8 + 1: return LMC_DOM_PICK(16, 0);
    case (64 + 1) * 8 + 1: return LMC_DOM_PICK(16, 1);
    default: break;
  }
#undef LMC_DOM_PICK
  if (!small_block) return nullptr;                 // speculative instantiations exist for blocks of <= 512 threads only
  switch ((lanes * 2 + (k_tab ? 1 : 0)) * 8 + speculate) {
    case (16 + 0) * 8 + 2: return reinterpret_cast<const void *>(cmc_domain_kernel<4, 0, 512, 2>);
    case (16 + 1) * 8 + 2: return reinterpret_cast<const void *>(cmc_domain_kernel<4, 1, 512, 2>);
    case (16 + 0) * 8 + 4: return reinterpret_cast<const void *>(cmc_domain_kernel<4, 0, 512, 4>);
    case (16 + 1) * 8 + 4: return reinterpret_cast<const void *>(cmc_domain_kernel<4, 1, 512, 4>);
    case (32 + 0) * 8 + 2: return reinterpret_cast<const void *>(cmc_domain_kernel<8, 0, 512, 2>);
    case (32 + 1) * 8 + 2: return reinterpret_cast<const void *>(cmc_domain_kernel<8, 1, 512, 2>);
    default: return nullptr;
  }
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

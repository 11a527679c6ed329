// grouping.h -- device-side grouping of batched evaluation requests by the lattice region they read, so that neighbouring
// threads of the evaluation kernels gather from the same cache lines (engine.cu decides when to use it; grouping.cu holds
// the kernels and the radix sort, compiled as its own translation unit).
#pragma once
#include <cstdint>

#include <cuda_runtime.h>

#include "lattice.h"

namespace lmc {

constexpr int kGroupKeyBits = 16;          // two 8-bit radix passes
constexpr int kGroupSample = 8192;         // requests inspected to see whether the caller's order is already local

struct GroupPlan {
  int shift;                               // key = (walker * padded_size + padded index of the first site) >> shift
  size_t temp_bytes;                       // radix-sort scratch
  size_t total_bytes;                      // keys in/out + permutation in/out + scratch + the locality counter
};

// workspace layout for n requests (n < 2^31)
GroupPlan group_plan(const LatticeDesc &lat, int n_walkers, int64_t n);
// keys of the first min(n, kGroupSample) requests; *d_local (device) receives the number of adjacent requests whose keys
// differ by at most one bucket -- the caller compares it with the sample size
void group_sample_locality(const LatticeDesc &lat, const GroupPlan &plan, int64_t n, const int32_t *walker, const int64_t *site, unsigned int *d_local,
                           cudaStream_t stream);
// permutation (request indices ordered by key, stable) into workspace; returns the device pointer to it
const uint32_t *group_requests(const LatticeDesc &lat, const GroupPlan &plan, int64_t n, const int32_t *walker, const int64_t *site, void *workspace,
                               cudaStream_t stream);

}  // namespace lmc

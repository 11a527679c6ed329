// grouping.cu -- see grouping.h.  The evaluation kernels (barrier_kernel, swap_de_rows_kernel) gather 60 / 2 x 43 bytes
// around the sites of a request; with requests in arbitrary order every lane of a warp touches its own sectors (one L1
// wavefront and one L1 -> L2 sector per lane and load instruction: the measured bound of both kernels).  Ordering the
// requests by the padded index of their first site makes neighbouring lanes read the same rows.  The order only has to be
// LOCAL, not exact: the key keeps the 16 leading bits of (walker, padded index), i.e. two radix passes; the sort is stable,
// so requests of one bucket keep the caller's order.  The kernels then process request perm[q] in thread q and write its
// results to slot perm[q]: outputs stay in the caller's order.
#include <cub/device/device_radix_sort.cuh>

#include "grouping.h"

namespace lmc {

namespace {

__device__ __forceinline__ uint32_t group_key(const LatticeDesc &lat, int shift, int64_t q, const int32_t *walker, const int64_t *site) {
  const int64_t id = site[q];
  if (id < 0 || id >= lat.num_sites) return 0u;                    // the evaluation kernel reports bad ids
  const uint64_t cell = static_cast<uint64_t>(walker ? walker[q] : 0) * static_cast<uint64_t>(lat.padded_size) +
                        static_cast<uint64_t>(lat.padded_index_of_id(id));
  return static_cast<uint32_t>(cell >> shift) & ((1u << kGroupKeyBits) - 1u);
}

__global__ void group_keys_kernel(LatticeDesc lat, int shift, int64_t n, const int32_t *__restrict__ walker, const int64_t *__restrict__ site,
                                  uint32_t *__restrict__ keys, uint32_t *__restrict__ index) {
  const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (q >= n) return;
  keys[q] = group_key(lat, shift, q, walker, site);
  index[q] = static_cast<uint32_t>(q);
}

__global__ void group_locality_kernel(LatticeDesc lat, int shift, int64_t n, const int32_t *__restrict__ walker, const int64_t *__restrict__ site,
                                      unsigned int *__restrict__ local) {
  const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  bool near = false;
  if (q + 1 < n) {
    const int64_t d = static_cast<int64_t>(group_key(lat, shift, q, walker, site)) - static_cast<int64_t>(group_key(lat, shift, q + 1, walker, site));
    near = d >= -1 && d <= 1;
  }
  const unsigned votes = __popc(__ballot_sync(0xFFFFFFFFu, near));
  if ((threadIdx.x & 31) == 0 && votes) atomicAdd(local, votes);
}

size_t align256(size_t v) { return (v + 255) & ~static_cast<size_t>(255); }

}  // namespace

GroupPlan group_plan(const LatticeDesc &lat, int n_walkers, int64_t n) {
  GroupPlan plan{};
  const uint64_t cells = static_cast<uint64_t>(n_walkers > 0 ? n_walkers : 1) * static_cast<uint64_t>(lat.padded_size);
  int bits = 0;
  while ((1ULL << bits) < cells) ++bits;
  plan.shift = bits > kGroupKeyBits ? bits - kGroupKeyBits : 0;
  uint32_t *nil = nullptr;
  cub::DeviceRadixSort::SortPairs(nullptr, plan.temp_bytes, nil, nil, nil, nil, static_cast<int>(n), 0, kGroupKeyBits);
  plan.total_bytes = 4 * align256(static_cast<size_t>(n) * 4) + align256(plan.temp_bytes) + 256;
  return plan;
}

void group_sample_locality(const LatticeDesc &lat, const GroupPlan &plan, int64_t n, const int32_t *walker, const int64_t *site, unsigned int *d_local,
                           cudaStream_t stream) {
  const int64_t m = n < kGroupSample ? n : kGroupSample;
  cudaMemsetAsync(d_local, 0, sizeof(unsigned int), stream);
  group_locality_kernel<<<static_cast<unsigned>((m + 255) / 256), 256, 0, stream>>>(lat, plan.shift, m, walker, site, d_local);
}

const uint32_t *group_requests(const LatticeDesc &lat, const GroupPlan &plan, int64_t n, const int32_t *walker, const int64_t *site, void *workspace,
                               cudaStream_t stream) {
  char *p = static_cast<char *>(workspace);
  const size_t stride = align256(static_cast<size_t>(n) * 4);
  uint32_t *keys_in = reinterpret_cast<uint32_t *>(p), *keys_out = reinterpret_cast<uint32_t *>(p + stride);
  uint32_t *index_in = reinterpret_cast<uint32_t *>(p + 2 * stride), *index_out = reinterpret_cast<uint32_t *>(p + 3 * stride);
  void *temp = p + 4 * stride;
  size_t temp_bytes = plan.temp_bytes;
  group_keys_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(lat, plan.shift, n, walker, site, keys_in, index_in);
  cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, index_in, index_out, static_cast<int>(n), 0, kGroupKeyBits, stream);
  return index_out;
}

}  // namespace lmc

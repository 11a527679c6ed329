// cmc_domain.h -- plain-data interface between the engine (engine.cu) and the domain-decomposed CMC / SA driver, which is
// compiled as its own translation unit (cmc_domain.cu: the sweep kernel has two dozen instantiations).
#pragma once
#include <cstdint>

#include "device_tables.h"
#include <cuda_runtime.h>

#include "cmc_state.h"
#include "lattice.h"

namespace lmc {

constexpr int kDomWorldMax = 8;
constexpr int kSiteEnvCount = 42;

constexpr int kDomMaxWalkers = 2048;     // replicas one launch can drive (per-replica temperature lives in shared memory)
constexpr int kDomMaxThreads = 1024;
constexpr int kDomPidxBytes = (kSiteEnvCount * kSiteEnvCount + 15) & ~15;

struct DomState {                        // per-replica chain state between sweeps
  double energy;
  unsigned long long steps, accepted;
  double temperature;                    // fixed temperature (CMC) or the schedule's current one (SA)
  SaSchedule sa;
};

struct DomLine { unsigned long long v[4]; unsigned long long flag; unsigned long long pad[3]; };   // 64 bytes

struct CmcDomainParams {
  int ndx, ndy, ndz;                     // domains per axis
  int tile_y, tile_zh;                   // tile strides: index = (tx * tile_y + ty) * tile_zh + (tz >> 1)
  int tile_cells;                        // bytes of the tile proper (multiple of 16)
  int tile_bytes;                        // shared memory per team: tile + species rows (96 B per group) + solute list (2 B per core site)
  int max_core;                          // largest core of any domain (sites)
  int rounds;                            // Metropolis rounds per domain and sweep (each up to kDomTries candidate draws)
  int n_walkers;
  int world, rank;
  uint8_t *occ[2];                       // double-buffered occupancy, padded layout, [walker][padded_size]
  uint8_t *peer_occ[2][kDomWorldMax];   // the same buffers of every rank (peer mappings; [.][rank] = own)
  DomState *state;                       // [2][n_walkers]
  unsigned long long *accum;             // [3][n_walkers][4]: sum dE (2^-44 eV fixed point), trials, accepted, errors
  DomLine *lines;                        // [2][world] sweep totals of every rank (multi-GPU; written by the peers)
  DomLine *peer_lines[kDomWorldMax];
  unsigned long long *barrier_counter;
  int *abort_flag;
  long long spin_limit;
  unsigned long long *sweep;             // sweeps done so far (persists over launches; Philox counter)
  unsigned long long *line_seq;          // inter-GPU line sequence (never reset while the peers are attached)
  unsigned int *queue;                   // [3] next domain of the sweep (dynamic distribution over the lane groups), by sweep % 3
  unsigned long long own_mask[16];       // environment positions whose table walk lane `sub` of a side owns (balanced by pair counts)
};

// first x index (inclusive) of rank r's slab of the domain grid
LMC_HD int domain_slab_begin(int ndx, int world, int r) { return static_cast<int>((static_cast<unsigned>(ndx) * static_cast<unsigned>(r)) / static_cast<unsigned>(world)); }
// lower bound (inclusive) of domain i along an axis of `period` half-units cut into nd parts (z: even bounds, see below)
LMC_HD int domain_lo(int i, int period, int nd) { return static_cast<int>((static_cast<unsigned>(i) * static_cast<unsigned>(period)) / static_cast<unsigned>(nd)); }   // i <= nd, period < 2^15: 32 bits
LMC_HD int domain_lo_z(int i, int fz, int nd) { return 2 * static_cast<int>((static_cast<unsigned>(i) * static_cast<unsigned>(fz)) / static_cast<unsigned>(nd)); }


// entry points of cmc_domain.cu
const void *cmc_domain_kernel_for(int lanes, int k_tab, int speculate, bool small_block);
void cmc_domain_state_init(int n_walkers, const CmcState &st, const double *temperatures, DomState *dst, cudaStream_t stream);
void cmc_domain_refresh_halo(const LatticeDesc &lat, uint8_t *padded, int n_walkers, cudaStream_t stream);

}  // namespace lmc

// device_tables.h -- POD view of the constant tables as the kernels see them (all pointers are device memory).
#pragma once
#include <cstdint>

#include "lattice.h"

namespace lmc {

constexpr int kPairDeltaStride = 61;   // 60 offsets + 1 pad word: rows of different (direction, z-parity) hit different banks

struct DevTables {
  // geometry
  const int32_t *pair_delta;     // [12*2][61]  padded-layout offset of state position t from the first site, row = k*2 + zpar
  const int32_t *site_delta;     // [2][43]     same for the 43-site neighbourhood, row = zpar
  const int8_t *dir_lut;         // [27]        (dx+1)*9 + (dy+1)*3 + (dz+1) -> direction k, -1 if not a first neighbour
  const int8_t *nn1;             // [12][4]     first-neighbour vectors (x,y,z,0)
  const uint8_t *kmc_slot;       // [64][12]    event order (rank of the neighbour's lattice id) of jump k by the vacancy's boundary / parity class
  const int8_t *frame_p;         // [12][4]     canonical perpendicular first neighbour of each direction
  const int8_t *pair_off;        // [12][2][60][4] lattice offsets of the ordered pair neighbourhood (debug taps)
  const int8_t *site_off;        // [43][4]
  int32_t pair_first_pos, pair_second_pos, site_centre_pos;
  // KMC box scan: the 7 x 7 x 4 padded cells around a vacancy that cover the neighbourhoods of all 12 jumps
  const int32_t *box_delta;      // [2][196]      padded-layout offset of box cell c from the vacancy, row = z parity
  const int8_t *box_envpos;      // [2][12][196]  env index (0..57) of box cell c in the neighbourhood of jump k; 58 / 59 =
                                 //               first / second site of the pair; -1 = not part of that neighbourhood
  // chemistry
  int32_t n_species;             // species without vacancy; vacancy code == n_species
  int32_t solvent;               // compact code of the expansion origin
  // contracted jump tables (delta form): Q = C[m] + sum A[m][t][e] + sum B[m][p][a][b], 3 quantities each
  const double *pair_C, *pair_A, *pair_B;
  // the same folded to the two numbers a barrier needs: (dE, log E0) with log E0 = logKs + 2 logD   [..][2]
  const double *pair_C2, *pair_A2, *pair_B2;
  int32_t barrier_model;         // 0: quartic closed form Ea(dE, E0 = Ks D^2); 1: E0 model, Ea = max(0, e0 + dE/2), log e0 in slot 2
  const uint64_t *pair_mask_hi;  // [58]
  const uint16_t *pair_base;     // [58]
  int32_t n_pair_pairs;          // 556
  // contracted site tables: H(x) = C[x] + sum A[x][t][e] + sum B[x][p][a][b]
  const double *site_C, *site_A, *site_B;
  const uint64_t *site_mask_hi;  // [42]
  const uint16_t *site_base;     // [42]
  int32_t n_site_pairs;          // 204
  // total energy
  const double *e_single, *e_pair, *e_triplet;
  const uint8_t *e_walk;         // [144][4]: (pos2, pos3, label, 0) positions in the 43-list
  const uint8_t *e_shell_pos;    // [42][2]: (pos, shell) of every neighbour in the 43-list
  int32_t n_e_walk;
  // debug taps: the reference's cluster mappings in flat form
  const int16_t *type_lut;       // [8][m][m][m]
  int32_t n_types;
  const int8_t *map_state_pair;  // [473][4]: label, pos1, pos2, pos3 (-1 padded)
  int32_t n_state_pair;
  const int8_t *map_state_site;  // [247][4]
  int32_t n_state_site;
  const int16_t *map_mmm;        // [614][4]: slot offset of the group, a, b (-1 for singlets), flags (bit0 symmetric)
  const int16_t *map_mm2;        // [614][4]
  int32_t n_avg_clusters, len_mmm, len_mm2;
  const int8_t *env_of_list;     // [4][58]: mmm, mm2, mm2 backward variant 0, variant 1 -> env index
  const int8_t *state_pos_of_env;  // [58]
};

}  // namespace lmc

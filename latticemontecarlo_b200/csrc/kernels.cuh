// kernels.cuh -- hand-written sm_100a kernels of the LatticeMC hot path.
//
// Design: the reference evaluates one candidate event by ~2800 string/hash-keyed cluster lookups
// (pred/src/VacancyMigrationPredictorQuartic.cpp:112-246).  Here every quantity of an event is the contracted form
//     Q = C[m] + sum_t A[m][t][e_t] + sum_{(t,u)} B[m][(t,u)][e_t][e_u]          (tables.h)
// over the 58 environment sites, stored relative to the solvent species so that only solute sites contribute.
// The work of an event is therefore (i) one gather of 60 occupancy bytes through constant offset tables -- the
// HBM/L2-bound part -- and (ii) a short data-dependent walk over the solute bits.  One thread owns one event.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "device_common.h"
#include "device_tables.h"

namespace lmc {

// structural constants of the ordered neighbourhoods (verified against tables.cpp at engine creation)
constexpr int kBoxRows = 48;            // KMC box scan: (dx, dy) rows of the 7 x 7 box around a vacancy that hold a neighbourhood site of
                                        // some jump (the 4 corner rows hold none: 45 rows, padded to 3 x 16 for the half-warp scan)
constexpr int kBoxCells = kBoxRows * 4; // 4 consecutive z slots (cells of the padded layout) per row
constexpr int kBoxCentreRow = 22;       // position of the vacancy's own row (dx = dy = 0) among the kept rows


// The environment of an event is summarised by ONE bit mask over the env index: bit t set <=> the species at env
// site t differs from the solvent.  Only those sites contribute to the contracted (delta-form) tables, and their
// species codes are re-read from the occupancy when they are visited (an L1 hit; solute sites are rare).

__device__ __forceinline__ int direction_of(const LatticeDesc &lat, const DevTables &tab, int xi, int yi, int zi, int xj, int yj,
                                            int zj) {
  int dx = xj - xi, dy = yj - yi, dz = zj - zi;
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  dx = dx > px / 2 ? dx - px : (dx < -px / 2 ? dx + px : dx);
  dy = dy > py / 2 ? dy - py : (dy < -py / 2 ? dy + py : dy);
  dz = dz > pz / 2 ? dz - pz : (dz < -pz / 2 ? dz + pz : dz);
  if (dx < -1 || dx > 1 || dy < -1 || dy > 1 || dz < -1 || dz > 1) return -1;
  return tab.dir_lut[(dx + 1) * 9 + (dy + 1) * 3 + (dz + 1)];
}

// Closed form of pred/src/VacancyMigrationPredictorQuartic.cpp:266-275, algebraically reduced.  With
//   b = 4 dE / D^3,  a = Ks / (4 D^2),  c = (9 b^2 - 16 a^2 D^2) / (32 a)   one has   9 b^2 - 32 a c = 16 a^2 D^2,
// so delta = 4 a D, and with  E0 = Ks D^2  and  x = b / (a D) = 16 dE / E0  the barrier is
//   Ea = E0 (3x + 4)^2 (8 + 4x - 1.5 x^2) / 8192        (dE = 0  =>  Ea = Ks D^2 / 64).
// log_e0 = logKs + 2 logD comes straight from the contracted tables, so one exp and one division suffice.
__device__ __forceinline__ double quartic_barrier_log(double dE, double log_e0) {
  const double e0 = exp(log_e0);
  // dE is exactly 0 for a solvent atom in a pure-solvent neighbourhood (9 % of the events of a 4 % alloy, i.e. some lane of
  // almost every warp), and a zero QUOTIENT sends the whole warp through the slow path of the fp64 division (~60
  // instructions): divide a non-zero stand-in and put the zero back
  const bool zero = dE == 0.0;
  double num = zero ? 1.0 : 16.0 * dE;
  asm volatile("" : "+d"(num));             // keep the stand-in: the compiler would otherwise divide 16 dE for every lane again
  double x = num / e0;
  x = zero ? 0.0 : x;
  const double s = 3.0 * x + 4.0;
  return e0 * (s * s) * (8.0 + 4.0 * x - 1.5 * (x * x)) * (1.0 / 8192.0);
}

// Barrier from the folded pair (dE, log E0) for either coefficient model.  model 1 = VacancyMigrationPredictorE0
// (pred/src/VacancyMigrationPredictorE0.cpp:153-159): Ea = max(0, e0 + dE / 2) with log e0 in the second slot.
__device__ __forceinline__ double barrier_from_folded(double dE, double log_e0, int model) {
  if (model == 0) return quartic_barrier_log(dE, log_e0);
  return fmax(0.0, exp(log_e0) + 0.5 * dE);
}

// env index of a state-list position (the jump pair sits at positions 21 and 38)
__host__ __device__ constexpr int env_of_pos(int t) { return t - (t > kFirstPos) - (t > kSecondPos); }
__device__ __forceinline__ int pos_of_env(int e) { return e + (e >= kFirstPos) + (e >= kSecondPos - 1); }

// Gather the 60 ordered sites of a jump (first site at padded index `base`; drow = offset row of its direction and
// z parity) and return the solute mask over the 58 env sites plus the species at the two pair sites.
__device__ __forceinline__ uint64_t gather_pair_env(const uint8_t *occ, int64_t base, const int32_t *__restrict__ drow,
                                                    unsigned solvent, unsigned *first, unsigned *mig) {
  unsigned codes[60];
#pragma unroll
  for (int t = 0; t < 60; ++t) codes[t] = occ[base + drow[t]];
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int t = 0; t < 60; ++t) {
    if (t == kFirstPos || t == kSecondPos) continue;
    const int e = env_of_pos(t);
    if (e < 32) lo |= (codes[t] != solvent) ? (1u << e) : 0u;
    else hi |= (codes[t] != solvent) ? (1u << (e - 32)) : 0u;
  }
  *first = codes[kFirstPos];
  *mig = codes[kSecondPos];
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// Walk the solute bits of a jump environment and accumulate the three contracted quantities (dE, logD, logKs).
// Returns false if a vacancy sits in the environment (the reference has no cluster type for two vacancies).
__device__ __forceinline__ bool accumulate_pair_tables(const DevTables &tab, int m, uint64_t sol, const uint8_t *occ, int64_t base,
                                                       const int32_t *__restrict__ drow, double acc[3]) {
  const int n = tab.n_species;
  const double *__restrict__ C = tab.pair_C + m * 3;
  acc[0] = __ldg(C); acc[1] = __ldg(C + 1); acc[2] = __ldg(C + 2);
  const double *__restrict__ A = tab.pair_A + static_cast<size_t>(m) * kEnvN * n * 3;
  const double *__restrict__ B = tab.pair_B + static_cast<size_t>(m) * tab.n_pair_pairs * n * n * 3;
  bool ok = true;
  while (sol) {
    const int t = __ffsll(static_cast<long long>(sol)) - 1;
    sol &= sol - 1;
    const int et = occ[base + drow[pos_of_env(t)]];
    if (et >= n) { ok = false; continue; }
    const double *a = A + (t * n + et) * 3;
    acc[0] += __ldg(a); acc[1] += __ldg(a + 1); acc[2] += __ldg(a + 2);
    const uint64_t hi = __ldg(tab.pair_mask_hi + t);
    uint64_t partners = hi & sol;
    const int pbase = __ldg(tab.pair_base + t);
    while (partners) {
      const int u = __ffsll(static_cast<long long>(partners)) - 1;
      partners &= partners - 1;
      const int eu = occ[base + drow[pos_of_env(u)]];
      if (eu >= n) { ok = false; continue; }
      const int p = pbase + __popcll(hi & ((1ULL << u) - 1ULL));
      const double *b = B + ((static_cast<size_t>(p) * n + et) * n + eu) * 3;
      acc[0] += __ldg(b); acc[1] += __ldg(b + 1); acc[2] += __ldg(b + 2);
    }
  }
  return ok;
}

// ----------------------------------------------------------------------------------------------- occupancy I/O
// element enum codes by lattice id  ->  compact codes in the padded layout (halo cells replicate their periodic image)
// blockIdx.y = walker (both arrays are walker-major)
__global__ void upload_occupancy_kernel(LatticeDesc lat, const uint8_t *__restrict__ occ_by_id, uint8_t *__restrict__ padded,
                                        const int8_t *__restrict__ code_of_enum, int *__restrict__ error) {
  const int64_t cell = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (cell >= lat.padded_size) return;
  occ_by_id += blockIdx.y * lat.num_sites;
  padded += blockIdx.y * lat.padded_size;
  const int zi = static_cast<int>(cell % lat.nz);
  const int64_t r = cell / lat.nz;
  const int yp = static_cast<int>(r % lat.ny);
  const int xp = static_cast<int>(r / lat.ny);
  const int X = wrap_coord(xp - kHalo, 2 * lat.fx), Y = wrap_coord(yp - kHalo, 2 * lat.fy);
  // z index zi holds Z + 4 in {2 zi, 2 zi + 1}; the parity that makes X+Y+Z even is the real site
  const int Zp = 2 * zi + (((xp - kHalo) + (yp - kHalo)) & 1);
  const int Z = wrap_coord(wrap_coord(Zp - kHaloZ, 2 * lat.fz), 2 * lat.fz);
  const int enum_code = occ_by_id[lat.id_of_coords(X, Y, Z)];
  const int code = enum_code < 16 ? code_of_enum[enum_code] : -1;
  if (code < 0) { atomicOr(error, kErrBadSite); padded[cell] = 0; return; }
  padded[cell] = static_cast<uint8_t>(code);
}

__global__ void download_occupancy_kernel(LatticeDesc lat, const uint8_t *__restrict__ padded, uint8_t *__restrict__ occ_by_id,
                                          const uint8_t *__restrict__ enum_of_code) {
  const int64_t id = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (id >= lat.num_sites) return;
  occ_by_id += blockIdx.y * lat.num_sites;
  padded += blockIdx.y * lat.padded_size;
  occ_by_id[id] = enum_of_code[padded[lat.padded_index_of_id(id)]];
}

// write one site and all its periodic halo images.  A coordinate within `halo` of a face has exactly one image along
// that axis (periods are >= 8 half-units > 2 * halo), so there are at most 7 images: the non-empty subsets of the axes.
__device__ __forceinline__ void store_site(const LatticeDesc &lat, uint8_t *occ, int X, int Y, int Z, uint8_t code) {
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  const int64_t base = lat.padded_index(X, Y, Z);
  occ[base] = code;
  const int sx = X < kHalo ? px : (X >= px - kHalo ? -px : 0);
  const int sy = Y < kHalo ? py : (Y >= py - kHalo ? -py : 0);
  const int sz = Z < kHaloZ ? pz : (Z >= pz - kHaloZ ? -pz : 0);
  if ((sx | sy | sz) == 0) return;                                      // interior site: no image
  const int64_t ix = static_cast<int64_t>(sx) * lat.ny * lat.nz, iy = static_cast<int64_t>(sy) * lat.nz, iz = sz / 2;
  if (sx) occ[base + ix] = code;
  if (sy) occ[base + iy] = code;
  if (sz) occ[base + iz] = code;
  if (sx && sy) occ[base + ix + iy] = code;
  if (sx && sz) occ[base + ix + iz] = code;
  if (sy && sz) occ[base + iy + iz] = code;
  if (sx && sy && sz) occ[base + ix + iy + iz] = code;
}

// one periodic image per caller: `which` in 0..7 selects the subset of axes (bit 0: x, bit 1: y, bit 2: z) whose image
// shift is applied; subsets that do not exist for this site write nothing.  which == 0 is the site itself.
__device__ __forceinline__ void store_site_image(const LatticeDesc &lat, uint8_t *occ, int X, int Y, int Z, int which, uint8_t code) {
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  const int sx = X < kHalo ? px : (X >= px - kHalo ? -px : 0);
  const int sy = Y < kHalo ? py : (Y >= py - kHalo ? -py : 0);
  const int sz = Z < kHaloZ ? pz : (Z >= pz - kHaloZ ? -pz : 0);
  const bool ok = (!(which & 1) || sx) && (!(which & 2) || sy) && (!(which & 4) || sz);
  if (!ok) return;
  const int64_t idx = lat.padded_index(X, Y, Z) + ((which & 1) ? static_cast<int64_t>(sx) * lat.ny * lat.nz : 0) +
                      ((which & 2) ? static_cast<int64_t>(sy) * lat.nz : 0) + ((which & 4) ? sz / 2 : 0);
  occ[idx] = code;
}

__global__ void lattice_jump_kernel(LatticeDesc lat, uint8_t *occ, int64_t a, int64_t b) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int xa, ya, za, xb, yb, zb;
  lat.coords_of_id(a, xa, ya, za);
  lat.coords_of_id(b, xb, yb, zb);
  const uint8_t ca = occ[lat.padded_index(xa, ya, za)], cb = occ[lat.padded_index(xb, yb, zb)];
  store_site(lat, occ, xa, ya, za, cb);
  store_site(lat, occ, xb, yb, zb, ca);
}

// ----------------------------------------------------------------------------------------------- barriers
// One thread per candidate event.  Algorithmic traffic per event: 60 B occupancy + 240 B of
// neighbour indices (here replaced by a 61-word offset row that lives in shared memory) + 16 B of output.
constexpr int kBarrierThreads = 128;

__global__ void __launch_bounds__(kBarrierThreads)
barrier_kernel(LatticeDesc lat, DevTables tab, const uint8_t *__restrict__ occ, int64_t walker_stride, int64_t n,
               const int32_t *__restrict__ walker, const int64_t *__restrict__ site_i, const int64_t *__restrict__ site_j,
               double *__restrict__ Ea, double *__restrict__ dE, double *__restrict__ D_out, double *__restrict__ Ks_out,
               int *__restrict__ error, const uint32_t *__restrict__ perm) {
  __shared__ int32_t s_delta[24 * kPairDeltaStride];
  for (int q = threadIdx.x; q < 24 * kPairDeltaStride; q += blockDim.x) s_delta[q] = tab.pair_delta[q];
  __syncthreads();
  const int64_t q_thread = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (q_thread >= n) return;
  const int64_t e = perm ? static_cast<int64_t>(perm[q_thread]) : q_thread;      // grouped by lattice region (grouping.h); results in caller order
  const int64_t i = site_i[e], j = site_j[e];
  const int w = walker ? walker[e] : 0;
  const double nan = CUDART_NAN;
  int err = 0;
  double out_ea = nan, out_de = nan, out_d = nan, out_ks = nan;
  if (i < 0 || i >= lat.num_sites || j < 0 || j >= lat.num_sites) {
    err = kErrBadSite;
  } else {
    int xi, yi, zi, xj, yj, zj;
    lat.coords_of_id(i, xi, yi, zi);
    lat.coords_of_id(j, xj, yj, zj);
    const int k = direction_of(lat, tab, xi, yi, zi, xj, yj, zj);
    if (k < 0) {
      err = kErrNotNeighbour;
    } else {
      const uint8_t *o = occ + w * walker_stride;
      const int64_t base = lat.padded_index(xi, yi, zi);
      const int32_t *drow = s_delta + (k * 2 + (zi & 1)) * kPairDeltaStride;
      unsigned first = 0, mig = 0;
      const uint64_t sol = gather_pair_env(o, base, drow, static_cast<unsigned>(tab.solvent), &first, &mig);
      const unsigned vac = static_cast<unsigned>(tab.n_species);
      if (first != vac || mig == vac) err = kErrNotVacancy;
      else {
        double acc[3];
        if (!accumulate_pair_tables(tab, static_cast<int>(mig), sol, o, base, drow, acc)) err = kErrExtraVacancy;
        else {
          out_de = acc[0];
          out_ea = barrier_from_folded(out_de, acc[2] + 2.0 * acc[1], tab.barrier_model);
          if (D_out) out_d = exp(acc[1]);
          if (Ks_out) out_ks = exp(acc[2]);
        }
      }
    }
  }
  if (err) atomicOr(error, err);
  Ea[e] = out_ea;
  dE[e] = out_de;
  if (D_out) D_out[e] = out_d;
  if (Ks_out) Ks_out[e] = out_ks;
}

// ----------------------------------------------------------------------------------------------- site / swap dE
// Site neighbourhood (43 ordered sites, centre at position 21): solute mask over the 42 env sites.  In the coupled
// swap case one env site (`override_index`, a padded index) is to be seen with `override_code` instead of memory.
__device__ __forceinline__ uint64_t gather_site_env(const uint8_t *occ, int64_t base, const int32_t *__restrict__ drow,
                                                    unsigned solvent, unsigned *centre, int64_t override_index, unsigned override_code) {
  unsigned codes[43];
#pragma unroll
  for (int t = 0; t < 43; ++t) codes[t] = occ[base + drow[t]];
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int t = 0; t < 43; ++t) {
    if (t == kCentrePos) continue;
    const int e = t - (t > kCentrePos);
    unsigned c = codes[t];
    if (base + drow[t] == override_index) c = override_code;
    if (e < 32) lo |= (c != solvent) ? (1u << e) : 0u;
    else hi |= (c != solvent) ? (1u << (e - 32)) : 0u;
  }
  *centre = codes[kCentrePos];
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// H(x_new, env) - H(x_old, env) with the contracted site tables
__device__ __forceinline__ double site_energy_change(const DevTables &tab, int x_old, int x_new, uint64_t sol, const uint8_t *occ,
                                                     int64_t base, const int32_t *__restrict__ drow, int64_t override_index,
                                                     unsigned override_code) {
  const int m = tab.n_species + 1;
  const size_t a_stride = static_cast<size_t>(kSiteEnvN) * m, b_stride = static_cast<size_t>(tab.n_site_pairs) * m * m;
  const double *__restrict__ A_new = tab.site_A + x_new * a_stride, *__restrict__ A_old = tab.site_A + x_old * a_stride;
  const double *__restrict__ B_new = tab.site_B + x_new * b_stride, *__restrict__ B_old = tab.site_B + x_old * b_stride;
  double acc = __ldg(tab.site_C + x_new) - __ldg(tab.site_C + x_old);
  auto code_at = [&](int e) -> int {
    const int64_t idx = base + drow[e + (e >= kCentrePos)];
    return idx == override_index ? static_cast<int>(override_code) : static_cast<int>(occ[idx]);
  };
  while (sol) {
    const int t = __ffsll(static_cast<long long>(sol)) - 1;
    sol &= sol - 1;
    const int et = code_at(t);
    acc += __ldg(A_new + t * m + et) - __ldg(A_old + t * m + et);
    const uint64_t hi = __ldg(tab.site_mask_hi + t);
    uint64_t partners = hi & sol;
    const int pbase = __ldg(tab.site_base + t);
    while (partners) {
      const int u = __ffsll(static_cast<long long>(partners)) - 1;
      partners &= partners - 1;
      const int eu = code_at(u);
      const size_t p = (static_cast<size_t>(pbase + __popcll(hi & ((1ULL << u) - 1ULL))) * m + et) * m + eu;
      acc += __ldg(B_new + p) - __ldg(B_old + p);
    }
  }
  return acc;
}

// EnergyChangePredictorPairSite::GetDeFromLatticeIdPair (pred/src/EnergyChangePredictorPairSite.cpp:70-146).
// Uncoupled pairs: dE_site(a -> e_b) + dE_site(b -> e_a) on the unchanged occupancy.  Coupled pairs (b within the
// third shell of a): the two single-site changes are applied one after the other (the second sees the first), which
// is the same sum over "clusters touching a or b" that the reference recounts on the fly (:96-134).
__device__ __forceinline__ double swap_energy_change(const LatticeDesc &lat, const DevTables &tab, const uint8_t *__restrict__ occ,
                                                     const int32_t *__restrict__ s_delta, int xa, int ya, int za, int xb, int yb,
                                                     int zb, int *err, bool first_neighbours_only = false) {
  const unsigned solvent = static_cast<unsigned>(tab.solvent), vac = static_cast<unsigned>(tab.n_species);
  (void)vac;
  int64_t base_a = lat.padded_index(xa, ya, za), base_b = lat.padded_index(xb, yb, zb);
  unsigned ea = occ[base_a], eb = occ[base_b];
  if (ea == eb) return 0.0;
  int dx = xb - xa, dy = yb - ya, dz = zb - za;
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  dx = dx > px / 2 ? dx - px : (dx < -px / 2 ? dx + px : dx);
  dy = dy > py / 2 ? dy - py : (dy < -py / 2 ? dy + py : dy);
  dz = dz > pz / 2 ? dz - pz : (dz < -pz / 2 ? dz + pz : dz);
  const int r2 = dx * dx + dy * dy + dz * dz;
  const bool coupled = r2 <= 6;
  if (first_neighbours_only && r2 != 2) {     // EnergyChangePredictorPair: only first-neighbour pairs have a list (.at throws, :83)
    *err |= kErrNotNeighbour;
    return CUDART_NAN;
  }
  int zpa = za & 1, zpb = zb & 1;
  if (coupled && eb == vac) {   // move the vacancy first so that no intermediate state holds two vacancies
    const int64_t tb = base_a; base_a = base_b; base_b = tb;
    const unsigned te = ea; ea = eb; eb = te;
    const int tz = zpa; zpa = zpb; zpb = tz;
    dx = -dx; dy = -dy; dz = -dz;
  }
  unsigned centre;
  const int32_t *row_a = s_delta + zpa * 43, *row_b = s_delta + zpb * 43;
  uint64_t sol = gather_site_env(occ, base_a, row_a, solvent, &centre, -1, 0);
  double de = site_energy_change(tab, static_cast<int>(ea), static_cast<int>(eb), sol, occ, base_a, row_a, -1, 0);
  // second site; in the coupled case it sees site a already holding e_b.  The halo images of a are not updated in
  // memory, so the override is applied by *position*: a sits at displacement (-dx,-dy,-dz) from b.
  int64_t override_index = -1;
  if (coupled) override_index = base_b + lat.padded_delta(-dx, -dy, -dz, zpb);
  sol = gather_site_env(occ, base_b, row_b, solvent, &centre, override_index, eb);
  de += site_energy_change(tab, static_cast<int>(eb), static_cast<int>(ea), sol, occ, base_b, row_b, override_index, eb);
  if (de != de) *err |= kErrExtraVacancy;   // NaN: a cluster type the reference has no index for
  return de;
}

// ----------------------------------------------------------------------------------------------- row-based site gather
// The 43-site neighbourhood in list order (GetSortedLatticeVectorStateOfSite: lexicographic in (dx, dy, dz)) is 21 rows
// of constant (dx, dy), each 1-3 consecutive cells of the padded layout (z is the fast axis with the parity squeezed out):
//   kind 0: dz = -2, 0, +2  -> cells cz-1, cz, cz+1          kind 1: dz = 0 -> cell cz
//   kind 2: dz = -1, +1     -> cells cz-1, cz (Z even) or cz, cz+1 (Z odd)
// so ONE load per row (an aligned 8-byte word; a second one only when the 1-3 bytes straddle it) replaces 43 byte loads,
// and list positions are compile-time constants.  Species codes are kept as 4-bit nibbles in registers (the table walk
// never goes back to memory), the solute mask is built from the same words.  The row table is verified against the host
// geometry (site_delta) at engine creation.
constexpr int kSiteRows = 21;
__host__ __device__ constexpr int site_row_dx(int r) { return r < 3 ? -2 : (r < 8 ? -1 : (r < 13 ? 0 : (r < 18 ? 1 : 2))); }
__host__ __device__ constexpr int site_row_dy(int r) { return r < 3 ? r - 1 : (r < 8 ? r - 5 : (r < 13 ? r - 10 : (r < 18 ? r - 15 : r - 19))); }
__host__ __device__ constexpr int site_row_kind(int r) {
  const int dx = site_row_dx(r), dy = site_row_dy(r);
  if ((dx + dy) & 1) return 2;
  return (dx == 2 || dx == -2 || dy == 2 || dy == -2) ? 1 : 0;
}
__host__ __device__ constexpr int site_row_count(int r) { return site_row_kind(r) == 0 ? 3 : (site_row_kind(r) == 1 ? 1 : 2); }
__host__ __device__ constexpr int site_row_pos(int r) { return r == 0 ? 0 : site_row_pos(r - 1) + site_row_count(r - 1); }
static_assert(site_row_pos(kSiteRows - 1) + site_row_count(kSiteRows - 1) == 43, "row table must cover the 43-site list");
static_assert(site_row_pos(10) + 1 == kCentrePos, "centre row");

struct SiteEnvRegs {
  uint64_t n0, n1, n2; // species code of list position t in nibble t (16 positions per word)
  uint64_t sol;        // bit e (env index, centre removed) set <=> species != solvent
  __device__ __forceinline__ unsigned code_at_pos(int t) const {
    const uint64_t w = t < 16 ? n0 : (t < 32 ? n1 : n2);
    return static_cast<unsigned>(w >> ((t & 15) * 4)) & 0xFu;
  }
  __device__ __forceinline__ unsigned code_at_env(int e) const { return code_at_pos(e + (e >= kCentrePos)); }
};

// one row (template parameter => every table value is an immediate)
template <int R>
__device__ __forceinline__ void gather_site_row(const uint8_t *__restrict__ occ, int64_t base, int zp, int sy, int sx, uint32_t solvent4,
                                                uint64_t &n0, uint64_t &n1, uint64_t &n2, uint32_t &sol_lo, uint32_t &sol_hi) {
  constexpr int kind = site_row_kind(R), cnt = site_row_count(R), pos = site_row_pos(R), dx = site_row_dx(R), dy = site_row_dy(R);
  const int64_t idx = base + dx * sx + dy * sy + (kind == 1 ? 0 : (kind == 0 ? -1 : zp - 1));
  const uintptr_t addr = reinterpret_cast<uintptr_t>(occ + idx);
  const uint64_t *p = reinterpret_cast<const uint64_t *>(addr & ~static_cast<uintptr_t>(7));
  const unsigned b = static_cast<unsigned>(addr & 7u), sh = b * 8u;
  uint32_t w = static_cast<uint32_t>(*p >> sh);
  if (cnt > 1 && b + cnt > 8) w |= static_cast<uint32_t>(p[1] << (64u - sh));   // the row's bytes straddle the aligned word
  uint32_t v = w & 0xFu;                                                        // codes as nibbles
  if (cnt > 1) v |= (w >> 4) & 0xF0u;
  if (cnt > 2) v |= (w >> 8) & 0xF00u;
  constexpr int word = pos >> 4, shift = (pos & 15) * 4;
  const uint64_t placed = static_cast<uint64_t>(v) << shift;
  if (word == 0) n0 |= placed; else if (word == 1) n1 |= placed; else n2 |= placed;
  if ((pos & 15) + cnt > 16) {
    const uint64_t spill = static_cast<uint64_t>(v) >> (64 - shift);
    if (word == 0) n1 |= spill; else n2 |= spill;
  }
  const uint32_t x = w ^ solvent4;                                              // solute bits
#pragma unroll
  for (int k = 0; k < cnt; ++k) {
    const int t = pos + k;
    const bool solute = (x & (0xFFu << (8 * k))) != 0;
    if (t < 32) sol_lo |= solute ? (1u << t) : 0u;
    else sol_hi |= solute ? (1u << (t - 32)) : 0u;
  }
}
template <int... R>
__device__ __forceinline__ void gather_site_rows_impl(const uint8_t *__restrict__ occ, int64_t base, int zp, int sy, int sx, uint32_t solvent4,
                                                      uint64_t &n0, uint64_t &n1, uint64_t &n2, uint32_t &sol_lo, uint32_t &sol_hi) {
  (gather_site_row<R>(occ, base, zp, sy, sx, solvent4, n0, n1, n2, sol_lo, sol_hi), ...);
}

__device__ __forceinline__ SiteEnvRegs gather_site_rows(const uint8_t *__restrict__ occ, int64_t base, int zp, int sy, int sx,
                                                       unsigned solvent) {
  SiteEnvRegs env;
  env.n0 = env.n1 = env.n2 = 0;
  uint32_t sol_lo = 0, sol_hi = 0;                          // over list positions t (centre included), split at 32
  gather_site_rows_impl<0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20>(occ, base, zp, sy, sx, solvent * 0x01010101u, env.n0,
                                                                                                  env.n1, env.n2, sol_lo, sol_hi);
  const uint64_t sol_t = (static_cast<uint64_t>(sol_hi) << 32) | sol_lo;
  env.sol = (sol_t & ((1ULL << kCentrePos) - 1ULL)) | ((sol_t >> (kCentrePos + 1)) << kCentrePos);
  return env;
}

// H(x_new, env) - H(x_old, env) from registers: A / mask / base tables in shared memory, B through the read-only path
struct SiteWalkTables {
  const double *s_A;            // [m][42][m]
  const uint64_t *s_mask_hi;    // [42]
  const uint16_t *s_base;       // [42]
  const double *B, *C;          // global
  int m, n_pairs;
};
__device__ __forceinline__ double site_energy_change_regs(const SiteWalkTables &T, int x_old, int x_new, const SiteEnvRegs &env) {
  const int m = T.m;
  const int a_stride = kSiteEnvN * m;
  const size_t b_stride = static_cast<size_t>(T.n_pairs) * m * m;
  const double *A_new = T.s_A + x_new * a_stride, *A_old = T.s_A + x_old * a_stride;
  const double *__restrict__ B_new = T.B + x_new * b_stride, *__restrict__ B_old = T.B + x_old * b_stride;
  double acc = __ldg(T.C + x_new) - __ldg(T.C + x_old);
  uint64_t sol = env.sol;
  while (sol) {
    const int t = __ffsll(static_cast<long long>(sol)) - 1;
    sol &= sol - 1;
    const int et = static_cast<int>(env.code_at_env(t));
    acc += A_new[t * m + et] - A_old[t * m + et];
    const uint64_t hi = T.s_mask_hi[t];
    uint64_t partners = hi & sol;
    if (partners) {
      const int pbase = T.s_base[t];
      do {
        const int u = __ffsll(static_cast<long long>(partners)) - 1;
        partners &= partners - 1;
        const int eu = static_cast<int>(env.code_at_env(u));
        const size_t p = (static_cast<size_t>(pbase + __popcll(hi & ((1ULL << u) - 1ULL))) * m + et) * m + eu;
        acc += __ldg(B_new + p) - __ldg(B_old + p);
      } while (partners);
    }
  }
  return acc;
}

constexpr int kSwapThreads = 128;
constexpr int kSwapMaxM = 8;     // species incl. vacancy (nibble codes < 16; the element set holds at most 7 species)
inline size_t swap_rows_smem_bytes(int m) { return static_cast<size_t>(m) * kSiteEnvN * m * 8 + kSiteEnvN * 8 + 2 * 43 * 4 + kSiteEnvN * 2 + 16; }

// Batched swap dE, one thread per pair, persistent blocks (grid-stride) so that the tables are staged once per block.
// Uncoupled pairs take the row-based register path; the rare coupled pairs (b within the third shell of a, ~42/N of
// random pairs) and bad ids take swap_energy_change() above.
template <int kMinBlocks>
__global__ void __launch_bounds__(kSwapThreads, kMinBlocks)
swap_de_rows_kernel(LatticeDesc lat, DevTables tab, const uint8_t *__restrict__ occ, int64_t walker_stride, int64_t n,
                    const int32_t *__restrict__ walker, const int64_t *__restrict__ site_a, const int64_t *__restrict__ site_b,
                    double *__restrict__ dE, int *__restrict__ error, int first_neighbours_only, const uint32_t *__restrict__ perm) {
  // dynamic shared memory sized for the actual species count (small footprint => the rest of the 256 KB stays L1)
  extern __shared__ __align__(16) unsigned char swap_smem[];
  const int m = tab.n_species + 1;
  double *s_A = reinterpret_cast<double *>(swap_smem);                                   // [m][42][m]
  uint64_t *s_mask_hi = reinterpret_cast<uint64_t *>(s_A + m * kSiteEnvN * m);           // [42]
  int32_t *s_delta = reinterpret_cast<int32_t *>(s_mask_hi + kSiteEnvN);                 // [2][43]
  uint16_t *s_base = reinterpret_cast<uint16_t *>(s_delta + 2 * 43);                     // [42]
  for (int q = threadIdx.x; q < m * kSiteEnvN * m; q += blockDim.x) s_A[q] = tab.site_A[q];
  for (int q = threadIdx.x; q < kSiteEnvN; q += blockDim.x) { s_mask_hi[q] = tab.site_mask_hi[q]; s_base[q] = tab.site_base[q]; }
  for (int q = threadIdx.x; q < 2 * 43; q += blockDim.x) s_delta[q] = tab.site_delta[q];
  __syncthreads();
  const SiteWalkTables T{s_A, s_mask_hi, s_base, tab.site_B, tab.site_C, m, tab.n_site_pairs};
  const unsigned solvent = static_cast<unsigned>(tab.solvent);
  const int sy = lat.nz, sx = lat.ny * lat.nz;
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  for (int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; q < n; q += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t e = perm ? static_cast<int64_t>(perm[q]) : q;      // grouped by the region of site a (grouping.h); results in caller order
    const int64_t a = site_a[e], b = site_b[e];
    if (a < 0 || a >= lat.num_sites || b < 0 || b >= lat.num_sites) {
      atomicOr(error, kErrBadSite);
      dE[e] = CUDART_NAN;
      continue;
    }
    const uint8_t *o = occ + (walker ? walker[e] : 0) * walker_stride;
    int xa, ya, za, xb, yb, zb;
    lat.coords_of_id(a, xa, ya, za);
    lat.coords_of_id(b, xb, yb, zb);
    const int64_t base_a = lat.padded_index(xa, ya, za), base_b = lat.padded_index(xb, yb, zb);
    const unsigned ea = o[base_a], eb = o[base_b];
    double de = 0.0;
    if (ea != eb) {
      int dx = xb - xa, dy = yb - ya, dz = zb - za;
      dx = dx > px / 2 ? dx - px : (dx < -px / 2 ? dx + px : dx);
      dy = dy > py / 2 ? dy - py : (dy < -py / 2 ? dy + py : dy);
      dz = dz > pz / 2 ? dz - pz : (dz < -pz / 2 ? dz + pz : dz);
      const int r2 = dx * dx + dy * dy + dz * dz;
      if (r2 <= 6 || first_neighbours_only) {          // coupled (or the first-neighbour-only predictor): general path
        int err = 0;
        de = swap_energy_change(lat, tab, o, s_delta, xa, ya, za, xb, yb, zb, &err, first_neighbours_only != 0);
        if (err) atomicOr(error, err);
      } else {
        const SiteEnvRegs env_a = gather_site_rows(o, base_a, za & 1, sy, sx, solvent);
        de = site_energy_change_regs(T, static_cast<int>(ea), static_cast<int>(eb), env_a);
        const SiteEnvRegs env_b = gather_site_rows(o, base_b, zb & 1, sy, sx, solvent);
        de += site_energy_change_regs(T, static_cast<int>(eb), static_cast<int>(ea), env_b);
        if (de != de) atomicOr(error, kErrExtraVacancy);
      }
    }
    dE[e] = de;
  }
}

__global__ void __launch_bounds__(kSwapThreads)
swap_de_kernel(LatticeDesc lat, DevTables tab, const uint8_t *__restrict__ occ, int64_t walker_stride, int64_t n,
               const int32_t *__restrict__ walker, const int64_t *__restrict__ site_a, const int64_t *__restrict__ site_b,
               double *__restrict__ dE, int *__restrict__ error, int first_neighbours_only) {
  __shared__ int32_t s_delta[2 * 43];
  for (int q = threadIdx.x; q < 2 * 43; q += blockDim.x) s_delta[q] = tab.site_delta[q];
  __syncthreads();
  const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (e >= n) return;
  const int64_t a = site_a[e], b = site_b[e];
  if (a < 0 || a >= lat.num_sites || b < 0 || b >= lat.num_sites) {
    atomicOr(error, kErrBadSite);
    dE[e] = CUDART_NAN;
    return;
  }
  int xa, ya, za, xb, yb, zb;
  lat.coords_of_id(a, xa, ya, za);
  lat.coords_of_id(b, xb, yb, zb);
  int err = 0;
  const double de = swap_energy_change(lat, tab, occ + (walker ? walker[e] : 0) * walker_stride, s_delta, xa, ya, za, xb, yb, zb, &err,
                                       first_neighbours_only != 0);
  if (err) atomicOr(error, err);
  dE[e] = de;
}

__global__ void __launch_bounds__(kSwapThreads)
site_de_kernel(LatticeDesc lat, DevTables tab, const uint8_t *__restrict__ occ, int64_t walker_stride, int64_t n,
               const int32_t *__restrict__ walker, const int64_t *__restrict__ site, const uint8_t *__restrict__ new_code,
               double *__restrict__ dE, int *__restrict__ error) {
  __shared__ int32_t s_delta[2 * 43];
  for (int q = threadIdx.x; q < 2 * 43; q += blockDim.x) s_delta[q] = tab.site_delta[q];
  __syncthreads();
  const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (e >= n) return;
  const int64_t s = site[e];
  if (s < 0 || s >= lat.num_sites || new_code[e] > tab.n_species) {
    atomicOr(error, kErrBadSite);
    dE[e] = CUDART_NAN;
    return;
  }
  int x, y, z;
  lat.coords_of_id(s, x, y, z);
  const uint8_t *o = occ + (walker ? walker[e] : 0) * walker_stride;
  unsigned centre = 0;
  const int64_t base = lat.padded_index(x, y, z);
  const int32_t *row = s_delta + (z & 1) * 43;
  const uint64_t sol = gather_site_env(o, base, row, static_cast<unsigned>(tab.solvent), &centre, -1, 0);
  double de = 0.0;
  if (centre != new_code[e]) {
    de = site_energy_change(tab, static_cast<int>(centre), static_cast<int>(new_code[e]), sol, o, base, row, -1, 0);
    if (de != de) atomicOr(error, kErrExtraVacancy);
  }
  dE[e] = de;
}

// ----------------------------------------------------------------------------------------------- total energy
// EnergyPredictor::GetEncode / GetEnergy (pred/src/EnergyPredictor.cpp:40-96,173-177): ordered-tuple counting,
// one thread per site; per-block partial sums are reduced in a fixed order (deterministic).
constexpr int kEnergyThreads = 128;

__global__ void __launch_bounds__(kEnergyThreads)
energy_kernel(LatticeDesc lat, DevTables tab, const uint8_t *__restrict__ occ, double *__restrict__ block_sums,
              unsigned long long *__restrict__ counts, const uint8_t *__restrict__ member) {
  // member != nullptr: EnergyPredictor::GetEncodeOfCluster (pred/src/EnergyPredictor.cpp:97-172) -- only clusters whose
  // sites all lie in the marked set (the listed atoms and their 1-3NN shells) are counted
  extern __shared__ unsigned char smem_raw[];
  const int m = tab.n_species + 1;
  double *s_single = reinterpret_cast<double *>(smem_raw);
  double *s_pair = s_single + m;
  double *s_trip = s_pair + 3 * m * m;
  int32_t *s_delta = reinterpret_cast<int32_t *>(s_trip + 4 * m * m * m);
  unsigned int *s_counts = reinterpret_cast<unsigned int *>(s_delta + 2 * 43);
  __shared__ double s_red[kEnergyThreads / 32];
  for (int q = threadIdx.x; q < m; q += blockDim.x) s_single[q] = tab.e_single[q];
  for (int q = threadIdx.x; q < 3 * m * m; q += blockDim.x) s_pair[q] = tab.e_pair[q];
  for (int q = threadIdx.x; q < 4 * m * m * m; q += blockDim.x) s_trip[q] = tab.e_triplet[q];
  for (int q = threadIdx.x; q < 2 * 43; q += blockDim.x) s_delta[q] = tab.site_delta[q];
  if (counts)
    for (int q = threadIdx.x; q < tab.n_types; q += blockDim.x) s_counts[q] = 0;
  __syncthreads();
  double acc = 0.0;
  const int64_t id = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (id < lat.num_sites) {
    int x, y, z;
    lat.coords_of_id(id, x, y, z);
    const int64_t base = lat.padded_index(x, y, z);
    const int32_t *drow = s_delta + (z & 1) * 43;
    unsigned char code[43];
    unsigned long long in_set = ~0ULL;                       // bit t: list position t belongs to the cluster's site set
#pragma unroll
    for (int t = 0; t < 43; ++t) code[t] = occ[base + drow[t]];
    if (member) {
      in_set = 0ULL;
#pragma unroll
      for (int t = 0; t < 43; ++t) in_set |= static_cast<unsigned long long>(member[base + drow[t]] != 0) << t;
    }
    if ((in_set >> kCentrePos) & 1ULL) {
      const int c1 = code[kCentrePos];
      acc += s_single[c1];
      if (counts) atomicAdd(&s_counts[tab.type_lut[c1 * m * m]], 1u);
      for (int q = 0; q < 42; ++q) {
        const int pos = tab.e_shell_pos[2 * q], shell = tab.e_shell_pos[2 * q + 1];
        if (!((in_set >> pos) & 1ULL)) continue;
        const int c2 = code[pos];
        acc += s_pair[((shell - 1) * m + c1) * m + c2];
        if (counts) atomicAdd(&s_counts[tab.type_lut[((shell * m + c1) * m + c2) * m]], 1u);
      }
      for (int q = 0; q < tab.n_e_walk; ++q) {
        const int p2 = tab.e_walk[4 * q], p3 = tab.e_walk[4 * q + 1], label = tab.e_walk[4 * q + 2];
        if (!((in_set >> p2) & (in_set >> p3) & 1ULL)) continue;
        const int c2 = code[p2], c3 = code[p3];
        acc += s_trip[(((label - 4) * m + c1) * m + c2) * m + c3];
        if (counts) atomicAdd(&s_counts[tab.type_lut[((label * m + c1) * m + c2) * m + c3]], 1u);
      }
    }
  }
  // block reduction in a fixed order
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int q = 0; q < kEnergyThreads / 32; ++q) s += s_red[q];
    block_sums[blockIdx.x] = s;
  }
  if (counts) {
    for (int q = threadIdx.x; q < tab.n_types; q += blockDim.x)
      if (s_counts[q]) atomicAdd(&counts[q], static_cast<unsigned long long>(s_counts[q]));
  }
}

// Config::GetElementAtLatticeId for a list of sites (element enum codes)
__global__ void gather_sites_kernel(LatticeDesc lat, const uint8_t *__restrict__ padded, const int64_t *__restrict__ sites, int64_t n,
                                    const uint8_t *__restrict__ enum_of_code, uint8_t *__restrict__ out, int *error) {
  const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (q >= n) return;
  const int64_t id = sites[q];
  if (id < 0 || id >= lat.num_sites) { atomicOr(error, kErrBadSite); out[q] = 0; return; }
  out[q] = enum_of_code[padded[lat.padded_index_of_id(id)]];
}
// Config::GetVacancyLatticeId generalised: the lowest lattice id holding compact code `code` (-1 if none), and the count
__global__ void find_element_kernel(LatticeDesc lat, const uint8_t *__restrict__ padded, int code, unsigned long long *first, unsigned long long *count) {
  const int64_t id = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const bool hit = id < lat.num_sites && padded[lat.padded_index_of_id(id)] == code;
  const unsigned bal = __ballot_sync(0xffffffffu, hit);
  if (bal && (threadIdx.x & 31) == 0) {
    atomicMin(first, static_cast<unsigned long long>(id + __ffs(static_cast<int>(bal)) - 1));
    atomicAdd(count, static_cast<unsigned long long>(__popc(bal)));
  }
}

// site set of GetEncodeOfCluster: every listed site and its 42 first- to third-neighbours, marked in a padded-layout byte
// array (halo images included, like the occupancy)
__global__ void mark_cluster_sites_kernel(LatticeDesc lat, DevTables tab, const int64_t *__restrict__ sites, int64_t n, uint8_t *member, int *error) {
  const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (q >= n * 43) return;
  const int64_t id = sites[q / 43];
  if (id < 0 || id >= lat.num_sites) { atomicOr(error, kErrBadSite); return; }
  int x, y, z;
  lat.coords_of_id(id, x, y, z);
  const int8_t *o = tab.site_off + 4 * (q % 43);
  store_site(lat, member, wrap_coord(x + o[0], 2 * lat.fx), wrap_coord(y + o[1], 2 * lat.fy), wrap_coord(z + o[2], 2 * lat.fz), 1);
}

// ----------------------------------------------------------------------------------------------- debug taps
// One block per call; recomputes the reference's integer artefacts for one jump pair straight from its
// definitions (ordered lists with the min-id frame rule, cluster-type histograms, one-hot numerators).
__device__ __forceinline__ int64_t wrapped_id(const LatticeDesc &lat, int x, int y, int z) {
  return lat.id_of_coords(wrap_coord(wrap_coord(x, 2 * lat.fx), 2 * lat.fx), wrap_coord(wrap_coord(y, 2 * lat.fy), 2 * lat.fy),
                          wrap_coord(wrap_coord(z, 2 * lat.fz), 2 * lat.fz));
}

// frame flag of jump (x,y,z) -> direction k: 0 if +frame_p[k] is the perpendicular first neighbour with the smaller
// lattice id (the one Config::GetLatticePairRotationMatrix meets first, cfg/src/Config.cpp:253-260), else 1
__device__ __forceinline__ int frame_flag(const LatticeDesc &lat, const DevTables &tab, int x, int y, int z, int k) {
  const int8_t *p = tab.frame_p + 4 * k;
  const int64_t plus = wrapped_id(lat, x + p[0], y + p[1], z + p[2]);
  const int64_t minus = wrapped_id(lat, x - p[0], y - p[1], z - p[2]);
  return plus < minus ? 0 : 1;
}

__global__ void debug_pair_kernel(LatticeDesc lat, DevTables tab, const uint8_t *__restrict__ occ, int64_t site_i, int64_t site_j,
                                  int64_t *__restrict__ lists /*60+58+58+58*/, int32_t *__restrict__ counts /*2*n_types*/,
                                  int32_t *__restrict__ enc /*len_mmm + 2*len_mm2*/, int *__restrict__ error) {
  __shared__ int64_t s_ids[60];
  __shared__ uint8_t s_code[60];
  __shared__ int s_k, s_flag_f, s_flag_b;
  int xi, yi, zi, xj, yj, zj;
  lat.coords_of_id(site_i, xi, yi, zi);
  lat.coords_of_id(site_j, xj, yj, zj);
  if (threadIdx.x == 0) {
    const int k = direction_of(lat, tab, xi, yi, zi, xj, yj, zj);
    s_k = k;
    if (k >= 0) {
      s_flag_f = frame_flag(lat, tab, xi, yi, zi, k);
      // reversed pair: direction -d has the same perpendicular pair; its flag is evaluated at site j.  frame_p of the
      // reversed direction may be stored with either sign, so compare actual vectors: backward variant 1 <=> the
      // backward y axis equals the forward y axis.
      int kb = -1;
      for (int q = 0; q < 12; ++q)
        if (tab.nn1[4 * q] == -tab.nn1[4 * k] && tab.nn1[4 * q + 1] == -tab.nn1[4 * k + 1] && tab.nn1[4 * q + 2] == -tab.nn1[4 * k + 2]) kb = q;
      const int fb = frame_flag(lat, tab, xj, yj, zj, kb);
      const int sf = s_flag_f ? -1 : 1, sb = fb ? -1 : 1;
      const bool same = (sf * tab.frame_p[4 * k] == sb * tab.frame_p[4 * kb]) && (sf * tab.frame_p[4 * k + 1] == sb * tab.frame_p[4 * kb + 1]) &&
                        (sf * tab.frame_p[4 * k + 2] == sb * tab.frame_p[4 * kb + 2]);
      s_flag_b = same ? 1 : 0;
    }
  }
  __syncthreads();
  if (s_k < 0) {
    if (threadIdx.x == 0) atomicOr(error, kErrNotNeighbour);
    return;
  }
  const int k = s_k;
  for (int t = threadIdx.x; t < 60; t += blockDim.x) {
    const int8_t *o = tab.pair_off + ((k * 2 + s_flag_f) * 60 + t) * 4;
    s_ids[t] = wrapped_id(lat, xi + o[0], yi + o[1], zi + o[2]);
    // occupancy is read through the frame-independent offset table used by the production kernels
    s_code[t] = 0;
  }
  __syncthreads();
  // species by lattice id (independent of the fast kernels' offset rows)
  for (int t = threadIdx.x; t < 60; t += blockDim.x) s_code[t] = occ[lat.padded_index_of_id(s_ids[t])];
  __syncthreads();
  const int n = tab.n_species, m = n + 1;
  if (lists) {
    for (int t = threadIdx.x; t < 60; t += blockDim.x) lists[t] = s_ids[t];
    for (int a = threadIdx.x; a < 58; a += blockDim.x) {
      lists[60 + a] = s_ids[tab.state_pos_of_env[tab.env_of_list[a]]];
      lists[118 + a] = s_ids[tab.state_pos_of_env[tab.env_of_list[58 + a]]];
      lists[176 + a] = s_ids[tab.state_pos_of_env[tab.env_of_list[(2 + s_flag_b) * 58 + a]]];
    }
  }
  if (counts) {
    for (int q = threadIdx.x; q < 2 * tab.n_types; q += blockDim.x) counts[q] = 0;
    __syncthreads();
    const int mig = s_code[kSecondPos];
    for (int c = threadIdx.x; c < tab.n_state_pair; c += blockDim.x) {
      const int8_t *cl = tab.map_state_pair + 4 * c;
      int start[3] = {0, 0, 0}, end[3] = {0, 0, 0};
      for (int q = 0; q < 3; ++q) {
        const int pos = cl[1 + q];
        if (pos < 0) break;
        start[q] = s_code[pos];
        end[q] = pos == kFirstPos ? mig : (pos == kSecondPos ? n : s_code[pos]);
      }
      const int ts = tab.type_lut[((cl[0] * m + start[0]) * m + start[1]) * m + start[2]];
      const int te = tab.type_lut[((cl[0] * m + end[0]) * m + end[1]) * m + end[2]];
      if (ts < 0 || te < 0) { atomicOr(error, kErrExtraVacancy); continue; }
      atomicAdd(&counts[ts], 1);
      atomicAdd(&counts[tab.n_types + te], 1);
    }
  }
  if (enc) {
    const int total = tab.len_mmm + 2 * tab.len_mm2;
    for (int q = threadIdx.x; q < total; q += blockDim.x) enc[q] = 0;
    __syncthreads();
    for (int which = 0; which < 3; ++which) {
      const int16_t *map = which == 0 ? tab.map_mmm : tab.map_mm2;
      const int8_t *env_of = tab.env_of_list + (which == 0 ? 0 : (which == 1 ? 58 : (2 + s_flag_b) * 58));
      int32_t *out = enc + (which == 0 ? 0 : (which == 1 ? tab.len_mmm : tab.len_mmm + tab.len_mm2));
      for (int c = threadIdx.x; c < tab.n_avg_clusters; c += blockDim.x) {
        const int16_t *cl = map + 4 * c;
        const int ea = s_code[tab.state_pos_of_env[env_of[cl[1]]]];
        if (ea >= n) { atomicOr(error, kErrExtraVacancy); continue; }
        int slot;
        if (cl[2] < 0) {
          slot = ea;
        } else {
          const int eb = s_code[tab.state_pos_of_env[env_of[cl[2]]]];
          if (eb >= n) { atomicOr(error, kErrExtraVacancy); continue; }
          if (cl[3] & 1) {
            const int lo = ea < eb ? ea : eb, hi = ea < eb ? eb : ea;
            slot = lo * n - lo * (lo - 1) / 2 + (hi - lo);
          } else {
            slot = ea * n + eb;
          }
        }
        atomicAdd(&out[cl[0] + slot], 1);
      }
    }
  }
}

__global__ void debug_site_kernel(LatticeDesc lat, DevTables tab, const uint8_t *__restrict__ occ, int64_t site, int new_code,
                                  int64_t *__restrict__ list43, int32_t *__restrict__ counts, int *__restrict__ error) {
  __shared__ int64_t s_ids[43];
  __shared__ uint8_t s_code[43];
  int x, y, z;
  lat.coords_of_id(site, x, y, z);
  for (int t = threadIdx.x; t < 43; t += blockDim.x) {
    const int8_t *o = tab.site_off + 4 * t;
    s_ids[t] = wrapped_id(lat, x + o[0], y + o[1], z + o[2]);
    s_code[t] = occ[lat.padded_index_of_id(s_ids[t])];
    if (list43) list43[t] = s_ids[t];
  }
  __syncthreads();
  if (!counts) return;
  const int m = tab.n_species + 1;
  for (int q = threadIdx.x; q < 2 * tab.n_types; q += blockDim.x) counts[q] = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < tab.n_state_site; c += blockDim.x) {
    const int8_t *cl = tab.map_state_site + 4 * c;
    int start[3] = {0, 0, 0}, end[3] = {0, 0, 0};
    for (int q = 0; q < 3; ++q) {
      const int pos = cl[1 + q];
      if (pos < 0) break;
      start[q] = s_code[pos];
      end[q] = pos == kCentrePos ? new_code : s_code[pos];
    }
    const int ts = tab.type_lut[((cl[0] * m + start[0]) * m + start[1]) * m + start[2]];
    const int te = tab.type_lut[((cl[0] * m + end[0]) * m + end[1]) * m + end[2]];
    if (ts < 0 || te < 0) { atomicOr(error, kErrExtraVacancy); continue; }
    atomicAdd(&counts[ts], 1);
    atomicAdd(&counts[tab.n_types + te], 1);
  }
}

}  // namespace lmc

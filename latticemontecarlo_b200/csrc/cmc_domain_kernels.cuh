// cmc_domain_kernels.cuh -- canonical MC / simulated annealing by SPATIAL DOMAIN DECOMPOSITION ("sublattice" driver).
//
// The batched drivers (cmc_kernels.cuh, cmc_grid_kernels.cuh) keep the reference's GLOBAL pair draw
// (mc/src/CanonicalMcAbstract.cpp:43-51) and pay for it with a claim round trip and two grid barriers around every batch of
// ~N/172 mutually non-interfering trials.  This driver changes the proposal instead (stated semantic change):
//
//   * the periodic lattice is cut into a grid of box-shaped DOMAINS of about `domain_edge` half lattice constants per axis;
//     the grid origin is moved by a random vector every SWEEP (Philox, same on every rank);
//   * a domain's ACTIVE CORE is the domain minus one plane of sites on every face.  Sites of different cores are at least
//     3 half-units apart in some coordinate, i.e. outside each other's 43-site neighbourhoods (max offset component 2), so
//     trials of different domains never interfere: the non-interference rule of CanonicalMcOmp.cpp:47-72 holds by
//     construction, with no claims, no atomics and no barrier;
//   * one LANE GROUP (8, 16 or 32 lanes) owns a domain for a sweep: it copies the domain plus a one-plane halo (the frozen
//     margin planes of its neighbours) into shared memory, runs `rounds` sequential Metropolis trials on pairs of sites drawn
//     uniformly from the core (redrawn while the two species are equal, as in the reference), all reads and writes in shared
//     memory, and writes the domain back;
//   * ONE grid barrier per sweep (thousands of trials per domain-sweep instead of ~7 kept trials per SM and batch), after
//     which the sweep totals (fixed-point energy, trial / accept counts) update the energy, the step counter and the
//     SimulatedAnnealing schedule (SimulatedAnnealing.cpp:99-139, applied per sweep).
//
// Every sweep is a composition of Metropolis kernels with symmetric proposals on the exact energy model (dE is evaluated
// on the full 43-site neighbourhoods, frozen halo included), so the canonical distribution is stationary; the random shifts
// make the chain ergodic.  Occupancy is double buffered (read buffer s & 1, write the other) so that halo reads and domain
// write-backs of one sweep never race.
//
// Several GPUs (one process per GPU): rank r owns a slab of the domain grid along x.  At write-back every domain row is
// stored into the local buffer AND, through NVLink peer mappings, into the buffer of every rank whose slab or one-plane
// halo will contain that x plane in the NEXT sweep (the next shift is known: counter-based RNG) -- a per-sweep halo /
// migration exchange written from inside the persistent kernel, no NCCL call and no host round trip.  The sweep totals
// travel as flag-in-data lines (cmc_grid_kernels.cuh), which double as the inter-GPU barrier.  All decisions are keyed by
// (seed, sweep, domain), so the trajectory is identical for every world size.
#pragma once
#include <cuda_runtime.h>

#include "cmc_domain.h"
#include "device_common.h"

namespace lmc {

// ---- Philox4x32-10 with a full 128-bit counter
__device__ __forceinline__ void philox4x32_10_c4(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c[4] = {c0, c1, c2, c3};
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// grid origin of a sweep: uniform over the lattice periods (any parity: tiles carry their own parity)
__device__ __forceinline__ void domain_shift(uint64_t seed, unsigned long long sweep, int px, int py, int pz, int &sx, int &sy, int &sz) {
  uint32_t r[4];
  philox4x32_10_c4(0xD0A11CE5u, 0u, static_cast<uint32_t>(sweep), static_cast<uint32_t>(sweep >> 32), static_cast<uint32_t>(seed),
                   static_cast<uint32_t>(seed >> 32) ^ 0x5F3759DFu, r);
  sx = static_cast<int>(__umulhi(r[0], static_cast<uint32_t>(px)));
  sy = static_cast<int>(__umulhi(r[1], static_cast<uint32_t>(py)));
  sz = static_cast<int>(__umulhi(r[2], static_cast<uint32_t>(pz)));
}

__device__ __forceinline__ bool dom_grid_barrier(const CmcDomainParams &dp, unsigned long long &target, unsigned n_cta) {
  __syncthreads();
  target += n_cta;
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(dp.barrier_counter) : "memory");
    const long long t0 = clock64();
    int ok = 1;
    for (;;) {
      unsigned long long seen;
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(dp.barrier_counter) : "memory");
      if (seen >= target) break;
      if (*reinterpret_cast<volatile int *>(dp.abort_flag)) { ok = 0; break; }
      if (clock64() - t0 > dp.spin_limit) { *reinterpret_cast<volatile int *>(dp.abort_flag) = 1; ok = 0; break; }
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

// SimulatedAnnealing::UpdateTemperature (mc/src/SimulatedAnnealing.cpp:99-139) applied once per sweep
__device__ __forceinline__ void sa_update_sweep(SaSchedule &sa, unsigned long long n_kept, unsigned long long n_acc, double energy,
                                                unsigned long long steps) {
  const double cool = exp(-3.0 / static_cast<double>(sa.maximum_steps > 0 ? sa.maximum_steps : 1ULL));
  const unsigned long long wt = static_cast<unsigned long long>(sa.window_trials) + n_kept, wa = static_cast<unsigned long long>(sa.window_accepts) + n_acc;
  sa.window_trials = wt > 0xFFFFFFFFULL ? 0xFFFFFFFFu : static_cast<unsigned int>(wt);
  sa.window_accepts = wa > 0xFFFFFFFFULL ? 0xFFFFFFFFu : static_cast<unsigned int>(wa);
  if (n_acc > 0 && energy < sa.recent_best_energy - kSaEpsilon) { sa.recent_best_energy = energy; sa.last_improvement_step = steps; }
  double acc_est = 1.0;
  if (sa.window_trials >= sa.window_size) {
    acc_est = static_cast<double>(sa.window_accepts) / static_cast<double>(sa.window_trials);
    if (acc_est > 0.50) sa.temperature *= 0.99;
    sa.window_trials = 0; sa.window_accepts = 0;
  } else if (sa.window_trials > 0u) {
    acc_est = static_cast<double>(sa.window_accepts) / static_cast<double>(sa.window_trials);
  }
  // a sweep may span whole windows: the acceptance of the window that just closed stands in for the running estimate
  if (sa.reheats_done < 5u && (steps - sa.last_improvement_step >= sa.reheat_trigger_steps) &&
      (steps - sa.last_reheat_step >= sa.reheat_cooldown_steps) && acc_est < 0.05) {
    sa.temperature *= 1.10; sa.last_improvement_step = steps; sa.last_reheat_step = steps;
    sa.recent_best_energy = energy; ++sa.reheats_done;
  }
  sa.temperature *= pow(cool, static_cast<double>(n_kept));
}

// state hand-over between the engine's per-replica arrays and the sweep-parity state buffers
__global__ void dom_state_init_kernel(int n_walkers, CmcState st, const double *__restrict__ temperatures, DomState *dst) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_walkers) return;
  DomState s;
  s.energy = st.energy[w]; s.steps = st.steps[w]; s.accepted = st.accepted[w]; s.sa = st.sa[w];
  s.temperature = s.sa.enabled ? s.sa.temperature : temperatures[w];
  dst[w] = s;
}

// periodic halo images of the padded layout from their canonical cells (the sweeps only maintain canonical cells)
__global__ void dom_refresh_halo_kernel(LatticeDesc lat, uint8_t *__restrict__ padded) {
  const int64_t cell = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (cell >= lat.padded_size) return;
  padded += blockIdx.y * lat.padded_size;
  const int zi = static_cast<int>(cell % lat.nz);
  const int64_t r = cell / lat.nz;
  const int yp = static_cast<int>(r % lat.ny), xp = static_cast<int>(r / lat.ny);
  const int xr = xp - kHalo, yr = yp - kHalo;
  const int zr = 2 * zi + ((xr + yr) & 1) - kHaloZ;           // the site of this cell (X + Y + Z even)
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  if (xr >= 0 && xr < px && yr >= 0 && yr < py && zr >= 0 && zr < pz) return;      // canonical cell
  const int X = wrap_coord(xr, px), Y = wrap_coord(yr, py), Z = wrap_coord(wrap_coord(zr, pz), pz);
  padded[cell] = padded[lat.padded_index(X, Y, Z)];
}

// chain state through L2 (written by CTA 0 one sweep earlier; plain loads could hit a stale L1 line)
__device__ __forceinline__ DomState dom_load_state(const DomState *src) {
  static_assert(sizeof(DomState) % 16 == 0, "DomState is read as 16-byte words");
  DomState s;
  const ulonglong2 *p = reinterpret_cast<const ulonglong2 *>(src);
  ulonglong2 *q = reinterpret_cast<ulonglong2 *>(&s);
#pragma unroll
  for (int i = 0; i < static_cast<int>(sizeof(DomState) / 16); ++i) q[i] = __ldcg(p + i);
  return s;
}

// Inter-GPU barrier + all-reduce of four 64-bit integers.  CTA 0 pushes this rank's values to every rank as a 64-byte line
// whose flag word is the sequence number (release store at system scope, after a system fence: the rank's earlier peer
// stores, ordered before by the local grid barrier, are performed first); thread 0 of every CTA then waits for the lines
// of all ranks in its own (local) buffer.  Lines are double buffered by sequence parity.
__device__ __forceinline__ bool dom_intergpu_sum(const CmcDomainParams &dp, unsigned long long seq, const unsigned long long *mine,
                                                 unsigned long long *s_sum, int *s_ok) {
  const int parity = static_cast<int>(seq & 1ULL);
  if (blockIdx.x == 0 && static_cast<int>(threadIdx.x) < dp.world) {
    const ulonglong2 v0 = __ldcg(reinterpret_cast<const ulonglong2 *>(mine)), v1 = __ldcg(reinterpret_cast<const ulonglong2 *>(mine) + 1);
    DomLine *dst = dp.peer_lines[threadIdx.x] + parity * kDomWorldMax + dp.rank;
    dst->v[0] = v0.x; dst->v[1] = v0.y; dst->v[2] = v1.x; dst->v[3] = v1.y;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&dst->flag), "l"(seq) : "memory");
  }
  if (threadIdx.x == 0) {
    unsigned long long sum[4] = {0ULL, 0ULL, 0ULL, 0ULL};
    int ok = 1;
    const long long t0 = clock64();
    for (int p = 0; p < dp.world && ok; ++p) {
      const DomLine *src = dp.lines + parity * kDomWorldMax + p;
      for (;;) {
        unsigned long long f;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(&src->flag) : "memory");
        if (f == seq) break;
        if (*reinterpret_cast<volatile int *>(dp.abort_flag)) { ok = 0; break; }
        if (clock64() - t0 > dp.spin_limit) { *reinterpret_cast<volatile int *>(dp.abort_flag) = 1; ok = 0; break; }
      }
      for (int q = 0; q < 4; ++q) sum[q] += *reinterpret_cast<const volatile unsigned long long *>(&src->v[q]);
    }
    for (int q = 0; q < 4; ++q) s_sum[q] = sum[q];
    *s_ok = ok;
  }
  __syncthreads();
  return *s_ok != 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// The sweep kernel.  L lanes per trial side, G = 2 L lanes per trial, S trials of one domain in flight (a TEAM of
// T = S G <= 32 lanes owns a domain).
//
// Speculation (S > 1).  A domain's rounds form one sequential Markov chain, so a lattice with few domains (40^3: 1000) runs
// at the latency of ONE dependent round per domain.  But only ~10 % of the trials are accepted, and a rejected trial leaves
// the state untouched: the S groups of a team evaluate rounds r0 .. r0 + S - 1 at once, all on the current state, and the
// team commits them up to and including the first accepted one; later rounds of the batch are discarded and drawn again
// (their random numbers are a function of the round index, so the redo sees the same draws).  This is exactly the
// sequential chain -- the trajectory does not depend on S.
//
// kTab == 1: DIFFERENCE tables in shared memory.  Both sites of a swap change between the same two species x_lo < x_hi, one
// in each direction, so the walk needs D = T[x_hi] - T[x_lo] only (one lookup and one add per term instead of two; the
// direction is a sign).  In the delta form solvent sites never enter a walk, so D is stored for non-solvent environment
// species only: m (m - 1) / 2 species pairs x (1 + 42 (m - 1) + 204 (m - 1)^2) doubles -- 94 KB for a ternary alloy with
// vacancy, less than the pair table itself.  The sums are the same terms in the same order (negation is exact), so both
// table forms give bit-identical dE.  kTab == 0: A staged, B through the read-only path (more species than fit).
//
// Proposal.  The reference draws two uniform sites until their species differ (CanonicalMcAbstract.cpp:43-51), i.e. every
// unordered unlike pair is equally likely.  In a dilute alloy 92 % of such draws are solvent-solvent, so the same
// distribution is sampled from the other end: every unlike pair holds at least one NON-SOLVENT site, hence
//     a from the domain's list of non-solvent core sites (built at tile load, updated on accepted swaps),
//     b uniform from all core sites, redrawn while species(a) == species(b),
//     and a pair of two (different) non-solvent species -- which can be drawn from either end -- kept with probability 1/2
// gives every unordered unlike pair of the core the same probability 1 / (S_list * ncore) per draw.  One Philox call per
// draw, keyed by (seed, sweep, domain, round, try): the random stream does not depend on the launch shape.
constexpr int kDomTries = 4;
#ifdef LMC_DOM_PROFILE
__device__ unsigned long long g_dom_hist[6][64];
#endif
template <int L, int kTab, int kMaxThreads, int S>
__global__ void __launch_bounds__(kMaxThreads, 1)
cmc_domain_kernel(LatticeDesc lat, DevTables tab, CmcDomainParams dp, CmcState st, uint64_t seed, unsigned long long target_steps) {
  constexpr int G = 2 * L, T = S * G;
  static_assert(T <= 32 && (32 % T) == 0, "a team is a power-of-two slice of a warp");
  constexpr unsigned kTeamBits = T >= 32 ? 0xFFFFFFFFu : ((1u << T) - 1u);
  constexpr unsigned kSideBits = (1u << L) - 1u;
  constexpr unsigned kLeaders = G >= 32 ? 1u : 0xFFFFFFFFu / ((1u << (G % 32)) - 1u);     // lane 0 of every group of a warp
  const int n_cta = static_cast<int>(gridDim.x), cta = static_cast<int>(blockIdx.x);
  const int tid = threadIdx.x, B = blockDim.x, lane = tid & 31;
  const int tl = lane % T, tb = lane - tl;                 // lane within the team, first lane of the team
  const int grp = tl / G, gl = tl % G;                     // speculative slot of this lane's group, lane within the group
  const int side = gl / L, sub = gl % L;
  const int side_shift = tb + grp * G + side * L;
  const int team_in_cta = tid / T;
  const int nw = dp.n_walkers;
  const int m = tab.n_species + 1, mm = m * m;
  const unsigned solvent = static_cast<unsigned>(tab.solvent), vac = static_cast<unsigned>(tab.n_species);

  __shared__ int s_shift[3];
  __shared__ unsigned long long s_tot[4];
  __shared__ int s_done, s_fail, s_ok2;
  __shared__ unsigned long long s_sum[4];
  // dynamic shared memory, fixed-size tables first (constant offsets): [mask 42 x 8] [pidx 42 x 42] [tdelta 2 x 44 x 2] [sp: 64]
  // then kTab == 0: [C: m] [A: m*42*m]   kTab == 1: [dC: n_sp] [DA: n_sp*42*ns] [DB: n_sp*204*ns*ns]
  // then [temperature: nw] [team tiles: tile | species rows S x 2 x 48 | solute list]
  extern __shared__ double s_dyn[];
  uint64_t *s_mask = reinterpret_cast<uint64_t *>(s_dyn);
  uint8_t *s_pidx = reinterpret_cast<uint8_t *>(s_mask + kSiteEnvN);
  int16_t *s_tdelta = reinterpret_cast<int16_t *>(s_pidx + kDomPidxBytes);
  uint8_t *s_sp = reinterpret_cast<uint8_t *>(s_tdelta + 2 * 44);   // [x_old * m + x_new] -> species-pair index | 0x80 if x_old > x_new
  double *s_C = reinterpret_cast<double *>(s_sp + 64);
  const int ns = m - 1, ns2 = ns * ns, n_sp = m * (m - 1) / 2;
  const int a_len = kTab ? n_sp * kSiteEnvN * ns : m * kSiteEnvN * m, b_len = kTab ? n_sp * tab.n_site_pairs * ns2 : 0;
  const int c_len = kTab ? n_sp : m;
  double *s_A = s_C + c_len;
  double *s_B = s_A + a_len;
  double *s_temp = s_B + b_len;
  uint8_t *s_tiles = reinterpret_cast<uint8_t *>(s_temp + nw);
  uint8_t *tile = s_tiles + static_cast<size_t>(team_in_cta) * dp.tile_bytes;
  uint8_t *row = tile + dp.tile_cells + (grp * 2 + side) * 48;   // species of this side's 42 environment sites
  uint16_t *list = reinterpret_cast<uint16_t *>(tile + dp.tile_cells + S * 96);   // core indices of the non-solvent core sites

  for (int q = tid; q < m * m; q += B) {
    const int xo = q / m, xn = q % m, lo = xo < xn ? xo : xn, hi = xo < xn ? xn : xo;
    s_sp[q] = static_cast<uint8_t>((lo * m - lo * (lo + 1) / 2 + (hi - lo - 1)) | (xo > xn ? 0x80 : 0));
  }
  if (kTab) {
    // species pair sp = (lo < hi); non-solvent species index e' = e - (e > solvent)
    for (int q = tid; q < n_sp + a_len + b_len; q += B) {
      int r = q, which = 0;
      if (r >= n_sp) { r -= n_sp; which = 1; }
      if (which == 1 && r >= a_len) { r -= a_len; which = 2; }
      const int per_sp = which == 0 ? 1 : (which == 1 ? kSiteEnvN * ns : tab.n_site_pairs * ns2);
      const int sp = r / per_sp, rest = r - sp * per_sp;
      int lo = 0, acc = 0;
      while (acc + (m - 1 - lo) <= sp) { acc += m - 1 - lo; ++lo; }
      const int hi = lo + 1 + (sp - acc);
      double v;
      if (which == 0) v = tab.site_C[hi] - tab.site_C[lo];
      else if (which == 1) {
        const int t = rest / ns, e1 = rest % ns, e = e1 + (e1 >= static_cast<int>(solvent));
        v = tab.site_A[(hi * kSiteEnvN + t) * m + e] - tab.site_A[(lo * kSiteEnvN + t) * m + e];
      } else {
        const int pr = rest / ns2, e12 = rest % ns2, e1 = e12 / ns, e2_ = e12 % ns;
        const int ea_ = e1 + (e1 >= static_cast<int>(solvent)), eb_ = e2_ + (e2_ >= static_cast<int>(solvent));
        const size_t b_stride = static_cast<size_t>(tab.n_site_pairs) * mm;
        v = tab.site_B[hi * b_stride + (pr * m + ea_) * m + eb_] - tab.site_B[lo * b_stride + (pr * m + ea_) * m + eb_];
      }
      s_C[q] = v;      // dC, DA, DB are contiguous
    }
  } else {
    for (int q = tid; q < m; q += B) s_C[q] = tab.site_C[q];
    for (int q = tid; q < a_len; q += B) s_A[q] = tab.site_A[q];
  }
  for (int q = tid; q < kSiteEnvN; q += B) s_mask[q] = tab.site_mask_hi[q];
  for (int q = tid; q < kSiteEnvN * kSiteEnvN; q += B) {
    const int t = q / kSiteEnvN, u = q % kSiteEnvN;
    const uint64_t hi = tab.site_mask_hi[t];
    s_pidx[q] = static_cast<uint8_t>(tab.site_base[t] + __popcll(hi & ((1ULL << u) - 1ULL)));
  }
  const int TY = dp.tile_y, TZH = dp.tile_zh;
  for (int q = tid; q < 2 * 43; q += B) {       // tile offset of neighbour t from a site whose tile z is odd (row 1) / even (row 0)
    const int zp = q / 43, t = q % 43;
    const int dx = tab.site_off[4 * t], dy = tab.site_off[4 * t + 1], dz = tab.site_off[4 * t + 2];
    s_tdelta[zp * 44 + t] = static_cast<int16_t>((dx * TY + dy) * TZH + (((zp + dz + 8) >> 1) - ((zp + 8) >> 1)));
  }
  const unsigned own_lo = static_cast<unsigned>(dp.own_mask[sub]), own_hi = static_cast<unsigned>(dp.own_mask[sub] >> 32);
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  const int world = dp.world, rank = dp.rank;
  const int ix_lo = domain_slab_begin(dp.ndx, world, rank), ix_hi = domain_slab_begin(dp.ndx, world, rank + 1);
  const int nd_yz = dp.ndy * dp.ndz, nd_rank = (ix_hi - ix_lo) * nd_yz, nd_total = dp.ndx * nd_yz;
  const int n_items = nw * nd_rank;
  const int safe_idx = (2 * TY + 2) * TZH + 1;            // a core cell of every tile: gather base of groups without a trial
  const int rounds = dp.rounds;
  unsigned long long sweep = *dp.sweep;
  unsigned long long line_seq = world > 1 ? *dp.line_seq : 0ULL;
  unsigned long long bar_target = 0;
  bool healthy = true;
  int err = 0;
  __syncthreads();

  for (;;) {
    // ---------------- prologue: state at the start of this sweep = update(state before, totals of the previous sweep)
    const DomState *s_prev = dp.state + static_cast<size_t>((sweep + 1) & 1ULL) * nw;
    DomState *s_next = dp.state + static_cast<size_t>(sweep & 1ULL) * nw;
    const unsigned long long *acc_prev = dp.accum + static_cast<size_t>((sweep + 2) % 3ULL) * nw * 4;
    if (tid == 0) { s_done = 1; s_fail = 0; s_tot[0] = s_tot[1] = s_tot[2] = s_tot[3] = 0ULL; }
    __syncthreads();
    for (int w = tid; w < nw; w += B) {
      DomState s = dom_load_state(s_prev + w);
      const ulonglong2 v0 = __ldcg(reinterpret_cast<const ulonglong2 *>(acc_prev + 4 * w)), v1 = __ldcg(reinterpret_cast<const ulonglong2 *>(acc_prev + 4 * w) + 1);
      const unsigned long long n_kept = v0.y, n_acc = v1.x;
      s.energy += static_cast<double>(static_cast<long long>(v0.x)) / kEnergyFixedScale;
      s.steps += n_kept;
      s.accepted += n_acc;
      if (s.sa.enabled && n_kept > 0) { sa_update_sweep(s.sa, n_kept, n_acc, s.energy, s.steps); s.temperature = s.sa.temperature; }
      if (v1.y) s_fail = 1;
      if (s.steps < target_steps) s_done = 0;
      s_temp[w] = s.temperature;
      if (cta == 0) s_next[w] = s;
    }
    if (tid == 0) {
      int sx, sy, sz;
      domain_shift(seed, sweep, px, py, pz, sx, sy, sz);
      s_shift[0] = sx; s_shift[1] = sy; s_shift[2] = sz;
    }
    __syncthreads();
    if (s_fail) { err |= kErrExtraVacancy; break; }
    if (s_done) {
      if (world > 1) {
        // final all-gather: every rank pushes its own slab (decomposition of the sweep that is not run) to all peers, so
        // that every replica is complete when the launch ends
        uint8_t *const *peers = dp.peer_occ[sweep & 1ULL];
        const uint8_t *own = dp.occ[sweep & 1ULL];
        const int b0 = domain_lo(ix_lo, px, dp.ndx), b1 = domain_lo(ix_hi, px, dp.ndx);
        const long long per_plane = static_cast<long long>(py) * lat.fz, n_cells = static_cast<long long>(b1 - b0) * per_plane;
        for (long long c = static_cast<long long>(cta) * B + tid; c < n_cells; c += static_cast<long long>(n_cta) * B) {
          const int xi = static_cast<int>(c / per_plane);
          const long long r2 = c - xi * per_plane;
          const int Y = static_cast<int>(r2 / lat.fz), zi = static_cast<int>(r2 - static_cast<long long>(Y) * lat.fz);
          int X = b0 + xi + s_shift[0]; X -= X >= px ? px : 0; X -= X >= px ? px : 0;
          const int64_t off = (static_cast<int64_t>(X + kHalo) * lat.ny + (Y + kHalo)) * lat.nz + zi + kHaloZ / 2;
          const uint8_t v = __ldcg(own + off);
          for (int p = 0; p < world; ++p)
            if (p != rank) peers[p][off] = v;
        }
        __threadfence_system();
        if (!dom_grid_barrier(dp, bar_target, n_cta)) { healthy = false; break; }
        ++line_seq;
        if (!dom_intergpu_sum(dp, line_seq, dp.accum + static_cast<size_t>((sweep + 1) % 3ULL) * nw * 4, s_sum, &s_ok2)) { healthy = false; break; }
      }
      break;
    }
    // the totals buffer (and the domain queue) of the sweep after this one were last used one sweep ago: clear them now
    if (cta == 0) {
      unsigned long long *acc_clear = dp.accum + static_cast<size_t>((sweep + 1) % 3ULL) * nw * 4;
      for (int q = tid; q < nw * 4; q += B) acc_clear[q] = 0ULL;
      if (tid == 0) dp.queue[(sweep + 1) % 3ULL] = 0u;
    }
    unsigned long long *acc_now = dp.accum + static_cast<size_t>(sweep % 3ULL) * nw * 4;
    const uint8_t *src_occ = dp.occ[sweep & 1ULL];
    const int dst_buf = static_cast<int>((sweep + 1) & 1ULL);
    uint8_t *dst_occ = dp.occ[dst_buf];
    const int shx = s_shift[0], shy = s_shift[1], shz = s_shift[2];
    int nsx = 0, nsy = 0, nsz = 0;
    if (world > 1) domain_shift(seed, sweep + 1, px, py, pz, nsx, nsy, nsz);
    (void)nsy; (void)nsz;

    long long my_fixed = 0;                     // team leader: totals of this team's domains (n_walkers == 1: summed per CTA)
    unsigned int my_kept = 0, my_acc = 0;
    unsigned int *queue = dp.queue + sweep % 3ULL;
    for (;;) {
      // next domain of the sweep: handed out dynamically (teams that drew cheap domains take more)
      int item = 0;
      if (tl == 0) item = static_cast<int>(atomicAdd(queue, 1u));
      item = __shfl_sync(0xffffffffu, item, tb);
      const bool has_item = item < n_items;
      if (!__any_sync(0xffffffffu, has_item)) break;
      int w = 0, ix = 0, iy = 0, iz = 0;
      if (has_item) {
        w = item / nd_rank;
        int d = item - w * nd_rank;
        ix = ix_lo + d / nd_yz; d %= nd_yz;
        iy = d / dp.ndz; iz = d % dp.ndz;
      }
      const int lox = domain_lo(ix, px, dp.ndx), loy = domain_lo(iy, py, dp.ndy), loz = domain_lo_z(iz, lat.fz, dp.ndz);
      const int Dx = domain_lo(ix + 1, px, dp.ndx) - lox, Dy = domain_lo(iy + 1, py, dp.ndy) - loy, Dz = domain_lo_z(iz + 1, lat.fz, dp.ndz) - loz;
      // global coordinates of tile cell (0, 0, 0): one plane below the domain
      const int gx0 = (lox + shx - 1 + px) % px, gy0 = (loy + shy - 1 + py) % py, gz0 = (loz + shz - 1 + pz) % pz;
      const int par_o = (gx0 + gy0 + gz0) & 1;
      const int64_t w_off = static_cast<int64_t>(w) * lat.padded_size;
      // ---- load the tile: (Dx + 2) x (Dy + 2) rows of (Dz + 2) / 2 sites
      if (has_item) {
        const int n_rows = (Dx + 2) * (Dy + 2), nk = (Dz + 2) >> 1;
        for (int r = tl; r < n_rows; r += T) {
          const int tx = r / (Dy + 2), ty = r - tx * (Dy + 2);
          int X = gx0 + tx; X -= X >= px ? px : 0; X -= X >= px ? px : 0;
          int Y = gy0 + ty; Y -= Y >= py ? py : 0; Y -= Y >= py ? py : 0;
          const int q = (par_o + tx + ty) & 1;
          const uint8_t *src = src_occ + w_off + (static_cast<int64_t>(X + kHalo) * lat.ny + (Y + kHalo)) * lat.nz;
          uint8_t *dst = tile + (tx * TY + ty) * TZH;
          int Z = gz0 + q; Z -= Z >= pz ? pz : 0;
          for (int k = 0; k < nk; ++k) {
            dst[k] = __ldcg(src + ((Z + kHaloZ) >> 1));
            Z += 2; Z -= Z >= pz ? pz : 0;
          }
        }
      }
      __syncwarp();
      const int nyc = Dy - 2, nzc = (Dz - 2) >> 1;
      const uint32_t ncore = has_item ? static_cast<uint32_t>((Dx - 2) * nyc * nzc) : 0u;
      const uint32_t magic_z = nzc > 1 ? 0xFFFFFFFFu / static_cast<uint32_t>(nzc) + 1u : 0u;     // floor(i / nzc) = umulhi(i, magic) for i < 2^16
      const uint32_t magic_y = nyc > 1 ? 0xFFFFFFFFu / static_cast<uint32_t>(nyc) + 1u : 0u;
      // core site index -> packed tile coordinates tx | ty << 8 | zi << 16 (z as pair index)
      auto decode = [&](uint32_t i) -> uint32_t {
        const uint32_t q1 = nzc > 1 ? __umulhi(i, magic_z) : i;
        const uint32_t q2 = nyc > 1 ? __umulhi(q1, magic_y) : q1;
        return (q2 + 2u) | ((q1 - q2 * nyc + 2u) << 8) | ((i - q1 * nzc + 1u) << 16);
      };
      auto tile_index = [&](uint32_t packed) -> int {
        return (static_cast<int>(packed & 0xFFu) * TY + static_cast<int>((packed >> 8) & 0xFFu)) * TZH + static_cast<int>(packed >> 16);
      };
      // ---- the non-solvent sites of the core, in core-index order
      uint32_t n_sol = 0;
      for (uint32_t i0 = 0; i0 < static_cast<uint32_t>(dp.max_core); i0 += T) {
        const uint32_t i = i0 + tl;
        bool sol = false;
        if (i < ncore) sol = tile[tile_index(decode(i))] != solvent;
        const unsigned bal = (__ballot_sync(0xffffffffu, sol) >> tb) & kTeamBits;
        if (sol) list[n_sol + __popc(bal & ((1u << tl) - 1u))] = static_cast<uint16_t>(i);
        n_sol += __popc(bal);
      }
      __syncwarp();
      // ---- Metropolis rounds
      const double beta = 1.0 / kBoltzmannEv / fmax(s_temp[w], 1e-12);
      const uint32_t item_global = static_cast<uint32_t>(w) * static_cast<uint32_t>(nd_total) + static_cast<uint32_t>((ix * dp.ndy + iy) * dp.ndz + iz);
      long long fixed = 0;
      unsigned int kept = 0, acc = 0;
      const bool can_draw = n_sol > 0u && ncore > 1u;
      // The first draw of a round is state independent up to the list lookup: lane j of the team draws for round blk + j,
      // T rounds at a time (one Philox call per lane and T rounds), and a round's values are broadcast when it comes up.
      uint32_t pre_kb = 0, pre_ulo = 0, pre_uhi = 0;
      int r0 = can_draw ? 0 : rounds, blk = -T;              // next round to commit; first round of the drawn block
#ifdef LMC_DOM_PROFILE
      const long long tp0 = clock64();
      int tp_steps = 0;
#endif
      while (__any_sync(0xffffffffu, r0 < rounds)) {
        if (r0 < rounds && r0 >= blk + T) {
          blk += T;
          uint32_t r[4];
          philox4x32_10_c4(static_cast<uint32_t>((blk + tl) * kDomTries), item_global, static_cast<uint32_t>(sweep), static_cast<uint32_t>(sweep >> 32),
                           static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
          pre_kb = __umulhi(r[0], n_sol) | (__umulhi(r[1], ncore) << 16);     // list slot | core index of b
          pre_ulo = r[2]; pre_uhi = r[3];
        }
        // this group's round of the batch; a batch does not run past the drawn block
        const int my_round = r0 + grp;
        const int limit = min(rounds, blk + T);
        const bool active = my_round < limit;
        // -- draw (identical in every lane of the group)
        bool found = false;
        uint32_t pa = 0, pb = 0, slot = 0, core_b = 0, u_lo = 0, u_hi = 0;
        unsigned ea = solvent, eb = solvent;
        {
          const int src = active ? tb + (my_round - blk) : lane;
          const uint32_t kb = __shfl_sync(0xffffffffu, pre_kb, src), w2 = __shfl_sync(0xffffffffu, pre_ulo, src), w3 = __shfl_sync(0xffffffffu, pre_uhi, src);
          if (active) {
            const uint32_t k = kb & 0xFFFFu, ib = kb >> 16;
            const uint32_t qa_ = decode(list[k]), qb_ = decode(ib);
            const unsigned ca = tile[tile_index(qa_)], cb = tile[tile_index(qb_)];
            // unlike species; a pair of two non-solvent species is reachable from both ends: keep it with probability 1/2
            // (bit 0 of the low word is below the 53 bits the uniform uses)
            if (ca != cb && (cb == solvent || (w2 & 1u))) {
              found = true; pa = qa_; pb = qb_; slot = k; core_b = ib; ea = ca; eb = cb; u_lo = w2; u_hi = w3;
            }
          }
        }
        for (int tr = 1; tr < kDomTries; ++tr) {    // redraws (like species, or a thinned non-solvent pair)
          if (__all_sync(0xffffffffu, found || !active)) break;
          if (active && !found) {
            uint32_t r[4];
            philox4x32_10_c4(static_cast<uint32_t>(my_round * kDomTries + tr), item_global, static_cast<uint32_t>(sweep), static_cast<uint32_t>(sweep >> 32),
                             static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
            const uint32_t k = __umulhi(r[0], n_sol), ib = __umulhi(r[1], ncore);
            const uint32_t qa_ = decode(list[k]), qb_ = decode(ib);
            const unsigned ca = tile[tile_index(qa_)], cb = tile[tile_index(qb_)];
            if (ca != cb && (cb == solvent || (r[2] & 1u))) {
              found = true; pa = qa_; pb = qb_; slot = k; core_b = ib; ea = ca; eb = cb; u_lo = r[2]; u_hi = r[3];
            }
          }
        }
        // -- evaluate: EnergyChangePredictorPairSite::GetDeFromLatticeIdPair on the tile (L lanes per site).  Warp-uniform
        // control flow: groups without a trial gather around a dummy cell and discard the result.
        int txa = pa & 0xFFu, tya = (pa >> 8) & 0xFFu, zia = pa >> 16, txb = pb & 0xFFu, tyb = (pb >> 8) & 0xFFu, zib = pb >> 16;
        int qa = (par_o + txa + tya) & 1, qb = (par_o + txb + tyb) & 1;
        const int idx_a0 = (txa * TY + tya) * TZH + zia, idx_b0 = (txb * TY + tyb) * TZH + zib;
        int idx_a = idx_a0, idx_b = idx_b0;
        // displacement b - a in half-units (both sites in one tile; core sites are never neighbours through the period)
        const int dx = txb - txa, dy = tyb - tya, dz = (2 * zib + qb) - (2 * zia + qa);
        const bool coupled = found && dx * dx + dy * dy + dz * dz <= 6;
        unsigned e1 = ea, e2 = eb;                  // species at the site evaluated first / second
        if (coupled && eb == vac) {                 // move the vacancy first: no intermediate state with two vacancies
          idx_a = idx_b0; idx_b = idx_a0;
          const int tq = qa; qa = qb; qb = tq;
          e1 = eb; e2 = ea;
        }
        const int base = found ? (side ? idx_b : idx_a) : safe_idx;
        const int16_t *drow = s_tdelta + (side ? qb : qa) * 44;
        const int override_index = (side && coupled) ? idx_a : -1;    // side 1 of a coupled pair sees the first site changed
        constexpr int kPerLane = (43 + L - 1) / L;
        unsigned m_lo = 0, m_hi = 0;                // non-solvent mask over the 43 list positions of this side
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {
          const int t = sub + i * L;
          unsigned c = solvent;
          if (t < 43) {
            const int at = base + drow[t];
            c = tile[at];
            if (at == override_index) c = e2;
            if (t != kCentrePos) row[t - (t > kCentrePos)] = static_cast<uint8_t>(kTab ? c - (c > solvent) : c);
            else c = solvent;
          }
          const unsigned bits = (__ballot_sync(0xffffffffu, c != solvent) >> side_shift) & kSideBits;   // positions i L .. i L + L - 1
          if (i * L < 32) m_lo |= bits << ((i * L) & 31); else m_hi |= bits << ((i * L) & 31);
        }
        // list position -> environment index: drop the (always clear) centre bit
        const uint64_t m43 = (static_cast<uint64_t>(m_hi) << 32) | m_lo;
        const uint64_t env = (m43 & 0x1FFFFFULL) | ((m43 >> 22) << 21);
        const unsigned lo = static_cast<unsigned>(env), hi = static_cast<unsigned>(env >> 32);
        __syncwarp();
        // Table walk over this lane's share of the non-solvent positions (tried and dropped: four partners at a time with
        // predication and two accumulators -- more instructions than latency saved: -10 % at 40^3, -17 % at 100^3).
        double de = 0.0;
        auto walk_all = [&](auto single, auto pair) {
          auto walk = [&](unsigned mine, int t0) {
            while (mine) {
              const int t = t0 + __ffs(static_cast<int>(mine)) - 1;
              mine &= mine - 1;
              const int et = row[t];
              de += single(t, et);
              const uint64_t mask = s_mask[t];
              const uint8_t *prow = s_pidx + t * kSiteEnvN;
              unsigned plo = static_cast<unsigned>(mask) & lo, phi = static_cast<unsigned>(mask >> 32) & hi;
              while (plo) {
                const int u = __ffs(static_cast<int>(plo)) - 1;
                plo &= plo - 1;
                de += pair(prow[u], et, row[u]);
              }
              while (phi) {
                const int u = 32 + __ffs(static_cast<int>(phi)) - 1;
                phi &= phi - 1;
                de += pair(prow[u], et, row[u]);
              }
            }
          };
          walk(lo & own_lo, 0);
          walk(hi & own_hi, 32);
        };
        if (kTab) {
          // difference tables: the species pair of the swap, the direction as a sign (side 1 changes the other way)
          const unsigned spw = s_sp[e1 * m + e2];
          const int sp = spw & 0x7F;
          const double *DA = s_A + sp * (kSiteEnvN * ns), *DB = s_B + sp * (tab.n_site_pairs * ns2);
          if (sub == 0) de = s_C[sp];
          if (found)
            walk_all([&](int t, int et) { return DA[t * ns + et]; },
                     [&](int p, int et, int eu) { return DB[(p * ns + et) * ns + eu]; });
          if (((spw >> 7) ^ static_cast<unsigned>(side)) & 1u) de = -de;
        } else {
          const int x_old = static_cast<int>(side ? e2 : e1), x_new = static_cast<int>(side ? e1 : e2);
          const int a_stride = kSiteEnvN * m, b_stride = tab.n_site_pairs * mm;
          const double *A_new = s_A + x_new * a_stride, *A_old = s_A + x_old * a_stride;
          const double *B_new = tab.site_B + x_new * b_stride, *B_old = tab.site_B + x_old * b_stride;
          if (sub == 0) de = s_C[x_new] - s_C[x_old];
          if (found)
            walk_all([&](int t, int et) { return A_new[t * m + et] - A_old[t * m + et]; },
                     [&](int p, int et, int eu) { const int q = (p * m + et) * m + eu; return __ldg(B_new + q) - __ldg(B_old + q); });
        }
        __syncwarp();
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) de += __shfl_xor_sync(0xffffffffu, de, off);
        bool accept = false;
        if (found) {
          if (de != de) err |= kErrExtraVacancy;
          // CanonicalMcAbstract::SelectEvent (:86-101): dE < 0 accepts, else u < exp(-dE beta).  The exponential is bracketed in
          // single precision first; the double-precision call only decides the (rare) draws inside the bracket, so the
          // decisions are those of the double-precision test
          accept = de < 0.0;
          if (!accept) {
            const double x = -de * beta, u = uniform53(u_lo, u_hi);
            const float ef = __expf(static_cast<float>(x));
            const double e_lo = static_cast<double>(ef) * (1.0 - 1e-4), e_hi = static_cast<double>(ef) * (1.0 + 1e-4);
            accept = u < e_lo;
            if (!accept && !(u > e_hi)) accept = u < exp(x);
          }
        }
        // -- commit the batch up to and including its first accepted round
        {
          const unsigned acc_team = (__ballot_sync(0xffffffffu, accept) >> tb) & kTeamBits;
          const unsigned found_team = (__ballot_sync(0xffffffffu, found) >> tb) & kTeamBits & kLeaders;
          const int first = acc_team ? (__ffs(static_cast<int>(acc_team)) - 1) / G : S;      // slot of the first accepted round
          const int n_active = max(0, min(S, limit - r0));
          const int committed = min(first + 1, n_active);
          kept += __popc(found_team & (committed * G >= 32 ? 0xFFFFFFFFu : ((1u << (committed * G)) - 1u)));
          if (S > 1) {
            const double de_first = __shfl_sync(0xffffffffu, de, tb + (first < S ? first : 0) * G);
            if (first < S) { ++acc; fixed += __double2ll_rn(de_first * kEnergyFixedScale); }
          } else if (first < S) { ++acc; fixed += __double2ll_rn(de * kEnergyFixedScale); }
          if (grp == first && gl == 0) {
            tile[idx_a0] = static_cast<uint8_t>(eb); tile[idx_b0] = static_cast<uint8_t>(ea);
            if (eb == solvent) list[slot] = static_cast<uint16_t>(core_b);     // the non-solvent atom now sits at b
          }
          if (r0 < rounds) r0 += committed;
#ifdef LMC_DOM_PROFILE
          if (n_active > 0) ++tp_steps;
#endif
        }
        __syncwarp();
      }
#ifdef LMC_DOM_PROFILE
      if (tl == 0 && has_item) {
        const long long cyc = clock64() - tp0;
        atomicAdd(&g_dom_hist[0][min(63, tp_steps / 4)], 1ULL);                 // speculative batches per domain-sweep
        atomicAdd(&g_dom_hist[1][min(63, static_cast<int>(cyc >> 15))], 1ULL);  // cycles per domain-sweep (2^15 per bin)
        atomicAdd(&g_dom_hist[2][min(63, static_cast<int>(n_sol))], 1ULL);      // non-solvent core sites
        atomicAdd(&g_dom_hist[3][min(63, static_cast<int>(acc))], 1ULL);        // accepted trials
        atomicAdd(&g_dom_hist[4][min(63, static_cast<int>(n_sol))], static_cast<unsigned long long>(cyc));   // cycles by n_sol
        atomicAdd(&g_dom_hist[5][min(63, static_cast<int>(n_sol))], static_cast<unsigned long long>(tp_steps));
      }
#endif
      // ---- write the domain back (canonical cells only) -- locally and to the ranks that will hold these planes next sweep
      if (has_item) {
        const int nk = Dz >> 1;
        const int n_rows = Dx * Dy;
        for (int r = tl; r < n_rows; r += T) {
          const int tx = r / Dy + 1, ty = r - (tx - 1) * Dy + 1;
          int X = gx0 + tx; X -= X >= px ? px : 0; X -= X >= px ? px : 0;
          int Y = gy0 + ty; Y -= Y >= py ? py : 0; Y -= Y >= py ? py : 0;
          const int q = (par_o + tx + ty) & 1;
          const int64_t row_off = w_off + (static_cast<int64_t>(X + kHalo) * lat.ny + (Y + kHalo)) * lat.nz;
          const int k0 = q ? 0 : 1;                 // tile z in [1, Dz]: pair 0 holds z = 1 only for odd-parity rows
          const uint8_t *srcp = tile + (tx * TY + ty) * TZH + k0;
          unsigned dest = 1u << rank;
          if (world > 1) {
            // ranks whose slab (plus one halo plane per side) holds plane X in the next sweep's decomposition
            for (int p = 0; p < world; ++p) {
              const int b0 = domain_lo(domain_slab_begin(dp.ndx, world, p), px, dp.ndx), b1 = domain_lo(domain_slab_begin(dp.ndx, world, p + 1), px, dp.ndx);
              if (b1 == b0) continue;
              int rel = X - (b0 + nsx - 1); rel %= px; rel += rel < 0 ? px : 0;
              if (rel < b1 - b0 + 2) dest |= 1u << p;
            }
          }
          int Z = gz0 + 2 * k0 + q; Z -= Z >= pz ? pz : 0; Z -= Z >= pz ? pz : 0;
          for (int k = 0; k < nk; ++k) {
            const int64_t off = row_off + ((Z + kHaloZ) >> 1);
            const uint8_t v = srcp[k];
            dst_occ[off] = v;
            if (world > 1) {
              unsigned rest = dest & ~(1u << rank);
              while (rest) {
                const int p = __ffs(static_cast<int>(rest)) - 1;
                rest &= rest - 1;
                dp.peer_occ[dst_buf][p][off] = v;
              }
            }
            Z += 2; Z -= Z >= pz ? pz : 0;
          }
        }
        if (tl == 0) {
          if (nw > 1) {
            if (fixed) atomicAdd(acc_now + 4 * w, static_cast<unsigned long long>(fixed));
            if (kept) atomicAdd(acc_now + 4 * w + 1, static_cast<unsigned long long>(kept));
            if (acc) atomicAdd(acc_now + 4 * w + 2, static_cast<unsigned long long>(acc));
          } else {
            my_fixed += fixed; my_kept += kept; my_acc += acc;
          }
        }
      }
      __syncwarp();
    }
    if (nw == 1) {
      // one lattice: totals per CTA through shared memory, then one set of global atomics
      long long f = tl == 0 ? my_fixed : 0LL;
      unsigned int k = tl == 0 ? my_kept : 0u, a = tl == 0 ? my_acc : 0u;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        f += __shfl_xor_sync(0xffffffffu, f, off);
        k += __shfl_xor_sync(0xffffffffu, k, off);
        a += __shfl_xor_sync(0xffffffffu, a, off);
      }
      if (lane == 0) {
        if (f) atomicAdd(&s_tot[0], static_cast<unsigned long long>(f));
        if (k) atomicAdd(&s_tot[1], static_cast<unsigned long long>(k));
        if (a) atomicAdd(&s_tot[2], static_cast<unsigned long long>(a));
      }
    }
    const int block_err = __syncthreads_or(err != 0);
    if (tid == 0) {
      if (nw == 1) {
        if (s_tot[0]) atomicAdd(acc_now, s_tot[0]);
        if (s_tot[1]) atomicAdd(acc_now + 1, s_tot[1]);
        if (s_tot[2]) atomicAdd(acc_now + 2, s_tot[2]);
      }
      if (block_err) atomicAdd(acc_now + 3, 1ULL);
    }
    if (world > 1) __threadfence_system();        // this thread's peer stores are performed before the CTA arrives
    if (!dom_grid_barrier(dp, bar_target, n_cta)) { healthy = false; break; }
    if (world > 1) {
      // every rank's totals to every rank; the lines double as the inter-GPU barrier of the sweep
      ++line_seq;
      if (!dom_intergpu_sum(dp, line_seq, acc_now, s_sum, &s_ok2)) { healthy = false; break; }
      // the next prologue reads the totals from acc_now: replace the local sums by the global ones (identical on every rank;
      // every CTA has passed the grid barrier, nobody adds to acc_now any more)
      if (cta == 0 && tid < 4) acc_now[tid] = s_sum[tid];
      if (!dom_grid_barrier(dp, bar_target, n_cta)) { healthy = false; break; }
    }
    ++sweep;
  }
#ifdef LMC_DOM_PROFILE
  __syncthreads();
  if (cta == 0 && tid == 0) {
    const char *names[4] = {"batches/4", "cycles>>15", "n_sol", "accepted"};
    for (int h = 0; h < 4; ++h) {
      printf("hist %s:", names[h]);
      for (int q = 0; q < 64; ++q) if (g_dom_hist[h][q]) printf(" %d:%llu", q, g_dom_hist[h][q]);
      printf("\n");
    }
    printf("mean cycles / batches by n_sol:");
    for (int q = 0; q < 64; ++q) if (g_dom_hist[2][q]) printf(" %d:%llu/%llu", q, g_dom_hist[4][q] / g_dom_hist[2][q], g_dom_hist[5][q] / g_dom_hist[2][q]);
    printf("\n");
  }
#endif
  __syncthreads();
  if (cta == 0) {
    // leaving through the "done" test: the prologue has just written the final state into the parity buffer of `sweep`
    const DomState *fin = dp.state + static_cast<size_t>(sweep & 1ULL) * nw;
    if (healthy && !err)
      for (int w = tid; w < nw; w += B) {
        const DomState s = fin[w];
        st.energy[w] = s.energy; st.steps[w] = s.steps; st.accepted[w] = s.accepted;
        SaSchedule sa = s.sa;
        if (sa.enabled) sa.temperature = s.temperature;
        st.sa[w] = sa;
      }
    if (tid == 0) { *dp.sweep = sweep; if (world > 1) *dp.line_seq = line_seq; }
  }
  if (!healthy) err |= kErrBadSite;
  if (err && tid == 0) atomicOr(&st.error[0], err);
}

}  // namespace lmc

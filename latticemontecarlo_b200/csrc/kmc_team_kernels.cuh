// kmc_team_kernels.cuh -- first-order KMC (mc::KineticMcFirstOmp semantics) for FEW walkers: the latency kernel.
// Reference: mc/src/KineticMcAbstract.cpp:140-188 (OneStepSimulation), mc/src/KineticMcFirstOmp.cpp:52-82
// (BuildEventList / CalculateTime), mc/src/JumpEvent.cpp:6-13.
//
// kmc_run_kernel (kmc_kernels.cuh) gives a walker one half-warp and is built for throughput: thousands of walkers hide
// the ~6000-cycle dependent chain of a step behind each other.  With a few walkers per SM (the 8192-walker job spread
// over 8 GPUs, or a single trajectory) that chain IS the step time.  Here a whole thread block owns one walker and the
// chain is cut instead:
//  * G lanes per candidate jump (12 G "event" threads) gather the 60 ordered sites of that jump directly (one byte per
//    lane and pass -- no box scan, no compaction, no cell -> env mapping loop), the non-solvent mask comes from ballots,
//    every lane contracts the table terms of the sites it loaded itself (A term + its pairs with higher partners, the
//    table loads of up to four partners in flight together), and a shuffle tree adds the partial sums.  The first lane
//    of each group files (dE, log E0) in the reference's event order (slot = rank of the neighbour's lattice id, a table
//    lookup by the vacancy's boundary / parity class).
//  * one more warp, the SELECTOR, owns the walker's clock: while the event threads work it prepares the step's
//    uniforms; when the events have arrived (named barrier A: the events only ARRIVE) it evaluates the 12 closed forms
//    and rates in one instruction stream, runs the reference's sequential total / division / running sum /
//    select, writes the jump, publishes the new vacancy site and arrives at barrier B, on which the event threads wait.
//    The residence time, the clock, T(t) and the rate corrector are updated after that, off the critical path.
//
// Same Philox stream (key = seed ^ walker, counter = step number), same (Ea, rate) chain and same select as kmc_run_kernel.
// The contracted sums are added in a different order (tree over lanes), but the folded tables sit on a binary grid
// (Engine::load_coefficients) on which every partial sum is exact, so (dE, log E0) -- and with them rates, clocks,
// energies and trajectories -- are bit-identical to kmc_run_kernel's.  That is what lets Engine::kmc_run hand the tail of
// a large launch from that kernel to this one (prm.steps_target / tail_order) without the result depending on when.
#pragma once
#include "kmc_kernels.cuh"
#ifdef LMC_KMC_TEAM_PROFILE
#include <cstdio>
#define LMC_TEAM_TICK(i) do { const long long now_ = clock64(); prof[i] += now_ - tick; tick = now_; } while (0)
#else
#define LMC_TEAM_TICK(i) do { } while (0)
#endif

namespace lmc {

// named barriers (PTX bar.sync / bar.arrive): producers arrive without waiting, consumers wait; the barrier orders the
// producers' earlier shared / global writes before the consumers' later reads
__device__ __forceinline__ void team_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void team_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// kSmemOcc: the walker's whole (padded) occupancy lives in shared memory for the launch -- small cells only (the 8 x 8 x 8
// cell of the batched workload is 5.8 KB).  The gather then never leaves the SM (shared-memory latency instead of L1 hits
// plus two L2 round trips for the lines the previous jump has just written); the selector writes each jump to both copies.
template <int G, bool kInstrumented, bool kSmemOcc>
__global__ void __launch_bounds__(12 * G + 32, G == 8 ? 7 : (G == 16 ? 3 : 1))     // resident blocks per SM the dispatch counts on
kmc_team_run_kernel(LatticeDesc lat, DevTables tab, uint8_t *occ, int64_t walker_stride, int n_walkers, KmcState st, KmcParams prm,
                    int64_t n_steps_all, const double *__restrict__ replay_u1, const double *__restrict__ replay_u2, KmcTraceDev tr) {
  // tail of a hybrid launch (prm.steps_target): block b takes the walker with the b-th most steps left, for exactly those
  // steps; the blocks beyond the list leave at once
  int64_t n_steps = n_steps_all;
  int w = blockIdx.x;
  if (prm.steps_target) {
    if (static_cast<int>(blockIdx.x) >= *prm.tail_count) return;
    w = prm.tail_order[blockIdx.x];
    n_steps = prm.steps_target[w] - st.steps[w];
    if (n_steps <= 0) return;
  }
  static_assert(G == 8 || G == 16 || G == 32, "lanes per candidate jump");
  constexpr int NJ = (60 + G - 1) / G;            // gather passes: lane `sub` loads the state positions sub, sub + G, ...
  constexpr int kEventThreads = 12 * G, kThreads = kEventThreads + 32;
  constexpr int kBarA = 1, kBarB = 2;
  constexpr unsigned full = 0xFFFFFFFFu;
  __shared__ int32_t s_delta[24 * kPairDeltaStride + 4];
  extern __shared__ double s_A2[];                // [n][58][n][2]: (dE, log E0) singlet terms; kSmemOcc: followed by the occupancy
  __shared__ uint64_t s_mask_hi[kEnvN];
  __shared__ uint16_t s_pbase[kEnvN];
  __shared__ uint8_t s_codes[12][64];             // species by env index, per candidate jump (non-solvent sites only)
  __shared__ double s_de[12], s_le[12];           // (dE, log E0) in event (slot) order
  __shared__ __align__(16) double s_p[12];
  __shared__ uint8_t s_dir[12], s_mig[12];
  __shared__ __align__(16) int s_sel[4];          // the new vacancy site (x, y, z) and the stop flag: one 16-byte load per event thread
  __shared__ int s_err;
  // The event order of the 12 jumps (ascending lattice id of the neighbour) depends on the vacancy site only through, per
  // axis, whether a neighbour wraps around the period (coordinate 0 or period - 1) and the coordinate's parity: 4 classes
  // per axis.  Slot of jump k for every class: DevTables::kmc_slot, ranked on the host on a representative site of the class.
  __shared__ uint8_t s_slot[64][12];
  for (int q = threadIdx.x; q < tab.n_species * kEnvN * tab.n_species * 2; q += blockDim.x) s_A2[q] = tab.pair_A2[q];
  for (int q = threadIdx.x; q < kEnvN; q += blockDim.x) { s_mask_hi[q] = tab.pair_mask_hi[q]; s_pbase[q] = tab.pair_base[q]; }
  for (int q = threadIdx.x; q < 24 * kPairDeltaStride; q += blockDim.x) s_delta[q] = tab.pair_delta[q];
  if (threadIdx.x == 0) s_err = 0;
  for (int q = threadIdx.x; q < 64 * 12; q += blockDim.x) (&s_slot[0][0])[q] = tab.kmc_slot[q];
  __syncthreads();
  if (w >= n_walkers || st.error[w] != 0 || st.vacancy[w] < 0) return;     // uniform over the block
  uint8_t *o = occ + w * walker_stride;
  uint8_t *s_occ = reinterpret_cast<uint8_t *>(s_A2 + tab.n_species * kEnvN * tab.n_species * 2);
  if (kSmemOcc) {
    for (int q = threadIdx.x; q < static_cast<int>(lat.padded_size); q += blockDim.x) s_occ[q] = o[q];
    __syncthreads();
  }
  int X, Y, Z;
  lat.coords_of_id(st.vacancy[w], X, Y, Z);
  const unsigned solvent = static_cast<unsigned>(tab.solvent), vac_code = static_cast<unsigned>(tab.n_species);
  const int n = tab.n_species;
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  const int lane = threadIdx.x & 31;
  // (Rotating the roles over the warps of a block from walker to walker, so that the selector warps of an SM do not all
  // sit at the same warp index, was measured at 7 walkers per SM: 2.771 vs 2.774 us per step -- no effect, not kept.)
  const int vtid = static_cast<int>(threadIdx.x);

  if (vtid < kEventThreads) {
    // ================================================================================================ event threads
    const int k = vtid / G;                          // candidate jump (first-neighbour direction) of this lane group
    const int sub = vtid % G;
    const int gshift = lane & ~(G - 1);              // position of the group inside its warp
    const unsigned gmask = G == 32 ? full : ((1u << G) - 1u);
    const double2 *__restrict__ B_all = reinterpret_cast<const double2 *>(tab.pair_B2);
    const double2 *A_all = reinterpret_cast<const double2 *>(s_A2);
    const uint2 *mask_hi2 = reinterpret_cast<const uint2 *>(s_mask_hi);
    const int b_stride = tab.n_pair_pairs * n * n;
    uint8_t *codes = s_codes[k];
#ifdef LMC_KMC_TEAM_PROFILE
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tick = clock64();
#endif
    for (int64_t s = 0; s < n_steps; ++s) {
      // ---- gather: the 60 ordered sites of the jump k, one byte per lane and pass
      const int zp = Z & 1;
      const int64_t base = lat.padded_index(X, Y, Z);
      const int32_t *drow = s_delta + (k * 2 + zp) * kPairDeltaStride;
      unsigned c[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int t = sub + G * j;
        if (kSmemOcc) c[j] = t < 60 ? static_cast<unsigned>(s_occ[static_cast<int>(base) + drow[t]]) : solvent;
        else c[j] = t < 60 ? static_cast<unsigned>(o[base + drow[t]]) : solvent;
      }
      // ---- event order (KineticMcFirstOmp.cpp:52-68): slot = rank of this jump's neighbour id among the 12 neighbour ids,
      // a function of the vacancy's boundary / parity class only (table built at block start)
      const int slot = s_slot[(coord_class(X, px) * 4 + coord_class(Y, py)) * 4 + coord_class(Z, pz)][k];
      // ---- non-solvent mask over the env index, species of the two pair sites
      uint64_t pm = 0;                               // by state position
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const unsigned b = (__ballot_sync(full, c[j] != solvent) >> gshift) & gmask;
        pm |= static_cast<uint64_t>(b) << (G * j);
      }
      const unsigned first = __shfl_sync(full, c[kFirstPos / G], kFirstPos % G, G);
      const unsigned mig = __shfl_sync(full, c[kSecondPos / G], kSecondPos % G, G);
      // env index = position minus the pair sites below it (positions 21 and 38)
      const uint64_t sol = (pm & ((1ULL << kFirstPos) - 1ULL)) | (((pm >> (kFirstPos + 1)) & ((1ULL << (kSecondPos - kFirstPos - 1)) - 1ULL)) << kFirstPos) |
                           ((pm >> (kSecondPos + 1)) << (kSecondPos - 1));
      LMC_TEAM_TICK(0);                              // gather + slot + masks
      const uint32_t sol_lo = static_cast<uint32_t>(sol), sol_hi = static_cast<uint32_t>(sol >> 32);
      int err = 0;
      if (first != vac_code || mig == vac_code) err |= kErrNotVacancy;
      // this lane's non-solvent env sites as a bit per gather pass + the packed species bytes; their species go to shared
      // memory for the partner lookups of the other lanes
      unsigned own = 0;
      uint64_t cw = 0;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int tp = sub + G * j;
        if (c[j] != solvent && tp < 60 && tp != kFirstPos && tp != kSecondPos) {
          own |= 1u << j;
          codes[env_of_pos(tp)] = static_cast<uint8_t>(c[j]);
        }
        cw |= static_cast<uint64_t>(c[j]) << (8 * j);
      }
      __syncwarp(full);
      // ---- contracted tables: every lane adds the terms of the sites it loaded (A) and of their pairs with higher partners (B)
      double a0 = 0.0, a1 = 0.0;
      if (mig < vac_code) {
        const int m = static_cast<int>(mig);
        const double2 *A = A_all + m * (kEnvN * n);
        const double2 *__restrict__ B = B_all + m * b_stride;
        if (sub == 0) { a0 = __ldg(tab.pair_C2 + m * 2); a1 = __ldg(tab.pair_C2 + m * 2 + 1); }
        while (own) {                                                     // ONE copy of the loop body for all passes
          const int j = __ffs(static_cast<int>(own)) - 1;
          own &= own - 1;
          const int et = static_cast<int>((cw >> (8 * j)) & 0xFFu);
          if (et >= n) { err |= kErrExtraVacancy; continue; }
          const int t = env_of_pos(sub + G * j);
          const double2 a = A[t * n + et];
          a0 += a.x; a1 += a.y;
          const uint2 hi = mask_hi2[t];
          uint32_t p_lo = hi.x & sol_lo, p_hi = hi.y & sol_hi;          // non-solvent partners u > t
          if ((p_lo | p_hi) == 0) continue;
          const int row = (s_pbase[t] * n + et) * n;
          do {                                                            // four partners per pass: the table loads of a pass are all
            double2 b[4];                                                 // in flight before the first one is added
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              b[q] = make_double2(0.0, 0.0);
              if ((p_lo | p_hi) == 0) continue;
              int u;
              if (p_lo) { u = __ffs(static_cast<int>(p_lo)) - 1; p_lo &= p_lo - 1; }
              else { u = 31 + __ffs(static_cast<int>(p_hi)); p_hi &= p_hi - 1; }
              const int eu = codes[u];
              if (eu >= n) continue;                                      // reported by the lane that owns u
              const int rank = u < 32 ? __popc(hi.x & ((1u << u) - 1u)) : __popc(hi.x) + __popc(hi.y & ((1u << (u - 32)) - 1u));
              b[q] = __ldg(B + row + rank * (n * n) + eu);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) { a0 += b[q].x; a1 += b[q].y; }
          } while (p_lo | p_hi);
        }
      }
      __syncwarp(full);
      LMC_TEAM_TICK(1);                              // table walk
#pragma unroll
      for (int off = G / 2; off > 0; off >>= 1) {
        a0 += __shfl_xor_sync(full, a0, off, G);
        a1 += __shfl_xor_sync(full, a1, off, G);
      }
      LMC_TEAM_TICK(2);                              // shuffle tree
      if (err) atomicOr(&s_err, err);
      else if (sub == 0) {                           // (dE, log E0) of this jump; the selector evaluates the 12 closed forms in one stream
        s_de[slot] = a0;
        s_le[slot] = a1;
        s_dir[slot] = static_cast<uint8_t>(k);
        s_mig[slot] = static_cast<uint8_t>(mig);
      }
      __syncwarp(full);
      LMC_TEAM_TICK(3);                              // filing
      team_bar_arrive(kBarA, kThreads);              // the 12 events of this step are filed
      team_bar_sync(kBarB, kThreads);                // the selector has jumped (or stopped the walker)
      const int4 sel = *reinterpret_cast<const int4 *>(s_sel);   // after the barrier (a memory clobber): re-read every step
      if (sel.w) break;
      LMC_TEAM_TICK(4);                              // wait for the selector (the load cannot pass the barrier)
      X = sel.x; Y = sel.y; Z = sel.z;
    }
#ifdef LMC_KMC_TEAM_PROFILE
    if (blockIdx.x == 0 && vtid == 0)
      printf("team profile G=%d event thread 0, cycles per step: gather %lld walk %lld tree %lld closed form %lld wait for selector %lld\n", G,
             prof[0] / n_steps, prof[1] / n_steps, prof[2] / n_steps, prof[3] / n_steps, prof[4] / n_steps);
#endif
    return;
  }

  // ==================================================================================================== selector warp
  double time = st.time[w], energy = st.energy[w], temperature = st.temperature[w];
  int64_t steps = st.steps[w];
  const double c_vac = st.c_vacancy[w], c_sol = st.c_solute[w];
  const int ql = lane < 12 ? lane : 0;             // lane q < 12 keeps direction q (the jump) and owns slot q (the select)
  const int barrier_model = tab.barrier_model;
  const int dxl = tab.nn1[4 * ql], dyl = tab.nn1[4 * ql + 1], dzl = tab.nn1[4 * ql + 2];
  const bool tracing = kInstrumented && (tr.from || tr.to || tr.slot || tr.dt || tr.Ea || tr.dE || tr.total_rate || tr.temperature);
  double beta = 1.0 / kBoltzmannEv / temperature;
  double corr = prm.rate_corrector ? rate_correction(c_vac, c_sol, temperature) : 1.0;
  double corr_over_prefactor = corr / kPrefactorHz;
  double ahead_neg_log_u1 = 0.0, ahead_u2 = 0.0;   // lane l < 16: -ln(u1) and u2 of step (s & ~15) + l
  bool failed = false;
#ifdef LMC_KMC_TEAM_PROFILE
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tick = clock64();
#endif
  for (int64_t s = 0; s < n_steps; ++s) {
    // ---- before the events arrive: temperature of this step, its two uniforms, the neighbour sites
    if (prm.n_tt > 0) {                            // UpdateTemperature (KineticMcAbstract.cpp:45-50)
      const double t_now = interpolate_temperature(prm, time);
      if (t_now != temperature) {
        temperature = t_now;
        beta = 1.0 / kBoltzmannEv / temperature;
        if (prm.rate_corrector) { corr = rate_correction(c_vac, c_sol, temperature); corr_over_prefactor = corr / kPrefactorHz; }
      }
    }
    double neg_log_u1, u2;
    if (kInstrumented && replay_u1) {
      neg_log_u1 = -log(replay_u1[static_cast<int64_t>(w) * n_steps + s]);
      u2 = replay_u2[static_cast<int64_t>(w) * n_steps + s];
    } else {
      if ((s & 15) == 0) {                         // lane l draws for step s + l (same stream as kmc_run_kernel)
        const int64_t ctr = steps + (lane & 15);
        uint32_t r[4];
        philox4x32_10(static_cast<uint32_t>(ctr), static_cast<uint32_t>(static_cast<uint64_t>(ctr) >> 32),
                      static_cast<uint32_t>(prm.seed) ^ static_cast<uint32_t>(w), static_cast<uint32_t>(prm.seed >> 32), r);
        ahead_neg_log_u1 = -log(uniform53(r[0], r[1]) + (1.0 / 9007199254740992.0));
        ahead_u2 = uniform53(r[2], r[3]);
      }
      neg_log_u1 = __shfl_sync(full, ahead_neg_log_u1, static_cast<int>(s & 15));
      u2 = __shfl_sync(full, ahead_u2, static_cast<int>(s & 15));
    }
    const int xq = wrap_coord(X + dxl, px), yq = wrap_coord(Y + dyl, py), zq = wrap_coord(Z + dzl, pz);   // lane q < 12: the neighbour in direction q
    LMC_TEAM_TICK(0);                              // preparation
    team_bar_sync(kBarA, kThreads);                // ---- the 12 (Ea, dE) of this step are in shared memory, in event order
    // an event group that found an error has filed nothing (its slot holds the previous step's pair): the flag is loaded
    // here and looked at just before the jump is written, so that the load does not sit in front of the rate chain
    const int step_err = *static_cast<volatile int *>(&s_err);
    LMC_TEAM_TICK(1);                              // wait for the events
    // CalculateTime + SelectEvent (KineticMcFirstOmp.cpp:55-77, KineticMcAbstract.cpp:106-116): lane q < 12 owns slot q
    const double my_de = s_de[ql];
    double my_ea, rate;                            // 12 closed forms and rates in one instruction stream (lanes >= 12 shadow slot 0)
    barrier_and_rate_chain(my_de, s_le[ql], barrier_model, beta, my_ea, rate);
    const int my_dir = s_dir[ql];
    const unsigned my_mig = s_mig[ql];
    // the neighbour site of the jump in slot q, while the exponential is in flight
    const int xs = __shfl_sync(full, xq, my_dir), ys = __shfl_sync(full, yq, my_dir), zs = __shfl_sync(full, zq, my_dir);
    if (lane < 12) s_p[lane] = rate;
    __syncwarp(full);
    LMC_TEAM_TICK(4);                              // rates
    const double2 *p2 = reinterpret_cast<const double2 *>(s_p);
    double r_ord[12];                              // the 12 rates in slot order
#pragma unroll
    for (int q = 0; q < 6; ++q) { const double2 v = p2[q]; r_ord[2 * q] = v.x; r_ord[2 * q + 1] = v.y; }
    // one-pass select (select_event_fast, kmc_kernels.cuh): the same slot as the reference's total / division / running sum
    // unless u2 lies within rounding distance of a boundary -- then (and only then) the sequential form runs
    bool below;
    double total = 0.0;
    bool have_total = false;
    const bool sure = select_event_fast<0>(r_ord, lane, u2, prm.select_margin, below);
    const unsigned unsure = __ballot_sync(full, lane < 12 && !sure);           // the two votes go out back to back
    unsigned hit = __ballot_sync(full, lane < 12 && !below) & 0xFFFu;
    if (unsure) {
      below = select_event_sequential(s_p, lane, lane < 12, u2, &total);
      have_total = true;
      hit = __ballot_sync(full, lane < 12 && !below) & 0xFFFu;
    }
    LMC_TEAM_TICK(7);                              // select
    if (step_err != 0) {
      failed = true;                               // the walker stops; its state is left as it was before this step
      if (lane == 0) *reinterpret_cast<int4 *>(s_sel) = make_int4(X, Y, Z, 1);
      __syncwarp(full);
      team_bar_arrive(kBarB, kThreads);
      break;
    }
    const int sel_slot = hit ? (__ffs(static_cast<int>(hit)) - 1) : 11;
    const unsigned sel_mig = __shfl_sync(full, my_mig, sel_slot);
    const int nx = __shfl_sync(full, xs, sel_slot), ny = __shfl_sync(full, ys, sel_slot), nz = __shfl_sync(full, zs, sel_slot);
    // Config::LatticeJump: lanes 0-7 write the images of the old vacancy site, lanes 8-15 those of the new one
    if (lane < 16) {
      if (kSmemOcc) store_site_image(lat, s_occ, lane < 8 ? X : nx, lane < 8 ? Y : ny, lane < 8 ? Z : nz, lane & 7, static_cast<uint8_t>(lane < 8 ? sel_mig : vac_code));
      store_site_image(lat, o, lane < 8 ? X : nx, lane < 8 ? Y : ny, lane < 8 ? Z : nz, lane & 7, static_cast<uint8_t>(lane < 8 ? sel_mig : vac_code));
    }
    if (lane == 0) *reinterpret_cast<int4 *>(s_sel) = make_int4(nx, ny, nz, 0);
    __syncwarp(full);
    team_bar_arrive(kBarB, kThreads);              // ---- the jump is visible: the event threads start the next step
    LMC_TEAM_TICK(2);                              // select + jump
    const double sel_de = __shfl_sync(full, my_de, sel_slot), sel_ea = __shfl_sync(full, my_ea, sel_slot);
    // ---- off the critical path: total rate in slot order (sequential, KineticMcFirstOmp.cpp:55-77), residence time, clock, energy, trace
    if (!have_total) {                             // s_p still holds the rates (re-read: nothing stays live across the select)
      double2 v[6];
#pragma unroll
      for (int q = 0; q < 6; ++q) v[q] = p2[q];
#pragma unroll
      for (int q = 0; q < 6; ++q) { total += v[q].x; total += v[q].y; }
    }
    const double dt = __dmul_rn(neg_log_u1 / total, corr_over_prefactor);
    if (kInstrumented && tracing && lane == 0) {
      const int64_t at = static_cast<int64_t>(w) * n_steps + s;
      if (tr.from) tr.from[at] = lat.id_of_coords(X, Y, Z);
      if (tr.to) tr.to[at] = lat.id_of_coords(nx, ny, nz);
      if (tr.slot) tr.slot[at] = sel_slot;
      if (tr.dt) tr.dt[at] = dt;
      if (tr.Ea) tr.Ea[at] = sel_ea;
      if (tr.dE) tr.dE[at] = sel_de;
      if (tr.total_rate) tr.total_rate[at] = total;
      if (tr.temperature) tr.temperature[at] = temperature;
    }
    time += dt;
    energy += sel_de;
    ++steps;
    X = nx; Y = ny; Z = nz;
    LMC_TEAM_TICK(3);                              // clock
  }
#ifdef LMC_KMC_TEAM_PROFILE
  if (blockIdx.x == 0 && lane == 0)
    printf("team profile G=%d selector, cycles per step: preparation %lld wait for events %lld rates %lld select %lld pick + jump %lld clock %lld\n", G, prof[0] / n_steps,
           prof[1] / n_steps, prof[4] / n_steps, prof[7] / n_steps, prof[2] / n_steps, prof[3] / n_steps);
#endif
  if (lane == 0) {
    const int err = *static_cast<volatile int *>(&s_err);
    if (err) atomicOr(&st.error[w], err);
    else if (!failed && st.error[w] == 0) {
      st.vacancy[w] = lat.id_of_coords(X, Y, Z);
      st.time[w] = time;
      st.energy[w] = energy;
      st.steps[w] = steps;
      st.temperature[w] = temperature;
      st.previous[w] = -1;        // a first-order run leaves no second-order history
    }
  }
}

}  // namespace lmc

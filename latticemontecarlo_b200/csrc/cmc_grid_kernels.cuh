// cmc_grid_kernels.cuh -- canonical MC / simulated annealing on ONE large lattice with the whole GPU, and with several
// GPUs (BASELINE configs[3]: SimulatedAnnealing factor=100, 4M sites; SURVEY 8(e) "large-lattice CMC / SA").
//
// The cluster kernel (cmc_kernels.cuh) gives one replica at most 16 SMs.  A lattice of N sites admits about N / 172
// mutually non-interfering trials per batch (2 sites x 43-site neighbourhoods x 2), which for >= 1M sites is more than
// 16 SMs can evaluate at once.  Here the batch is spread over a persistent cooperative GRID (one CTA per SM); the phases
// are those of the cluster kernel, the cluster barriers become grid barriers and the DSMEM exchange becomes a small
// global-memory exchange.
//
// Several GPUs (one process per GPU): every rank holds the WHOLE occupancy (a 4M-site lattice is 4 MB) and draws the
// same Philox proposals, so proposals, claims (marks) and the batch composition are identical everywhere without
// communication; only the expensive part -- the two 43-site gathers + table walks per trial -- is sharded (warp w of
// every CTA is evaluated by rank w mod world).  Owners write the kept / accept bit masks and their partial dE sums
// straight into every peer's exchange buffer through NVLink peer mappings, fence, and raise a per-rank flag; every rank
// then applies ALL accepted swaps to its own replica.  The exchange is a few KB per batch and sits inside the persistent
// kernel (no host round trip, no NCCL call on the data path).  Because all decisions are keyed by the trial's position
// in the batch, the trajectory is bit-identical for every world size.
#pragma once
#include "cmc_kernels.cuh"

namespace lmc {

constexpr int kGridMaxWorld = 8;
constexpr int kGridMaxCtas = 160;                       // >= SM count of a B200 (148)
constexpr int kGridMaxWarps = kCmcMaxThreads / 32;      // warps (= groups of 16 trials) per CTA
constexpr int kGridPartialDoubles = 4;                  // sum, kept, accepted, error (as doubles: one 32-byte store)

// Exchange buffer of one rank (device memory, IPC-shared with the peers).  Everything travels as 16-byte "lines" in the
// style of NCCL's LL protocol: two 32-bit payload words, each followed by a 32-bit flag (the low half of the batch sequence
// number), written with ONE vector store.  8-byte halves arrive atomically, so a reader that sees both flags equal to the
// sequence number it waits for has the payload: no fence, no separate flag, one NVLink traversal per message.  Lines are
// double buffered by batch parity (a peer can be at most one batch ahead).
struct CmcLine { unsigned int d0, f0, d1, f1; };
struct CmcExchange {
  CmcLine masks[2][kGridMaxCtas][kGridMaxWarps];                       // [parity][cta][warp] {kept mask, accept mask}
  CmcLine partials[2][kGridMaxWorld][kGridMaxCtas][2];                 // [parity][rank][cta] {sum lo, sum hi}, {kept, accepted | err << 31}
};

__device__ __forceinline__ void line_store(CmcLine *dst, unsigned d0, unsigned d1, unsigned flag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "r"(d0), "r"(flag), "r"(d1), "r"(flag) : "memory");
}
// optimistic read at L2 (where local and peer writes land); not ordered against other loads, so many can be in flight.
// A line whose flags do not match yet is re-read by line_wait.
__device__ __forceinline__ uint4 line_load(const CmcLine *src) {      // {d0, f0, d1, f1}
  return __ldcg(reinterpret_cast<const uint4 *>(src));
}
// spin until both flags of the line equal `flag`; false on timeout / abort
__device__ __forceinline__ bool line_wait(const CmcLine *src, unsigned flag, unsigned &d0, unsigned &d1, int *abort_flag, long long spin_limit) {
  const long long t0 = clock64();
  for (;;) {
    unsigned a, fa, b, fb;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(fa), "=r"(b), "=r"(fb) : "l"(src) : "memory");
    if (fa == flag && fb == flag) { d0 = a; d1 = b; return true; }
    if (*reinterpret_cast<volatile int *>(abort_flag)) return false;
    if (clock64() - t0 > spin_limit) { *reinterpret_cast<volatile int *>(abort_flag) = 1; return false; }
  }
}


struct CmcGridParams {
  int world, rank;
  CmcExchange *xchg[kGridMaxWorld];       // xchg[rank] is the local buffer; the others are peer mappings
  unsigned long long *barrier_counter;    // local: monotonically increasing arrival counter (zeroed before the launch)
  int *abort_flag;                        // local: set when a spin loop times out (a peer died): every CTA leaves
  unsigned long long *sequence;           // local: exchange sequence number, never reset (flags compare against it)
  unsigned long long *accum;              // local: [2][4] batch totals by parity {sum dE (fixed point), kept, accepted, errors}
  long long spin_limit;                   // clock64 ticks a spin loop may wait
};

// grid-wide barrier over co-resident CTAs (cooperative launch).  `target` counts arrivals expected so far.
__device__ __forceinline__ bool grid_barrier(const CmcGridParams &gp, unsigned long long &target, unsigned n_cta) {
  __syncthreads();
  target += n_cta;
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    // arrive with release semantics (cumulative over the CTA's writes through the bar.sync above), wait with acquire loads:
    // one fence less on each side than __threadfence + atomicAdd + volatile spin + __threadfence
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(gp.barrier_counter) : "memory");
    const long long t0 = clock64();
    int ok = 1;
    for (;;) {
      unsigned long long seen;
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(gp.barrier_counter) : "memory");
      if (seen >= target) break;
      if (*reinterpret_cast<volatile int *>(gp.abort_flag)) { ok = 0; break; }
      if (clock64() - t0 > gp.spin_limit) { *reinterpret_cast<volatile int *>(gp.abort_flag) = 1; ok = 0; break; }
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

__device__ __forceinline__ void sa_update_batch(SaSchedule &sa, unsigned int n_kept, unsigned int n_acc, double energy, unsigned long long steps,
                                                double cool) {
  // batch-granular SimulatedAnnealing::UpdateTemperature (SimulatedAnnealing.cpp:99-139): see cmc_run_kernel
  sa.window_trials += n_kept;
  sa.window_accepts += n_acc;
  if (n_acc > 0 && energy < sa.recent_best_energy - kSaEpsilon) { sa.recent_best_energy = energy; sa.last_improvement_step = steps; }
  if (sa.window_trials >= sa.window_size) {
    if (static_cast<double>(sa.window_accepts) / static_cast<double>(sa.window_trials) > 0.50) sa.temperature *= 0.99;
    sa.window_trials = 0; sa.window_accepts = 0;
  }
  const double acc_est = sa.window_trials > 0u ? static_cast<double>(sa.window_accepts) / static_cast<double>(sa.window_trials) : 1.0;
  if (sa.reheats_done < 5u && (steps - sa.last_improvement_step >= sa.reheat_trigger_steps) &&
      (steps - sa.last_reheat_step >= sa.reheat_cooldown_steps) && acc_est < 0.05) {
    sa.temperature *= 1.10; sa.last_improvement_step = steps; sa.last_reheat_step = steps;
    sa.recent_best_energy = energy; ++sa.reheats_done;
  }
  sa.temperature *= pow(cool, static_cast<double>(n_kept));
}

// ---------------------------------------------------------------------------------------------------------------------
// Lane-GROUP evaluation of one swap trial.  The reference's non-interference rule (CanonicalMcOmp.cpp:47-72) caps a batch at
// about N / 172 trials, so a 40^3 lattice keeps an SM busy with only ~7 kept trials per batch: with one lane per trial
// side the evaluation is a single-warp latency chain whose length is set by the most solute-rich neighbourhood of the
// whole batch.  Here a trial owns 2 L lanes (L per side): the 43 cell loads of a side are dealt round-robin to its L lanes,
// the solute mask and the highest claim mark are combined with warp REDUX over the group, conflicting trials leave before
// the table walk, and the walk is split by neighbour position (lane `sub` owns the solutes t with t mod L == sub and their
// partner lists).  Partial sums meet in a fixed XOR tree, so the result depends on nothing but the trial itself.
// Tables: s_pidx[t * 42 + u] (u > t) is the row of the (t, u) site pair in the B table (mask_hi / base folded into one byte).
#ifdef LMC_CMC_PROFILE
__device__ long long g_group_prof[4];     // thread 0 of CTA 0: cycles in {cell loads, mask / mark reduction, walk + tree}, calls
#define LMC_GROUP_TICK(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long t_now = clock64(); g_group_prof[k] += t_now - t_prev; t_prev = t_now; } } while (0)
#else
#define LMC_GROUP_TICK(k) do { } while (0)
#endif
template <int L, bool kStagedB>
__device__ __forceinline__ double swap_energy_change_group(const LatticeDesc &lat, const DevTables &tab, const double *s_C, const double *s_A,
                                                           const double *Bt, const uint64_t *s_mask, const uint8_t *s_pidx,
                                                           const unsigned int *cells, const int32_t *__restrict__ s_delta, uint8_t *row,
                                                           int side, int sub, unsigned trial_mask, unsigned side_mask, int xa, int ya, int za,
                                                           int xb, int yb, int zb, unsigned my_mark, bool *conflict) {
#ifdef LMC_CMC_PROFILE
  long long t_prev = clock64();
#endif
  const unsigned solvent = static_cast<unsigned>(tab.solvent), vac = static_cast<unsigned>(tab.n_species);
  int64_t base_a = lat.padded_index(xa, ya, za), base_b = lat.padded_index(xb, yb, zb);
  unsigned ea = cells[base_a] & kCellSpeciesMask, eb = cells[base_b] & kCellSpeciesMask;
  int dx = xb - xa, dy = yb - ya, dz = zb - za;
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  dx = dx > px / 2 ? dx - px : (dx < -px / 2 ? dx + px : dx);
  dy = dy > py / 2 ? dy - py : (dy < -py / 2 ? dy + py : dy);
  dz = dz > pz / 2 ? dz - pz : (dz < -pz / 2 ? dz + pz : dz);
  const bool coupled = dx * dx + dy * dy + dz * dz <= 6;
  int zpa = za & 1, zpb = zb & 1;
  if (coupled && eb == vac) {   // move the vacancy first so that no intermediate state holds two vacancies (swap_energy_change)
    const int64_t tb = base_a; base_a = base_b; base_b = tb;
    const unsigned te = ea; ea = eb; eb = te;
    const int tz = zpa; zpa = zpb; zpb = tz;
    dx = -dx; dy = -dy; dz = -dz;
  }
  const int64_t base = side ? base_b : base_a;
  const int32_t *drow = s_delta + (side ? zpb : zpa) * 43;
  // side 1 of a coupled pair sees the first site already changed (its halo images are not updated: override by position)
  const int64_t override_index = (side && coupled) ? base_b + lat.padded_delta(-dx, -dy, -dz, zpb) : -1;
  constexpr int kPerLane = (43 + L - 1) / L;
  unsigned cell[kPerLane];
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    const int t = sub + i * L;
    cell[i] = t < 43 ? cells[base + drow[t]] : 0u;            // plain loads (L1 was invalidated by the barrier's acquire)
  }
  unsigned worst = 0, lo = 0, hi = 0;
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    const int t = sub + i * L;
    if (t < 43 && t != kCentrePos) {
      const unsigned mk = cell[i] >> 8;
      worst = mk > worst ? mk : worst;
      unsigned c = cell[i] & kCellSpeciesMask;
      if (base + drow[t] == override_index) c = eb;
      const int e = t - (t > kCentrePos);
      row[e] = static_cast<uint8_t>(c);
      const unsigned bit = (c != solvent) ? 1u : 0u;
      if (e < 32) lo |= bit << e; else hi |= bit << (e - 32);
    } else if (t == kCentrePos) {
      const unsigned mk = cell[i] >> 8;
      worst = mk > worst ? mk : worst;
    }
  }
  LMC_GROUP_TICK(0);
  // marks of older epochs are numerically smaller than any mark of the current epoch
  worst = __reduce_max_sync(trial_mask, worst);
  if (L > 1) {
    __syncwarp(side_mask);                                  // row[] is complete for every lane of the side
    lo = __reduce_or_sync(side_mask, lo);
    hi = __reduce_or_sync(side_mask, hi);
  }
  LMC_GROUP_TICK(1);
  if (worst > my_mark) { *conflict = true; return 0.0; }
  if (ea == eb) return 0.0;
  const int m = tab.n_species + 1, mm = m * m;
  const int x_old = static_cast<int>(side ? eb : ea), x_new = static_cast<int>(side ? ea : eb);
  const int a_stride = kSiteEnvN * m, b_stride = tab.n_site_pairs * mm;
  const double *A_new = s_A + x_new * a_stride, *A_old = s_A + x_old * a_stride;
  const double *B_new = Bt + static_cast<size_t>(x_new) * b_stride, *B_old = Bt + static_cast<size_t>(x_old) * b_stride;
  double acc = sub == 0 ? s_C[x_new] - s_C[x_old] : 0.0;
  constexpr unsigned kStripe = L >= 32 ? 1u : 0xFFFFFFFFu / ((1u << (L % 32)) - 1u);   // bits 0, L, 2L, ...
  auto walk = [&](unsigned mine, int t0) {
    while (mine) {
      const int t = t0 + __ffs(static_cast<int>(mine)) - 1;
      mine &= mine - 1;
      const int et = row[t];
      acc += A_new[t * m + et] - A_old[t * m + et];
      const uint64_t mask = s_mask[t];
      unsigned plo = static_cast<unsigned>(mask) & lo, phi = static_cast<unsigned>(mask >> 32) & hi;
      const uint8_t *prow = s_pidx + t * kSiteEnvN;
      const int col = et * m;
      while (plo) {
        const int u = __ffs(static_cast<int>(plo)) - 1;
        plo &= plo - 1;
        const int p = prow[u] * mm + col + row[u];
        acc += kStagedB ? B_new[p] - B_old[p] : __ldg(B_new + p) - __ldg(B_old + p);
      }
      while (phi) {
        const int u = 32 + __ffs(static_cast<int>(phi)) - 1;
        phi &= phi - 1;
        const int p = prow[u] * mm + col + row[u];
        acc += kStagedB ? B_new[p] - B_old[p] : __ldg(B_new + p) - __ldg(B_old + p);
      }
    }
  };
  walk(lo & (kStripe << sub), 0);
  walk(hi & (kStripe << sub), 32);          // 32 is a multiple of L: the stripe continues unbroken
#pragma unroll
  for (int off = L / 2; off > 0; off >>= 1) acc += __shfl_xor_sync(trial_mask, acc, off);
  LMC_GROUP_TICK(2);
#ifdef LMC_CMC_PROFILE
  if (blockIdx.x == 0 && threadIdx.x == 0) g_group_prof[3] += 1;
#endif
  return acc;
}

template <int L, bool kStagedB>
__global__ void __launch_bounds__(kCmcMaxThreads)
cmc_grid_kernel(LatticeDesc lat, DevTables tab, uint8_t *o, uint8_t *by_id, unsigned int *cells, CmcState st, const double *__restrict__ temperatures,
                uint64_t seed, unsigned long long target_steps, CmcGridParams gp) {
  const int n_cta = static_cast<int>(gridDim.x), cta = static_cast<int>(blockIdx.x);
  __shared__ int32_t s_delta[2 * 43];
  __shared__ long long s_warp_fixed[kGridMaxWarps];
  __shared__ unsigned int s_warp_cnt[kGridMaxWarps], s_warp_acc[kGridMaxWarps], s_warp_live[kGridMaxWarps];
  __shared__ double s_energy, s_temperature;
  __shared__ unsigned long long s_steps, s_accepted, s_proposals, s_epoch, s_sequence;
  __shared__ SaSchedule s_sa;
  __shared__ int32_t s_live_a[kCmcMaxThreads], s_live_b[kCmcMaxThreads];
  __shared__ uint16_t s_live_sp[kCmcMaxThreads];   // species of the two sites at batch start (a | b << 8), from the proposal's mirror reads
  __shared__ int s_flag_ok;
  __shared__ long long s_part_e[kGridMaxWorld];
  __shared__ unsigned int s_part_k[kGridMaxWorld], s_part_a[kGridMaxWorld];
  extern __shared__ double s_dyn[];                // [C: m] [A: m*42*m] [B: m*204*m*m, optional] [mask: 42] [base] [pidx: 42*42] [codes]

  const int tid = threadIdx.x, B = blockDim.x;
  const int warp = tid >> 5, n_warps = B >> 5;
  const int stage_b_table = kStagedB ? 1 : 0;
  for (int q = tid; q < 2 * 43; q += B) s_delta[q] = tab.site_delta[q];
  const int m = tab.n_species + 1;
  double *s_C = s_dyn;
  double *s_A = s_C + m;
  const int a_len = m * kSiteEnvN * m, b_len = m * tab.n_site_pairs * m * m;
  double *s_B = s_A + a_len;
  uint64_t *s_mask = reinterpret_cast<uint64_t *>(s_B + (stage_b_table ? b_len : 0));
  uint16_t *s_base = reinterpret_cast<uint16_t *>(s_mask + kSiteEnvN);
  uint8_t *s_pidx = reinterpret_cast<uint8_t *>(s_base + 44);            // [42][42]: B-table row of the site pair (t, u), u > t
  uint8_t *s_codes = s_pidx + kSiteEnvN * kSiteEnvN + 4;                  // L == 1: one column per thread; L > 1: one 48-byte row per trial side
  for (int q = tid; q < m; q += B) s_C[q] = tab.site_C[q];
  for (int q = tid; q < a_len; q += B) s_A[q] = tab.site_A[q];
  if (stage_b_table)
    for (int q = tid; q < b_len; q += B) s_B[q] = tab.site_B[q];
  for (int q = tid; q < kSiteEnvN; q += B) { s_mask[q] = tab.site_mask_hi[q]; s_base[q] = tab.site_base[q]; }
  for (int q = tid; q < kSiteEnvN * kSiteEnvN; q += B) {
    const int t = q / kSiteEnvN, u = q % kSiteEnvN;
    const uint64_t hi = tab.site_mask_hi[t];
    s_pidx[q] = static_cast<uint8_t>(tab.site_base[t] + __popcll(hi & ((1ULL << u) - 1ULL)));   // meaningful where bit u of hi is set
  }
  const SiteTablesView tv{s_A, stage_b_table ? s_B : tab.site_B, s_C, s_mask, s_base, m, tab.n_site_pairs};
  if (tid == 0) {
    s_energy = st.energy[0]; s_steps = st.steps[0]; s_accepted = st.accepted[0]; s_proposals = st.proposals[0];
    s_epoch = st.epoch[0]; s_sa = st.sa[0]; s_sequence = *gp.sequence;
    s_temperature = s_sa.enabled ? s_sa.temperature : temperatures[0];
  }
  __syncthreads();
  const uint32_t n_sites = static_cast<uint32_t>(lat.num_sites);
  const double cool = s_sa.enabled ? exp(-3.0 / static_cast<double>(s_sa.maximum_steps > 0 ? s_sa.maximum_steps : 1ULL)) : 1.0;
  const int gtid = cta * B + tid;
  constexpr int G = 2 * L;                         // lanes per trial (L per side)
  const int half = B / G;                          // trials a CTA can evaluate per batch
  const int n_prop = 2 * half;                     // proposals per CTA and batch: twice the capacity (about half survive the draw / claim)
  constexpr int S = L >= 4 ? 4 : L;                // lanes that share the kCmcDraws draws of one proposal (n_prop * S <= B)
  constexpr int kMyDraws = kCmcDraws / S;          // consecutive draws of this lane (one Philox call yields two draws)
  const int prop_id = tid / S, prop_lane = tid % S;
  const bool proposer = prop_id < n_prop;
  const int gprop = cta * n_prop + prop_id;
  const int window = n_prop * n_cta;               // proposals per batch
  const int pair_id = tid / G, side = (tid / L) & 1, sub = tid % L;
  const unsigned trial_mask = (G >= 32 ? 0xFFFFFFFFu : ((1u << G) - 1u)) << ((tid & 31) / G * G);
  const unsigned side_mask = ((1u << L) - 1u) << ((tid & 31) / L * L);
  constexpr unsigned kLeaders = G >= 32 ? 1u : 0xFFFFFFFFu / ((1u << (G % 32)) - 1u);      // lane 0 of every trial of a warp
  const int world = gp.world, rank = gp.rank;
  CmcExchange *mine = gp.xchg[rank];
  unsigned long long bar_target = 0;
  int err = 0;
  bool healthy = true;
  // lattice ids of this lane's draws of the coming batch (state-independent: drawn while the previous batch's barrier is pending)
  uint32_t ida[kMyDraws], idb[kMyDraws];
  auto draw_ids = [&](unsigned long long prop_base) {
    if (!proposer) return;
#pragma unroll
    for (int d = 0; d < kMyDraws; d += 2) {
      uint32_t r[4];
      const unsigned long long g = (prop_base + gprop) * (kCmcDraws / 2) + ((prop_lane * kMyDraws + d) >> 1);
      philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
      ida[d] = __umulhi(r[0], n_sites); idb[d] = __umulhi(r[1], n_sites);
      ida[d + 1] = __umulhi(r[2], n_sites); idb[d + 1] = __umulhi(r[3], n_sites);
    }
  };
#pragma unroll
  for (int d = 0; d < kMyDraws; ++d) { ida[d] = 0; idb[d] = 0; }
  draw_ids(s_proposals);
#ifdef LMC_CMC_PROFILE
  __shared__ long long s_gprof[8], s_eprof, s_rprof[3];
  if (tid == 0) { for (int q = 0; q < 8; ++q) s_gprof[q] = 0; s_eprof = 0; s_rprof[0] = s_rprof[1] = s_rprof[2] = 0; }
  long long tg_prev = clock64();
  unsigned long long n_batches = 0;
#define LMC_GTICK(k) do { if (tid == 0) { const long long t_now = clock64(); s_gprof[k] += t_now - tg_prev; tg_prev = t_now; } } while (0)
#else
#define LMC_GTICK(k) do { } while (0)
#endif

  for (;;) {
    const unsigned long long steps0 = s_steps, epoch = s_epoch + 1, prop0 = s_proposals, seq = s_sequence + 1;
    const double t_batch = s_temperature, energy0 = s_energy;
    const int parity = static_cast<int>(seq & 1ULL);
    const unsigned flag = static_cast<unsigned>(seq);            // never 0: the buffers start zeroed, seq starts at 1
    if (steps0 >= target_steps) break;
    __syncthreads();
    if ((epoch & 0xFFULL) == 0) {                   // 8-bit epoch wrapped: forget all marks (keep the species bytes)
      for (int64_t q = gtid; q < lat.padded_size; q += static_cast<int64_t>(B) * n_cta) cells[q] &= kCellSpeciesMask;
      if (!grid_barrier(gp, bar_target, n_cta)) { healthy = false; break; }
    }
    const unsigned int epoch8 = static_cast<unsigned int>(epoch & 0xFFULL);
    // ---------------- proposals: identical on every rank (same counters), first unlike-species pair of 8 draws.  The lattice
    // ids of this batch were drawn (Philox) before the previous batch's closing barrier; S lanes share a proposal.
    int32_t a = -1, b = -1;
    unsigned sp = 0;
    {
      uint8_t sa_[kMyDraws], sb_[kMyDraws];
#pragma unroll
      for (int d = 0; d < kMyDraws; ++d) {
        sa_[d] = proposer ? __ldcg(by_id + ida[d]) : 0;
        sb_[d] = proposer ? __ldcg(by_id + idb[d]) : 0;
      }
      int first = kCmcDraws;                        // index (over the proposal's 8 draws) of this lane's first unlike pair
#pragma unroll
      for (int d = kMyDraws - 1; d >= 0; --d)
        if (proposer && sa_[d] != sb_[d]) { first = prop_lane * kMyDraws + d; a = static_cast<int32_t>(ida[d]); b = static_cast<int32_t>(idb[d]); sp = sa_[d] | (sb_[d] << 8); }
      int best = first;
#pragma unroll
      for (int off = S / 2; off > 0; off >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, off));
      const int src = ((tid & 31) / S) * S + (best < kCmcDraws ? best / kMyDraws : 0);
      a = __shfl_sync(0xffffffffu, a, src);
      b = __shfl_sync(0xffffffffu, b, src);
      sp = __shfl_sync(0xffffffffu, sp, src);
      if (best >= kCmcDraws || prop_lane != 0) a = -1;         // one lane per proposal enters the compaction
    }
    LMC_GTICK(0);
    // ---------------- compaction of this CTA's live trials
    const bool has = a >= 0;
    int n_live = 0;
    {
      const unsigned bal = __ballot_sync(0xffffffffu, has);
      if ((tid & 31) == 0) s_warp_live[warp] = __popc(bal);
      __syncthreads();
      int my_off = 0;
      for (int q = 0; q < n_warps; ++q) {
        if (q == warp) my_off = n_live;
        n_live += s_warp_live[q];
      }
      if (has) {
        const int slot = my_off + __popc(bal & ((1u << (tid & 31)) - 1u));
        s_live_a[slot] = a;
        s_live_b[slot] = b;
        s_live_sp[slot] = static_cast<uint16_t>(sp);
      }
    }
    __syncthreads();
    if (n_live > half) n_live = half;               // surplus proposals are dropped (redrawn in a later batch)
    LMC_GTICK(1);
    // ---------------- claims: every rank marks ALL live trials (lane pair i = trial i, one site per lane)
    bool live = pair_id < n_live;
    int xa = 0, ya = 0, za = 0, xb = 0, yb = 0, zb = 0;
    const unsigned int gpair = static_cast<unsigned int>(cta * half + pair_id);   // position of the trial in the batch = priority
    const unsigned int my_mark = (epoch8 << 16) | (0xFFFFu - gpair);      // 24 bits
    if (live) {
      a = s_live_a[pair_id]; b = s_live_b[pair_id];
      sp = s_live_sp[pair_id];
      lat.coords_of_id(a, xa, ya, za);
      lat.coords_of_id(b, xb, yb, zb);
      if (sub == 0) {
        const int mx = side ? xb : xa, my = side ? yb : ya, mz = side ? zb : za;
        mark_site(lat, cells, mx, my, mz, (my_mark << 8) | (side ? sp >> 8 : sp & 0xFFu));
      }
    }
    if (!grid_barrier(gp, bar_target, n_cta)) { healthy = false; break; }
    LMC_GTICK(2);
    // ---------------- evaluation: warp w belongs to rank (w mod world)
    const bool owner = (warp % world) == rank;
    bool conflict = false, same = false;
    double de = 0.0;
#ifdef LMC_CMC_PROFILE
    const long long te0 = clock64();
#endif
    if (live && owner) {
      if constexpr (L == 1)
        de = swap_side_energy_change_marked(lat, tab, tv, cells, s_delta, s_codes + tid, B, side, xa, ya, za, xb, yb, zb, my_mark, &conflict, &same);
      else
        de = swap_energy_change_group<L, kStagedB>(lat, tab, s_C, s_A, kStagedB ? s_B : tab.site_B, s_mask, s_pidx, cells, s_delta,
                                                   s_codes + (tid / L) * 48, side, sub, trial_mask, side_mask, xa, ya, za, xb, yb, zb, my_mark,
                                                   &conflict);
    }
#ifdef LMC_CMC_PROFILE
    const long long te1 = clock64();
#endif
    conflict = __shfl_xor_sync(0xffffffffu, conflict ? 1 : 0, L) || conflict;     // the other side of the trial
    de += __shfl_xor_sync(0xffffffffu, de, L);
    bool kept = live && owner && !conflict;
    if (kept && de != de) err |= kErrExtraVacancy;
    bool accept = false;
    if (kept) {
      accept = de < 0.0;                            // CanonicalMcAbstract::SelectEvent (:86-101)
      if (!accept) {
        uint32_t r2[4];
        const unsigned long long g = prop0 + gpair;
        philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(seed),
                      static_cast<uint32_t>(seed >> 32) ^ 0x9E3779B9u, r2);
        const double beta = 1.0 / kBoltzmannEv / fmax(t_batch, 1e-12);
        accept = uniform53(r2[0], r2[1]) < exp(-de * beta);
      }
    }
    unsigned kept_mask = __ballot_sync(0xffffffffu, kept), acc_mask = __ballot_sync(0xffffffffu, accept);
    {
      double sum = (accept && (tid % G) == 0) ? de : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
      if ((tid & 31) == 0) {
        s_warp_fixed[warp] = __double2ll_rn(sum * kEnergyFixedScale);   // rounded per warp group = per unit of ownership:
                                                                        // the integer total is the same for every world size
        s_warp_cnt[warp] = __popc(kept_mask & kLeaders);         // one lane per trial
        s_warp_acc[warp] = __popc(acc_mask & kLeaders);
      }
      // the owner warp publishes its group's masks in every peer's buffer right away (lane d -> rank d)
      if (world > 1 && owner && (tid & 31) < world && (tid & 31) != rank)
        line_store(&gp.xchg[tid & 31]->masks[parity][cta][warp], kept_mask, acc_mask, flag);
    }
#ifdef LMC_CMC_PROFILE
    const long long te2 = clock64();
    if (tid == 0) { s_gprof[7] += te1 - te0; s_eprof += te2 - te1; }
#endif
    const int block_err = __syncthreads_or(err != 0);
    LMC_GTICK(3);
    if (warp == 0) {
      // this CTA's partial sums (own warps only; integer adds: any order).  Thread d ships them to rank d as two lines.
      long long e = tid < n_warps ? s_warp_fixed[tid] : 0LL;
      unsigned int n_kept = tid < n_warps ? s_warp_cnt[tid] : 0u, n_acc = tid < n_warps ? s_warp_acc[tid] : 0u;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, off);
        n_kept += __shfl_xor_sync(0xffffffffu, n_kept, off);
        n_acc += __shfl_xor_sync(0xffffffffu, n_acc, off);
      }
      if (tid >= world) { }
      else if (tid == rank) { s_part_e[rank] = e; s_part_k[rank] = n_kept; s_part_a[rank] = n_acc | (block_err ? 0x80000000u : 0u); }
      else {
        CmcLine *dst = gp.xchg[tid]->partials[parity][rank][cta];
        const unsigned long long bits = static_cast<unsigned long long>(e);
        line_store(dst, static_cast<unsigned>(bits), static_cast<unsigned>(bits >> 32), flag);
        line_store(dst + 1, n_kept, n_acc | (block_err ? 0x80000000u : 0u), flag);
      }
    }
    {
      // a warp whose group was evaluated elsewhere waits for that rank's masks; threads d != rank of warp 0 wait for the
      // partial sums of CTA c of rank d (CTA c only ever needs the CTAs c of its peers: no grid barrier in the exchange)
      bool ok = true;
      if (world > 1) {
        if (!owner) {
          unsigned k = 0, a2 = 0;
          if ((tid & 31) == 0) ok = line_wait(&mine->masks[parity][cta][warp], flag, k, a2, gp.abort_flag, gp.spin_limit);
          kept_mask = __shfl_sync(0xffffffffu, k, 0);
          acc_mask = __shfl_sync(0xffffffffu, a2, 0);
        }
        if (tid < world && tid != rank) {
          const CmcLine *src = mine->partials[parity][tid][cta];
          unsigned lo = 0, hi = 0, nk = 0, na = 0;
          ok = line_wait(src, flag, lo, hi, gp.abort_flag, gp.spin_limit) && ok;
          ok = line_wait(src + 1, flag, nk, na, gp.abort_flag, gp.spin_limit) && ok;
          s_part_e[tid] = static_cast<long long>((static_cast<unsigned long long>(hi) << 32) | lo);
          s_part_k[tid] = nk;
          s_part_a[tid] = na;
        }
      }
      if (__syncthreads_or(!ok)) { healthy = false; break; }
      if (tid == 0) {
        // contributions of CTA c of every rank go into this rank's batch accumulators (before the closing grid barrier)
        long long fixed = 0;
        unsigned long long n_kept = 0, n_acc = 0, n_err = 0;
        for (int r = 0; r < world; ++r) {
          fixed += s_part_e[r];
          n_kept += s_part_k[r];
          n_acc += s_part_a[r] & 0x7FFFFFFFu;
          n_err += s_part_a[r] >> 31;
        }
        unsigned long long *acc = gp.accum + 4 * parity;
        if (fixed) atomicAdd(acc, static_cast<unsigned long long>(fixed));
        if (n_kept) atomicAdd(acc + 1, n_kept);
        if (n_acc) atomicAdd(acc + 2, n_acc);
        if (n_err) atomicAdd(acc + 3, n_err);
      }
    }
    LMC_GTICK(4);
    // ---------------- apply every accepted swap of this CTA's trials (all ranks alike)
    if (live && sub == 0 && ((acc_mask >> (tid & 31)) & 1u)) {
      const uint8_t ea = static_cast<uint8_t>(sp & 0xFFu), eb = static_cast<uint8_t>(sp >> 8);   // no kept trial of the batch touched them
      if (side == 0) { store_site(lat, o, xa, ya, za, eb); store_site_cells(lat, cells, xa, ya, za, eb); by_id[a] = eb; }
      else { store_site(lat, o, xb, yb, zb, ea); store_site_cells(lat, cells, xb, yb, zb, ea); by_id[b] = ea; }
    }
    draw_ids(prop0 + window);                       // the next batch's ids: hidden behind the barrier
    if (!grid_barrier(gp, bar_target, n_cta)) { healthy = false; break; }
    LMC_GTICK(5);
    // ---------------- totals: one 32-byte read per CTA (the accumulators are complete after the grid barrier)
    int any_err = 0;
    if (tid == 0) {
      const ulonglong2 *acc = reinterpret_cast<const ulonglong2 *>(gp.accum + 4 * parity);
      const ulonglong2 v0 = __ldcg(acc), v1 = __ldcg(acc + 1);
      const double e = static_cast<double>(static_cast<long long>(v0.x)) / kEnergyFixedScale;
      const unsigned int n_kept = static_cast<unsigned int>(v0.y), n_acc = static_cast<unsigned int>(v1.x);
      s_energy = energy0 + e;
      s_steps = steps0 + n_kept;
      s_accepted += n_acc;
      s_proposals = prop0 + window;
      if (s_sa.enabled && n_kept > 0) {
        SaSchedule sa = s_sa;
        sa_update_batch(sa, n_kept, n_acc, s_energy, s_steps, cool);
        s_sa = sa;
        s_temperature = sa.temperature;
      }
      s_epoch = epoch;
      s_sequence = seq;
      s_flag_ok = v1.y != 0ULL ? 0 : 1;
      // the other parity's accumulators were last read one batch ago by CTAs that have all passed two barriers since
      if (cta == 0) {
        unsigned long long *next = gp.accum + 4 * (parity ^ 1);
        next[0] = 0ULL; next[1] = 0ULL; next[2] = 0ULL; next[3] = 0ULL;
      }
    }
    __syncthreads();
    LMC_GTICK(6);
#ifdef LMC_CMC_PROFILE
    ++n_batches;
#endif
    any_err = s_flag_ok == 0;
    if (any_err) { err |= kErrExtraVacancy; break; }
  }
#ifdef LMC_CMC_PROFILE
  if (tid == 0 && cta == 0) {
    printf("group profile thread 0 (cycles/batch): cell loads %lld reductions %lld walk+tree %lld (walked %lld of %llu batches)\n", g_group_prof[0] / (long long)max(1ULL, n_batches),
           g_group_prof[1] / (long long)max(1ULL, n_batches), g_group_prof[2] / (long long)max(1LL, g_group_prof[3]), g_group_prof[3], n_batches);
    g_group_prof[0] = g_group_prof[1] = g_group_prof[2] = g_group_prof[3] = 0;
  }
  if (tid == 0 && cta == 0)
    printf("grid cmc profile rank %d (cycles/batch): propose %lld compact %lld mark+barrier %lld evaluate %lld exchange %lld apply+barrier %lld reduce %lld | "
           "thread0: dE call %lld accept+reduce %lld | reduce parts: loads %lld shuffles %lld update %lld | batches %llu ctas %d threads %d\n", rank, s_gprof[0] / (long long)max(1ULL, n_batches), s_gprof[1] / (long long)max(1ULL, n_batches),
           s_gprof[2] / (long long)max(1ULL, n_batches), s_gprof[3] / (long long)max(1ULL, n_batches), s_gprof[4] / (long long)max(1ULL, n_batches),
           s_gprof[5] / (long long)max(1ULL, n_batches), s_gprof[6] / (long long)max(1ULL, n_batches), s_gprof[7] / (long long)max(1ULL, n_batches),
           s_eprof / (long long)max(1ULL, n_batches), s_rprof[0] / (long long)max(1ULL, n_batches), s_rprof[1] / (long long)max(1ULL, n_batches),
           s_rprof[2] / (long long)max(1ULL, n_batches), n_batches, n_cta, B);
#endif
  __syncthreads();
  if (cta == 0 && tid == 0) {
    st.energy[0] = s_energy; st.steps[0] = s_steps; st.accepted[0] = s_accepted; st.proposals[0] = s_proposals;
    st.epoch[0] = s_epoch;
    if (s_sa.enabled) s_sa.temperature = s_temperature;
    st.sa[0] = s_sa;
    *gp.sequence = s_sequence;
  }
  if (!healthy) err |= kErrBadSite;                 // reported as a lost peer / barrier timeout by the host
  if (err) atomicOr(&st.error[0], err);
}

}  // namespace lmc

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

// exchange buffer of one rank (device memory, IPC-shared with the peers)
struct CmcExchange {
  unsigned long long flags[kGridMaxWorld][kGridMaxCtas];                          // flags[r][c]: last sequence number CTA c of rank r finished writing
  unsigned int masks[2][kGridMaxCtas][kGridMaxWarps][2];                         // [parity][cta][warp]{kept, accept}
  double partials[2][kGridMaxWorld][kGridMaxCtas][kGridPartialDoubles];          // [parity][rank][cta]
};

struct CmcGridParams {
  int world, rank;
  CmcExchange *xchg[kGridMaxWorld];       // xchg[rank] is the local buffer; the others are peer mappings
  unsigned long long *barrier_counter;    // local: monotonically increasing arrival counter (zeroed before the launch)
  int *abort_flag;                        // local: set when a spin loop times out (a peer died): every CTA leaves
  unsigned long long *sequence;           // local: exchange sequence number, never reset (flags compare against it)
  long long spin_limit;                   // clock64 ticks a spin loop may wait
};

// grid-wide barrier over co-resident CTAs (cooperative launch).  `target` counts arrivals expected so far.
__device__ __forceinline__ bool grid_barrier(const CmcGridParams &gp, unsigned long long &target, unsigned n_cta) {
  __syncthreads();
  target += n_cta;
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(gp.barrier_counter, 1ULL);
    const long long t0 = clock64();
    int ok = 1;
    while (*reinterpret_cast<volatile unsigned long long *>(gp.barrier_counter) < target) {
      if (*reinterpret_cast<volatile int *>(gp.abort_flag)) { ok = 0; break; }
      if (clock64() - t0 > gp.spin_limit) { *reinterpret_cast<volatile int *>(gp.abort_flag) = 1; ok = 0; break; }
    }
    __threadfence();
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

__global__ void __launch_bounds__(kCmcMaxThreads)
cmc_grid_kernel(LatticeDesc lat, DevTables tab, uint8_t *o, uint8_t *by_id, unsigned int *cells, CmcState st, const double *__restrict__ temperatures,
                uint64_t seed, unsigned long long target_steps, CmcGridParams gp, int stage_b_table) {
  const int n_cta = static_cast<int>(gridDim.x), cta = static_cast<int>(blockIdx.x);
  __shared__ int32_t s_delta[2 * 43];
  __shared__ double s_warp_sum[kGridMaxWarps];
  __shared__ unsigned int s_warp_cnt[kGridMaxWarps], s_warp_acc[kGridMaxWarps], s_warp_live[kGridMaxWarps];
  __shared__ double s_energy, s_temperature;
  __shared__ unsigned long long s_steps, s_accepted, s_proposals, s_epoch, s_sequence;
  __shared__ SaSchedule s_sa;
  __shared__ int32_t s_live_a[kCmcMaxThreads], s_live_b[kCmcMaxThreads];
  __shared__ int s_flag_ok;
  __shared__ unsigned int s_kept_mask[kGridMaxWarps], s_acc_mask[kGridMaxWarps];
  extern __shared__ double s_dyn[];                // [C: m] [A: m*42*m] [B: m*204*m*m, optional] [mask: 42] [base] [codes]

  const int tid = threadIdx.x, B = blockDim.x;
  const int warp = tid >> 5, n_warps = B >> 5;
  for (int q = tid; q < 2 * 43; q += B) s_delta[q] = tab.site_delta[q];
  const int m = tab.n_species + 1;
  double *s_C = s_dyn;
  double *s_A = s_C + m;
  const int a_len = m * kSiteEnvN * m, b_len = m * tab.n_site_pairs * m * m;
  double *s_B = s_A + a_len;
  uint64_t *s_mask = reinterpret_cast<uint64_t *>(s_B + (stage_b_table ? b_len : 0));
  uint16_t *s_base = reinterpret_cast<uint16_t *>(s_mask + kSiteEnvN);
  uint8_t *s_codes = reinterpret_cast<uint8_t *>(s_base + 44);
  for (int q = tid; q < m; q += B) s_C[q] = tab.site_C[q];
  for (int q = tid; q < a_len; q += B) s_A[q] = tab.site_A[q];
  if (stage_b_table)
    for (int q = tid; q < b_len; q += B) s_B[q] = tab.site_B[q];
  for (int q = tid; q < kSiteEnvN; q += B) { s_mask[q] = tab.site_mask_hi[q]; s_base[q] = tab.site_base[q]; }
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
  const int window = B * n_cta;                    // proposals (threads) per batch
  const int half = B / 2;
  const int pair_id = tid >> 1, side = tid & 1;
  const int world = gp.world, rank = gp.rank;
  CmcExchange *mine = gp.xchg[rank];
  unsigned long long bar_target = 0;
  int err = 0;
  bool healthy = true;
#ifdef LMC_CMC_PROFILE
  __shared__ long long s_gprof[8], s_eprof;
  if (tid == 0) { for (int q = 0; q < 8; ++q) s_gprof[q] = 0; s_eprof = 0; }
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
    if (steps0 >= target_steps) break;
    __syncthreads();
    if ((epoch & 0xFFULL) == 0) {                   // 8-bit epoch wrapped: forget all marks (keep the species bytes)
      for (int64_t q = gtid; q < lat.padded_size; q += window) cells[q] &= kCellSpeciesMask;
      if (!grid_barrier(gp, bar_target, n_cta)) { healthy = false; break; }
    }
    const unsigned int epoch8 = static_cast<unsigned int>(epoch & 0xFFULL);
    // ---------------- proposals: identical on every rank (same counters), first unlike-species pair of 8 draws
    int32_t a = -1, b = -1;
    {
      uint32_t ida[kCmcDraws], idb[kCmcDraws];
      uint8_t sa_[kCmcDraws], sb_[kCmcDraws];
#pragma unroll
      for (int d = 0; d < kCmcDraws; d += 2) {
        uint32_t r[4];
        const unsigned long long g = (prop0 + gtid) * (kCmcDraws / 2) + (d >> 1);
        philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
        ida[d] = __umulhi(r[0], n_sites); idb[d] = __umulhi(r[1], n_sites);
        ida[d + 1] = __umulhi(r[2], n_sites); idb[d + 1] = __umulhi(r[3], n_sites);
      }
#pragma unroll
      for (int d = 0; d < kCmcDraws; ++d) { sa_[d] = __ldcg(by_id + ida[d]); sb_[d] = __ldcg(by_id + idb[d]); }
#pragma unroll
      for (int d = kCmcDraws - 1; d >= 0; --d)
        if (sa_[d] != sb_[d]) { a = static_cast<int32_t>(ida[d]); b = static_cast<int32_t>(idb[d]); }
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
      lat.coords_of_id(a, xa, ya, za);
      lat.coords_of_id(b, xb, yb, zb);
      const int mx = side ? xb : xa, my = side ? yb : ya, mz = side ? zb : za;
      const unsigned species = cells[lat.padded_index(mx, my, mz)] & kCellSpeciesMask;
      mark_site(lat, cells, mx, my, mz, (my_mark << 8) | species);
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
    if (live && owner) de = swap_side_energy_change_marked(lat, tab, tv, cells, s_delta, s_codes + tid, B, side, xa, ya, za, xb, yb, zb, my_mark,
                                                           &conflict, &same);
#ifdef LMC_CMC_PROFILE
    const long long te1 = clock64();
#endif
    conflict = __shfl_xor_sync(0xffffffffu, conflict ? 1 : 0, 1) || conflict;
    de += __shfl_xor_sync(0xffffffffu, de, 1);
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
      double sum = (accept && side == 0) ? de : 0.0;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
      if ((tid & 31) == 0) {
        s_warp_sum[warp] = sum;
        s_warp_cnt[warp] = __popc(kept_mask & 0x55555555u);      // one lane per pair
        s_warp_acc[warp] = __popc(acc_mask & 0x55555555u);
      }
      if ((tid & 31) == 0) { s_kept_mask[warp] = kept_mask; s_acc_mask[warp] = acc_mask; }
    }
#ifdef LMC_CMC_PROFILE
    const long long te2 = clock64();
    if (tid == 0) { s_gprof[7] += te1 - te0; s_eprof += te2 - te1; }
#endif
    const int block_err = __syncthreads_or(err != 0);
    LMC_GTICK(3);
    if (tid < world) {
      // thread d ships this CTA's results to rank d: partial sums (own warps only; fixed order), the masks of the warp
      // groups this rank evaluated, then -- after a system-scope fence -- the CTA's flag.  CTA c of rank d only waits for
      // the CTAs c of the other ranks (it applies its own trials), so no grid barrier is needed for the exchange.
      double e = 0.0;
      unsigned int n_kept = 0, n_acc = 0;
      for (int q = 0; q < n_warps; ++q) { e += s_warp_sum[q]; n_kept += s_warp_cnt[q]; n_acc += s_warp_acc[q]; }
      CmcExchange *dst = gp.xchg[tid];
      volatile double *vd = dst->partials[parity][rank][cta];
      vd[0] = e; vd[1] = static_cast<double>(n_kept); vd[2] = static_cast<double>(n_acc); vd[3] = static_cast<double>(block_err);
      if (world > 1 && tid != rank) {
        for (int q = rank; q < n_warps; q += world) {
          volatile unsigned int *mk = dst->masks[parity][cta][q];
          mk[0] = s_kept_mask[q];
          mk[1] = s_acc_mask[q];
        }
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(&dst->flags[rank][cta]) = seq;
      }
    }
    if (world > 1) {
      if (tid == 0) s_flag_ok = 1;
      __syncthreads();
      if (tid < world && tid != rank) {
        const long long t0 = clock64();
        while (*reinterpret_cast<volatile unsigned long long *>(&mine->flags[tid][cta]) < seq) {
          if (*reinterpret_cast<volatile int *>(gp.abort_flag) || clock64() - t0 > gp.spin_limit) {
            *reinterpret_cast<volatile int *>(gp.abort_flag) = 1;
            s_flag_ok = 0;
            break;
          }
        }
        __threadfence_system();
      }
      __syncthreads();
      if (!s_flag_ok) { healthy = false; break; }
      if (!owner) {
        const volatile unsigned int *mk = mine->masks[parity][cta][warp];
        kept_mask = mk[0]; acc_mask = mk[1];
      }
    }
    LMC_GTICK(4);
    // ---------------- apply every accepted swap of this CTA's trials (all ranks alike)
    if (live && ((acc_mask >> (tid & 31)) & 1u)) {
      const int64_t base_a = lat.padded_index(xa, ya, za), base_b = lat.padded_index(xb, yb, zb);
      const uint8_t ea = __ldcg(o + base_a), eb = __ldcg(o + base_b);
      __syncwarp(__activemask());                   // both lanes have read the old species before either writes
      if (side == 0) { store_site(lat, o, xa, ya, za, eb); store_site_cells(lat, cells, xa, ya, za, eb); by_id[a] = eb; }
      else { store_site(lat, o, xb, yb, zb, ea); store_site_cells(lat, cells, xb, yb, zb, ea); by_id[b] = ea; }
    }
    if (!grid_barrier(gp, bar_target, n_cta)) { healthy = false; break; }
    LMC_GTICK(5);
    // ---------------- totals: partials of all ranks and CTAs in one fixed order (identical in every CTA of every rank)
    int any_err = 0;
    if (warp == 0) {
      double e = 0.0, k = 0.0, ac = 0.0, er = 0.0;
      const int n_part = world * n_cta;
      // L2 loads (the entries were written by other SMs / other GPUs before the barrier), eight entries in flight per lane
      for (int q0 = tid; q0 < n_part; q0 += 32 * 8) {
        double2 lo[8], hi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int q = q0 + 32 * j;
          if (q < n_part) {
            const double2 *src = reinterpret_cast<const double2 *>(mine->partials[parity][q / n_cta][q % n_cta]);
            lo[j] = __ldcg(src);
            hi[j] = __ldcg(src + 1);
          } else {
            lo[j] = make_double2(0.0, 0.0);
            hi[j] = make_double2(0.0, 0.0);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { e += lo[j].x; k += lo[j].y; ac += hi[j].x; er += hi[j].y; }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        e += __shfl_xor_sync(0xffffffffu, e, off); k += __shfl_xor_sync(0xffffffffu, k, off);
        ac += __shfl_xor_sync(0xffffffffu, ac, off); er += __shfl_xor_sync(0xffffffffu, er, off);
      }
      if (tid == 0) {
        const unsigned int n_kept = static_cast<unsigned int>(k), n_acc = static_cast<unsigned int>(ac);
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
        s_flag_ok = er != 0.0 ? 0 : 1;
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
  if (tid == 0 && cta == 0)
    printf("grid cmc profile rank %d (cycles/batch): propose %lld compact %lld mark+barrier %lld evaluate %lld exchange %lld apply+barrier %lld reduce %lld | "
           "thread0: dE call %lld accept+reduce %lld | batches %llu ctas %d threads %d\n", rank, s_gprof[0] / (long long)max(1ULL, n_batches), s_gprof[1] / (long long)max(1ULL, n_batches),
           s_gprof[2] / (long long)max(1ULL, n_batches), s_gprof[3] / (long long)max(1ULL, n_batches), s_gprof[4] / (long long)max(1ULL, n_batches),
           s_gprof[5] / (long long)max(1ULL, n_batches), s_gprof[6] / (long long)max(1ULL, n_batches), s_gprof[7] / (long long)max(1ULL, n_batches),
           s_eprof / (long long)max(1ULL, n_batches), n_batches, n_cta, B);
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

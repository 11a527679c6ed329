// cmc_kernels.cuh -- batched canonical Monte Carlo / simulated annealing driver.
// Reference: mc/src/CanonicalMcAbstract.cpp:43-51,86-101 (GenerateLatticeIdJumpPair / SelectEvent),
// mc/src/CanonicalMcOmp.cpp:40-92 (batches of mutually non-interfering trials, evaluated in parallel, applied in
// order), mc/src/CanonicalMcSerial.cpp:40-51, mc/src/SimulatedAnnealing.cpp:99-185 (schedule).
//
// One thread block owns one replica ("walker") for the whole launch and loops over batches:
//   phase 1  every thread draws a few proposals (two uniform lattice ids each; Philox4x32-10) and keeps the first
//            unlike-species pair; the live trials are compacted to the low thread ids (dense warps) and each marks its
//            two sites in a per-replica mark array (atomicMax of epoch | priority);
//   phase 2  each live trial gathers the 43-site neighbourhoods of its two sites -- species for dE and marks for the
//            conflict test in the same pass.  A trial survives if no higher-priority trial of the batch marked a site
//            of its neighbourhoods; this is CanonicalMcOmp's `unavailable_position_` rule evaluated in parallel.  No
//            surviving trial reads a site another survivor may write, so dE evaluated on the batch-start occupancy is
//            the dE the serial chain would see; survivors draw the Metropolis uniform and apply their swap at once.
// Proposals whose two sites hold the same species, and proposals that lose the claim, are not trials (the reference
// redraws them: CanonicalMcAbstract.cpp:45-50, CanonicalMcOmp.cpp:47-58), so they do not advance `steps`.
//
// Replay mode (validation): trials (a, b, u) come from the host in the reference's serial order; a batch is the
// longest prefix of the pending trials without interference, dE is evaluated in parallel and the accept decisions,
// energy and the annealing schedule are then applied strictly in order by one thread -- this reproduces
// CanonicalMcSerial / SimulatedAnnealing traces exactly.
#pragma once
#include <cstdio>
#include <cooperative_groups.h>
#include "kmc_kernels.cuh"
#include "cmc_state.h"

namespace lmc {
namespace cg = cooperative_groups;

struct CmcReplay {              // host-ordered trial stream for replica 0 (device copies)
  const int64_t *a, *b;
  const double *u;
  double *dE, *energy_before, *temperature_before;   // per-trial outputs
  uint8_t *accepted;
};


// SimulatedAnnealing::UpdateTemperature (mc/src/SimulatedAnnealing.cpp:99-139), one trial
__device__ __forceinline__ void sa_update(SaSchedule &s, bool accepted, double energy, unsigned long long step, double cool_factor) {
  ++s.window_trials;
  if (accepted) {
    ++s.window_accepts;
    if (energy < s.recent_best_energy - kSaEpsilon) {
      s.recent_best_energy = energy;
      s.last_improvement_step = step;
    }
  }
  if (s.window_trials >= s.window_size) {
    const double acc = static_cast<double>(s.window_accepts) / static_cast<double>(s.window_trials);
    if (acc > 0.50) s.temperature *= 0.99;
    s.window_trials = 0;
    s.window_accepts = 0;
  }
  const double acc_est = s.window_trials > 0u ? static_cast<double>(s.window_accepts) / static_cast<double>(s.window_trials) : 1.0;
  if (s.reheats_done < 5u && (step - s.last_improvement_step >= s.reheat_trigger_steps) &&
      (step - s.last_reheat_step >= s.reheat_cooldown_steps) && (acc_est < 0.05)) {
    s.temperature *= 1.10;
    s.last_improvement_step = step;
    s.last_reheat_step = step;
    s.recent_best_energy = energy;
    ++s.reheats_done;
  }
  s.temperature *= cool_factor;
}

// CMC cell array: one 32-bit word per padded cell = claim mark (bits 31..8: batch epoch (8) | inverted priority (16)) over the
// species code (bits 7..0), so that ONE load per neighbour delivers both the species for dE and the mark for the conflict
// test.  The byte array `occ` stays the authoritative occupancy for every other entry point; both are written on accept.
constexpr unsigned kCellSpeciesMask = 0xFFu;

// write a claim on a site and all its periodic halo images (the checkers read cells through base + offset).
// `mark` = (mark24 << 8) | species of that site: atomicMax replaces the word only if the mark is higher, and then writes
// back the same species, so no compare-and-swap loop is needed.
__device__ __forceinline__ void mark_site(const LatticeDesc &lat, unsigned int *marks, int X, int Y, int Z, unsigned int mark) {
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  atomicMax(marks + lat.padded_index(X, Y, Z), mark);
  const bool edge = X < kHalo || X >= px - kHalo || Y < kHalo || Y >= py - kHalo || Z < kHaloZ || Z >= pz - kHaloZ;
  if (!edge) return;
  for (int a = -1; a <= 1; ++a) {
    const int x = X + a * px;
    if (x < -kHalo || x >= px + kHalo) continue;
    for (int b = -1; b <= 1; ++b) {
      const int y = Y + b * py;
      if (y < -kHalo || y >= py + kHalo) continue;
      for (int c = -1; c <= 1; ++c) {
        const int z = Z + c * pz;
        if (z < -kHaloZ || z >= pz + kHaloZ) continue;
        if (a | b | c) atomicMax(marks + lat.padded_index(x, y, z), mark);
      }
    }
  }
}

// species byte of a site and all its periodic halo images in the cell array (little endian: byte 0 of the word)
__device__ __forceinline__ void store_site_cells(const LatticeDesc &lat, unsigned int *cells, int X, int Y, int Z, uint8_t code) {
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  const int64_t base = lat.padded_index(X, Y, Z);
  auto put = [&](int64_t idx) { reinterpret_cast<uint8_t *>(cells + idx)[0] = code; };
  put(base);
  const int sx = X < kHalo ? px : (X >= px - kHalo ? -px : 0);
  const int sy = Y < kHalo ? py : (Y >= py - kHalo ? -py : 0);
  const int sz = Z < kHaloZ ? pz : (Z >= pz - kHaloZ ? -pz : 0);
  if ((sx | sy | sz) == 0) return;
  const int64_t ix = static_cast<int64_t>(sx) * lat.ny * lat.nz, iy = static_cast<int64_t>(sy) * lat.nz, iz = sz / 2;
  if (sx) put(base + ix);
  if (sy) put(base + iy);
  if (sz) put(base + iz);
  if (sx && sy) put(base + ix + iy);
  if (sx && sz) put(base + ix + iz);
  if (sy && sz) put(base + iy + iz);
  if (sx && sy && sz) put(base + ix + iy + iz);
}

// Tables of the single-site energy model staged in shared memory (a CMC replica is one thread block with few warps,
// so the dependent table walk must not pay global-memory latency).  B stays in global memory if it does not fit.
struct SiteTablesView {
  const double *A, *B, *C;        // [x][42][m], [x][204][m][m], [x]
  const uint64_t *mask_hi;        // [42]
  const uint16_t *base;           // [42]
  int m, n_pairs;
};

// 43-site gather of species AND claim marks: one cell word per neighbour.  Species codes are staged in shared memory
// (codes[t * stride], one column per thread) for the table walk; returns the solute mask; *conflict is set if any site
// of the neighbourhood carries a mark of this epoch with higher priority than `my_mark` (CanonicalMcOmp.cpp:47-72: a
// trial may not touch the neighbourhood of an earlier trial of the batch).
__device__ __forceinline__ uint64_t gather_site_env_marked(const unsigned int *cells, int64_t base, const int32_t *__restrict__ drow,
                                                           unsigned solvent, uint8_t *codes, int stride, int64_t override_index,
                                                           unsigned override_code, unsigned my_mark, bool *conflict) {
  uint32_t lo = 0, hi = 0;
  unsigned worst = 0;
  // plain (L1-cached) loads: the cells were last written before the preceding cluster / grid barrier, whose acquire
  // invalidates this SM's L1; neighbouring sites share 32-byte sectors.  All 43 loads are in flight at once.
  unsigned cell[43];
#pragma unroll
  for (int t = 0; t < 43; ++t) cell[t] = cells[base + drow[t]];
#pragma unroll
  for (int t = 0; t < 43; ++t) {
    const unsigned mk = cell[t] >> 8;
    worst = mk > worst ? mk : worst;
    unsigned c = cell[t] & kCellSpeciesMask;
    if (base + drow[t] == override_index) c = override_code;
    codes[t * stride] = static_cast<uint8_t>(c);
    if (t == kCentrePos) continue;
    const int e = t - (t > kCentrePos);
    if (e < 32) lo |= (c != solvent) ? (1u << e) : 0u;
    else hi |= (c != solvent) ? (1u << (e - 32)) : 0u;
  }
  // marks of older epochs are numerically smaller than any mark of the current epoch
  if (worst > my_mark) *conflict = true;
  return (static_cast<uint64_t>(hi) << 32) | lo;
}

// H(x_new, env) - H(x_old, env) from staged codes and (shared-memory) tables
__device__ __forceinline__ double site_energy_change_staged(const SiteTablesView &tv, int x_old, int x_new, uint64_t sol,
                                                            const uint8_t *codes, int stride) {
  const int m = tv.m;
  const size_t a_stride = static_cast<size_t>(kSiteEnvN) * m, b_stride = static_cast<size_t>(tv.n_pairs) * m * m;
  const double *A_new = tv.A + x_new * a_stride, *A_old = tv.A + x_old * a_stride;
  const double *B_new = tv.B + x_new * b_stride, *B_old = tv.B + x_old * b_stride;
  double acc = tv.C[x_new] - tv.C[x_old];
  while (sol) {
    const int t = __ffsll(static_cast<long long>(sol)) - 1;
    sol &= sol - 1;
    const int et = codes[(t + (t >= kCentrePos)) * stride];
    acc += A_new[t * m + et] - A_old[t * m + et];
    const uint64_t hi = tv.mask_hi[t];
    uint64_t partners = hi & sol;
    const int pbase = tv.base[t];
    while (partners) {
      const int u = __ffsll(static_cast<long long>(partners)) - 1;
      partners &= partners - 1;
      const int eu = codes[(u + (u >= kCentrePos)) * stride];
      const size_t p = (static_cast<size_t>(pbase + __popcll(hi & ((1ULL << u) - 1ULL))) * m + et) * m + eu;
      acc += B_new[p] - B_old[p];
    }
  }
  return acc;
}

// One side of swap_energy_change (kernels.cuh): the two single-site changes of a swap are evaluated by two adjacent
// lanes (side 0: the site that changes first, side 1: the other one, which sees the first already changed when the pair
// is coupled), each with the claim check fused into its 43-site gather.  Returns this side's H(new) - H(old).
__device__ __forceinline__ double swap_side_energy_change_marked(const LatticeDesc &lat, const DevTables &tab, const SiteTablesView &tv,
                                                                 const unsigned int *cells,
                                                                 const int32_t *__restrict__ s_delta, uint8_t *codes, int stride, int side,
                                                                 int xa, int ya, int za, int xb, int yb, int zb, unsigned my_mark,
                                                                 bool *conflict, bool *same_species) {
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
  if (coupled && eb == vac) {   // move the vacancy first so that no intermediate state holds two vacancies
    const int64_t tb = base_a; base_a = base_b; base_b = tb;
    const unsigned te = ea; ea = eb; eb = te;
    const int tz = zpa; zpa = zpb; zpb = tz;
    dx = -dx; dy = -dy; dz = -dz;
  }
  *same_species = ea == eb;
  // both lanes of a pair run ONE instruction stream: the side only selects the operands (no divergent copies of the gather)
  // side 1: the halo images of the first site are not updated in memory, so the override is applied by *position*: the
  // first site sits at displacement (-dx,-dy,-dz) from the second
  const int64_t base = side ? base_b : base_a;
  const int zp = side ? zpb : zpa;
  const int64_t override_index = (side && coupled) ? base_b + lat.padded_delta(-dx, -dy, -dz, zpb) : -1;
  const uint64_t sol = gather_site_env_marked(cells, base, s_delta + zp * 43, solvent, codes, stride, override_index, eb, my_mark, conflict);
  // a trial that lost the claim on either side is dropped by the caller: skip its table walk (both lanes of the pair are here)
  const unsigned pair_mask = 3u << (threadIdx.x & 30u);
  const bool lost = __shfl_xor_sync(pair_mask, *conflict ? 1 : 0, 1) || *conflict;
  if (lost) { *conflict = true; return 0.0; }
  if (ea == eb) return 0.0;
  return site_energy_change_staged(tv, static_cast<int>(side ? eb : ea), static_cast<int>(side ? ea : eb), sol, codes, stride);
}

constexpr int kCmcMaxThreads = 512;
constexpr int kCmcDraws = 8;          // proposals drawn per thread and batch; the first unlike-species pair is the thread's trial

// species by lattice id (compact codes): lets a proposal test "same species?" with two byte loads
__global__ void cmc_mirror_kernel(LatticeDesc lat, const uint8_t *__restrict__ padded, uint8_t *__restrict__ by_id) {
  const int64_t id = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (id >= lat.num_sites) return;
  by_id[blockIdx.y * lat.num_sites + id] = padded[blockIdx.y * lat.padded_size + lat.padded_index_of_id(id)];
}

// cell array from the padded occupancy (all cells incl. the halo): species byte, no marks
__global__ void cmc_cells_init_kernel(int64_t n, const uint8_t *__restrict__ padded, unsigned int *__restrict__ cells) {
  const int64_t q = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (q < n) cells[q] = padded[q];
}

// One thread-block CLUSTER per replica: the CTAs of the cluster share the batch (proposals, marks, evaluation) so that the
// scattered 43-site gathers of a batch are spread over several SMs' L1/LSU pipes; batch bookkeeping is replicated in every
// CTA and kept identical by exchanging per-CTA partial sums through distributed shared memory after each batch.
struct CmcPartial {
  double sum;
  unsigned int kept, accepted, live, first_conflict;
  int err;
};

__global__ void __launch_bounds__(kCmcMaxThreads)
cmc_run_kernel(LatticeDesc lat, DevTables tab, uint8_t *occ, int64_t walker_stride, uint8_t *mirror, unsigned int *cells_all,
               CmcState st, const double *__restrict__ temperatures, uint64_t seed, unsigned long long target_steps, CmcReplay replay,
               unsigned long long n_replay, int stage_b_table) {
  cg::cluster_group cluster = cg::this_cluster();
  const int n_cta = static_cast<int>(cluster.num_blocks()), rank = static_cast<int>(cluster.block_rank());
  __shared__ int32_t s_delta[2 * 43];
  __shared__ double s_warp_sum[kCmcMaxThreads / 32];
  __shared__ unsigned int s_warp_cnt[kCmcMaxThreads / 32], s_warp_acc[kCmcMaxThreads / 32], s_warp_live[kCmcMaxThreads / 32];
  __shared__ unsigned int s_first_conflict;
  __shared__ CmcPartial s_partial2[2];            // double buffered by batch parity: a CTA may run one phase ahead
  __shared__ double s_energy, s_temperature;
  __shared__ unsigned long long s_steps, s_accepted, s_proposals, s_epoch, s_replay_pos;
  __shared__ SaSchedule s_sa;
  __shared__ int32_t s_live_a[kCmcMaxThreads], s_live_b[kCmcMaxThreads];
  extern __shared__ double s_dyn[];                // [replay dE: B*n_cta] [C: m] [A: m*42*m] [B: m*204*m*m, optional] [mask: 42] [base] [codes]

  const int w = blockIdx.x / n_cta;
  const int tid = threadIdx.x, B = blockDim.x;
  uint8_t *o = occ + w * walker_stride;
  uint8_t *by_id = mirror + static_cast<size_t>(w) * lat.num_sites;
  unsigned int *cells = cells_all + static_cast<size_t>(w) * lat.padded_size;
  for (int q = tid; q < 2 * 43; q += B) s_delta[q] = tab.site_delta[q];
  // ---- stage the site tables in shared memory
  const int m = tab.n_species + 1;
  double *s_replay_de = s_dyn;                      // rank 0 holds the dE of the whole replay window
  double *s_C = s_replay_de + B * n_cta;          // (only B/2 * n_cta entries are used)
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
    s_energy = st.energy[w]; s_steps = st.steps[w]; s_accepted = st.accepted[w]; s_proposals = st.proposals[w];
    s_epoch = st.epoch[w]; s_sa = st.sa[w]; s_replay_pos = 0;
    s_temperature = s_sa.enabled ? s_sa.temperature : temperatures[w];
  }
  __syncthreads();
#ifdef LMC_CMC_PROFILE
  __shared__ long long s_prof[8];
  if (tid == 0) for (int q = 0; q < 8; ++q) s_prof[q] = 0;
  long long t_prev = clock64();
#define LMC_TICK(k) do { if (tid == 0) { const long long t_now = clock64(); s_prof[k] += t_now - t_prev; t_prev = t_now; } } while (0)
#else
#define LMC_TICK(k) do { } while (0)
#endif
  const bool replaying = replay.a != nullptr;
  const uint32_t n_sites = static_cast<uint32_t>(lat.num_sites);
  const double cool = s_sa.enabled ? exp(-3.0 / static_cast<double>(s_sa.maximum_steps > 0 ? s_sa.maximum_steps : 1ULL)) : 1.0;
  const int gtid = rank * B + tid;                 // index of this thread within the replica's cluster
  const int window = B * n_cta;                    // proposals (threads) per batch
  const int half = B / 2;                          // live trials a CTA evaluates per batch (one lane pair each)
  const int pair_id = tid >> 1, side = tid & 1;
  int err = 0;

  for (;;) {
    // ---------------- batch bookkeeping (identical in every CTA of the cluster)
    const unsigned long long steps0 = s_steps, epoch = s_epoch + 1, prop0 = s_proposals, rpos = s_replay_pos;
    CmcPartial &s_partial = s_partial2[epoch & 1ULL];
    const double t_batch = s_temperature, energy0 = s_energy;
    if (replaying ? (rpos >= n_replay) : (steps0 >= target_steps)) break;
    __syncthreads();
    if (tid == 0) s_first_conflict = 0xFFFFFFFFu;
    if ((epoch & 0xFFULL) == 0) {                   // 8-bit epoch wrapped: forget all marks (keep the species bytes)
      for (int64_t q = gtid; q < lat.padded_size; q += window) cells[q] &= kCellSpeciesMask;
      cluster.sync();
    }
    const unsigned int epoch8 = static_cast<unsigned int>(epoch & 0xFFULL);
    // ---------------- phase 1a: proposals (GenerateLatticeIdJumpPair: uniform ids, redrawn while the species are equal)
    int32_t a = -1, b = -1;
    if (replaying) {
      // replay trials are dealt round-robin so that priorities (= positions in the stream) interleave over the CTAs
      // thread q < B/2 of CTA `rank` fetches trial rank * B/2 + q of the window (stream order = priority order)
      const unsigned long long ridx = rpos + static_cast<unsigned long long>(rank) * half + tid;
      if (tid < half && ridx < n_replay) {
        const int64_t ra = replay.a[ridx], rb = replay.b[ridx];
        if (ra < 0 || rb < 0 || ra >= lat.num_sites || rb >= lat.num_sites) err |= kErrBadSite;
        else { a = static_cast<int32_t>(ra); b = static_cast<int32_t>(rb); }
      }
    } else {
      uint32_t ida[kCmcDraws], idb[kCmcDraws];
      uint8_t sa_[kCmcDraws], sb_[kCmcDraws];
#pragma unroll
      for (int d = 0; d < kCmcDraws; d += 2) {
        uint32_t r[4];
        const unsigned long long g = (prop0 + gtid) * (kCmcDraws / 2) + (d >> 1);
        philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(seed) ^ static_cast<uint32_t>(w),
                      static_cast<uint32_t>(seed >> 32), r);
        ida[d] = __umulhi(r[0], n_sites); idb[d] = __umulhi(r[1], n_sites);
        ida[d + 1] = __umulhi(r[2], n_sites); idb[d + 1] = __umulhi(r[3], n_sites);
      }
#pragma unroll
      for (int d = 0; d < kCmcDraws; ++d) { sa_[d] = __ldcg(by_id + ida[d]); sb_[d] = __ldcg(by_id + idb[d]); }
#pragma unroll
      for (int d = kCmcDraws - 1; d >= 0; --d)
        if (sa_[d] != sb_[d]) { a = static_cast<int32_t>(ida[d]); b = static_cast<int32_t>(idb[d]); }
    }
    LMC_TICK(0);
    // ---------------- compaction of this CTA's live trials (deterministic: ballot + prefix over warps)
    const bool has = a >= 0;
    int n_live = 0, slot = -1;
    if (replaying) {                                // keep stream order: no compaction, priority = stream position
      n_live = half;
      slot = tid;
      s_live_a[tid] = a;
      s_live_b[tid] = b;
    } else {
      const unsigned bal = __ballot_sync(0xffffffffu, has);
      if ((tid & 31) == 0) s_warp_live[tid >> 5] = __popc(bal);
      __syncthreads();
      int my_off = 0;
      for (int q = 0; q < (B + 31) / 32; ++q) {
        if (q == (tid >> 5)) my_off = n_live;
        n_live += s_warp_live[q];
      }
      if (has) {
        slot = my_off + __popc(bal & ((1u << (tid & 31)) - 1u));
        s_live_a[slot] = a;
        s_live_b[slot] = b;
      }
    }
    __syncthreads();
    LMC_TICK(1);
    // ---------------- phase 1b: lane pair i < min(n_live, B/2) owns live trial i of this CTA (side 0 / 1 = its two sites)
    // priority: device RNG: (CTA rank, slot); replay: position in the stream
    if (n_live > half) n_live = half;               // surplus proposals are dropped (they are redrawn in a later batch)
    bool live = pair_id < n_live;
    int xa = 0, ya = 0, za = 0, xb = 0, yb = 0, zb = 0;
    const unsigned int prio = static_cast<unsigned int>(rank * half + pair_id);
    const unsigned int my_mark = (epoch8 << 16) | (0xFFFFu - prio);       // 24 bits
    if (live) {
      a = s_live_a[pair_id]; b = s_live_b[pair_id];
      live = a >= 0;
    }
    if (live) {
      lat.coords_of_id(a, xa, ya, za);
      lat.coords_of_id(b, xb, yb, zb);
      const int mx = side ? xb : xa, my = side ? yb : ya, mz = side ? zb : za;
      const unsigned species = cells[lat.padded_index(mx, my, mz)] & kCellSpeciesMask;
      mark_site(lat, cells, mx, my, mz, (my_mark << 8) | species);
    }
    cluster.sync();
    LMC_TICK(2);
    // ---------------- phase 2: conflict check fused with the dE gathers, one site per lane
    bool conflict = false, same = false;
    double de = 0.0;
    if (live) de = swap_side_energy_change_marked(lat, tab, tv, cells, s_delta, s_codes + tid, B, side, xa, ya, za, xb, yb, zb, my_mark,
                                                  &conflict, &same);
    conflict = __shfl_xor_sync(0xffffffffu, conflict ? 1 : 0, 1) || conflict;
    de += __shfl_xor_sync(0xffffffffu, de, 1);      // both lanes of the pair now hold the swap's dE
    bool kept = live && !conflict;
    LMC_TICK(3);
    const unsigned int gpair = static_cast<unsigned int>(rank * half + pair_id);   // position of the trial in the window
    if (replaying) {
      // serial semantics: only the conflict-free prefix of the window forms the batch; every CTA learns the position of
      // the first conflict through distributed shared memory
      if (live && conflict && side == 0) atomicMin(&s_first_conflict, gpair);
      __syncthreads();
      if (tid == 0) s_partial.first_conflict = s_first_conflict;
      cluster.sync();
      unsigned int limit = 0xFFFFFFFFu;
      for (int r = 0; r < n_cta; ++r) {
        const unsigned int v = cluster.map_shared_rank(&s_partial, r)->first_conflict;
        limit = v < limit ? v : limit;
      }
      kept = live && gpair < limit;
      if (kept && side == 0) cluster.map_shared_rank(s_replay_de, 0)[gpair] = de;      // rank 0 applies the prefix serially
      if (tid == 0) s_first_conflict = limit;
    }
    if (kept && de != de) err |= kErrExtraVacancy;
    bool accept = false;
    if (kept && !replaying) {
      // CanonicalMcAbstract::SelectEvent (:86-101): dE < 0 accepts, else u < exp(-dE beta); both lanes of the pair draw
      // the same uniform (same counter), so they agree without communication
      accept = de < 0.0;
      if (!accept) {
        uint32_t r2[4];
        const unsigned long long g = prop0 + gpair;
        philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(seed) ^ static_cast<uint32_t>(w),
                      static_cast<uint32_t>(seed >> 32) ^ 0x9E3779B9u, r2);
        const double beta = 1.0 / kBoltzmannEv / fmax(t_batch, 1e-12);
        accept = uniform53(r2[0], r2[1]) < exp(-de * beta);
      }
      if (accept) {
        const int64_t base_a = lat.padded_index(xa, ya, za), base_b = lat.padded_index(xb, yb, zb);
        const uint8_t ea = __ldcg(o + base_a), eb = __ldcg(o + base_b);
        __syncwarp(__activemask());                 // both lanes have read the old species before either writes
        if (side == 0) { store_site(lat, o, xa, ya, za, eb); store_site_cells(lat, cells, xa, ya, za, eb); by_id[a] = eb; }
        else { store_site(lat, o, xb, yb, zb, ea); store_site_cells(lat, cells, xb, yb, zb, ea); by_id[b] = ea; }
      }
    }
    kept = kept && side == 0;                       // count every trial once
    accept = accept && side == 0;
    LMC_TICK(4);
    // ---------------- reductions (fixed order: deterministic)
    const unsigned kept_mask = __ballot_sync(0xffffffffu, kept), acc_mask = __ballot_sync(0xffffffffu, accept);
    double sum = accept ? de : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
    if ((tid & 31) == 0) { s_warp_sum[tid >> 5] = sum; s_warp_cnt[tid >> 5] = __popc(kept_mask); s_warp_acc[tid >> 5] = __popc(acc_mask); }
    const int block_err = __syncthreads_or(err != 0);
    if (tid == 0) {
      double e = 0.0;
      unsigned int n_kept = 0, n_acc = 0;
      for (int q = 0; q < (B + 31) / 32; ++q) { e += s_warp_sum[q]; n_kept += s_warp_cnt[q]; n_acc += s_warp_acc[q]; }
      s_partial.sum = e; s_partial.kept = n_kept; s_partial.accepted = n_acc; s_partial.live = static_cast<unsigned int>(n_live);
      s_partial.err = block_err;
    }
    cluster.sync();
    int any_err = 0;
    for (int r = 0; r < n_cta; ++r) any_err |= cluster.map_shared_rank(&s_partial, r)->err;   // uniform over the cluster
    if (tid == 0) {
      if (!replaying) {
        double e = 0.0;
        unsigned int n_kept = 0, n_acc = 0;
        for (int r = 0; r < n_cta; ++r) {                                  // rank order: every CTA computes the same totals
          const CmcPartial *pp = cluster.map_shared_rank(&s_partial, r);
          e += pp->sum; n_kept += pp->kept; n_acc += pp->accepted;
        }
        s_energy = energy0 + e;
        s_steps = steps0 + n_kept;
        s_accepted += n_acc;
        s_proposals = prop0 + window;
        if (s_sa.enabled && n_kept > 0) {
          // batch-granular schedule: all trials of a batch see the batch-start temperature; the geometric cooling of
          // n_kept trials and the acceptance-window / reheat logic (SimulatedAnnealing.cpp:99-139) are applied per batch
          // (window sizes are >> batch sizes, SimulatedAnnealing.h:52-57)
          SaSchedule sa = s_sa;
          sa.window_trials += n_kept;
          sa.window_accepts += n_acc;
          if (n_acc > 0 && s_energy < sa.recent_best_energy - kSaEpsilon) { sa.recent_best_energy = s_energy; sa.last_improvement_step = s_steps; }
          if (sa.window_trials >= sa.window_size) {
            if (static_cast<double>(sa.window_accepts) / static_cast<double>(sa.window_trials) > 0.50) sa.temperature *= 0.99;
            sa.window_trials = 0; sa.window_accepts = 0;
          }
          const double acc_est = sa.window_trials > 0u ? static_cast<double>(sa.window_accepts) / static_cast<double>(sa.window_trials) : 1.0;
          if (sa.reheats_done < 5u && (s_steps - sa.last_improvement_step >= sa.reheat_trigger_steps) &&
              (s_steps - sa.last_reheat_step >= sa.reheat_cooldown_steps) && acc_est < 0.05) {
            sa.temperature *= 1.10; sa.last_improvement_step = s_steps; sa.last_reheat_step = s_steps;
            sa.recent_best_energy = s_energy; ++sa.reheats_done;
          }
          sa.temperature *= pow(cool, static_cast<double>(n_kept));
          s_sa = sa;
          s_temperature = sa.temperature;
        }
      } else {
        // strictly serial accept / schedule over the conflict-free prefix.  Every CTA runs the same loop on the same
        // inputs (dE window of rank 0, host stream) so the replicated state stays identical; only rank 0 writes.
        const double *de_window = cluster.map_shared_rank(s_replay_de, 0);
        const unsigned long long remaining = n_replay - rpos;
        const unsigned int batch_cap = static_cast<unsigned int>(half * n_cta);
        unsigned int limit = s_first_conflict < batch_cap ? s_first_conflict : batch_cap;
        if (limit > remaining) limit = static_cast<unsigned int>(remaining);
        unsigned long long pos = rpos;
        double energy = energy0;
        unsigned long long step = steps0;
        for (unsigned int q = 0; q < limit; ++q, ++pos) {
          const double d = de_window[q];
          const double temp = s_sa.enabled ? s_sa.temperature : t_batch;
          const double beta = s_sa.enabled ? 1.0 / kBoltzmannEv / fmax(temp, 1e-12) : 1.0 / kBoltzmannEv / temp;
          const bool acc = d < 0.0 || replay.u[pos] < exp(-d * beta);
          if (rank == 0) {
            if (replay.dE) replay.dE[pos] = d;
            if (replay.energy_before) replay.energy_before[pos] = energy;
            if (replay.temperature_before) replay.temperature_before[pos] = temp;
            if (replay.accepted) replay.accepted[pos] = acc ? 1 : 0;
          }
          if (acc) {
            if (rank == 0) {
              int x1, y1, z1, x2, y2, z2;
              lat.coords_of_id(replay.a[pos], x1, y1, z1);
              lat.coords_of_id(replay.b[pos], x2, y2, z2);
              const uint8_t e1 = __ldcg(o + lat.padded_index(x1, y1, z1)), e2 = __ldcg(o + lat.padded_index(x2, y2, z2));
              store_site(lat, o, x1, y1, z1, e2);
              store_site(lat, o, x2, y2, z2, e1);
              store_site_cells(lat, cells, x1, y1, z1, e2);
              store_site_cells(lat, cells, x2, y2, z2, e1);
              by_id[replay.a[pos]] = e2;
              by_id[replay.b[pos]] = e1;
            }
            energy += d;
            ++s_accepted;
          }
          if (s_sa.enabled) sa_update(s_sa, acc, energy, step, cool);
          ++step;
        }
        s_energy = energy;
        s_steps = step;
        s_replay_pos = limit == 0 ? n_replay : pos;   // limit == 0 cannot happen (trial 0 never conflicts); guards against livelock
        if (s_sa.enabled) s_temperature = s_sa.temperature;
      }
      s_epoch = epoch;
    }
    // No third barrier: the swaps of this batch were stored before the barrier above (visible to every CTA's next
    // proposals), the partial sums are double buffered, and the dE window of a replay batch is only rewritten after the
    // next batch's first barrier, which a CTA cannot pass before every CTA has finished reading here.
    __syncthreads();
    LMC_TICK(5);
    if (any_err) break;
  }
#ifdef LMC_CMC_PROFILE
  if (tid == 0 && w == 0 && rank == 0)
    printf("cmc profile (cycles): propose %lld compact %lld mark+sync %lld evaluate %lld accept %lld reduce+sync %lld | batches %llu ctas %d\n",
           s_prof[0], s_prof[1], s_prof[2], s_prof[3], s_prof[4], s_prof[5], s_epoch - st.epoch[w], n_cta);
#endif
  cluster.sync();                                   // nobody may exit while others still read its shared memory
  if (tid == 0 && rank == 0) {
    st.energy[w] = s_energy; st.steps[w] = s_steps; st.accepted[w] = s_accepted; st.proposals[w] = s_proposals;
    st.epoch[w] = s_epoch;
    if (s_sa.enabled) s_sa.temperature = s_temperature;
    st.sa[w] = s_sa;
  }
  if (err) atomicOr(&st.error[w], err);
}

}  // namespace lmc

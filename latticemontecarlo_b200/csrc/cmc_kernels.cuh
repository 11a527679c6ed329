// cmc_kernels.cuh -- batched canonical Monte Carlo / simulated annealing driver.
// Reference: mc/src/CanonicalMcAbstract.cpp:43-51,86-101 (GenerateLatticeIdJumpPair / SelectEvent),
// mc/src/CanonicalMcOmp.cpp:40-92 (batches of mutually non-interfering trials, evaluated in parallel, applied in
// order), mc/src/CanonicalMcSerial.cpp:40-51, mc/src/SimulatedAnnealing.cpp:99-185 (schedule).
//
// One thread block owns one replica ("walker") for the whole launch and loops over batches:
//   phase 1  every thread proposes one swap trial (two uniform lattice ids; Philox4x32-10, counter = trial number) and
//            claims the 43-site neighbourhoods of both sites in a per-replica claim array (64-bit atomicMax of a
//            tag = batch epoch | priority); this is CanonicalMcOmp's `unavailable_position_` set, built in parallel;
//   phase 2  a trial survives if no higher-priority trial of the batch claimed one of its two sites (so no surviving
//            trial reads a site another surviving trial may write: dE evaluated on the batch-start occupancy is the
//            dE the serial chain would see); survivors evaluate dE (swap_energy_change), draw the Metropolis uniform,
//            and apply their swap immediately.
// Proposals whose two sites hold the same species, and proposals that lose the claim, are not trials (the reference
// redraws them: CanonicalMcAbstract.cpp:45-50, CanonicalMcOmp.cpp:47-58), so they do not advance `steps`.
//
// Replay mode (validation): trials (a, b, u) come from the host in the reference's serial order; a batch is the
// longest prefix of the pending trials without interference, dE is evaluated in parallel and the accept decisions,
// energy and the annealing schedule are then applied strictly in order by one thread -- this reproduces
// CanonicalMcSerial / SimulatedAnnealing traces exactly.
#pragma once
#include "kmc_kernels.cuh"

namespace lmc {

struct SaSchedule {             // SimulatedAnnealing members (mc/include/SimulatedAnnealing.h:33-68)
  double temperature;
  double recent_best_energy;
  unsigned long long last_improvement_step, last_reheat_step;
  unsigned long long maximum_steps, reheat_trigger_steps, reheat_cooldown_steps, window_size;
  unsigned int window_trials, window_accepts, reheats_done;
  int enabled;
};

struct CmcState {               // per-replica arrays
  double *energy;               // energy_ (relative to the start, like McAbstract::energy_ with restart_energy 0)
  unsigned long long *steps;    // effective trials so far (steps_)
  unsigned long long *accepted;
  unsigned long long *proposals;   // Philox counter: proposals drawn so far
  unsigned long long *epoch;    // batch counter for the claim tags
  SaSchedule *sa;
  int32_t *error;
};

struct CmcReplay {              // host-ordered trial stream for replica 0 (device copies)
  const int64_t *a, *b;
  const double *u;
  double *dE, *energy_before, *temperature_before;   // per-trial outputs
  uint8_t *accepted;
};

constexpr double kSaEpsilon = 1e-4;   // kEpsilon (cfg/include/VectorMatrix.hpp:65) used by SimulatedAnnealing.cpp:85,104

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

__device__ __forceinline__ uint64_t mul_hi_u64(uint64_t x, uint64_t n) { return __umul64hi(x, n); }

// claim the 43-site neighbourhood of a site (cells addressed as base + offset row, i.e. possibly halo images)
__device__ __forceinline__ void claim_site(unsigned long long *claim, int64_t base, const int32_t *__restrict__ drow, unsigned long long tag) {
#pragma unroll 1
  for (int t = 0; t < 43; ++t) atomicMax(claim + base + drow[t], tag);
}

// highest tag found on any periodic image of the site (interior cell + halo images)
__device__ __forceinline__ unsigned long long strongest_claim(const LatticeDesc &lat, const unsigned long long *claim, int X, int Y, int Z) {
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  unsigned long long best = claim[lat.padded_index(X, Y, Z)];
  const bool edge = X < kHalo || X >= px - kHalo || Y < kHalo || Y >= py - kHalo || Z < kHaloZ || Z >= pz - kHaloZ;
  if (!edge) return best;
  for (int a = -1; a <= 1; ++a) {
    const int x = X + a * px;
    if (x < -kHalo || x >= px + kHalo) continue;
    for (int b = -1; b <= 1; ++b) {
      const int y = Y + b * py;
      if (y < -kHalo || y >= py + kHalo) continue;
      for (int c = -1; c <= 1; ++c) {
        const int z = Z + c * pz;
        if (z < -kHaloZ || z >= pz + kHaloZ) continue;
        const unsigned long long v = claim[lat.padded_index(x, y, z)];
        best = v > best ? v : best;
      }
    }
  }
  return best;
}

constexpr int kCmcMaxThreads = 1024;

__global__ void __launch_bounds__(kCmcMaxThreads)
cmc_run_kernel(LatticeDesc lat, DevTables tab, uint8_t *occ, int64_t walker_stride, unsigned long long *claims, CmcState st,
               const double *__restrict__ temperatures, uint64_t seed, unsigned long long target_steps, CmcReplay replay,
               unsigned long long n_replay) {
  __shared__ int32_t s_delta[2 * 43];
  __shared__ double s_warp_sum[kCmcMaxThreads / 32];
  __shared__ unsigned int s_warp_cnt[kCmcMaxThreads / 32], s_warp_acc[kCmcMaxThreads / 32];
  __shared__ unsigned int s_first_conflict;
  __shared__ double s_energy, s_temperature;
  __shared__ unsigned long long s_steps, s_accepted, s_proposals, s_epoch, s_replay_pos;
  __shared__ SaSchedule s_sa;
  extern __shared__ double s_replay_de[];          // replay mode: dE of the batch, blockDim.x doubles

  const int w = blockIdx.x;
  const int tid = threadIdx.x, B = blockDim.x;
  uint8_t *o = occ + w * walker_stride;
  unsigned long long *claim = claims + static_cast<size_t>(w) * lat.padded_size;
  for (int q = tid; q < 2 * 43; q += B) s_delta[q] = tab.site_delta[q];
  if (tid == 0) {
    s_energy = st.energy[w]; s_steps = st.steps[w]; s_accepted = st.accepted[w]; s_proposals = st.proposals[w];
    s_epoch = st.epoch[w]; s_sa = st.sa[w]; s_replay_pos = 0;
    s_temperature = s_sa.enabled ? s_sa.temperature : temperatures[w];
  }
  __syncthreads();
  const bool replaying = replay.a != nullptr;
  const uint64_t n_sites = static_cast<uint64_t>(lat.num_sites);
  const double cool = s_sa.enabled ? exp(-3.0 / static_cast<double>(s_sa.maximum_steps > 0 ? s_sa.maximum_steps : 1ULL)) : 1.0;
  int err = 0;

  for (;;) {
    // ---------------- batch bookkeeping (uniform across the block)
    const unsigned long long steps0 = s_steps, epoch = s_epoch + 1, prop0 = s_proposals, rpos = s_replay_pos;
    const double t_batch = s_temperature, energy0 = s_energy;
    if (replaying ? (rpos >= n_replay) : (steps0 >= target_steps)) break;
    __syncthreads();
    if (tid == 0) s_first_conflict = 0xFFFFFFFFu;
    // ---------------- phase 1: propose + claim
    int64_t a = -1, b = -1;
    uint32_t r[4] = {0, 0, 0, 0};
    if (replaying) {
      if (rpos + tid < n_replay) { a = replay.a[rpos + tid]; b = replay.b[rpos + tid]; }
    } else {
      const unsigned long long g = prop0 + tid;
      philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(seed) ^ static_cast<uint32_t>(w),
                    static_cast<uint32_t>(seed >> 32), r);
      a = static_cast<int64_t>(mul_hi_u64((static_cast<uint64_t>(r[1]) << 32) | r[0], n_sites));
      b = static_cast<int64_t>(mul_hi_u64((static_cast<uint64_t>(r[3]) << 32) | r[2], n_sites));
    }
    bool live = a >= 0 && b >= 0 && a < lat.num_sites && b < lat.num_sites;
    if (replaying && rpos + tid < n_replay && !live) err |= kErrBadSite;
    int xa = 0, ya = 0, za = 0, xb = 0, yb = 0, zb = 0;
    int64_t base_a = 0, base_b = 0;
    const unsigned long long tag = (epoch << 16) | static_cast<unsigned long long>(0xFFFF - tid);
    if (live) {
      lat.coords_of_id(a, xa, ya, za);
      lat.coords_of_id(b, xb, yb, zb);
      base_a = lat.padded_index(xa, ya, za);
      base_b = lat.padded_index(xb, yb, zb);
      if (!replaying && o[base_a] == o[base_b]) live = false;      // same species: not a trial (redrawn by the reference)
    }
    if (live) {
      claim_site(claim, base_a, s_delta + (za & 1) * 43, tag);
      claim_site(claim, base_b, s_delta + (zb & 1) * 43, tag);
    }
    __syncthreads();
    // ---------------- phase 2: survivors
    bool kept = false;
    if (live) {
      const unsigned long long ca = strongest_claim(lat, claim, xa, ya, za), cb = strongest_claim(lat, claim, xb, yb, zb);
      kept = ca <= tag && cb <= tag;       // tags of this epoch from lower thread ids are larger; stale epochs are smaller
      if (!kept && replaying) atomicMin(&s_first_conflict, static_cast<unsigned int>(tid));
    }
    __syncthreads();
    if (replaying) {
      // serial semantics: only the conflict-free prefix of the pending trials forms the batch
      const unsigned int limit = s_first_conflict;
      kept = live && static_cast<unsigned int>(tid) < limit;
    }
    double de = 0.0;
    bool accept = false;
    if (kept) {
      de = swap_energy_change(lat, tab, o, s_delta, xa, ya, za, xb, yb, zb, &err);
      if (!replaying) {
        // CanonicalMcAbstract::SelectEvent (:86-101): dE < 0 accepts, else u < exp(-dE beta); SA uses the temperature
        // the trial would see after the geometric cooling of the trials before it in this batch
        accept = de < 0.0;
        if (!accept) {
          uint32_t r2[4];
          const unsigned long long g = prop0 + tid;
          philox4x32_10(static_cast<uint32_t>(g), static_cast<uint32_t>(g >> 32), static_cast<uint32_t>(seed) ^ static_cast<uint32_t>(w),
                        static_cast<uint32_t>(seed >> 32) ^ 0x9E3779B9u, r2);
          const double u = uniform53(r2[0], r2[1]);
          const double beta = 1.0 / kBoltzmannEv / fmax(t_batch, 1e-12);
          accept = u < exp(-de * beta);
        }
        if (accept) {
          const uint8_t ea = o[base_a], eb = o[base_b];
          store_site(lat, o, xa, ya, za, eb);
          store_site(lat, o, xb, yb, zb, ea);
        }
      } else {
        s_replay_de[tid] = de;
      }
    }
    // ---------------- reductions (fixed order: deterministic)
    const unsigned kept_mask = __ballot_sync(0xffffffffu, kept), acc_mask = __ballot_sync(0xffffffffu, accept);
    double sum = accept ? de : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, off);
    if ((tid & 31) == 0) { s_warp_sum[tid >> 5] = sum; s_warp_cnt[tid >> 5] = __popc(kept_mask); s_warp_acc[tid >> 5] = __popc(acc_mask); }
    __syncthreads();
    if (tid == 0) {
      if (!replaying) {
        double e = 0.0;
        unsigned int n_kept = 0, n_acc = 0;
        for (int q = 0; q < (B + 31) / 32; ++q) { e += s_warp_sum[q]; n_kept += s_warp_cnt[q]; n_acc += s_warp_acc[q]; }
        s_energy = energy0 + e;
        s_steps = steps0 + n_kept;
        s_accepted += n_acc;
        s_proposals = prop0 + B;
        if (s_sa.enabled && n_kept > 0) {
          // batch-granular schedule: the per-trial geometric cooling is exact; acceptance-window and reheat logic see the
          // batch as one block of trials (window sizes are >> batch sizes, SimulatedAnnealing.h:52-57)
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
        // strictly serial accept / schedule over the conflict-free prefix
        const unsigned int limit = s_first_conflict < static_cast<unsigned int>(B) ? s_first_conflict : static_cast<unsigned int>(B);
        unsigned long long pos = rpos;
        double energy = energy0;
        unsigned long long step = steps0;
        for (unsigned int q = 0; q < limit && pos < n_replay; ++q, ++pos) {
          const double d = s_replay_de[q];
          const double temp = s_sa.enabled ? s_sa.temperature : t_batch;
          const double beta = s_sa.enabled ? 1.0 / kBoltzmannEv / fmax(temp, 1e-12) : 1.0 / kBoltzmannEv / temp;
          if (replay.dE) replay.dE[pos] = d;
          if (replay.energy_before) replay.energy_before[pos] = energy;
          if (replay.temperature_before) replay.temperature_before[pos] = temp;
          const bool acc = d < 0.0 || replay.u[pos] < exp(-d * beta);
          if (replay.accepted) replay.accepted[pos] = acc ? 1 : 0;
          if (acc) {
            int x1, y1, z1, x2, y2, z2;
            lat.coords_of_id(replay.a[pos], x1, y1, z1);
            lat.coords_of_id(replay.b[pos], x2, y2, z2);
            const uint8_t e1 = o[lat.padded_index(x1, y1, z1)], e2 = o[lat.padded_index(x2, y2, z2)];
            store_site(lat, o, x1, y1, z1, e2);
            store_site(lat, o, x2, y2, z2, e1);
            energy += d;
            ++s_accepted;
          }
          if (s_sa.enabled) sa_update(s_sa, acc, energy, step, cool);
          ++step;
        }
        s_energy = energy;
        s_steps = step;
        s_replay_pos = pos;
        if (s_sa.enabled) s_temperature = s_sa.temperature;
        if (limit == 0) s_replay_pos = n_replay;     // cannot happen (a trial never conflicts with itself); guards against livelock
      }
      s_epoch = epoch;
    }
    __syncthreads();
    if (__syncthreads_or(err != 0)) break;
  }
  if (tid == 0) {
    st.energy[w] = s_energy; st.steps[w] = s_steps; st.accepted[w] = s_accepted; st.proposals[w] = s_proposals;
    st.epoch[w] = s_epoch;
    if (s_sa.enabled) s_sa.temperature = s_temperature;
    st.sa[w] = s_sa;
  }
  if (err) atomicOr(&st.error[w], err);
}

}  // namespace lmc

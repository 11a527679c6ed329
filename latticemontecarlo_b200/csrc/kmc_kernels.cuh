// kmc_kernels.cuh -- batched first-order kinetic Monte Carlo driver (mc::KineticMcFirstOmp semantics) over many
// independent walkers.  Reference: mc/src/KineticMcAbstract.cpp:140-188 (OneStepSimulation / Simulate),
// mc/src/KineticMcFirstOmp.cpp:52-82 (BuildEventList / CalculateTime), mc/src/JumpEvent.cpp:6-13.
//
// One half-warp owns one walker for the whole launch: lanes 0..11 evaluate the 12 candidate jumps of the walker's
// vacancy (barrier kernel body), the 12 events are put in the reference's order (ascending neighbour lattice id),
// the total rate and the cumulative probabilities of the select are accumulated in the reference's (sequential) order,
// and the jump is written back to the walker's occupancy.  The residence time is (-ln u1 / total) * (corr / 1e13) -- one
// division per step; the reference divides by 1e13 first and multiplies by corr afterwards, so the two agree to the
// last bit or one ulp per step (the tests bound the accumulated time at 1e-9 relative).  Random numbers come from
// Philox4x32-10 (key = seed ^ walker, counter = step; u1 is shifted to (0, 1] so that the logarithm stays finite, the
// reference's generate_canonical delivers [0, 1)) or, in replay mode, from host-supplied (u1, u2) streams so that the
// reference's event sequence can be reproduced.
#pragma once
#include "kernels.cuh"

namespace lmc {

constexpr double kPrefactorHz = 1e13;                // Constants.hpp:32

// pair-table loads of the walk in flight together (same summation order as one at a time: bit-identical results).
// Measured on B200, 8192 walkers x 2048 hops x 8 launches: 1 -> 128.0 ms, 2 -> 125.6 ms, 3 -> 129.6 ms
#ifndef LMC_KMC_BATCH_B
#define LMC_KMC_BATCH_B 2
#endif

struct KmcState {            // per-walker arrays in device memory
  int64_t *vacancy;          // lattice id of the vacancy
  double *time, *energy;     // time_ / energy_ of McAbstract
  int64_t *steps;            // steps_
  double *temperature;       // temperature_ (constant per walker unless a T(t) table is given)
  double *c_vacancy, *c_solute;   // RateCorrector inputs (pred/include/RateCorrector.hpp)
  int32_t *error;            // sticky per-walker EventError bits
  int64_t *previous;         // second-order KMC: lattice id the vacancy came from (previous_j_lattice_id_); -1 = not set
};

struct KmcParams {
  int32_t n_tt;              // number of (time, temperature) points; 0 = constant temperature
  const double *tt_time, *tt_temp;
  int32_t rate_corrector;
  uint64_t seed;
  // Tail hand-off (Engine::kmc_run): kmc_run_kernel counts the walkers that have done all their steps in *handoff_done
  // and, once handoff_threshold of them are through, the half-warps still running stop at the next 16-step boundary and
  // save their state; kmc_team_run_kernel then takes every walker from st.steps[w] to steps_target[w].  All null / 0 = off.
  int *handoff_done;
  int32_t handoff_threshold;
  const int64_t *steps_target;
  const int32_t *tail_order, *tail_count;   // tail: block b takes walker tail_order[b] (most steps left first), b < *tail_count
  unsigned long long *finish_ns;   // diagnostics (LMC_KMC_FINISH_TIMES=1): %globaltimer when a walker's half-warp leaves kmc_run_kernel; else null
  double select_margin;      // latency kernel: rounding margin of its one-pass event selection (select_event_fast); > 1 = always the sequential form
};

// One-pass event selection of the latency kernel (kmc_team_kernels.cuh).
// SelectEvent (KineticMcAbstract.cpp:106-116) picks the first slot whose running sum of p_q = rate_q / total is not below
// u2, with `total` and the running sum accumulated one after the other: three dependent passes over the 12 events.
// The same slot follows from ONE pass whenever u2 is not within rounding distance of a boundary.  With the exact
// partial sums S_i of the (non-negative) rates and S = S_11, the reference's running sum c_i equals S_i / S up to a
// relative error below 24 ulp (11 additions in `total`, one division, 11 additions in the sum), i.e. |c_i - S_i / S| <
// 3e-15.  Here lane i adds the rates of slots 0..i as a tree, P_i (relative error <= 4 ulp), T = P_11, and compares
// P_i with u2 T (one more rounding): |(P_i - u2 T) / T - (S_i / S - u2)| < 2e-15.  If |P_i - u2 T| > margin T with
// margin = 1e-12 for every slot, both forms order every c_i and u2 alike and select the same slot.  Otherwise (about
// 2e-11 of the steps, or a total that is zero, infinite or NaN) the caller runs the reference's sequential form.
// Returns whether `below` (this lane's "c_i < u2") is decided; the caller votes over the lanes that own a slot.
constexpr double kSelectMargin = 1e-12;
// kWidth = 0: every lane adds T itself (a latency-bound caller); 16 / 32: T is shuffled from lane 11 of each kWidth-lane group
template <int kWidth>
__device__ __forceinline__ bool select_event_fast(const double (&r)[12], int lane, double u2, double margin, bool &below) {
  double m[12];
#pragma unroll
  for (int q = 0; q < 12; ++q) m[q] = q <= lane ? r[q] : 0.0;
  const double P = ((m[0] + m[1]) + (m[2] + m[3])) + ((m[4] + m[5]) + (m[6] + m[7])) + ((m[8] + m[9]) + (m[10] + m[11]));
  double T;
  if (kWidth == 0) T = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7])) + ((r[8] + r[9]) + (r[10] + r[11]));
  else T = __shfl_sync(0xFFFFFFFFu, P, 11, kWidth == 0 ? 32 : kWidth);
  const double d = fma(-u2, T, P);
  below = d < 0.0;
  return fabs(d) > margin * T;                                             // false for NaN and for T = 0 or infinity
}

// The reference's form (KineticMcFirstOmp.cpp:55-77, KineticMcAbstract.cpp:106-116): total, p_q = rate_q / total and the
// running sum, each accumulated in slot order.  `rates` (shared memory: the 12 rates of this lane group in slot order) is
// overwritten with the probabilities.  Called by all lanes of a warp together; out of line, because it runs about once
// in 1e10 steps and must not cost the callers registers.  Returns this lane's "c_lane < u2".
__device__ __noinline__ bool select_event_sequential(double *rates, int lane, bool owns_slot, double u2, double *total_out) {
  double total = 0.0;
#pragma unroll
  for (int q = 0; q < 12; ++q) total += rates[q];
  const double mine = rates[owns_slot ? lane : 0];
  __syncwarp(0xFFFFFFFFu);
  if (owns_slot) rates[lane] = mine / total;
  __syncwarp(0xFFFFFFFFu);
  double cumulative = 0.0;                      // ((p0 + p1) + p2) + ... + p_lane
#pragma unroll
  for (int q = 0; q < 12; ++q)
    if (q <= lane) cumulative += rates[q];
  __syncwarp(0xFFFFFFFFu);
  *total_out = total;
  return cumulative < u2;
}

// exp(x) for the dependent chain of a KMC step (all first- and second-order KMC kernels use it, so that every launch
// shape computes bit-identical rates): k = round(x / ln 2), r = x - k ln 2 (two-constant Cody-Waite), a degree-13
// Taylor polynomial in Estrin form (4 dependent FMA levels; truncation 4e-18 for |r| <= 0.347, rounding <= 2 ulp), 2^k
// added to the exponent field.  Branch-free, about 25 instructions, 11 of them on the dependent path (the library's
// exp: 47 and two branches).  Valid for |x| < 700 only: the caller checks the arguments once and otherwise takes the
// library functions.
__device__ __forceinline__ double exp_chain(double x) {
  constexpr double kMagic = 6755399441055744.0;                        // 1.5 * 2^52: the add rounds to the nearest integer
  const double t = fma(x, 1.4426950408889634074, kMagic);
  const int k = __double2loint(t);
  const double kf = t - kMagic;
  double r = fma(kf, -6.93147180369123816490e-01, x);
  r = fma(kf, -1.90821492927058770002e-10, r);
  const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
  const double a0 = fma(r, 1.0, 1.0), a1 = fma(r, 1.0 / 6.0, 0.5), a2 = fma(r, 1.0 / 120.0, 1.0 / 24.0), a3 = fma(r, 1.0 / 5040.0, 1.0 / 720.0),
               a4 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0), a5 = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0),
               a6 = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600.0);
  const double b0 = fma(a1, r2, a0), b1 = fma(a3, r2, a2), b2 = fma(a5, r2, a4);
  const double d0 = fma(b1, r4, b0), d1 = fma(a6, r4, b2);
  const double p = fma(d1, r8, d0);
  return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// (Ea, rate) of a jump from the folded pair (dE, log E0): barrier_from_folded (kernels.cuh) and JumpEvent.cpp:13 as one
// dependent chain.  The division of the quartic form, x = 16 dE / E0, becomes a second exponential that runs beside the
// first (x = 16 dE exp(-log E0)); one range check at the end covers the three exponentials.  Agrees with
// barrier_from_folded / exp to a few ulp.
__device__ __forceinline__ void barrier_and_rate_chain(double dE, double log_e0, int model, double beta, double &ea_out, double &rate_out) {
  const double e0 = exp_chain(log_e0), inv_e0 = exp_chain(-log_e0);
  const double x = 16.0 * dE * inv_e0;
  const double s = 3.0 * x + 4.0;
  double ea = e0 * (s * s) * (8.0 + 4.0 * x - 1.5 * (x * x)) * (1.0 / 8192.0);
  if (model != 0) ea = fmax(0.0, e0 + 0.5 * dE);
  const double arg = -ea * beta;
  double rate = exp_chain(arg);
  if (!(fabs(log_e0) < 700.0 && fabs(arg) < 700.0)) {          // barriers of tens of eV, NaN: the library functions
    ea = barrier_from_folded(dE, log_e0, model);
    rate = exp(-ea * beta);
  }
  ea_out = ea;
  rate_out = rate;
}

// Hybrid launch, between the two kernels: the walkers that still have steps to do, those with the most steps left first
// (64 classes; the latency kernel's blocks start in this order, so the longest remainder is not left for last).  One block.
__global__ void kmc_tail_order_kernel(const int64_t *__restrict__ steps, const int64_t *__restrict__ target, const int32_t *__restrict__ error, int n,
                                      int64_t n_steps, int32_t *__restrict__ order, int32_t *__restrict__ count) {
  __shared__ int s_hist[64], s_start[64];
  for (int q = threadIdx.x; q < 64; q += blockDim.x) s_hist[q] = 0;
  __syncthreads();
  auto cls = [&](int w) -> int {             // 0 = most steps left; -1 = nothing to do
    const int64_t left = error[w] ? 0 : target[w] - steps[w];
    if (left <= 0) return -1;
    const int64_t c = 63 - (left * 63) / (n_steps > 0 ? n_steps : 1);
    return static_cast<int>(c < 0 ? 0 : (c > 63 ? 63 : c));
  };
  for (int w = threadIdx.x; w < n; w += blockDim.x) {
    const int c = cls(w);
    if (c >= 0) atomicAdd(&s_hist[c], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int at = 0;
    for (int q = 0; q < 64; ++q) { s_start[q] = at; at += s_hist[q]; }
    *count = at;
  }
  __syncthreads();
  for (int w = threadIdx.x; w < n; w += blockDim.x) {
    const int c = cls(w);
    if (c >= 0) order[atomicAdd(&s_start[c], 1)] = w;
  }
}

// boundary / parity class of a half-unit coordinate (periods are even and >= 8): 0 = coordinate 0, 3 = period - 1, else 1 + parity
__device__ __forceinline__ int coord_class(int v, int period) { return v == 0 ? 0 : (v == period - 1 ? 3 : 1 + (v & 1)); }

struct KmcTraceDev {         // optional per-step records, [walker][n_steps]; any pointer may be null
  int64_t *from, *to;
  int32_t *slot;
  double *dt, *Ea, *dE, *total_rate, *temperature;
};

// ---- Philox4x32-10 (Salmon et al. 2011), counter = (c0,c1,0,0), key = (k0,k1)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c[4] = {c0, c1, 0u, 0u};
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

// pred::TimeTemperatureInterpolator::GetTemperature (pred/src/TimeTemperatureInterpolator.cpp:43-65)
__device__ __forceinline__ double interpolate_temperature(const KmcParams &p, double time) {
  int k = 0;                                   // lower_bound: first point with time_k >= time
  while (k < p.n_tt && p.tt_time[k] < time) ++k;
  if (k == p.n_tt) return p.tt_temp[p.n_tt - 1];
  if (k == 0 && time <= p.tt_time[0]) return p.tt_temp[0];
  const double x1 = p.tt_time[k], y1 = p.tt_temp[k], x0 = p.tt_time[k - 1], y0 = p.tt_temp[k - 1];
  return y0 + ((time - x0) / (x1 - x0)) * (y1 - y0);
}

// pred::RateCorrector::GetTimeCorrectionFactor (pred/include/RateCorrector.hpp:17-24)
__device__ __forceinline__ double rate_correction(double c_vac, double c_sol, double temperature) {
  const double correct = 1.64 * exp(-(0.66 / kBoltzmannEv / temperature - 0.7));
  return c_vac / correct / (1.0 - 13.0 * c_sol);
}

// One block per walker: locate the vacancy, count species (Config::GetVacancyLatticeId, GetVacancyConcentration,
// GetSoluteConcentration(Al): cfg/src/Config.cpp:296-310,373-393), reset time / energy / steps.
__global__ void kmc_init_kernel(LatticeDesc lat, const uint8_t *__restrict__ occ, int64_t walker_stride, KmcState st,
                                int vac_code, int al_code, int reset_clock) {
  const int w = blockIdx.x;
  const uint8_t *o = occ + w * walker_stride;
  __shared__ unsigned long long s_vac_id;
  __shared__ unsigned int s_nvac, s_nsol;
  if (threadIdx.x == 0) { s_vac_id = ~0ULL; s_nvac = 0; s_nsol = 0; }
  __syncthreads();
  unsigned nvac = 0, nsol = 0;
  unsigned long long first = ~0ULL;
  for (int64_t id = threadIdx.x; id < lat.num_sites; id += blockDim.x) {
    const int c = o[lat.padded_index_of_id(id)];
    if (c == vac_code) { ++nvac; if (static_cast<unsigned long long>(id) < first) first = id; }
    else if (c != al_code) ++nsol;
  }
  atomicAdd(&s_nvac, nvac);
  atomicAdd(&s_nsol, nsol);
  atomicMin(&s_vac_id, first);
  __syncthreads();
  if (threadIdx.x == 0) {
    st.vacancy[w] = s_nvac ? static_cast<int64_t>(s_vac_id) : -1;
    st.c_vacancy[w] = static_cast<double>(s_nvac) / static_cast<double>(lat.num_sites);
    st.c_solute[w] = static_cast<double>(s_nsol) / static_cast<double>(lat.num_sites);
    st.error[w] = s_nvac == 1 ? 0 : kErrNotVacancy;
    if (reset_clock) { st.time[w] = 0.0; st.energy[w] = 0.0; st.steps[w] = 0; }
    st.previous[w] = -1;
  }
}


constexpr int kKmcThreads = 128;   // 8 walkers per block
constexpr int kKmcWalkersPerBlock = kKmcThreads / 16;
// the vacancy sits in row kBoxCentreRow of the box, slot 2 (even Z) or 1 (odd Z)

// Shared-memory tables and constants every event evaluation needs (set up once per block)
struct KmcEvalContext {
  const int32_t *s_box;            // [2][196]
  const int8_t *s_envpos;          // [2][12][196]
  const double2 *s_A2v;            // [n][58][n] (dE, log E0)
  const uint2 *s_mask_hi2;         // [58]
  const uint16_t *s_pbase;         // [58]
  const double2 *B_all;            // global: [n][556][n][n]
  const double *C2;                // global: [n][2]
  int b_stride, n_species;
  unsigned solvent, vac_code;
  int barrier_model;               // DevTables::barrier_model
};

// The 12 jumps of the vacancy at (X, Y, Z) of occupancy `o`, one jump per lane (lanes 0..11 of a half-warp; all 16 lanes
// take part in the scan).  The box around the vacancy is scanned ONCE: each lane reads up to 4 rows of 4 consecutive
// bytes, hits (non-solvent cells) are kept as a bit mask + packed codes and compacted into the half-warp's list; every
// event lane then maps that list into its own symmetry-ordered environment and contracts the folded (dE, log E0) tables.
//
// kHypothetical (second-order KMC, KineticMcChainOmpi.cpp:68-86): the configuration is read "as if" the vacancy had just
// jumped here from the cell at absolute padded index `hypo_index`: the centre (which really holds the migrated atom) is
// the vacancy, the cell at hypo_index holds `hypo_code`, and the jump in direction `hypo_dir` (back to where the vacancy
// came from) moves that atom.  Nothing is written to memory.
template <bool kHypothetical, bool kGlobalA = false>
__device__ __forceinline__ void kmc_scan_and_evaluate(const LatticeDesc &lat, const KmcEvalContext &ctx, const uint8_t *o, int X, int Y,
                                                      int Z, int lane, bool active, int k, int32_t dmig0, int32_t dmig1,
                                                      uint16_t *list, uint8_t *my_codes, double beta,
                                                      int64_t hypo_index, unsigned hypo_code, int hypo_dir, int &err, double &ea,
                                                      double &de, double &rate, unsigned &mig) {
  constexpr unsigned hmask = 0xFFFFFFFFu;
  const unsigned solvent = ctx.solvent, vac_code = ctx.vac_code;
  const int n_species = ctx.n_species;
  const int32_t *s_box = ctx.s_box;
  const int8_t *s_envpos = ctx.s_envpos;
  const double2 *s_A2v = ctx.s_A2v;
  const uint2 *s_mask_hi2 = ctx.s_mask_hi2;
  const uint16_t *s_pbase = ctx.s_pbase;
  const double2 *__restrict__ B_all = ctx.B_all;
  const int b_stride = ctx.b_stride;
  const int barrier_model = ctx.barrier_model;
  {
    // ---- box scan: the non-solvent cells around the vacancy, in cell order
    const int zp = Z & 1;
    const int64_t base = lat.padded_index(X, Y, Z);
    const int32_t *box = s_box + zp * kBoxCells;
    // each lane reads up to 4 rows of the box (4 consecutive bytes each: one address computation per row); hits are
    // kept as a bit mask + packed codes (index = 4 * iteration + slot) and compacted once at the end
    unsigned hit_mask = 0;
    unsigned row_word0 = 0, row_word1 = 0, row_word2 = 0;     // the scanned rows of this lane (kBoxRows / 16 == 3)
    const unsigned solvent4 = solvent * 0x01010101u;
    const int centre_slot = zp ? 1 : 2;
#pragma unroll
    static_assert(kBoxRows % 16 == 0, "the half-warp scans 16 rows per iteration");
    for (int it = 0; it < kBoxRows / 16; ++it) {
      const int row = it * 16 + lane;
      {
        // 4 consecutive bytes at arbitrary alignment: two aligned 32-bit loads + one funnel shift (the occupancy buffer
        // carries 16 bytes of slack at its end for the second word)
        const uintptr_t addr = reinterpret_cast<uintptr_t>(o + base + box[row * 4]);
        const unsigned *aligned = reinterpret_cast<const unsigned *>(addr & ~static_cast<uintptr_t>(3));
        unsigned word = __funnelshift_r(aligned[0], aligned[1], 8u * static_cast<unsigned>(addr & 3u));
        if (kHypothetical) {
          const int64_t off = hypo_index - (base + box[row * 4]);      // the cell the vacancy came from holds the moved atom
          if (off >= 0 && off < 4) word = (word & ~(0xFFu << (8 * static_cast<int>(off)))) | (hypo_code << (8 * static_cast<int>(off)));
        }
        if (row == kBoxCentreRow) {                                    // the row through the vacancy itself
          if (!kHypothetical && ((word >> (8 * centre_slot)) & 0xFFu) != vac_code) err |= kErrNotVacancy;
          word = (word & ~(0xFFu << (8 * centre_slot))) | (solvent << (8 * centre_slot));
        }
        // non-solvent bytes of the row, branch-free: bit 7 of every byte of `nz` that differs from the solvent code, packed
        // into 4 mask bits; the row word itself is kept for the (rare) list entries
        const unsigned diff = word ^ solvent4;
        const unsigned nz = (diff | ((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu)) & 0x80808080u;
        hit_mask |= ((((nz >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * it);   // bytes 0..3 -> bits 0..3
        if (it == 0) row_word0 = word; else if (it == 1) row_word1 = word; else row_word2 = word;
      }
    }
    // exclusive prefix of the per-lane hit counts over the half-warp
    const int mine = __popc(hit_mask);
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 16; off <<= 1) {
      const int v = __shfl_up_sync(hmask, incl, off, 16);
      if (lane >= off) incl += v;
    }
    const int count = __shfl_sync(hmask, incl, 15, 16);
    int pos = incl - mine;
    while (hit_mask) {
      const int it = __ffs(static_cast<int>(hit_mask)) - 1;
      hit_mask &= hit_mask - 1;
      // one 16-bit entry per non-solvent cell: cell index (row * 4 + slot) | species code << 8
      const unsigned rw = (it >> 2) == 0 ? row_word0 : ((it >> 2) == 1 ? row_word1 : row_word2);
      list[pos] = static_cast<uint16_t>((((it >> 2) * 16 + lane) * 4 + (it & 3)) | (((rw >> (8 * (it & 3))) & 0xFFu) << 8));
      ++pos;
    }
    __syncwarp(hmask);
    if (active) {
      mig = o[base + (zp ? dmig1 : dmig0)];
      if (kHypothetical && k == hypo_dir) mig = hypo_code;
      if (mig == vac_code) err |= kErrNotVacancy;
      else {
        // map the box list into this jump's environment: solute mask + species by env index
        const int8_t *envpos = s_envpos + (zp * 12 + k) * kBoxCells;
        uint64_t sol = 0;
        for (int q = 0; q < count; ++q) {
          const unsigned ent = list[q];
          const int t = envpos[ent & 0xFFu];
          if (t >= 0 && t < kEnvN) {
            sol |= 1ULL << t;
            my_codes[t] = static_cast<uint8_t>(ent >> 8);
          }
        }
        // contracted tables: Q = C[m] + sum_t A[m][t][e_t] + sum_(t,u) B[m][(t,u)][e_t][e_u] over the solute sites
        const int m = static_cast<int>(mig), n = n_species;
        const double *__restrict__ C = ctx.C2 + m * 2;
        double a0 = __ldg(C), a1 = __ldg(C + 1);                      // (dE, log E0)
        const double2 *A = s_A2v + m * (kEnvN * n);                   // 32-bit index arithmetic throughout
        const double2 *__restrict__ B = B_all + m * b_stride;
        bool ok = true;
        // the solute mask is walked as two 32-bit words (single-instruction ffs / popc)
        uint32_t w_lo = static_cast<uint32_t>(sol), w_hi = static_cast<uint32_t>(sol >> 32);
        while (w_lo | w_hi) {
          int t;
          if (w_lo) { t = __ffs(static_cast<int>(w_lo)) - 1; w_lo &= w_lo - 1; }
          else { t = 31 + __ffs(static_cast<int>(w_hi)); w_hi &= w_hi - 1; }
          const int et = my_codes[t];
          if (et >= n) { ok = false; continue; }
          const double2 a = kGlobalA ? __ldg(A + t * n + et) : A[t * n + et];
          a0 += a.x; a1 += a.y;
          const uint2 hi = s_mask_hi2[t];
          uint32_t p_lo = hi.x & w_lo, p_hi = hi.y & w_hi;             // partners u > t still to visit
          if ((p_lo | p_hi) == 0) continue;
          const int row = (s_pbase[t] * n + et) * n;
#if LMC_KMC_BATCH_B > 1
          do {                                       // LMC_KMC_BATCH_B partners per pass: their table loads are in flight together
            double2 b[LMC_KMC_BATCH_B];
#pragma unroll
            for (int q = 0; q < LMC_KMC_BATCH_B; ++q) {
              b[q] = make_double2(0.0, 0.0);
              if ((p_lo | p_hi) == 0) continue;
              int u;
              if (p_lo) { u = __ffs(static_cast<int>(p_lo)) - 1; p_lo &= p_lo - 1; }
              else { u = 31 + __ffs(static_cast<int>(p_hi)); p_hi &= p_hi - 1; }
              const int eu = my_codes[u];
              if (eu >= n) { ok = false; continue; }
              const int rank = u < 32 ? __popc(hi.x & ((1u << u) - 1u)) : __popc(hi.x) + __popc(hi.y & ((1u << (u - 32)) - 1u));
              b[q] = __ldg(B + row + rank * (n * n) + eu);
            }
#pragma unroll
            for (int q = 0; q < LMC_KMC_BATCH_B; ++q) { a0 += b[q].x; a1 += b[q].y; }
          } while (p_lo | p_hi);
#else
          do {
            int u;
            if (p_lo) { u = __ffs(static_cast<int>(p_lo)) - 1; p_lo &= p_lo - 1; }
            else { u = 31 + __ffs(static_cast<int>(p_hi)); p_hi &= p_hi - 1; }
            const int eu = my_codes[u];
            if (eu >= n) { ok = false; continue; }
            // index of pair (t,u) among the pairs of t = number of mask bits below u
            const int rank = u < 32 ? __popc(hi.x & ((1u << u) - 1u)) : __popc(hi.x) + __popc(hi.y & ((1u << (u - 32)) - 1u));
            const double2 b = __ldg(B + row + rank * (n * n) + eu);
            a0 += b.x; a1 += b.y;
          } while (p_lo | p_hi);
#endif
        }
        if (!ok) err |= kErrExtraVacancy;
        else {
          de = a0;
          barrier_and_rate_chain(de, a1, barrier_model, beta, ea, rate);    // closed form + JumpEvent.cpp:13
        }
      }
    }
  }
}

// The 12 jumps of one vacancy share a 7 x 7 x 7 half-unit box (196 padded cells).  Per step the half-warp scans the box
// ONCE (13 byte loads per lane instead of 60 per event), compacts the non-solvent cells into a short list, and every
// event lane then maps that list into its own symmetry-ordered environment through a constant cell -> env-index table.
// kInstrumented: replayed uniforms and / or per-step traces (validation runs); the production instantiation carries neither
// 7 blocks per SM: the 8192-walker workload puts at most 7 on an SM (1024 blocks / 148 SMs), and the bound of 8 cost 8
// registers (spills) and, through the shared-memory carve-out for an 8th block, 32 KB of L1
template <bool kInstrumented>
__global__ void __launch_bounds__(kKmcThreads, 7)
kmc_run_kernel(LatticeDesc lat, DevTables tab, uint8_t *occ, int64_t walker_stride, int n_walkers, KmcState st, KmcParams prm,
               int64_t n_steps, const double *__restrict__ replay_u1, const double *__restrict__ replay_u2, KmcTraceDev tr) {
  __shared__ int32_t s_box[2 * kBoxCells];
  __shared__ int8_t s_envpos[2 * 12 * kBoxCells];
  __shared__ uint16_t s_list[kKmcWalkersPerBlock][kBoxCells];
  __shared__ uint8_t s_codes[kKmcWalkersPerBlock * 12][kEnvN + 2];     // one row per EVENT lane (lanes 12..15 of a half-warp have none)
  __shared__ double s_ord_rate[kKmcWalkersPerBlock][12];
  __shared__ uint8_t s_ord_lane[kKmcWalkersPerBlock][12];
  __shared__ uint64_t s_mask_hi[kEnvN];
  __shared__ uint16_t s_pbase[kEnvN];
  __shared__ uint8_t s_slot[64][12];               // event order by the vacancy's boundary / parity class (DevTables::kmc_slot)
  for (int q = threadIdx.x; q < 64 * 12; q += blockDim.x) (&s_slot[0][0])[q] = tab.kmc_slot[q];
  for (int q = threadIdx.x; q < kEnvN; q += blockDim.x) { s_mask_hi[q] = tab.pair_mask_hi[q]; s_pbase[q] = tab.pair_base[q]; }
  for (int q = threadIdx.x; q < 2 * kBoxCells; q += blockDim.x) s_box[q] = tab.box_delta[q];
  for (int q = threadIdx.x; q < 2 * 12 * kBoxCells; q += blockDim.x) s_envpos[q] = tab.box_envpos[q];
  __syncthreads();
  const int lane = threadIdx.x & 15;
  const int wl = threadIdx.x >> 4;                 // walker slot within the block
  // Both half-warps of a warp run the same control flow (a dead walker keeps stepping without side effects), so every
  // shuffle / ballot / syncwarp can use the full mask -- no MATCH.ANY convergence code in the loop.
  const int w_raw = static_cast<int>((blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 4);
  const int w = w_raw < n_walkers ? w_raw : n_walkers - 1;
  const int hshift = threadIdx.x & 16;
  constexpr unsigned hmask = 0xFFFFFFFFu;
  uint8_t *o = occ + w * walker_stride;
  bool alive = w_raw < n_walkers && st.error[w] == 0 && st.vacancy[w] >= 0;
  if (!__any_sync(hmask, alive)) return;

  int X, Y, Z;
  lat.coords_of_id(st.vacancy[w] >= 0 ? st.vacancy[w] : 0, X, Y, Z);
  double time = st.time[w], energy = st.energy[w], temperature = st.temperature[w];
  int64_t steps = st.steps[w];
  const double c_vac = st.c_vacancy[w], c_sol = st.c_solute[w];
  const unsigned solvent = static_cast<unsigned>(tab.solvent), vac_code = static_cast<unsigned>(tab.n_species);
  const int n_species = tab.n_species;
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  const bool active = lane < 12;
  const int k = active ? lane : 0;
  const int dxk = tab.nn1[4 * k], dyk = tab.nn1[4 * k + 1], dzk = tab.nn1[4 * k + 2];
  const int32_t dmig0 = lat.padded_delta(dxk, dyk, dzk, 0), dmig1 = lat.padded_delta(dxk, dyk, dzk, 1);
  double *ord_rate = s_ord_rate[wl];
  uint8_t *ord_lane = s_ord_lane[wl];
  uint16_t *list = s_list[wl];
  uint8_t *my_codes = s_codes[wl * 12 + k];
  const double2 *__restrict__ B_all = reinterpret_cast<const double2 *>(tab.pair_B2);
  // The singlet table A (8.3 KB for 3 species) is read through L1, not staged: with 8 blocks per SM every staged KB costs
  // 8 KB of the 256 KB that shared memory and L1 split, and the walkers' boxes only stay L1-resident if L1 gets ~100 KB
  const double2 *s_A2v = reinterpret_cast<const double2 *>(tab.pair_A2);
  constexpr bool kGlobalA = true;
  const uint2 *s_mask_hi2 = reinterpret_cast<const uint2 *>(s_mask_hi);
  const KmcEvalContext ctx{s_box, s_envpos, s_A2v, s_mask_hi2, s_pbase, B_all, tab.pair_C2,
                           tab.n_pair_pairs * tab.n_species * tab.n_species, n_species, solvent, vac_code, tab.barrier_model};
  const bool tracing = kInstrumented && (tr.from || tr.to || tr.slot || tr.dt || tr.Ea || tr.dE || tr.total_rate || tr.temperature);
  int err = 0;

  double beta = 1.0 / kBoltzmannEv / temperature;
  double corr = prm.rate_corrector ? rate_correction(c_vac, c_sol, temperature) : 1.0;
  double corr_over_prefactor = corr / kPrefactorHz;     // dt = -ln(u1) / total / 1e13 * corr with one division per step
  double ahead_neg_log_u1 = 0.0, ahead_u2 = 0.0;        // lane l: -ln(u1) and u2 of step (s & ~15) + l
  bool handed_off = false;
  for (int64_t s = 0; s < n_steps; ++s) {
    if (prm.handoff_done && (s & 15) == 0 && s > 0) {      // one L2 read per warp and 16 steps; the value is warp-uniform
      int through = 0;
      if ((threadIdx.x & 31) == 0) through = *static_cast<volatile int *>(prm.handoff_done);
      through = __shfl_sync(hmask, through, 0);
      if (through >= prm.handoff_threshold) { handed_off = true; break; }
    }
    // 1. UpdateTemperature (KineticMcAbstract.cpp:45-50): only a T(t) table changes the temperature during a run
    if (prm.n_tt > 0) {
      const double t_now = interpolate_temperature(prm, time);
      if (t_now != temperature) {                   // beyond the ends of the table T(t) is constant: nothing to recompute
        temperature = t_now;
        beta = 1.0 / kBoltzmannEv / temperature;
        if (prm.rate_corrector) { corr = rate_correction(c_vac, c_sol, temperature); corr_over_prefactor = corr / kPrefactorHz; }
      }
    }
    // 2. BuildEventList: event order = ascending lattice id of the neighbour (adjacency lists are sorted) -- a function of the
    // vacancy's boundary / parity class only (table ranked by the host)
    const int xj = wrap_coord(X + dxk, px), yj = wrap_coord(Y + dyk, py), zj = wrap_coord(Z + dzk, pz);
    const int slot = s_slot[(coord_class(X, px) * 4 + coord_class(Y, py)) * 4 + coord_class(Z, pz)][k];
    double ea = 0.0, de = 0.0, rate = 0.0;
    unsigned mig = 0;
    kmc_scan_and_evaluate<false, kGlobalA>(lat, ctx, o, X, Y, Z, lane, active, k, dmig0, dmig1, list, my_codes, beta, 0, 0u, -1,
                                           err, ea, de, rate, mig);
    if ((__ballot_sync(hmask, err != 0) >> hshift) & 0xFFFFu) alive = false;    // this walker stops; its state is left untouched
    if (!__any_sync(hmask, alive)) break;
    // events in the reference's order through shared memory: s_rate[slot] = rate, s_lane[slot] = lane
    if (active) { ord_rate[slot] = rate; ord_lane[slot] = static_cast<uint8_t>(lane); }
    __syncwarp(hmask);
    // total rate and cumulative probabilities in slot order, sequentially (KineticMcFirstOmp.cpp:55-77)
    double total = 0.0;
#pragma unroll
    for (int q = 0; q < 12; ++q) total += ord_rate[q];
    // p_q = rate_q / total, then the running sum in slot order (same operands and order as the reference)
    __syncwarp(hmask);
    if (active) ord_rate[lane] = ord_rate[lane] / total;        // lane q owns slot q from here on
    __syncwarp(hmask);
    double my_cumulative = 0.0;                 // ((p0 + p1) + p2) + ... + p_lane: the reference's running sum at slot `lane`
#pragma unroll
    for (int q = 0; q < 12; ++q)
      if (q <= lane) my_cumulative += ord_rate[q];
    // 3./4. random numbers: u1 -> residence time, u2 -> event (CalculateTime then SelectEvent)
    double neg_log_u1, u2;
    if (kInstrumented && replay_u1) {
      neg_log_u1 = -log(replay_u1[static_cast<int64_t>(w) * n_steps + s]);
      u2 = replay_u2[static_cast<int64_t>(w) * n_steps + s];
    } else {
      // the Philox counter of a step is the walker's step number, so the numbers of the next 16 steps are known in
      // advance: every 16th step lane l draws (and takes the logarithm) for step s + l, and each step then picks its
      // pair with two shuffles instead of running Philox + log on all lanes
      if ((s & 15) == 0) {
        const int64_t ctr = steps + lane;
        uint32_t r[4];
        philox4x32_10(static_cast<uint32_t>(ctr), static_cast<uint32_t>(static_cast<uint64_t>(ctr) >> 32),
                      static_cast<uint32_t>(prm.seed) ^ static_cast<uint32_t>(w), static_cast<uint32_t>(prm.seed >> 32), r);
        ahead_neg_log_u1 = -log(uniform53(r[0], r[1]) + (1.0 / 9007199254740992.0));   // u1 in (0, 1]: finite
        ahead_u2 = uniform53(r[2], r[3]);
      }
      neg_log_u1 = __shfl_sync(hmask, ahead_neg_log_u1, static_cast<int>(s & 15), 16);
      u2 = __shfl_sync(hmask, ahead_u2, static_cast<int>(s & 15), 16);
    }
    const double dt = __dmul_rn(neg_log_u1 / total, corr_over_prefactor);   // never contracted into `time += dt`: both instantiations round alike
    // first slot whose cumulative probability is not < u2, else the last one (KineticMcAbstract.cpp:106-116)
    const unsigned hit = (__ballot_sync(hmask, lane < 12 && !(my_cumulative < u2)) >> hshift) & 0xFFFu;
    const int sel_slot = hit ? (__ffs(static_cast<int>(hit)) - 1) : 11;
    const int sel_lane = ord_lane[sel_slot];
    const double sel_ea = __shfl_sync(hmask, ea, sel_lane, 16), sel_de = __shfl_sync(hmask, de, sel_lane, 16);
    const int nx = __shfl_sync(hmask, xj, sel_lane, 16), ny = __shfl_sync(hmask, yj, sel_lane, 16),
              nz = __shfl_sync(hmask, zj, sel_lane, 16);
    const unsigned sel_mig = __shfl_sync(hmask, mig, sel_lane, 16);
    if (kInstrumented && lane == 0 && alive) {
      if (tracing) {
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
    }
    // 7. Config::LatticeJump: the atom moves into the vacancy, the vacancy into the atom's site.  Lanes 0-7 write the (up to
    // 8) periodic images of the old vacancy site, lanes 8-15 those of the new one.
    if (alive) store_site_image(lat, o, lane < 8 ? X : nx, lane < 8 ? Y : ny, lane < 8 ? Z : nz, lane & 7,
                                static_cast<uint8_t>(lane < 8 ? sel_mig : vac_code));
    if (alive) {
      time += dt;
      energy += sel_de;
      ++steps;
      X = nx; Y = ny; Z = nz;
    }
    __syncwarp(hmask);
  }
  if (prm.handoff_done && !handed_off && lane == 0 && w_raw < n_walkers) atomicAdd(prm.handoff_done, 1);   // through (or stopped by an error)
  if (prm.finish_ns && lane == 0 && w_raw < n_walkers) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    prm.finish_ns[w] = now;
  }
  if (lane == 0 && w_raw < n_walkers && (alive || err == 0) && st.error[w] == 0) {
    st.vacancy[w] = lat.id_of_coords(X, Y, Z);
    st.time[w] = time;
    st.energy[w] = energy;
    st.steps[w] = steps;
    st.temperature[w] = temperature;
    st.previous[w] = -1;        // a first-order run leaves no second-order history
  }
  if (err && w_raw < n_walkers) atomicOr(&st.error[w], err);
}

// ------------------------------------------------------------------------------------------------ batched event lists
// KineticMcFirstOmp::BuildEventList (mc/src/KineticMcFirstOmp.cpp:52-68) for many vacancies at once: item q is the vacancy
// at lattice id vacancy[q] of replica walker[q]; a half-warp evaluates its 12 jumps with ONE scan of the surrounding box
// (kmc_scan_and_evaluate) instead of 12 independent 60-site gathers.  Outputs per item, in the reference's event order
// (ascending neighbour lattice id): neighbour id, Ea, dE.  Blocks loop over items so that the shared-memory tables are
// staged once per block.
__global__ void __launch_bounds__(kKmcThreads, 7)
vacancy_events_kernel(LatticeDesc lat, DevTables tab, const uint8_t *__restrict__ occ, int64_t walker_stride, int64_t n_items,
                      const int32_t *__restrict__ walker, const int64_t *__restrict__ vacancy, int64_t *__restrict__ neighbour,
                      double *__restrict__ Ea, double *__restrict__ dE, int *error) {
  __shared__ int32_t s_box[2 * kBoxCells];
  __shared__ int8_t s_envpos[2 * 12 * kBoxCells];
  __shared__ uint16_t s_list[kKmcWalkersPerBlock][kBoxCells];
  __shared__ uint8_t s_codes[kKmcThreads][kEnvN + 2];
  extern __shared__ double s_A2[];
  __shared__ uint64_t s_mask_hi[kEnvN];
  __shared__ uint16_t s_pbase[kEnvN];
  __shared__ __align__(16) uint32_t s_ids[kKmcWalkersPerBlock][12];
  for (int q = threadIdx.x; q < tab.n_species * kEnvN * tab.n_species * 2; q += blockDim.x) s_A2[q] = tab.pair_A2[q];
  for (int q = threadIdx.x; q < kEnvN; q += blockDim.x) { s_mask_hi[q] = tab.pair_mask_hi[q]; s_pbase[q] = tab.pair_base[q]; }
  for (int q = threadIdx.x; q < 2 * kBoxCells; q += blockDim.x) s_box[q] = tab.box_delta[q];
  for (int q = threadIdx.x; q < 2 * 12 * kBoxCells; q += blockDim.x) s_envpos[q] = tab.box_envpos[q];
  __syncthreads();
  const int lane = threadIdx.x & 15, wl = threadIdx.x >> 4;
  constexpr unsigned hmask = 0xFFFFFFFFu;
  const unsigned solvent = static_cast<unsigned>(tab.solvent), vac_code = static_cast<unsigned>(tab.n_species);
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  const bool active = lane < 12;
  const int k = active ? lane : 0;
  const int dxk = tab.nn1[4 * k], dyk = tab.nn1[4 * k + 1], dzk = tab.nn1[4 * k + 2];
  const int32_t dmig0 = lat.padded_delta(dxk, dyk, dzk, 0), dmig1 = lat.padded_delta(dxk, dyk, dzk, 1);
  const KmcEvalContext ctx{s_box, s_envpos, reinterpret_cast<const double2 *>(s_A2), reinterpret_cast<const uint2 *>(s_mask_hi), s_pbase,
                           reinterpret_cast<const double2 *>(tab.pair_B2), tab.pair_C2, tab.n_pair_pairs * tab.n_species * tab.n_species,
                           tab.n_species, solvent, vac_code, tab.barrier_model};
  const int64_t slots = static_cast<int64_t>(gridDim.x) * kKmcWalkersPerBlock;
  const int64_t rounds = (n_items + slots - 1) / slots;          // every half-warp runs the same number of rounds (full-mask shuffles)
  for (int64_t r = 0; r < rounds; ++r) {
    const int64_t item_raw = r * slots + static_cast<int64_t>(blockIdx.x) * kKmcWalkersPerBlock + wl;
    const bool real = item_raw < n_items;
    const int64_t item = real ? item_raw : n_items - 1;
    const int64_t vac_id = vacancy[item];
    int err = 0;
    const bool id_ok = vac_id >= 0 && vac_id < lat.num_sites;
    if (!id_ok) err |= kErrBadSite;
    const uint8_t *o = occ + (walker ? walker[item] : 0) * walker_stride;
    int X, Y, Z;
    lat.coords_of_id(id_ok ? vac_id : 0, X, Y, Z);
    const int xj = wrap_coord(X + dxk, px), yj = wrap_coord(Y + dyk, py), zj = wrap_coord(Z + dzk, pz);
    const uint32_t id_j = static_cast<uint32_t>(lat.id_of_coords(xj, yj, zj));
    if (active) s_ids[wl][lane] = id_j;
    __syncwarp(hmask);
    int slot = 0;
    {
      const uint4 *idv = reinterpret_cast<const uint4 *>(s_ids[wl]);    // 12 ids = three 16-byte loads
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const uint4 v = idv[q];
        slot += (v.x < id_j ? 1 : 0) + (v.y < id_j ? 1 : 0) + (v.z < id_j ? 1 : 0) + (v.w < id_j ? 1 : 0);
      }
    }
    double ea = 0.0, de = 0.0, rate = 0.0;
    unsigned mig = 0;
    kmc_scan_and_evaluate<false>(lat, ctx, o, X, Y, Z, lane, active, k, dmig0, dmig1, s_list[wl], s_codes[threadIdx.x], 0.0,
                                 0, 0u, -1, err, ea, de, rate, mig);
    const bool bad = ((__ballot_sync(hmask, err != 0) >> (threadIdx.x & 16)) & 0xFFFFu) != 0;     // any lane of this item
    if (real && active) {
      const int64_t at = item * 12 + slot;
      neighbour[at] = id_j;
      Ea[at] = bad ? nan("") : ea;
      dE[at] = bad ? nan("") : de;
    }
    if (real && err) atomicOr(error, err);
    __syncwarp(hmask);
  }
}

// ------------------------------------------------------------------------------------------------ second-order KMC
// mc::KineticMcChainOmpi (mc/src/KineticMcChainOmpi.cpp:56-152, mc/include/KineticMcAbstract.h:65-143): the reference
// spreads the 12 first neighbours i of the vacancy site k over 12 MPI ranks; rank r moves the vacancy to i, evaluates
// the 12 jumps i -> l there (one of them is the jump back, whose reverse is the event k -> i), moves it back, and the
// ranks combine eight sums (MpiData) into the second-order residence time t_2 and event probabilities that remove the
// k <-> i flicker.  Here ONE BLOCK of 12 half-warps owns one walker: half-warp h plays rank "direction h" (lane = jump
// i -> l), reads the moved configuration through the hypothetical view of kmc_scan_and_evaluate (no memory writes), and
// the per-rank results meet in shared memory in the reference's rank order (ascending lattice id of i).  One uniform
// per step (SelectEvent; the second-order time is an expectation, not a sample).
constexpr int kChainThreads = 192;

__global__ void __launch_bounds__(kChainThreads, 4)
kmc_chain_run_kernel(LatticeDesc lat, DevTables tab, uint8_t *occ, int64_t walker_stride, int n_walkers, KmcState st, KmcParams prm,
                     int64_t n_steps, const double *__restrict__ replay_u, KmcTraceDev tr) {
  __shared__ int32_t s_box[2 * kBoxCells];
  __shared__ int8_t s_envpos[2 * 12 * kBoxCells];
  __shared__ uint16_t s_list[12][kBoxCells];
  __shared__ uint8_t s_codes[kChainThreads][kEnvN + 2];
  __shared__ double s_ord_rate[12][12];
  extern __shared__ double s_A2[];
  __shared__ uint64_t s_mask_hi[kEnvN];
  __shared__ uint16_t s_pbase[kEnvN];
  __shared__ __align__(16) uint32_t s_ids[12][12];
  // per-rank results in rank (slot) order
  __shared__ double s_fwd[12], s_bwd[12], s_total_i[12], s_barrier[12], s_de[12];
  __shared__ uint8_t s_dir[12], s_is_prev[12], s_mig[12];
  __shared__ double s_terms[8][12];
  __shared__ int s_sel_dir;
  __shared__ double s_sel_dt, s_sel_de;
  for (int q = threadIdx.x; q < tab.n_species * kEnvN * tab.n_species * 2; q += blockDim.x) s_A2[q] = tab.pair_A2[q];
  for (int q = threadIdx.x; q < kEnvN; q += blockDim.x) { s_mask_hi[q] = tab.pair_mask_hi[q]; s_pbase[q] = tab.pair_base[q]; }
  for (int q = threadIdx.x; q < 2 * kBoxCells; q += blockDim.x) s_box[q] = tab.box_delta[q];
  for (int q = threadIdx.x; q < 2 * 12 * kBoxCells; q += blockDim.x) s_envpos[q] = tab.box_envpos[q];
  __syncthreads();
  const int w = blockIdx.x;
  if (w >= n_walkers || st.error[w] != 0 || st.vacancy[w] < 0) return;     // uniform over the block
  const int lane = threadIdx.x & 15;
  const int h = threadIdx.x >> 4;                  // rank of this half-warp = direction k -> i
  const int hshift = threadIdx.x & 16;
  constexpr unsigned hmask = 0xFFFFFFFFu;
  uint8_t *o = occ + w * walker_stride;

  int X, Y, Z;
  lat.coords_of_id(st.vacancy[w], X, Y, Z);
  double time = st.time[w], energy = st.energy[w], temperature = st.temperature[w];
  int64_t steps = st.steps[w];
  int64_t previous = st.previous[w];
  const double c_vac = st.c_vacancy[w], c_sol = st.c_solute[w];
  const unsigned solvent = static_cast<unsigned>(tab.solvent), vac_code = static_cast<unsigned>(tab.n_species);
  const int n_species = tab.n_species;
  const int px = 2 * lat.fx, py = 2 * lat.fy, pz = 2 * lat.fz;
  const bool active = lane < 12;
  const int k = active ? lane : 0;
  const int dxk = tab.nn1[4 * k], dyk = tab.nn1[4 * k + 1], dzk = tab.nn1[4 * k + 2];            // lane's direction
  const int dxh = tab.nn1[4 * h], dyh = tab.nn1[4 * h + 1], dzh = tab.nn1[4 * h + 2];            // half-warp's direction
  const int back_dir = tab.dir_lut[(-dxh + 1) * 9 + (-dyh + 1) * 3 + (-dzh + 1)];                // the jump i -> k
  const int32_t dmig0 = lat.padded_delta(dxk, dyk, dzk, 0), dmig1 = lat.padded_delta(dxk, dyk, dzk, 1);
  const double2 *__restrict__ B_all = reinterpret_cast<const double2 *>(tab.pair_B2);
  const KmcEvalContext ctx{s_box, s_envpos, reinterpret_cast<const double2 *>(s_A2), reinterpret_cast<const uint2 *>(s_mask_hi), s_pbase,
                           B_all, tab.pair_C2, tab.n_pair_pairs * tab.n_species * tab.n_species, n_species, solvent, vac_code, tab.barrier_model};
  const bool tracing = tr.from || tr.to || tr.slot || tr.dt || tr.Ea || tr.dE || tr.total_rate || tr.temperature;
  int err = 0;

  double beta = 1.0 / kBoltzmannEv / temperature;
  double corr = prm.rate_corrector ? rate_correction(c_vac, c_sol, temperature) : 1.0;
  double ahead_u = 0.0;                            // first half-warp, lane l: the uniform of step (s & ~15) + l
  for (int64_t s = 0; s < n_steps; ++s) {
    if (prm.n_tt > 0) {                            // UpdateTemperature (KineticMcAbstract.cpp:45-50)
      const double t_now = interpolate_temperature(prm, time);
      if (t_now != temperature) {                   // beyond the ends of the table T(t) is constant: nothing to recompute
        temperature = t_now;
        beta = 1.0 / kBoltzmannEv / temperature;
        if (prm.rate_corrector) corr = rate_correction(c_vac, c_sol, temperature);
      }
    }
    // ---- ranks: the 12 neighbours i of k in ascending lattice-id order
    const int xn = wrap_coord(X + dxk, px), yn = wrap_coord(Y + dyk, py), zn = wrap_coord(Z + dzk, pz);
    const uint32_t id_n = static_cast<uint32_t>(lat.id_of_coords(xn, yn, zn));   // lane's neighbour of k
    const uint32_t id_i = __shfl_sync(hmask, id_n, h, 16);
    const int rank = __popc((__ballot_sync(hmask, active && id_n < id_i) >> hshift) & 0xFFFu);
    if (previous < 0) {                            // previous_j_ starts as first neighbour 0 of the vacancy (KineticMcAbstract.cpp:255)
      uint32_t lo = active ? id_n : 0xFFFFFFFFu;
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) lo = min(lo, __shfl_xor_sync(hmask, lo, off, 16));
      previous = lo;
    }
    // ---- this rank's hypothetical state: vacancy at i, the atom of i at k
    const int xi = wrap_coord(X + dxh, px), yi = wrap_coord(Y + dyh, py), zi = wrap_coord(Z + dzh, pz);
    const int zpk = Z & 1, zpi = zi & 1;
    const int64_t base_k = lat.padded_index(X, Y, Z), base_i = lat.padded_index(xi, yi, zi);
    const unsigned mig_i = o[base_k + lat.padded_delta(dxh, dyh, dzh, zpk)];
    if (mig_i == vac_code) err |= kErrNotVacancy;
    if (lane == 0 && o[base_k] != vac_code) err |= kErrNotVacancy;
    const int64_t hypo_index = base_i + lat.padded_delta(-dxh, -dyh, -dzh, zpi);
    // event order of the jumps i -> l: ascending lattice id of l
    const int xl = wrap_coord(xi + dxk, px), yl = wrap_coord(yi + dyk, py), zl = wrap_coord(zi + dzk, pz);
    const uint32_t id_l = static_cast<uint32_t>(lat.id_of_coords(xl, yl, zl));
    if (active) s_ids[h][lane] = id_l;
    __syncwarp(hmask);
    int slot = 0;
    {
      const uint4 *idv = reinterpret_cast<const uint4 *>(s_ids[h]);     // 12 ids = three 16-byte loads
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const uint4 v = idv[q];
        slot += (v.x < id_l ? 1 : 0) + (v.y < id_l ? 1 : 0) + (v.z < id_l ? 1 : 0) + (v.w < id_l ? 1 : 0);
      }
    }
    double ea = 0.0, de = 0.0, rate = 0.0;
    unsigned mig = 0;
    kmc_scan_and_evaluate<true>(lat, ctx, o, xi, yi, zi, lane, active, k, dmig0, dmig1, s_list[h],
                                s_codes[threadIdx.x], beta, hypo_index, mig_i, back_dir, err, ea, de, rate, mig);
    // total_rate_i in event order (KineticMcChainOmpi.cpp:71-85)
    if (active) s_ord_rate[h][slot] = rate;
    __syncwarp(hmask);
    double total_i = 0.0;
#pragma unroll
    for (int q = 0; q < 12; ++q) total_i += s_ord_rate[h][q];
    // event k -> i = reverse of i -> k (JumpEvent::GetReverseJumpEvent, JumpEvent.cpp:55-58)
    const double ea_ik = __shfl_sync(hmask, ea, back_dir, 16), de_ik = __shfl_sync(hmask, de, back_dir, 16);
    if (lane == 0) {
      const double barrier = ea_ik - de_ik, change = -de_ik;
      s_barrier[rank] = barrier;
      s_de[rank] = change;
      s_total_i[rank] = total_i;
      s_dir[rank] = static_cast<uint8_t>(h);
      s_is_prev[rank] = static_cast<int64_t>(id_i) == previous ? 1 : 0;
      s_mig[rank] = static_cast<uint8_t>(mig_i);
    }
    if (__syncthreads_or(err != 0)) break;                       // the walker stops; its state is left untouched
    // ---- CalculateTime + SelectEvent (KineticMcChainOmpi.cpp:93-151, KineticMcAbstract.cpp:106-123) in the first
    // half-warp: lane r owns rank r; every sum runs sequentially in rank order like the reference's reductions
    if (threadIdx.x < 16) {
      // the 12 forward / backward rates of the k -> i events in ONE instruction stream (lane r = rank r) instead of one
      // lane-0 stream in each of the six warps
      if (active) {
        const double barrier = s_barrier[lane], change = s_de[lane];
        s_fwd[lane] = exp(-barrier * beta);                      // forward_rate_ (JumpEvent.cpp:13)
        s_bwd[lane] = exp((change - barrier) * beta);            // GetBackwardRate (:28-30)
      }
      __syncwarp(0xFFFFu);
      double total_k = 0.0;
#pragma unroll
      for (int r = 0; r < 12; ++r) total_k += s_fwd[r];
      const double t_1 = 1.0 / total_k / kPrefactorHz;
      const int r = active ? lane : 0;
      const double p_ki = s_fwd[r] / total_k, p_ik = s_bwd[r] / s_total_i[r];
      const double bb = p_ki * p_ik, b = p_ki * (1 - p_ik);
      const bool prev = s_is_prev[r] != 0;
      const double t_i = 1.0 / s_total_i[r] / kPrefactorHz;
      const double ts_term = (t_1 + t_i) * bb;
      if (active) {                                              // MpiData of rank r (KineticMcChainOmpi.cpp:104-113)
        s_terms[0][r] = bb; s_terms[1][r] = b; s_terms[2][r] = prev ? 0.0 : bb; s_terms[3][r] = prev ? 0.0 : b;
        s_terms[4][r] = prev ? b : 0.0; s_terms[5][r] = prev ? p_ki : 0.0; s_terms[6][r] = ts_term; s_terms[7][r] = prev ? 0.0 : ts_term;
      }
      __syncwarp(0xFFFFu);
      double mine = 0.0;                                         // lane q < 8 reduces field q in rank order (DataSum)
      if (lane < 8) {
        mine = s_terms[lane][0];
#pragma unroll
        for (int q = 1; q < 12; ++q) mine += s_terms[lane][q];
      }
      const double beta_bar_k = __shfl_sync(0xFFFFu, mine, 0, 16), beta_k = __shfl_sync(0xFFFFu, mine, 1, 16),
                   gamma_bar_k_j = __shfl_sync(0xFFFFu, mine, 2, 16), gamma_k_j = __shfl_sync(0xFFFFu, mine, 3, 16),
                   beta_k_j = __shfl_sync(0xFFFFu, mine, 4, 16), alpha_k_j = __shfl_sync(0xFFFFu, mine, 5, 16),
                   ts_num = __shfl_sync(0xFFFFu, mine, 6, 16), ts_j_num = __shfl_sync(0xFFFFu, mine, 7, 16);
      const double ts = ts_num / beta_bar_k, ts_j = ts_j_num / gamma_bar_k_j;
      const double inv = 1 / (1 - alpha_k_j);
      const double t_2 = inv * (gamma_k_j * t_1 + gamma_bar_k_j * (ts_j + t_1 + beta_bar_k / beta_k * ts));
      // second-order probability of rank r, then the running sum in rank order
      const double prob = prev ? inv * (gamma_bar_k_j / beta_k) * beta_k_j : inv * (1 + gamma_bar_k_j / beta_k) * b;
      __syncwarp(0xFFFFu);
      if (active) s_terms[0][r] = prob;
      __syncwarp(0xFFFFu);
      double cumulative = 0.0;
#pragma unroll
      for (int q = 0; q < 12; ++q)
        if (q <= lane) cumulative += s_terms[0][q];
      double u;
      if (replay_u) u = replay_u[static_cast<int64_t>(w) * n_steps + s];
      else {
        // the Philox counter is the step number: lane l draws for step s + l every 16th step (kmc_run_kernel does the same)
        if ((s & 15) == 0) {
          const int64_t ctr = steps + lane;
          uint32_t r4[4];
          philox4x32_10(static_cast<uint32_t>(ctr), static_cast<uint32_t>(static_cast<uint64_t>(ctr) >> 32),
                        static_cast<uint32_t>(prm.seed) ^ static_cast<uint32_t>(w), static_cast<uint32_t>(prm.seed >> 32), r4);
          ahead_u = uniform53(r4[2], r4[3]);
        }
        u = __shfl_sync(0xFFFFu, ahead_u, static_cast<int>(s & 15), 16);
      }
      const unsigned hit = __ballot_sync(0xFFFFu, active && !(cumulative < u)) & 0xFFFu;
      const int sel = hit ? (__ffs(static_cast<int>(hit)) - 1) : 11;
      const int sel_dir = s_dir[sel];
      const int nx = wrap_coord(X + tab.nn1[4 * sel_dir], px), ny = wrap_coord(Y + tab.nn1[4 * sel_dir + 1], py),
                nz = wrap_coord(Z + tab.nn1[4 * sel_dir + 2], pz);
      const double dt = t_2 * corr;
      if (lane == 0) {
        s_sel_dir = sel_dir;
        s_sel_dt = dt;
        s_sel_de = s_de[sel];
        if (tracing) {
          const int64_t at = static_cast<int64_t>(w) * n_steps + s;
          if (tr.from) tr.from[at] = lat.id_of_coords(X, Y, Z);
          if (tr.to) tr.to[at] = lat.id_of_coords(nx, ny, nz);
          if (tr.slot) tr.slot[at] = sel;
          if (tr.dt) tr.dt[at] = dt;
          if (tr.Ea) tr.Ea[at] = s_barrier[sel];
          if (tr.dE) tr.dE[at] = s_de[sel];
          if (tr.total_rate) tr.total_rate[at] = total_k;
          if (tr.temperature) tr.temperature[at] = temperature;
        }
      }
      // Config::LatticeJump: lanes 0-7 write the images of the old vacancy site, lanes 8-15 those of the new one
      store_site_image(lat, o, lane < 8 ? X : nx, lane < 8 ? Y : ny, lane < 8 ? Z : nz, lane & 7,
                       static_cast<uint8_t>(lane < 8 ? s_mig[sel] : vac_code));
    }
    __syncthreads();                               // the selection and the jump are visible to every half-warp
    const int sel_dir = s_sel_dir;
    previous = lat.id_of_coords(X, Y, Z);          // KineticMcChainAbstract::OneStepSimulation (KineticMcAbstract.cpp:260-263)
    time += s_sel_dt;
    energy += s_sel_de;
    ++steps;
    X = wrap_coord(X + tab.nn1[4 * sel_dir], px);
    Y = wrap_coord(Y + tab.nn1[4 * sel_dir + 1], py);
    Z = wrap_coord(Z + tab.nn1[4 * sel_dir + 2], pz);
  }
  if (err) atomicOr(&st.error[w], err);
  else if (threadIdx.x == 0 && st.error[w] == 0) {
    st.vacancy[w] = lat.id_of_coords(X, Y, Z);
    st.time[w] = time;
    st.energy[w] = energy;
    st.steps[w] = steps;
    st.temperature[w] = temperature;
    st.previous[w] = previous;
  }
}

}  // namespace lmc

/* lmc_b200.h -- C ABI of the B200-native LatticeMC hot-path engine (liblmc_b200.so).
 *
 * The reference (zhucongx/LatticeMonteCarlo) has no plugin / FFI boundary: its hot path is reached through
 * three C++ predictor classes and the mc:: drivers.  This header is the boundary a
 * reference-side binding would use instead; every entry point names the reference interface it replaces.
 * INTEGRATION.md shows the C++ adapter classes (same names / signatures as the reference's) that forward here.
 *
 * Conventions
 *  - all functions return 0 on success, a negative lmc_status on failure; lmc_last_error() gives the message
 *    (thread local).  The C++ adapters re-throw the reference's exception types from these codes.
 *  - element codes are the reference's ElementName enum values (lmc/cfg/include/Element.hpp:7):
 *    X(vacancy)=0, Al=1, Mg=2, Zn=3, Cu=4, Sn=5, pAl=6 ... pSn=10.
 *  - lattice ids are the reference's lattice ids in one of its two id orders (lmc_id_order).
 *  - host pointers unless the name ends in _dev; buffers are caller-owned; calls are synchronous.
 *  - one engine per GPU; an engine is not thread safe (callers serialise), engines are independent.
 *  - there is NO CPU fallback: every compute entry point fails with LMC_ERR_NO_DEVICE if the engine was created
 *    without a CUDA device (device < 0 is allowed only for the lmc_tables_* / geometry queries below).
 */
#ifndef LMC_B200_H_
#define LMC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lmc_engine lmc_engine;

typedef enum lmc_status {
  LMC_OK = 0,
  LMC_ERR_INVALID_ARGUMENT = -1, /* std::invalid_argument in the adapters */
  LMC_ERR_OUT_OF_RANGE = -2,     /* std::out_of_range: non-neighbour jump pair (VacancyMigrationPredictorQuartic.cpp:281-283),
                                    cluster type without index, e.g. two vacancies in range (EnergyUtility.cpp:815-817) */
  LMC_ERR_RUNTIME = -3,          /* std::runtime_error: "Cannot open <file>", JSON errors */
  LMC_ERR_NO_DEVICE = -4,        /* compute call on a host-only engine / CUDA extension unusable */
  LMC_ERR_CUDA = -5              /* CUDA runtime failure */
} lmc_status;

typedef enum lmc_id_order {
  LMC_ID_ORDER_GENERATE = 0,   /* cfg::GenerateFCC order (cfg/src/Config.cpp:1073-1090); used by SimulatedAnnealing */
  LMC_ID_ORDER_REASSIGNED = 1  /* Config::ReassignLatticeVector order (Config.cpp:466-552); every run started from a .cfg */
} lmc_id_order;

const char *lmc_last_error(void);
int lmc_abi_version(void);
/* number of CUDA devices visible to the library (0 if none) */
int lmc_device_count(void);

/* ------------------------------------------------------------------------------------------------ engine
 * Replaces construction of cfg::Config neighbour lists (Config::UpdateNeighbors, Config.cpp:955-1045) and of the
 * per-site tables in the predictor constructors (VacancyMigrationPredictorQuartic.cpp:64-100,
 * EnergyChangePredictorPairSite.cpp:41-57): geometry becomes constant offset tables, occupancy one uint8 per site.
 *   factors        supercell size in conventional FCC cells per axis (each >= 4, like the reference's cell list)
 *   id_order       lmc_id_order
 *   element_set    the `element_set` of the parameter file (ElementName codes, vacancy excluded), n_elements <= 7
 *   solvent        ElementName code of the majority species (expansion origin of the contracted tables; any member of
 *                  element_set gives identical results, the majority species gives the best speed); 0 = first in set
 *   n_walkers      number of independent replicas of the lattice held on the device (>= 1)
 *   device         CUDA device ordinal, or -1 for a host-only engine (table / geometry queries only)
 */
int lmc_engine_create(lmc_engine **out, const int32_t factors[3], int32_t id_order, const int32_t *element_set,
                      int32_t n_elements, int32_t solvent, int32_t n_walkers, int32_t device);
void lmc_engine_destroy(lmc_engine *engine);
int64_t lmc_engine_num_sites(const lmc_engine *engine);
int32_t lmc_engine_num_walkers(const lmc_engine *engine);

/* JSON coefficient file in the reference's format: {"Base":{"theta":[...]}, "<El>":{"mu_x_mmm":..,"U_mmm":..,..}}
 * (parsed like pred/src/VacancyMigrationPredictorQuartic.cpp:38-63, EnergyChangePredictorPairSite.cpp:29-40,
 * EnergyPredictor.cpp:27-38).  Missing file -> LMC_ERR_RUNTIME "Cannot open <file>". */
int lmc_engine_load_coefficients(lmc_engine *engine, const char *json_path);
/* The same with the barrier model named: the reference selects it by predictor class.
 *   LMC_BARRIER_QUARTIC  pred::VacancyMigrationPredictorQuartic[Lru] (the model every live driver uses, KineticMcAbstract.h:50);
 *                        keys mu_x_mmm sigma_x_mmm U_mmm mu_x_mm2 sigma_x_mm2 U_mm2 theta_D theta_Ks mu_D sigma_D mu_Ks sigma_Ks
 *   LMC_BARRIER_E0       pred::VacancyMigrationPredictorE0[Lru] (pred/src/VacancyMigrationPredictorE0.cpp:9-160, JSON keys :30-37
 *                        mu_x_mmm sigma_x_mmm U_mmm theta_e0 mu_e0 sigma_e0): dE as above, e0 = exp(mu_e0 + sigma_e0 theta_e0 . U_mmm x^),
 *                        Ea = max(0, e0 + dE / 2) (:153-159).  lmc_eval_barriers then returns D = 1 and Ks = e0; the KMC drivers
 *                        run on either model.
 * "Base".theta (the dE part, also used by lmc_eval_swap_de / lmc_eval_site_de / lmc_total_energy, i.e. by
 * pred::EnergyChangePredictorPairSite / Pair / Site and pred::EnergyPredictor) is read in both cases. */
typedef enum lmc_barrier_model { LMC_BARRIER_QUARTIC = 0, LMC_BARRIER_E0 = 1 } lmc_barrier_model;
int lmc_engine_load_coefficients_model(lmc_engine *engine, const char *json_path, int32_t model);
/* the barrier model of the coefficients currently loaded (an engine holds ONE at a time), or -1 if the file had no
 * per-element barrier blocks */
int lmc_engine_barrier_model(const lmc_engine *engine);

/* Occupancy by lattice id (ElementName codes), one replica ("walker") at a time or all at once (n_walkers*N bytes).
 * Replaces cfg::Config::{SetAtomElementTypeAtLattice, GetElementAtLatticeId} (Config.cpp:132-135,462-464). */
int lmc_engine_set_occupancy(lmc_engine *engine, int32_t walker, const uint8_t *occupancy, int64_t n);
int lmc_engine_get_occupancy(lmc_engine *engine, int32_t walker, uint8_t *occupancy, int64_t n);
int lmc_engine_set_occupancy_all(lmc_engine *engine, const uint8_t *occupancy, int64_t n_total);
int lmc_engine_get_occupancy_all(lmc_engine *engine, uint8_t *occupancy, int64_t n_total);
/* Config::LatticeJump (Config.cpp:431-456) as far as occupancy is concerned: exchange the species of two sites */
int lmc_engine_lattice_jump(lmc_engine *engine, int32_t walker, int64_t site_a, int64_t site_b);

/* ------------------------------------------------------------------------------------------------ hot path
 * VacancyMigrationPredictorQuartic::GetBarrierAndDiffFromLatticeIdPair (pred/include/VacancyMigrationPredictorQuartic.h:25-27,
 * src :247-276) for a batch of candidate events: event e is the exchange of the vacancy at site_i[e] with the atom at
 * site_j[e] on replica walker[e] (walker == NULL: replica 0).  D / Ks (GetD :216-246, GetKs :167-215) are optional.
 * Replaces the LRU-cached predictor (VacancyMigrationPredictorQuarticLru.cpp:11-49) by batch recomputation.
 * A non-neighbour pair, a first site that is not a vacancy, or a second vacancy within range makes the call
 * return LMC_ERR_OUT_OF_RANGE (outputs of the offending events are NaN, all others are valid). */
int lmc_eval_barriers(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site_i, const int64_t *site_j,
                      double *Ea, double *dE, double *D, double *Ks);
/* same with every pointer in device memory (inputs already resident in HBM; asynchronous on the engine stream) */
int lmc_eval_barriers_dev(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site_i,
                          const int64_t *site_j, double *Ea, double *dE, double *D, double *Ks);

/* The event list of a vacancy in one call: KineticMcFirstOmp::BuildEventList (mc/src/KineticMcFirstOmp.cpp:52-68) evaluates
 * the 12 jumps (vacancy, first neighbour) of the current vacancy site; here for a batch of n vacancies (vacancy_site[q] on
 * replica walker[q]; walker == NULL: replica 0).  Outputs are [n][12] in the reference's event order (ascending neighbour
 * lattice id = Config::GetFirstNeighborsAdjacencyList order): the neighbour ids, Ea and dE.  The 12 jumps share one scan of
 * the 7x7x7 half-unit box around the vacancy instead of 12 separate 60-site gathers (the fast path of the KMC drivers).
 * Errors as lmc_eval_barriers (offending items are NaN). */
int lmc_eval_vacancy_events(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *vacancy_site, int64_t *neighbour_site,
                            double *Ea, double *dE);
/* same with every pointer in device memory (asynchronous on the engine stream) */
int lmc_eval_vacancy_events_dev(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *vacancy_site,
                                int64_t *neighbour_site, double *Ea, double *dE);

/* EnergyChangePredictorPairSite::GetDeFromLatticeIdPair (pred/include/EnergyChangePredictorPairSite.h:20-23,
 * src :70-146): energy change of exchanging the species at site_a[e] and site_b[e]; 0 for equal species; coupled
 * pairs (b within the third shell of a) are evaluated exactly. */
int lmc_eval_swap_de(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site_a, const int64_t *site_b,
                     double *dE);
int lmc_eval_swap_de_dev(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site_a,
                         const int64_t *site_b, double *dE);
/* EnergyChangePredictorPair::GetDeFromLatticeIdPair (pred/include/EnergyChangePredictorPair.h:22-25, src :69-122): the same
 * exchange energy restricted to FIRST-NEIGHBOUR pairs, the only pairs that predictor holds a cluster list for: equal species
 * give 0 (:71-75); an unlike pair that is not a first-neighbour pair returns LMC_ERR_OUT_OF_RANGE (the reference's
 * unordered_map::at throws std::out_of_range, :83) with NaN in the offending outputs. */
int lmc_eval_pair_de(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site_a, const int64_t *site_b,
                     double *dE);
/* EnergyChangePredictorPairSite::GetDeFromLatticeIdSite (:154-192) == EnergyChangePredictorSite::GetDeFromLatticeIdSite
 * (pred/src/EnergyChangePredictorSite.cpp:56-98): species at `site` replaced by new_element */
int lmc_eval_site_de(lmc_engine *engine, int64_t n, const int32_t *walker, const int64_t *site,
                     const uint8_t *new_element, double *dE);
/* Config::GetElementAtLatticeId (cfg/src/Config.cpp:132-135) for a list of sites: ElementName codes */
int lmc_engine_get_elements(lmc_engine *engine, int32_t walker, int64_t n, const int64_t *lattice_ids, uint8_t *elements);
/* Config::GetVacancyLatticeId (cfg/src/Config.cpp:296-310) generalised: the lowest lattice id that holds `element` (-1 if
 * none; values below -1000 are error codes - 1000); *count (optional) receives the number of such sites */
int64_t lmc_engine_find_element(lmc_engine *engine, int32_t walker, int32_t element, int64_t *count);
/* path of the coefficient file whose tables are on the engine ("" if none): lets several predictor objects that share an
 * engine notice that another file has been loaded since */
const char *lmc_engine_coefficients_path(const lmc_engine *engine);
/* EnergyPredictor::GetEnergy / GetEncode (pred/src/EnergyPredictor.cpp:40-96,173-177) of one replica.
 * counts (optional, n_types int64) receives the exact integer cluster counts of GetEncode before normalisation. */
int lmc_total_energy(lmc_engine *engine, int32_t walker, double *energy, int64_t *counts, int32_t n_types);
/* EnergyPredictor::GetEnergyOfCluster / GetEncodeOfCluster (pred/src/EnergyPredictor.cpp:97-172,178-184): the energy of the
 * clusters that lie entirely inside the site set { listed sites and their first- to third-neighbour shells }.  The sites
 * are LATTICE ids (the reference takes atom ids and maps them with Config::GetLatticeIdFromAtomId; the adapters do that).
 * n == 0 gives 0.  counts as in lmc_total_energy. */
int lmc_energy_of_cluster(lmc_engine *engine, int32_t walker, const int64_t *lattice_ids, int64_t n, double *energy, int64_t *counts,
                          int32_t n_types);
/* EnergyPredictor::GetEncode (lattice_ids == NULL and n < 0: the whole configuration) / GetEncodeOfCluster: the cluster counts
 * divided by the per-label normalisers {256, 3072, 1536, 6144, 12288, 6144, 12288, 6144, ...} (EnergyPredictor.cpp:8), in
 * ClusterIndexer order (n_types doubles) -- the vector the reference dots with Base.theta. */
int lmc_energy_encode(lmc_engine *engine, int32_t walker, const int64_t *lattice_ids, int64_t n, double *encode, int32_t n_types);
/* EnergyPredictor::GetChemicalPotential(solvent) (pred/src/EnergyPredictor.cpp:196-214): for every element of the set and the
 * vacancy X, E(15 x 15 x 15 solvent cell with atom 0 replaced by the element) - E(pure solvent cell); 0 for the solvent.
 * Entries in element-name order (std::map<Element, double>: Element::operator< compares names).  Returns the number of
 * entries (call with elements == NULL for the length), negative on error. */
int32_t lmc_chemical_potential(lmc_engine *engine, int32_t solvent_element, int32_t *elements, double *mu, int32_t capacity);

/* ------------------------------------------------------------------------------------------------ measurement
 * the engine's cudaStream_t (all engine work is enqueued on it; e.g. to record CUDA events around calls) */
void *lmc_engine_cuda_stream(lmc_engine *engine);
int lmc_engine_synchronize(lmc_engine *engine);
/* device time (CUDA events on the engine stream) of the most recent hot-path kernel launch:
 * barrier_kernel / swap_de_kernel / kmc_run_kernel / cmc_run_kernel.  Blocks until that kernel has finished. */
double lmc_engine_last_kernel_ms(lmc_engine *engine);
/* number of kernels this engine has launched so far */
int64_t lmc_engine_launch_count(const lmc_engine *engine);

/* ------------------------------------------------------------------------------------------------ KMC driver
 * mc::KineticMcFirstOmp::Simulate (mc/src/KineticMcAbstract.cpp:140-188, mc/src/KineticMcFirstOmp.cpp:52-82) run for
 * every replica ("walker") of the engine at once: per step the 12 jumps of the walker's vacancy are evaluated,
 * ordered by ascending neighbour lattice id, rate_i = exp(-Ea_i/kT), dt = -ln(u1)/sum(rate)/1e13 * correction,
 * the event is the first slot whose cumulative probability is >= u2, then time/energy/occupancy are updated.
 * Walkers are independent (own occupancy, own temperature, own random stream).
 */
typedef struct lmc_kmc_params {
  double temperature;              /* `temperature` of the parameter file; used for walkers when temperatures == NULL */
  const double *temperatures;      /* optional per-walker temperatures [n_walkers] */
  int32_t n_time_temperature;      /* `time_temperature_filename` table (pred/src/TimeTemperatureInterpolator.cpp): */
  const double *tt_time;           /*   points sorted by time; 0 points = constant temperature */
  const double *tt_temperature;
  int32_t rate_corrector;          /* `rate_corrector` (pred/include/RateCorrector.hpp) */
  uint64_t seed;                   /* Philox4x32-10 key; walker w uses key (seed ^ w), counter = its step number */
} lmc_kmc_params;

typedef struct lmc_kmc_trace {     /* optional per-step records, each [n_walkers][n_steps] (host), any may be NULL */
  int64_t *from, *to;              /* vacancy lattice id before / after the step */
  int32_t *slot;                   /* selected event index in the reference's event order */
  double *dt, *Ea, *dE, *total_rate, *temperature;
} lmc_kmc_trace;

/* (re)initialise the per-walker KMC state from the current occupancy: locate the vacancy (Config::GetVacancyLatticeId),
 * vacancy / solute concentrations for the rate corrector (KineticMcAbstract.cpp:35), time = energy = steps = 0.
 * Fails with LMC_ERR_OUT_OF_RANGE unless every walker holds exactly one vacancy. */
int lmc_kmc_reset(lmc_engine *engine);
/* advance every walker by n_steps.  replay_u1 / replay_u2 (both or neither; host, [n_walkers][n_steps]) replace the
 * Philox stream by caller-supplied uniforms: u1 -> residence time, u2 -> event selection, in the reference's draw order. */
int lmc_kmc_run(lmc_engine *engine, const lmc_kmc_params *params, int64_t n_steps, const double *replay_u1,
                const double *replay_u2, const lmc_kmc_trace *trace);
/* launch shape of the most recent lmc_kmc_run: 0 = throughput kernel (half-warp per walker, kmc_run_kernel); 8 / 16 / 32 =
 * latency kernel (thread block per walker, that many lanes per candidate jump, kmc_team_run_kernel), which the library
 * picks when the engine holds only a few walkers per SM.  Both walk the same Philox trajectories. */
int lmc_kmc_last_launch_lanes(const lmc_engine *engine);
/* 1 if that launch was a latency-kernel launch that kept every walker's occupancy in shared memory (cells up to 48 KB
 * padded, e.g. the 8 x 8 x 8 cell of the batched workload: 5.8 KB; LMC_KMC_TEAM_SMEM=0 switches it off), else 0. */
int lmc_kmc_last_launch_resident_occupancy(const lmc_engine *engine);
/* 1 if that launch was a throughput-kernel launch whose tail -- the walkers still running when all but ~19 per SM (at most
 * 80 % of the launch) were through -- was handed to the latency kernel (>= 128 steps, no trace / replay;
 * LMC_KMC_HANDOFF=0 switches it off, a value in (0, 1) sets the fraction handed over).  Results are bit-identical with and
 * without the hand-off. */
int lmc_kmc_last_launch_handoff(const lmc_engine *engine);
/* Second-order ("chain") KMC: mc::KineticMcChainOmpi::Simulate (mc/src/KineticMcChainOmpi.cpp:56-152,
 * mc/include/KineticMcAbstract.h:65-143; the method script/kmc_param.txt:1 selects).  Per step and walker, for each of
 * the 12 neighbours i of the vacancy site k (the reference's 12 MPI ranks, ascending lattice id): the 12 jumps i -> l in
 * the state "vacancy at i" (144 barrier evaluations), the event k -> i as the reverse of i -> k, the eight MpiData sums in
 * rank order, the second-order residence time t_2 (an expectation: no uniform is consumed for it) and the second-order
 * event probabilities that exclude an immediate return to the site the vacancy came from (previous_j_lattice_id_, which
 * starts as first neighbour 0 of the vacancy and is kept per walker between calls; lmc_kmc_reset and lmc_kmc_run clear
 * it).  One uniform per step: replay_u (host, [n_walkers][n_steps], optional) replaces the Philox stream.  The trace has
 * the meaning of lmc_kmc_run's (Ea / dE of the chosen k -> i event, total_rate = total_rate_k_). */
int lmc_kmc_chain_run(lmc_engine *engine, const lmc_kmc_params *params, int64_t n_steps, const double *replay_u,
                      const lmc_kmc_trace *trace);
/* Restart (McAbstract constructor, mc/src/McAbstract.cpp:24-30: steps_ = restart_steps, energy_ = restart_energy, time_ =
 * restart_time): overwrite the per-walker clocks after lmc_kmc_reset (host arrays [n_walkers], any may be NULL).  The time
 * feeds T(t) and the rate corrector; the step number is the Philox counter, so a restarted run continues the random stream
 * of the run it resumes instead of repeating its first uniforms.
 * Note: lmc_engine_set_occupancy / lmc_engine_lattice_jump and any KMC error invalidate the KMC state -- the next run
 * starts from lmc_kmc_reset (clocks at zero) unless this call follows the reset. */
int lmc_kmc_set_state(lmc_engine *engine, const double *time, const double *energy, const int64_t *steps);
/* per-walker state after the last run (host arrays [n_walkers], any may be NULL): McAbstract::time_, energy_, steps_ */
int lmc_kmc_get_state(lmc_engine *engine, double *time, double *energy, int64_t *steps, int64_t *vacancy,
                      double *temperature);

/* ------------------------------------------------------------------------------------------------ CMC / SA driver
 * mc::CanonicalMcOmp / CanonicalMcSerial::Simulate (mc/src/CanonicalMcOmp.cpp:40-92, CanonicalMcSerial.cpp:40-51) and
 * mc::SimulatedAnnealing::Simulate (mc/src/SimulatedAnnealing.cpp:99-185) for every replica of the engine: random
 * unlike-species pairs (CanonicalMcAbstract.cpp:43-51), swap dE (EnergyChangePredictorPairSite), Metropolis accept
 * (CanonicalMcAbstract.cpp:86-101).  Like CanonicalMcOmp, trials are processed in batches of mutually
 * non-interfering pairs (no trial site inside the 43-site neighbourhood of another trial's sites), evaluated in
 * parallel and applied; rejected-by-interference proposals are redrawn, as in the reference.
 */
typedef struct lmc_cmc_params {
  double temperature;              /* `temperature` (CMC) */
  const double *temperatures;      /* optional per-replica temperatures [n_walkers] */
  uint64_t seed;                   /* Philox4x32-10 key */
  int32_t batch_size;              /* proposals per thread block and batch (power of two, 32..512; half as many trials
                                    * are evaluated); 0 = chosen from the lattice size */
} lmc_cmc_params;

/* reset steps / energy / counters of every replica; with sa_maximum_steps > 0 the SimulatedAnnealing schedule is armed:
 * T0 = sa_initial_temperature, T *= exp(-3/max) per trial, acceptance window 0.001 max, reheats (SimulatedAnnealing.h:42-68) */
int lmc_cmc_reset(lmc_engine *engine, double sa_initial_temperature, uint64_t sa_maximum_steps);
/* run until every replica has done at least n_trials more effective trials (device RNG).  An engine with ONE replica of
 * >= 32000 sites (or one attached to peers) is driven by the whole-GPU kernel of lmc_cmc_grid_run. */
int lmc_cmc_run(lmc_engine *engine, const lmc_cmc_params *params, int64_t n_trials);
/* ONE large lattice on the whole GPU, and on several GPUs (BASELINE configs[3]).  Same Markov-chain rules as
 * lmc_cmc_run, but the batch is spread over a persistent cooperative grid (one thread block per SM) instead of one
 * thread-block cluster; requires an engine with n_walkers == 1.  A trial is evaluated by 2 L lanes (L = 8 for small
 * batches, down to 1 at 512 proposals per block; environment variable LMC_CMC_GRID_LANES overrides it for tuning / A-B runs;
 * the trajectory does not depend on L up to the rounding of dE).
 *
 * Multi-GPU (one process per GPU, same occupancy / coefficients / seed / reset on every rank): every rank keeps the whole
 * lattice and draws the same proposals; the dE evaluation of a batch is sharded over the ranks, and kept / accept masks
 * and partial sums are written directly into the peers' exchange buffers over NVLink inside the kernel.  The trajectory
 * is bit-identical for every world size.  Set-up, collective over the ranks:
 *   1. each rank: lmc_cmc_exchange_handle(engine, h)      -- 64-byte CUDA IPC handle of its exchange buffer
 *   2. all-gather the handles (any transport; bench.py uses torch.distributed)
 *   3. each rank: lmc_cmc_attach_peers(engine, rank, world, handles[world][64], grid_ctas)
 *      grid_ctas = thread blocks per GPU, identical on all ranks (0 = this device's SM count; pass the minimum over ranks)
 *   4. barrier, then every rank calls lmc_cmc_grid_run with identical arguments.
 * A rank that waits longer than ~5 s for a peer gives up with LMC_ERR_RUNTIME instead of hanging the GPU. */
int lmc_cmc_grid_run(lmc_engine *engine, const lmc_cmc_params *params, int64_t n_trials);
int lmc_cmc_exchange_handle(lmc_engine *engine, void *handle64);
int lmc_cmc_attach_peers(lmc_engine *engine, int32_t rank, int32_t world, const void *handles, int32_t grid_ctas);
/* Domain-decomposed ("sublattice") CMC / SA driver -- BASELINE configs[1] / [3], SURVEY 8(e) "large-lattice CMC / SA".
 * SEMANTIC CHANGE against mc/src/CanonicalMcAbstract.cpp:43-51 (global random unlike pairs): the periodic lattice is cut
 * into box-shaped domains of about `domain_edge` half lattice constants per axis whose grid origin moves by a random
 * vector every sweep; within a sweep a trial swaps two sites drawn uniformly from ONE domain's active core (the domain
 * minus one plane of sites per face), redrawn while the species are equal as in the reference.  Cores of different
 * domains are outside each other's 43-site neighbourhoods, so the non-interference rule of CanonicalMcOmp.cpp:47-72 holds
 * by construction: domains run concurrently (one lane group each, the domain resident in shared memory) with no claims
 * and one grid barrier per sweep; inside a domain trials are sequential Metropolis steps (CanonicalMcAbstract.cpp:86-101) on
 * the exact dE of EnergyChangePredictorPairSite::GetDeFromLatticeIdPair.  Proposals are symmetric and every sweep leaves
 * the canonical distribution invariant; the moving origin makes the chain ergodic (ensemble parity with
 * mc::CanonicalMcSerial: tests/test_gpu_cmc_stat.py).  Energy / steps / SA schedule are updated once per sweep.
 * Works for any number of replicas (all replicas sweep in lock step until every one has done n_trials more).
 *
 * Multi-GPU (ONE lattice, n_walkers == 1; one process per GPU, same occupancy / coefficients / seed / reset everywhere):
 * rank r owns a slab of the domain grid along x; at write-back every row goes to the local buffer and, through NVLink
 * peer mappings, to every rank whose slab or halo holds that plane in the next sweep (halo + migration exchange inside
 * the persistent kernel); sweep totals travel as flag-carrying lines that double as the inter-GPU barrier.  The
 * trajectory is identical for every world size.  Set-up, collective over the ranks:
 *   1. each rank: lmc_cmc_domain_handles(engine, h)       -- three 64-byte CUDA IPC handles (192 bytes)
 *   2. all-gather the handles
 *   3. each rank: lmc_cmc_domain_attach_peers(engine, rank, world, handles[world][192])
 *   4. barrier, then every rank calls lmc_cmc_domain_run with identical arguments. */
typedef struct lmc_cmc_domain_params {
  int32_t domain_edge;             /* target domain edge in half lattice constants, 4..48 (0 = 6: cores of 32 sites) */
  int32_t rounds_per_sweep;        /* Metropolis rounds per domain and sweep (0 = 216) */
  int32_t speculate;               /* rounds of one domain evaluated at once on the current state and committed up to the first
                                    * accepted one: 1, 2 (with 8 or 16 lanes) or 4 (with 8 lanes); 0 = 32 / lanes when the lattice
                                    * has fewer domains than the GPU has warp slots, else 1.  Exactly the sequential chain */
  int32_t lanes;                   /* lanes per trial: 8, 16 or 32 (0 = from the number of domains).  Launch shape only:
                                    * the random stream is keyed by (seed, sweep, domain, round), the trajectory does not
                                    * depend on it (up to the rounding of dE in the last bit) */
  double passes;                   /* domains per lane group and sweep, handed out dynamically (0 = 1; > 1 trades resident
                                    * groups for load balance when the domains differ in cost).  Launch shape only */
} lmc_cmc_domain_params;
int lmc_cmc_domain_run(lmc_engine *engine, const lmc_cmc_params *params, const lmc_cmc_domain_params *domain, int64_t n_trials);
int lmc_cmc_domain_handles(lmc_engine *engine, void *handles192);
int lmc_cmc_domain_attach_peers(lmc_engine *engine, int32_t rank, int32_t world, const void *handles);
/* launch shape of the last lmc_cmc_domain_run: {domain_edge, domains, lanes, threads per block, blocks, rounds per sweep,
 * speculate} */
int lmc_cmc_domain_last_shape(const lmc_engine *engine, int32_t *shape7);
/* replay mode on replica `walker`: the n trials (site_a, site_b, u) are applied in the given order with the reference's
 * serial semantics (u is consumed only when dE >= 0).  Outputs (host, [n], optional): dE, energy and temperature before
 * each trial, accept flags. */
int lmc_cmc_replay(lmc_engine *engine, int32_t walker, const lmc_cmc_params *params, int64_t n, const int64_t *site_a,
                   const int64_t *site_b, const double *u, double *dE, double *energy_before, double *temperature_before,
                   uint8_t *accepted);
/* per-replica state (host arrays [n_walkers], any may be NULL): energy_ (relative to the reset), steps_, accepted trials,
 * current temperature (SA) */
int lmc_cmc_get_state(lmc_engine *engine, double *energy, int64_t *steps, int64_t *accepted, double *temperature);

/* ------------------------------------------------------------------------------------------------ debug taps
 * Integer artefacts of the reference's algorithm, recomputed on the device, for bit-exact parity checks.
 * lists: the symmetry-ordered lattice-id lists of pair (site_i, site_j)
 *   state[60]  GetSortedLatticeVectorStateOfPair       (pred/src/EnergyUtility.cpp:261-287)
 *   mmm[58]    GetSymmetricallySortedLatticeVectorMMM  (:45-74)
 *   mm2[58]    GetSymmetricallySortedLatticeVectorMM2  (:75-104)
 *   mm2_backward[58]  the same for the reversed pair (site_j, site_i), as used by GetKs (:187-188)
 * counts: start/end cluster-type histograms of GetDe (VacancyMigrationPredictorQuartic.cpp:124-154), n_types each
 * encodes: integer numerators of GetOneHotParametersFromMap (EnergyUtility.cpp:743-796): per slot the number of
 *   clusters of that type (the reference's encode is this divided by the group size, see lmc_tables_group_sizes). */
int lmc_debug_pair(lmc_engine *engine, int32_t walker, int64_t site_i, int64_t site_j, int64_t *state60, int64_t *mmm58,
                   int64_t *mm2_58, int64_t *mm2_backward58, int32_t *start_counts, int32_t *end_counts,
                   int32_t *enc_mmm, int32_t *enc_mm2_forward, int32_t *enc_mm2_backward);
/* GetSortedLatticeVectorStateOfSite (:288-313) + the counts of GetDeFromLatticeIdSite */
int lmc_debug_site(lmc_engine *engine, int32_t walker, int64_t site, int32_t new_element, int64_t *state43,
                   int32_t *start_counts, int32_t *end_counts);

/* ------------------------------------------------------------------------------------------------ host-side tables
 * (no device needed).  Neighbour lists in ascending lattice id like Config::Get{First,Second,Third}NeighborsAdjacencyList. */
int lmc_engine_neighbors(const lmc_engine *engine, int32_t shell, int64_t site, int64_t *out);
/* the 12 first neighbours of `site` in the order the first-order KMC kernels file their events (a 64 x 12 table by the
 * site's boundary / parity class, ranked on the host): must equal lmc_engine_neighbors(1, site), i.e. the event order of
 * KineticMcFirstOmp::BuildEventList (mc/src/KineticMcFirstOmp.cpp:52-68), for every site.  Host side; used by the tests. */
int lmc_engine_kmc_event_order(const lmc_engine *engine, int64_t site, int64_t *out12);
int lmc_engine_site_coords(const lmc_engine *engine, int64_t site, int32_t xyz_half_units[3]);
/* host evaluation of the ordered id lists (same definition as lmc_debug_pair / lmc_debug_site) */
int lmc_engine_pair_lists(const lmc_engine *engine, int64_t site_i, int64_t site_j, int64_t *state60, int64_t *mmm58,
                          int64_t *mm2_58, int64_t *mm2_backward58);
int lmc_engine_site_list(const lmc_engine *engine, int64_t site, int64_t *state43);
/* cluster mappings in list positions, flattened as: G, then per group: C, L, C*L entries (-1 = SIZE_MAX marker).
 * which: 0 GetClusterParametersMappingStatePair, 1 GetAverageClusterParametersMappingMMM, 2 ...MM2,
 * 3 GetClusterParametersMappingStateSite (pred/src/EnergyUtility.cpp:169-259,393-581). Returns the length needed. */
int64_t lmc_tables_mapping(int32_t which, int64_t *out, int64_t capacity);
/* cluster types in pred::ClusterIndexer order: rows [label, size, e1, e2, e3] (ElementName codes, -1 padded).
 * Returns the number of types (= required length of Base.theta). */
int32_t lmc_tables_cluster_types(const int32_t *element_set, int32_t n_elements, int32_t *rows5, int32_t capacity_rows);
/* encode vector lengths and per-slot group sizes of the mmm (which=1) / mm2 (which=2) mappings */
int32_t lmc_tables_group_sizes(int32_t which, int32_t n_elements, int32_t *sizes, int32_t capacity);

/* Introspection of the contracted coefficient tables (host copies; engine must have coefficients loaded).
 * which: 0 pair_C [m][3], 1 pair_A [m][58][n][3], 2 pair_B [m][556][n][n][3]   (jump tables: dE, logD, logKs)
 *        3 site_C [x],    4 site_A [x][42][n+1],  5 site_B [x][204][n+1][n+1]  (single-site tables, codes incl. vacancy)
 *        6 "Base".theta as read (one entry per cluster type in ClusterIndexer order)
 *        7 / 8 / 9 the folded tables the KMC kernels walk: [..][2] = (dE, log E0 = logKs + 2 logD) of pair_C / pair_A / pair_B,
 *        every entry rounded to a multiple of 2^-bits (lmc_engine_kmc_table_grid_bits) so that sums are exact in any order
 * Species codes are positions in the element set sorted by name; the vacancy is code n.  Returns the length. */
int64_t lmc_engine_get_tables(const lmc_engine *engine, int32_t which, double *out, int64_t capacity);
/* binary grid of the folded KMC tables: bits[c] for component c (0: dE, 1: log E0); chosen so that the largest sum any
 * environment can produce, |C| + sum_t max|A_t| + sum_pairs max|B|, stays below 2^(51 - bits[c]) */
int lmc_engine_kmc_table_grid_bits(const lmc_engine *engine, int32_t bits[2]);
/* environment pairs (t,u), t<u, as indices into the environment (ordered state list without the centre site(s)):
 * which 0: the 556 pairs of the jump environment, 1: the 204 pairs of the site environment. Returns the pair count. */
int32_t lmc_tables_env_pairs(int32_t which, int16_t *pairs, int32_t capacity_pairs);

#ifdef __cplusplus
}
#endif
#endif /* LMC_B200_H_ */

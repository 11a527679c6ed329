// engine.h -- the engine object behind the C ABI (include/lmc_b200.h): one per GPU, owns the device-resident
// occupancy (packed uint8, padded layout) and the constant tables.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <map>
#include <vector>

#include <cuda_runtime.h>

#include "device_tables.h"
#include "lattice.h"
#include "tables.h"

struct lmc_kmc_params;
struct lmc_kmc_trace;
struct lmc_cmc_params;
struct lmc_cmc_domain_params;

namespace lmc {

extern thread_local std::string g_last_error;
int guard(const std::function<void()> &fn);

class Engine {
 public:
  Engine(const int32_t factors[3], int32_t id_order, const int32_t *element_set, int32_t n_elements, int32_t solvent,
         int32_t n_walkers, int32_t device);
  ~Engine();
  Engine(const Engine &) = delete;
  Engine &operator=(const Engine &) = delete;

  void load_coefficients(const std::string &json_path, int model = kModelQuartic);
  void load_coefficients(const Coefficients &co, int model);
  void set_occupancy(int32_t walker, const uint8_t *occ, int64_t n, int32_t count);
  void get_occupancy(int32_t walker, uint8_t *occ, int64_t n, int32_t count);
  void lattice_jump(int32_t walker, int64_t a, int64_t b);
  void get_elements(int32_t walker, int64_t n, const int64_t *sites, uint8_t *out);
  int64_t find_element(int32_t walker, int32_t element, int64_t *count);

  void eval_barriers(int64_t n, const int32_t *walker, const int64_t *site_i, const int64_t *site_j, double *Ea, double *dE,
                     double *D, double *Ks);
  void eval_barriers_dev(int64_t n, const int32_t *walker, const int64_t *site_i, const int64_t *site_j, double *Ea,
                         double *dE, double *D, double *Ks);
  void eval_swap_de(int64_t n, const int32_t *walker, const int64_t *a, const int64_t *b, double *dE, bool first_neighbours_only = false);
  void eval_swap_de_dev(int64_t n, const int32_t *walker, const int64_t *a, const int64_t *b, double *dE, bool first_neighbours_only = false);
  void eval_site_de(int64_t n, const int32_t *walker, const int64_t *site, const uint8_t *new_element, double *dE);
  double total_energy(int32_t walker, int64_t *counts, int32_t n_types);
  double cluster_energy(int32_t walker, const int64_t *sites, int64_t n, int64_t *counts, int32_t n_types);
  void debug_pair(int32_t walker, int64_t i, int64_t j, int64_t *state, int64_t *mmm, int64_t *mm2, int64_t *mm2b, int32_t *sc,
                  int32_t *ec, int32_t *enc_mmm, int32_t *enc_f, int32_t *enc_b);
  void debug_site(int32_t walker, int64_t site, int32_t new_element, int64_t *state43, int32_t *sc, int32_t *ec);

  // drivers
  void eval_vacancy_events(int64_t n, const int32_t *walker, const int64_t *vacancy, int64_t *neighbour, double *Ea, double *dE);
  void eval_vacancy_events_dev(int64_t n, const int32_t *walker, const int64_t *vacancy, int64_t *neighbour, double *Ea, double *dE);
  void kmc_reset();
  void kmc_run(const lmc_kmc_params &params, int64_t n_steps, const double *u1, const double *u2, const lmc_kmc_trace *trace,
               bool second_order);
  void kmc_set_state(const double *time, const double *energy, const int64_t *steps);
  void kmc_get_state(double *time, double *energy, int64_t *steps, int64_t *vacancy, double *temperature);

  void cmc_reset(double sa_initial_temperature, uint64_t sa_maximum_steps);
  void cmc_run(const lmc_cmc_params &params, int64_t n_trials, int32_t replay_walker, int64_t n_replay, const int64_t *a,
               const int64_t *b, const double *u, double *dE, double *energy_before, double *temperature_before, uint8_t *accepted);
  void cmc_get_state(double *energy, int64_t *steps, int64_t *accepted, double *temperature);
  // whole-GPU / multi-GPU single-lattice driver (cmc_grid_kernels.cuh)
  void cmc_grid_prepare();
  void cmc_exchange_handle(void *handle64);
  void cmc_attach_peers(int32_t rank, int32_t world, const void *handles, int32_t grid_ctas);
  void cmc_grid_run(const lmc_cmc_params &params, int64_t n_trials);
  // domain-decomposed ("sublattice") driver (cmc_domain_kernels.cuh)
  void cmc_domain_prepare();
  void cmc_domain_handles(void *handles192);
  void cmc_domain_attach_peers(int32_t rank, int32_t world, const void *handles);
  void cmc_domain_run(const lmc_cmc_params &params, const lmc_cmc_domain_params *dom, int64_t n_trials);
  void cmc_sync_cells();                           // rebuild the batched drivers' mirror / cell arrays after the occupancy changed under them

  // host-side geometry (no device needed)
  void neighbors(int32_t shell, int64_t site, int64_t *out) const;
  void pair_lists(int64_t i, int64_t j, int64_t *state, int64_t *mmm, int64_t *mm2, int64_t *mm2b) const;
  void site_list(int64_t site, int64_t *state43) const;
  int64_t wrapped_id(int x, int y, int z) const;
  int host_direction(int64_t i, int64_t j, int *xyz_i) const;
  int host_frame_flag(const int *xyz, int k) const;

  void time_begin();
  void time_end();
  double last_kernel_ms();
  void require_device() const;
  void require_coefficients() const;
  void check_event_errors(const char *what);
  void *scratch(size_t bytes);
  template <class T> const T *to_device(const std::vector<T> &v);

  LatticeDesc lat{};
  Species species;
  int32_t n_walkers{1};
  int32_t device{-1};
  bool has_coefficients{false};
  Coefficients coefficients;
  std::string coefficients_path;
  PairTables pair_tables;
  SiteTables site_tables;
  EnergyTables energy_tables;

  cudaStream_t stream{nullptr};
  uint8_t *d_occ{nullptr};
  int *d_error{nullptr};
  DevTables tab{};
  const int8_t *d_code_of_enum{nullptr};
  const uint8_t *d_enum_of_code{nullptr};
  std::vector<void *> device_allocs;
  void *d_scratch{nullptr};
  void *h_pinned{nullptr};
  size_t scratch_bytes{0};
  // request grouping (grouping.h): own workspace (the staging scratch above holds the caller's arrays meanwhile)
  void *d_group{nullptr};
  size_t group_bytes{0};
  unsigned int *h_group_local{nullptr};
  bool last_grouped{false};
  const uint32_t *group_if_scattered(int64_t n, const int32_t *walker, const int64_t *site);
  // KMC per-walker state (device)
  int64_t *d_kmc_vacancy{nullptr}, *d_kmc_steps{nullptr};
  double *d_kmc_time{nullptr}, *d_kmc_energy{nullptr}, *d_kmc_temperature{nullptr}, *d_kmc_cvac{nullptr}, *d_kmc_csol{nullptr};
  int32_t *d_kmc_error{nullptr};
  int64_t *d_kmc_previous{nullptr};
  bool kmc_ready{false};
  int kmc_team_lanes{0};       // lanes per candidate jump of the last first-order KMC launch (0: half-warp kernel)
  std::vector<double> folded_C, folded_A, folded_B;   // the KMC kernels' tables: (dE, log E0) per entry, on the binary grid (host copies)
  int pair_grid_bits[2]{0, 0};         // the folded KMC tables are multiples of 2^-bits (dE, log E0): exact sums in any order
  bool kmc_team_resident_occ{false};   // ... and whether that launch kept the walkers' occupancy in shared memory
  bool kmc_handoff{false};             // the last half-warp launch handed its tail to the latency kernel
  int64_t *d_kmc_target{nullptr};      // hybrid launch: the step number every walker has to reach
  int *d_kmc_done{nullptr};            // hybrid launch: [0] walkers that are through, [1] walkers in the tail
  int32_t *d_kmc_tail_order{nullptr};  // hybrid launch: the tail's walkers, most steps left first
  std::vector<uint8_t> kmc_event_order_table() const;                        // [64][12], see engine.cu
  void kmc_event_order(int64_t site, int64_t *neighbours_in_slot_order) const;
  const void *kmc_team_kernel_choice(bool instrumented, size_t table_smem, int *lanes_out, size_t *smem_out, bool for_tail = false);
  // CMC / SA per-replica state (device)
  double *d_cmc_energy{nullptr};
  unsigned long long *d_cmc_steps{nullptr}, *d_cmc_accepted{nullptr}, *d_cmc_proposals{nullptr}, *d_cmc_epoch{nullptr};
  unsigned int *d_cmc_marks{nullptr};
  uint8_t *d_cmc_mirror{nullptr};
  void *d_cmc_sa{nullptr};
  int32_t *d_cmc_error{nullptr};
  double *d_cmc_temperature{nullptr};
  bool cmc_ready{false};
  void *d_cmc_xchg{nullptr};                       // CmcExchange (IPC-shareable)
  void *cmc_peer_xchg[8]{};                        // peer mappings (cudaIpcOpenMemHandle), [rank] = own buffer
  unsigned long long *d_cmc_grid_counter{nullptr}, *d_cmc_sequence{nullptr}, *d_cmc_accum{nullptr};
  int *d_cmc_abort{nullptr};
  int cmc_world{1}, cmc_rank{0}, cmc_grid_ctas{0};
  int cmc_grid_checked_threads{0};
  size_t cmc_grid_checked_smem{0};
  const void *cmc_grid_checked_kernel{nullptr};
  // domain-decomposed driver: second occupancy buffer, sweep-parity state, totals, inter-GPU lines
  uint8_t *d_occ_buf[2]{nullptr, nullptr};         // d_occ is always d_occ_buf[occ_cur] once the driver has been prepared
  int occ_cur{0};
  void *d_dom_state{nullptr};                      // DomState[2][n_walkers]
  unsigned long long *d_dom_accum{nullptr};        // [3][n_walkers][4]
  void *d_dom_lines{nullptr};                      // DomLine[2][8] (IPC-shareable)
  unsigned long long *d_dom_counters{nullptr};     // [0] grid barrier counter, [1] inter-GPU line sequence
  int *d_dom_abort{nullptr};
  uint8_t *dom_peer_occ[2][8]{};                   // peer mappings of both occupancy buffers ([.][rank] = own)
  void *dom_peer_lines[8]{};
  int dom_world{1}, dom_rank{0};
  bool cmc_cells_stale{false};
  int dom_last_spec{0}, dom_last_lanes{0}, dom_last_threads{0}, dom_last_ctas{0}, dom_last_domains{0}, dom_last_rounds{0}, dom_last_edge{0};
  std::map<int, int> attr_cache;
  int device_attr(int attr);
  // measurement: CUDA events around the last hot kernel on the engine stream, and a count of our kernel launches
  cudaEvent_t ev_begin{nullptr}, ev_end{nullptr};
  bool timing_pending{false};
  double last_ms{0.0};
  int64_t launch_count{0};

 private:
  void upload_geometry_tables();
};

}  // namespace lmc

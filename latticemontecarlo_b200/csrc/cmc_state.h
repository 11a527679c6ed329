// cmc_state.h -- per-replica CMC / SA chain state shared by the batched, grid and domain drivers.
#pragma once
#include <cstdint>

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

}  // namespace lmc

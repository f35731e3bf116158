// gfr_emu.cpp - TEST INFRASTRUCTURE ONLY.  Compiles the LANES = 1 instantiation of the device
// functions (grid-fed-rl-gym_b200/csrc/gfr_device.cuh) for the host so that control flow and
// arithmetic can be checked against the oracle in the build container, which has no GPU.
// Nothing in the package loads this; the product path is the CUDA library and nothing else.
// Built by tests/host_emu/build.py into tests/host_emu/_build/ (git-ignored).
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../grid-fed-rl-gym_b200/csrc/gfr_image.hpp"

using namespace gfr;

struct emu_env {
  FeederImage fi;
  EnvCfg cfg{};
  int solver = SOLVER_NEWTON;
  long long B = 0;
  std::vector<double> state, obs, bat_soc0;
};

struct EmuSlot {
  std::vector<double> work;    // shared-memory slot
  std::vector<D2> mg;          // the global scratch of the slot
  explicit EmuSlot(const Layout& lay) : work(8 + (newton_slot_bytes(lay.n, lay.n_pool, 0) + sweep_slot_bytes(lay.n, 0)) / 8, 0.0),
                                        mg(newton_scratch_doubles(lay.n) / 2 + 1) {}
  template <class G> G group(const Layout& lay) {
    G g;
    g.lane = 0; g.mask = 1u;
    bind_slot(g, reinterpret_cast<unsigned char*>(work.data()), lay.n, lay.n_pool, mg.data());
    return g;
  }
};

extern "C" {

const char* emu_error(void) { static std::string s; return s.c_str(); }

emu_env* emu_create(const gfr_feeder_desc* d, long long B, const gfr_env_cfg* c) {
  auto* e = new emu_env();
  std::string err = build_feeder_image(d, &e->fi);
  if (!err.empty()) { delete e; return nullptr; }
  const Layout& lay = e->fi.lay;
  e->B = B;
  e->solver = c->solver.solver == GFR_SOLVER_NEWTON ? SOLVER_NEWTON : SOLVER_SWEEP;
  EnvCfg& k = e->cfg;
  k.dt = c->timestep; k.v_min = c->v_min; k.v_max = c->v_max; k.f_min = c->f_min; k.f_max = c->f_max;
  k.penalty = c->safety_penalty; k.load_noise = c->load_noise; k.tol = c->solver.tolerance;
  k.accel = c->solver.acceleration != 0.0 ? c->solver.acceleration : 1.0;
  k.episode_length = c->episode_length; k.stochastic_loads = c->stochastic_loads != 0;
  k.weather_variation = c->weather_variation != 0; k.max_it = c->solver.max_iterations;
  e->state.assign((size_t)B * lay.R, 0.0);
  e->obs.assign((size_t)B * lay.D, 0.0);
  e->bat_soc0.assign(d->bat_soc0, d->bat_soc0 + lay.Bt);
  Lanes<1> g; g.lane = 0; g.mask = 1u;
  for (long long i = 0; i < B; ++i)
    reset_instance<1>(g, lay, (const int*)e->fi.img.data(), (const double*)e->fi.img.data(), k, i,
                      e->state.data(), e->obs.data(), e->fi.load_pq.data(), e->bat_soc0.data(),
                      nullptr, nullptr, 0.0, true, (long long)c->env_id_offset);
  return e;
}

void emu_destroy(emu_env* e) { delete e; }
double* emu_obs(emu_env* e) { return e->obs.data(); }
int emu_obs_dim(emu_env* e) { return e->fi.lay.D; }

void emu_reset(emu_env* e, const uint64_t* seeds, const uint8_t* mask, const double* noise,
               double start_time) {
  const Layout& lay = e->fi.lay;
  Lanes<1> g; g.lane = 0; g.mask = 1u;
  for (long long i = 0; i < e->B; ++i) {
    if (mask && !mask[i]) continue;
    reset_instance<1>(g, lay, (const int*)e->fi.img.data(), (const double*)e->fi.img.data(), e->cfg, i,
                      e->state.data(), e->obs.data(), e->fi.load_pq.data(), e->bat_soc0.data(), seeds,
                      noise, start_time, false, 0);
  }
}

void emu_step(emu_env* e, const double* actions, const double* noise, const gfr_step_out* out) {
  const Layout& lay = e->fi.lay;
  StepOut o{};
  o.reward = out->reward; o.terminated = out->terminated; o.truncated = out->truncated;
  o.error = out->error; o.converged = out->converged; o.iterations = out->iterations;
  o.max_voltage = out->max_voltage; o.min_voltage = out->min_voltage; o.losses = out->losses;
  o.max_mismatch = out->max_mismatch; o.violations = out->violations;
  o.violation_count = out->violation_count; o.current_step = out->current_step;
  o.episode_reward = out->episode_reward; o.noise_used = out->noise_used;
  const int* simg = (const int*)e->fi.img.data();
  const double* dimg = (const double*)e->fi.img.data();
  EmuSlot slot(lay);
  for (long long i = 0; i < e->B; ++i) {
    if (e->solver == SOLVER_NEWTON)
      step_instance<1, SOLVER_NEWTON>(slot.group<NGrp<1>>(lay), lay, simg, dimg, e->cfg, i,
                                      e->state.data(), e->obs.data(), actions, noise, o);
    else
      step_instance<1, SOLVER_SWEEP>(slot.group<SGrp<1>>(lay), lay, simg, dimg, e->cfg, i,
                                     e->state.data(), e->obs.data(), actions, noise, o);
  }
}

int emu_solve(const gfr_feeder_desc* d, long long B, const double* p_inj, const gfr_solver_cfg* c,
              const gfr_sol_out* out) {
  FeederImage fi;
  std::string err = build_feeder_image(d, &fi);
  if (!err.empty()) return -1;
  const Layout& lay = fi.lay;
  const int solver = c->solver == GFR_SOLVER_NEWTON ? SOLVER_NEWTON : SOLVER_SWEEP;
  EmuSlot slot(lay);
  EnvCfg k{};
  k.tol = c->tolerance; k.max_it = c->max_iterations; k.accel = c->acceleration != 0.0 ? c->acceleration : 1.0;
  SolOut o{};
  o.converged = out->converged; o.iterations = out->iterations; o.bus_voltages = out->bus_voltages;
  o.bus_angles = out->bus_angles; o.line_flows = out->line_flows; o.line_loadings = out->line_loadings;
  o.losses = out->losses; o.max_mismatch = out->max_mismatch;
  const int* simg = (const int*)fi.img.data();
  const double* dimg = (const double*)fi.img.data();
  for (long long i = 0; i < B; ++i) {
    if (solver == SOLVER_NEWTON)
      solve_instance<1, SOLVER_NEWTON>(slot.group<NGrp<1>>(lay), lay, simg, dimg, k, i, p_inj, o);
    else
      solve_instance<1, SOLVER_SWEEP>(slot.group<SGrp<1>>(lay), lay, simg, dimg, k, i, p_inj, o);
  }
  return 0;
}

void emu_noise_fill(long long B, int n_slots, const uint64_t* seeds, const uint64_t* draws, double* out) {
  for (long long i = 0; i < B; ++i)
    for (int s = 0; s < n_slots; ++s) out[i * n_slots + s] = noise_slot(seeds[i], draws[i], s);
}

}  // extern "C"

#ifdef GFR_EMU_STATS
extern "C" long long* emu_stats(void) { return gfr::gfr_emu_stats; }
#endif

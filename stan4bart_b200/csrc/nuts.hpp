// stan4bart_b200/csrc/nuts.hpp -- host-side NUTS control.
// Tree building, U-turn checks, dual averaging and windowed variance adaptation stay on the
// host exactly as north_star prescribes; every potential / gradient evaluation goes through
// GlmmModel::log_prob_grad, i.e. one device pass.
//   /root/reference/src/interruptable_sampler.hpp:118-210 (driver),
//   src/include/stan/mcmc/hmc/nuts/base_nuts.hpp:78-352, adapt_diag_e_nuts.hpp:25-49,
//   hmc/base_hmc.hpp:81-182, stepsize_adaptation.hpp:49-73, var_adaptation.hpp:17-46,
//   windowed_adaptation.hpp:23-110
#pragma once

#include "glmm.hpp"

#include <vector>

namespace s4b {

class NutsSampler {
 public:
  NutsSampler(GlmmModel& model, const s4b_stan_control& ctl, int chain_id, int num_warmup);
  int num_pars() const { return 7 + model_.num_constrained(); }
  // `skip` transitions; the last one is written to out[num_pars] (may be NULL)
  void run(bool warmup, double* out);
  void disengage_adaptation();
  double stepsize() const { return nom_epsilon_; }
  const std::vector<double>& inv_metric() const { return inv_metric_; }
  const std::vector<double>& q() const { return cont_params_; }
  int last_n_leapfrog() const { return n_leapfrog_; }

 private:
  struct Point { std::vector<double> q, p, g; double V = 0.0; };
  using Vec = std::vector<double>;

  void update_potential_gradient(Point& z);
  double kinetic(const Point& z) const;
  double hamiltonian(const Point& z) const { return kinetic(z) + z.V; }
  void sample_p(Point& z);
  void sharp(const Point& z, Vec& out) const;
  void evolve(Point& z, double eps);
  void init_stepsize();
  void transition();
  bool build_tree(int depth, Point& z_propose, Vec& p_sharp_beg, Vec& p_sharp_end, Vec& rho, Vec& p_beg, Vec& p_end, double H0,
                  double sign, int& n_leapfrog, double& log_sum_weight, double& sum_metro_prob);
  void learn_stepsize(double adapt_stat);
  bool learn_variance();
  void window_restart();
  void set_window_params(uint32_t num_warmup, uint32_t init_buffer, uint32_t term_buffer, uint32_t base_window);
  void compute_next_window();

  // scratch of one recursion level of build_tree (depth is unique along the call stack): allocated once, so that the
  // ~1000 leapfrogs of a deep transition do not touch the heap
  struct Level { Vec p_init_end, p_sharp_init_end, rho_init, p_final_beg, p_sharp_final_beg, rho_final, rho_subtree; Point z_propose_final; };
  std::vector<Level> levels_;

  GlmmModel& model_;
  s4b_stan_control ctl_;
  int d_;
  RngState rng_;
  Point z_;
  Vec inv_metric_, cont_params_, grad_tmp_;
  double nom_epsilon_ = 0.1, epsilon_ = 0.1, epsilon_jitter_ = 0.0;
  int max_depth_ = 5; double max_deltaH_ = 1000.0;
  int depth_ = 0, n_leapfrog_ = 0; bool divergent_ = false; double energy_ = 0.0;
  double lp_ = 0.0, accept_stat_ = 0.0;
  // dual averaging
  double sa_counter_ = 0, sa_s_bar_ = 0, sa_x_bar_ = 0, sa_mu_ = 0.5, sa_delta_ = 0.5, sa_gamma_ = 0.05, sa_kappa_ = 0.75, sa_t0_ = 10;
  // windowed variance adaptation; unsigned arithmetic on purpose (matches the reference's types)
  uint32_t num_warmup_ = 0, init_buffer_ = 0, term_buffer_ = 0, base_window_ = 0, window_counter_ = 0, next_window_ = 0, window_size_ = 0;
  double wf_n_ = 0; Vec wf_m_, wf_m2_;
  bool adapt_flag_ = true;
  double prof_lp_ns_ = 0, prof_tr_ns_ = 0; long long prof_lp_calls_ = 0, prof_tr_calls_ = 0;   // S4B_HOST_PROF=1 diagnostics
};

}  // namespace s4b

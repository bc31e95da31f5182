// stan4bart_b200/csrc/nuts.cu -- see nuts.hpp.  Host code only (compiled by nvcc for the shared headers).
#include "nuts.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace s4b {

static const double kInf = std::numeric_limits<double>::infinity();

static inline double log_sum_exp(double a, double b)
{
  if (a == -kInf) return b;
  if (a == kInf && b == kInf) return kInf;
  return a > b ? a + std::log1p(std::exp(b - a)) : b + std::log1p(std::exp(a - b));
}
static inline bool no_u_turn(const std::vector<double>& sharp_minus, const std::vector<double>& sharp_plus, const std::vector<double>& rho)
{
  double a = 0.0, b = 0.0;
  for (size_t i = 0; i < rho.size(); ++i) { a += sharp_plus[i] * rho[i]; b += sharp_minus[i] * rho[i]; }
  return a > 0 && b > 0;
}

NutsSampler::NutsSampler(GlmmModel& model, const s4b_stan_control& ctl, int chain_id, int num_warmup)
    : model_(model), ctl_(ctl), d_(model.num_params())
{
  std::memset(&rng_, 0, sizeof rng_);
  rng_.key0 = ctl.seed; rng_.key1 = (uint32_t) chain_id; rng_.stream = S4B_STREAM_STAN;
  const size_t d = (size_t) d_;
  z_.q.assign(d, 0.0); z_.p.assign(d, 0.0); z_.g.assign(d, 0.0);
  inv_metric_.assign(d, 1.0); cont_params_.assign(d, 0.0); grad_tmp_.assign(d + 1, 0.0);
  wf_m_.assign(d, 0.0); wf_m2_.assign(d, 0.0);
  // services/util/initialize.hpp:86-216: uniform(-R, R) inits, at most 100 attempts
  bool initialised = false;
  for (int attempt = 0; attempt < 100 && !initialised; ++attempt) {
    for (size_t i = 0; i < d; ++i) cont_params_[i] = ctl.init_radius == 0.0 ? 0.0 : -ctl.init_radius + 2.0 * ctl.init_radius * rng_uniform(rng_);
    double lp;
    initialised = model_.log_prob_grad(cont_params_.data(), &lp, grad_tmp_.data()) == 0;
  }
  // initialize.hpp:186-195: after 100 rejected points Stan gives up with "Initialization failed."
  if (!initialised) throw std::domain_error("Initialization failed. (100 attempts: no point with a finite log density and gradient)");
  if (ctl.stepsize > 0) nom_epsilon_ = ctl.stepsize;
  if (ctl.stepsize_jitter > 0 && ctl.stepsize_jitter < 1) epsilon_jitter_ = ctl.stepsize_jitter;
  if (ctl.max_treedepth > 0) max_depth_ = ctl.max_treedepth;
  levels_.resize((size_t) max_depth_ + 2);       // never resized while build_tree holds references into it
  sa_mu_ = std::log(10 * ctl.stepsize);
  if (ctl.adapt_delta > 0 && ctl.adapt_delta < 1) sa_delta_ = ctl.adapt_delta;
  if (ctl.adapt_gamma > 0) sa_gamma_ = ctl.adapt_gamma;
  if (ctl.adapt_kappa > 0) sa_kappa_ = ctl.adapt_kappa;
  if (ctl.adapt_t0 > 0) sa_t0_ = ctl.adapt_t0;
  window_restart();
  set_window_params((uint32_t) (num_warmup * ctl.skip), ctl.adapt_init_buffer, ctl.adapt_term_buffer, ctl.adapt_window);
  z_.q = cont_params_;
  init_stepsize();
}

void NutsSampler::update_potential_gradient(Point& z)
{
  double lp = 0.0;
  static const bool host_prof = getenv("S4B_HOST_PROF") != nullptr;
  std::chrono::steady_clock::time_point t0;
  if (host_prof) t0 = std::chrono::steady_clock::now();
  int status = model_.log_prob_grad(z.q.data(), &lp, grad_tmp_.data());
  if (host_prof) { prof_lp_ns_ += std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count(); ++prof_lp_calls_; }
  if (status == 0) { z.V = -lp; for (int i = 0; i < d_; ++i) z.g[(size_t) i] = -grad_tmp_[(size_t) i]; }
  else { z.V = kInf; for (int i = 0; i < d_; ++i) z.g[(size_t) i] = -z.g[(size_t) i]; }   // base_hamiltonian.hpp:61-70
}
double NutsSampler::kinetic(const Point& z) const { double t = 0.0; for (int i = 0; i < d_; ++i) t += z.p[(size_t) i] * inv_metric_[(size_t) i] * z.p[(size_t) i]; return 0.5 * t; }
void NutsSampler::sample_p(Point& z) { for (int i = 0; i < d_; ++i) z.p[(size_t) i] = rng_normal(rng_) / std::sqrt(inv_metric_[(size_t) i]); }
void NutsSampler::sharp(const Point& z, Vec& out) const { out.resize((size_t) d_); for (int i = 0; i < d_; ++i) out[(size_t) i] = inv_metric_[(size_t) i] * z.p[(size_t) i]; }

void NutsSampler::evolve(Point& z, double eps)
{
  for (int i = 0; i < d_; ++i) z.p[(size_t) i] -= 0.5 * eps * z.g[(size_t) i];
  for (int i = 0; i < d_; ++i) z.q[(size_t) i] += eps * inv_metric_[(size_t) i] * z.p[(size_t) i];
  update_potential_gradient(z);
  for (int i = 0; i < d_; ++i) z.p[(size_t) i] -= 0.5 * eps * z.g[(size_t) i];
}

void NutsSampler::init_stepsize()
{
  Point z_init = z_;
  if (nom_epsilon_ == 0 || nom_epsilon_ > 1e7 || std::isnan(nom_epsilon_)) return;
  auto trial = [&]() {
    sample_p(z_); update_potential_gradient(z_);
    double H0 = hamiltonian(z_);
    evolve(z_, nom_epsilon_);
    double h = hamiltonian(z_); if (std::isnan(h)) h = kInf;
    return H0 - h;
  };
  const double log08 = std::log(0.8);
  int direction = trial() > log08 ? 1 : -1;
  for (;;) {
    z_ = z_init;
    double delta_H = trial();
    if (direction == 1 && !(delta_H > log08)) break;
    else if (direction == -1 && !(delta_H < log08)) break;
    else nom_epsilon_ = direction == 1 ? 2.0 * nom_epsilon_ : 0.5 * nom_epsilon_;
    // base_hmc.hpp:131-139
    if (nom_epsilon_ > 1e7) throw std::runtime_error("Posterior is improper. Please check your model.");
    if (nom_epsilon_ == 0) throw std::runtime_error("No acceptably small step size could be found. Perhaps the posterior is not continuous?");
  }
  z_ = z_init;
}

bool NutsSampler::build_tree(int depth, Point& z_propose, Vec& p_sharp_beg, Vec& p_sharp_end, Vec& rho, Vec& p_beg, Vec& p_end, double H0,
                             double sign, int& n_leapfrog, double& log_sum_weight, double& sum_metro_prob)
{
  if (depth == 0) {
    evolve(z_, sign * epsilon_);
    ++n_leapfrog;
    double h = hamiltonian(z_); if (std::isnan(h)) h = kInf;
    if ((h - H0) > max_deltaH_) divergent_ = true;
    log_sum_weight = log_sum_exp(log_sum_weight, H0 - h);
    sum_metro_prob += (H0 - h > 0) ? 1.0 : std::exp(H0 - h);
    z_propose = z_;
    sharp(z_, p_sharp_beg); p_sharp_end = p_sharp_beg;
    for (int i = 0; i < d_; ++i) rho[(size_t) i] += z_.p[(size_t) i];
    p_beg = z_.p; p_end = p_beg;
    return !divergent_;
  }
  double lsw_init = -kInf;
  if ((int) levels_.size() <= depth) levels_.resize((size_t) depth + 1);
  Level& W = levels_[(size_t) depth];
  const size_t d = (size_t) d_;
  Vec& p_init_end = W.p_init_end; Vec& p_sharp_init_end = W.p_sharp_init_end; Vec& rho_init = W.rho_init;
  p_init_end.resize(d); p_sharp_init_end.resize(d); rho_init.assign(d, 0.0);
  if (!build_tree(depth - 1, z_propose, p_sharp_beg, p_sharp_init_end, rho_init, p_beg, p_init_end, H0, sign, n_leapfrog, lsw_init, sum_metro_prob)) return false;
  Point& z_propose_final = W.z_propose_final;
  z_propose_final = z_;
  double lsw_final = -kInf;
  Vec& p_final_beg = W.p_final_beg; Vec& p_sharp_final_beg = W.p_sharp_final_beg; Vec& rho_final = W.rho_final;
  p_final_beg.resize(d); p_sharp_final_beg.resize(d); rho_final.assign(d, 0.0);
  if (!build_tree(depth - 1, z_propose_final, p_sharp_final_beg, p_sharp_end, rho_final, p_final_beg, p_end, H0, sign, n_leapfrog, lsw_final, sum_metro_prob)) return false;
  double lsw_subtree = log_sum_exp(lsw_init, lsw_final);
  log_sum_weight = log_sum_exp(log_sum_weight, lsw_subtree);
  if (lsw_final > lsw_subtree) z_propose = z_propose_final;
  else if (rng_uniform(rng_) < std::exp(lsw_final - lsw_subtree)) z_propose = z_propose_final;
  Vec& rho_subtree = W.rho_subtree;
  rho_subtree.resize(d);
  for (int i = 0; i < d_; ++i) { rho_subtree[(size_t) i] = rho_init[(size_t) i] + rho_final[(size_t) i]; rho[(size_t) i] += rho_subtree[(size_t) i]; }
  bool persist = no_u_turn(p_sharp_beg, p_sharp_end, rho_subtree);
  for (int i = 0; i < d_; ++i) rho_subtree[(size_t) i] = rho_init[(size_t) i] + p_final_beg[(size_t) i];
  persist &= no_u_turn(p_sharp_beg, p_sharp_final_beg, rho_subtree);
  for (int i = 0; i < d_; ++i) rho_subtree[(size_t) i] = rho_final[(size_t) i] + p_init_end[(size_t) i];
  persist &= no_u_turn(p_sharp_init_end, p_sharp_end, rho_subtree);
  return persist;
}

void NutsSampler::transition()
{
  epsilon_ = nom_epsilon_;
  if (epsilon_jitter_) epsilon_ *= 1.0 + epsilon_jitter_ * (2.0 * rng_uniform(rng_) - 1.0);
  z_.q = cont_params_;
  sample_p(z_);
  update_potential_gradient(z_);
  Point z_fwd = z_, z_bck = z_, z_sample = z_, z_propose = z_;
  Vec p_fwd_fwd = z_.p, p_sharp_fwd_fwd; sharp(z_, p_sharp_fwd_fwd);
  Vec p_fwd_bck = z_.p, p_sharp_fwd_bck = p_sharp_fwd_fwd;
  Vec p_bck_fwd = z_.p, p_sharp_bck_fwd = p_sharp_fwd_fwd;
  Vec p_bck_bck = z_.p, p_sharp_bck_bck = p_sharp_fwd_fwd;
  Vec rho = z_.p, rho_fwd((size_t) d_), rho_bck((size_t) d_), rho_ext((size_t) d_);
  double log_sum_weight = 0.0;
  const double H0 = hamiltonian(z_);
  int n_leapfrog = 0; double sum_metro_prob = 0.0;
  depth_ = 0; divergent_ = false;
  while (depth_ < max_depth_) {
    std::fill(rho_fwd.begin(), rho_fwd.end(), 0.0); std::fill(rho_bck.begin(), rho_bck.end(), 0.0);
    bool valid_subtree;
    double lsw_subtree = -kInf;
    if (rng_uniform(rng_) > 0.5) {
      z_ = z_fwd; rho_bck = rho; p_bck_fwd = p_fwd_fwd; p_sharp_bck_fwd = p_sharp_fwd_fwd;
      valid_subtree = build_tree(depth_, z_propose, p_sharp_fwd_bck, p_sharp_fwd_fwd, rho_fwd, p_fwd_bck, p_fwd_fwd, H0, 1.0, n_leapfrog, lsw_subtree, sum_metro_prob);
      z_fwd = z_;
    } else {
      z_ = z_bck; rho_fwd = rho; p_fwd_bck = p_bck_bck; p_sharp_fwd_bck = p_sharp_bck_bck;
      valid_subtree = build_tree(depth_, z_propose, p_sharp_bck_fwd, p_sharp_bck_bck, rho_bck, p_bck_fwd, p_bck_bck, H0, -1.0, n_leapfrog, lsw_subtree, sum_metro_prob);
      z_bck = z_;
    }
    if (!valid_subtree) break;
    ++depth_;
    if (lsw_subtree > log_sum_weight) z_sample = z_propose;
    else if (rng_uniform(rng_) < std::exp(lsw_subtree - log_sum_weight)) z_sample = z_propose;
    log_sum_weight = log_sum_exp(log_sum_weight, lsw_subtree);
    for (int i = 0; i < d_; ++i) rho[(size_t) i] = rho_bck[(size_t) i] + rho_fwd[(size_t) i];
    bool persist = no_u_turn(p_sharp_bck_bck, p_sharp_fwd_fwd, rho);
    for (int i = 0; i < d_; ++i) rho_ext[(size_t) i] = rho_bck[(size_t) i] + p_fwd_bck[(size_t) i];
    persist &= no_u_turn(p_sharp_bck_bck, p_sharp_fwd_bck, rho_ext);
    for (int i = 0; i < d_; ++i) rho_ext[(size_t) i] = rho_fwd[(size_t) i] + p_bck_fwd[(size_t) i];
    persist &= no_u_turn(p_sharp_bck_fwd, p_sharp_fwd_fwd, rho_ext);
    if (!persist) break;
  }
  n_leapfrog_ = n_leapfrog;
  accept_stat_ = sum_metro_prob / (double) n_leapfrog;
  z_ = z_sample;
  energy_ = hamiltonian(z_);
  cont_params_ = z_.q;
  lp_ = -z_.V;
}

void NutsSampler::window_restart() { window_counter_ = 0; window_size_ = base_window_; next_window_ = init_buffer_ + window_size_ - 1u; }
void NutsSampler::set_window_params(uint32_t num_warmup, uint32_t init_buffer, uint32_t term_buffer, uint32_t base_window)
{
  if (num_warmup < 20u) return;
  if (init_buffer + base_window + term_buffer > num_warmup) {
    num_warmup_ = num_warmup;
    init_buffer_ = (uint32_t) (0.15 * num_warmup);
    term_buffer_ = (uint32_t) (0.10 * num_warmup);
    base_window_ = num_warmup - (init_buffer_ + term_buffer_);
    return;   // the reference returns here without restart() (windowed_adaptation.hpp:49-75)
  }
  num_warmup_ = num_warmup; init_buffer_ = init_buffer; term_buffer_ = term_buffer; base_window_ = base_window;
  window_restart();
}
void NutsSampler::compute_next_window()
{
  if (next_window_ == num_warmup_ - term_buffer_ - 1u) return;
  window_size_ *= 2u;
  next_window_ = window_counter_ + window_size_;
  if (next_window_ == num_warmup_ - term_buffer_ - 1u) return;
  uint32_t boundary = next_window_ + 2u * window_size_;
  if (boundary >= num_warmup_ - term_buffer_) next_window_ = num_warmup_ - term_buffer_ - 1u;
}
bool NutsSampler::learn_variance()
{
  const bool in_window = window_counter_ >= init_buffer_ && window_counter_ < num_warmup_ - term_buffer_ && window_counter_ != num_warmup_;
  if (in_window) {
    wf_n_ += 1.0;
    for (int i = 0; i < d_; ++i) {
      double delta = z_.q[(size_t) i] - wf_m_[(size_t) i];
      wf_m_[(size_t) i] += delta / wf_n_;
      wf_m2_[(size_t) i] += delta * (z_.q[(size_t) i] - wf_m_[(size_t) i]);
    }
  }
  const bool end_window = window_counter_ == next_window_ && window_counter_ != num_warmup_;
  if (end_window) {
    compute_next_window();
    const double n = wf_n_;
    for (int i = 0; i < d_; ++i) {
      double var = inv_metric_[(size_t) i];
      if (n > 1.0) var = wf_m2_[(size_t) i] / (n - 1.0);
      inv_metric_[(size_t) i] = (n / (n + 5.0)) * var + 1e-3 * (5.0 / (n + 5.0));
    }
    wf_n_ = 0.0; std::fill(wf_m_.begin(), wf_m_.end(), 0.0); std::fill(wf_m2_.begin(), wf_m2_.end(), 0.0);
    ++window_counter_;
    return true;
  }
  ++window_counter_;
  return false;
}
void NutsSampler::learn_stepsize(double adapt_stat)
{
  sa_counter_ += 1.0;
  adapt_stat = adapt_stat > 1 ? 1 : adapt_stat;
  const double eta = 1.0 / (sa_counter_ + sa_t0_);
  sa_s_bar_ = (1.0 - eta) * sa_s_bar_ + eta * (sa_delta_ - adapt_stat);
  const double x = sa_mu_ - sa_s_bar_ * std::sqrt(sa_counter_) / sa_gamma_;
  const double x_eta = std::pow(sa_counter_, -sa_kappa_);
  sa_x_bar_ = (1.0 - x_eta) * sa_x_bar_ + x_eta * x;
  nom_epsilon_ = std::exp(x);
}

void NutsSampler::run(bool warmup, double* out)
{
  (void) warmup;
  static const bool host_prof = getenv("S4B_HOST_PROF") != nullptr;
  for (int m = 0; m < ctl_.skip; ++m) {
    std::chrono::steady_clock::time_point t0;
    if (host_prof) t0 = std::chrono::steady_clock::now();
    transition();
    if (host_prof) {
      prof_tr_ns_ += std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t0).count();
      if (++prof_tr_calls_ % 50 == 0)
        fprintf(stderr, "[s4b host prof] transitions %lld: %.1f us each, %.1f evaluations each, %.0f ns per evaluation inside log_prob_grad, %.0f ns per evaluation around it\n",
                prof_tr_calls_, prof_tr_ns_ / prof_tr_calls_ / 1e3, (double) prof_lp_calls_ / prof_tr_calls_, prof_lp_ns_ / prof_lp_calls_,
                (prof_tr_ns_ - prof_lp_ns_) / prof_lp_calls_);
    }
    if (adapt_flag_) {
      learn_stepsize(accept_stat_);
      if (learn_variance()) {
        init_stepsize();
        sa_mu_ = std::log(10 * nom_epsilon_);
        sa_counter_ = 0; sa_s_bar_ = 0; sa_x_bar_ = 0;
      }
    }
  }
  if (out) {
    out[0] = lp_; out[1] = accept_stat_; out[2] = epsilon_; out[3] = depth_; out[4] = n_leapfrog_; out[5] = divergent_ ? 1.0 : 0.0; out[6] = energy_;
    model_.write_array(cont_params_.data(), out + 7);
  }
}

void NutsSampler::disengage_adaptation() { adapt_flag_ = false; nom_epsilon_ = std::exp(sa_x_bar_); }

}  // namespace s4b

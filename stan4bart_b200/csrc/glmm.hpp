// stan4bart_b200/csrc/glmm.hpp -- device-backed GLMM density (continuous.stan) for the host NUTS.
#pragma once

#include "../../include/stan4bart_b200.h"
#include "s4b_common.cuh"
#include "shard.hpp"

#include <complex>
#include <string>
#include <vector>

namespace s4b {

struct GlmmDev {
  long long N, npad;
  int K, q, slots, bins;
  const double* X;        // [K][npad]
  const double* r;        // y - offset, [npad]
  const double* wt;       // observation weights [npad], nullptr = unweighted
  const int* zidx;        // [slots][npad]   (ELL; the padding of a short row points at columns the row does not use, with value 0)
  const double* zval;     // [slots][npad]
  unsigned int ones_mask; // bit s set => every stored value of slot s is exactly 1.0
  int row_distinct;       // every row's `slots` column indices are pairwise distinct (padding included): a row's bin updates are independent
  const double* theta;    // [K + q] : beta then b
  double* partials;       // [(1 + K + q)][G]
  double* result;         // [1 + K + q]
  unsigned int* ticket;
};

class GlmmModel {
 public:
  // `shard` != nullptr: d holds this rank's rows of an observation-sharded model; the (1 + K + q) data-term reductions and
  // the Gram matrix are summed over the ranks in rank order (shard.hpp), so every rank evaluates the same density
  GlmmModel(const s4b_glmm_data& d, cudaStream_t stream, ShardContext* shard = nullptr);
  ~GlmmModel();
  GlmmModel(const GlmmModel&) = delete;
  GlmmModel& operator=(const GlmmModel&) = delete;

  int num_params() const { return num_params_; }
  int num_constrained() const { return num_params_ + has_aux_ + K_ + q_ + len_theta_L_; }
  long long N() const { return N_; }
  int K() const { return K_; }
  int q() const { return q_; }
  bool has_aux() const { return has_aux_ != 0; }

  void set_offset_host(const double* offset);
  void set_response_host(const double* y);
  void set_offset_device(const double* d_offset);     // copies
  void set_response_device(const double* d_y);        // copies
  void set_inputs_device(const double* d_offset, const double* d_y);   // either may be NULL (unchanged); one fused pass
  // returns status: 0 ok, 1 non-finite lp / gradient (maps to V = +inf in the sampler)
  int log_prob_grad(const double* q, double* lp, double* grad);
  void write_array(const double* q, double* out) const;
  // names of the rows of write_array (constrained_param_names, continuous.hpp:3115-3204), in order
  std::vector<std::string> param_names() const;
  // device result; beta / b are host pointers
  void parametric_mean_device(const double* beta, const double* b, double* d_out, bool include_fixed, bool include_random);
  void parametric_mean_host(const double* constrained, double* out, bool include_fixed, bool include_random);
  // device pass: S = sum e^2, X'e, Z'e at (beta, b)
  void data_terms(const double* beta, const double* b, double* S, double* gbeta, double* gb);
  void launch_data_pass();
  double time_data_pass(int reps, int flush_l2);
  bool bulk_pass() const { return bulk_; }
  // the same three quantities through the sweep-level expansion (mode 1) or the device pass (mode 0)
  void data_terms_auto(const double* beta, const double* b, double* S, double* gbeta, double* gb);
  // 0: one device pass per evaluation; 1: one device pass per Gibbs sweep + exact quadratic expansion (default when K + q is small)
  void set_mode(int mode);
  int mode() const { return mode_; }
  long long num_device_passes() const { return num_passes_; }
  long long num_grad_evals() const { return num_grad_; }
  cudaStream_t stream() const { return stream_; }

 private:
  struct Params;
  void transform(const double* q, Params& P) const;
  Params* scratch_ = nullptr;                 // reused by log_prob_grad: no heap traffic per evaluation
  std::vector<double> gbeta_, gb_, lg_delta_, lg_shape_, lg_reg_, lg_reg2_;
  void refresh_r();

  bool sharded() const { return shard_ != nullptr && shard_->world() > 1; }

  cudaStream_t stream_;
  ShardContext* shard_ = nullptr;
  long long N_ = 0, npad_ = 0, N_total_ = 0;
  std::vector<double> prior_df_, d_extra_; std::vector<int> num_normals_;
  double global_prior_df_ = 0, global_prior_scale_ = 0, slab_df_ = 0, slab_scale_ = 0;
  int hs_ = 0, len_zbeta_ = 0, len_extra_ = 0;
  void coef_beta(const std::complex<double>* z, const std::complex<double>* ex, std::complex<double> aux, std::complex<double>* beta) const;
  int K_ = 0, q_ = 0, t_ = 0, len_theta_L_ = 0, len_rho_ = 0, len_z_T_ = 0, len_conc_ = 0, num_params_ = 0, has_aux_ = 0;
  int is_binary_ = 0, prior_dist_ = 0, prior_dist_for_aux_ = 0;
  double prior_scale_for_aux_ = 0, prior_mean_for_aux_ = 0, prior_df_for_aux_ = 0;
  std::vector<double> prior_scale_, prior_mean_, shape_, scale_, delta_, regularization_;
  std::vector<int> p_, l_;
  int slots_ = 0, grid_ = 1, block_ = 256, num_sms_ = 1;
  unsigned int ones_mask_ = 0;
  int row_distinct_ = 0;
  size_t smem_bytes_ = 0;
  bool bulk_ = false;                       // binned data pass staged through shared memory by bulk copies (k_glmm_data_terms_bulk)
  int bulk_tile_ = 0, bulk_stages_ = 0, bulk_tile_bytes_ = 0, bulk_grid_ = 1;
  size_t bulk_smem_ = 0;
  long long num_grad_ = 0, num_passes_ = 0;
  // sweep-level sufficient statistics: with e = r - A theta (A = [X Z]) the data terms are an exact quadratic in theta,
  //   S(theta0 + d) = S0 - 2 g0'd + d'G d,   A'e(theta0 + d) = g0 - G d,   G = A'A (fixed), (S0, g0) from one pass at theta0
  int mode_ = 0;
  bool expansion_valid_ = false;
  std::vector<double> gram_, theta0_, g0_, dl_, Gd_;
  // K + q > 512: the Gram matrix in pieces -- X'WX, X'WZ dense, Z'WZ as compressed sparse rows
  bool sparse_gram_ = false;
  double* d_gram_ = nullptr; double* d_dl_ = nullptr; double* h_dl_ = nullptr;    // dense Gram matrix on the device (filled-in Z'WZ), its operand / result
  long long num_gram_products_ = 0;
  void expand_sparse(const double* beta, const double* b, double* S, double* gbeta, double* gb);
  std::vector<double> gxx_, gxz_, gz_val_; std::vector<long long> gz_ptr_; std::vector<int> gz_col_;
  double S0_ = 0.0;

  double *d_X_ = nullptr, *d_y_ = nullptr, *d_offset_ = nullptr, *d_r_ = nullptr, *d_wt_ = nullptr, *d_zval_ = nullptr;
  int* d_zidx_ = nullptr;
  // column path of the data pass (K + q beyond the shared-memory bins): w e, and Z as compressed sparse columns
  bool columns_ = false;
  double *d_we_ = nullptr, *d_col_val_ = nullptr; long long* d_col_ptr_ = nullptr; int* d_col_obs_ = nullptr;
  double *d_theta_ = nullptr, *d_partials_ = nullptr, *d_result_ = nullptr, *d_tmp_ = nullptr;
  unsigned int* d_ticket_ = nullptr;
  double* h_pinned_ = nullptr;   // [2 * (1 + K + q)] : theta out, result in
};

}  // namespace s4b

// stan4bart_b200/csrc/bart.hpp -- host object for the device-resident BART sampler.
#pragma once

#include "../../include/stan4bart_b200.h"
#include "s4b_common.cuh"
#include "shard.hpp"

#include <string>
#include <vector>

namespace s4b {

// everything a kernel needs, passed by value
struct BartDev {
  long long n, npad;
  long long obs_offset;       // global index of local row 0 (sharded chains; keys the latent draws)
  const uint8_t* xt;
  double* R; double* yresc; const double* y; double* offset;
  StepDesc* desc; DTree* trees; BartParams* params; const double* pgrow; RngState* rng;
  double* partials; unsigned int* ticket;
  uint2* packs;               // streamed sweep: cached node indices of every quad, two buffers of nquad entries
  const double* wt;           // observation weights [npad] (zero padded), nullptr = unweighted
  double* trace; unsigned long long trace_cap; unsigned long long* trace_len;
  double* stats_out;
  unsigned long long* prof;   // cycle counters of the controller phases (last block, thread 0)
};

class BartFit {
 public:
  // `shard` != nullptr: this fit holds rows [obs_offset, obs_offset + n) of an observation-sharded chain (shard.hpp);
  // every rank must then make the same sequence of calls
  BartFit(const s4b_bart_config& cfg, const double* y, const double* x, const double* x_test, cudaStream_t stream, ShardContext* shard = nullptr);
  ~BartFit();
  BartFit(const BartFit&) = delete;
  BartFit& operator=(const BartFit&) = delete;

  void set_offset_host(const double* offset, bool update_scale);
  void set_offset_device(const double* d_offset, bool update_scale);
  void set_sigma(double sigma);
  // setResponse (dbarts table entry, init.cpp:68): a new response vector, fits and offset unchanged
  void set_response_host(const double* y);
  void sample_trees_from_prior();
  // k of the leaf prior (a draw per sweep when cfg.k_df > 0)
  double current_k();
  const double* d_k_ptr() const { return &d_params_->k; }
  void draw_k();
  bool k_modeled() const { return cfg_.k_df > 0.0; }
  double config_k() const { return cfg_.k; }
  // `thin` sweeps; results stay on device (train_out / test_out / latent_out)
  void run_sweeps();
  static void run_sweeps_batched(BartFit* const* fits, int count);       // several chains, one launch (grid.y = chain)
  static void batched_sweeps_on(BartFit* const* fits, int count, cudaStream_t st);
  void after_batched_sweeps();
  void draw_k_on(cudaStream_t st);
  int num_sms() const { return num_sms_; }
  void collect_results(double* train, double* test, uint32_t* varcount, double* sigma);
  // runSamplerWithResults with host result buffers (any may be NULL)
  void run(double* train, double* test, uint32_t* varcount, double* sigma);
  void store_latents(double* out);
  void predict(const double* x_test, long long rows, const double* test_offset, double* out);
  void get_trees(int32_t* tree_no, long long* n_obs, int32_t* var, double* value);
  long long num_nodes();
  // printInitialSummary (the dbarts table entry stan4bart calls at init.cpp:981): the lines the reference's tests parse
  std::string summary() const;
  // keepTrees (dbarts control$keepTrees; stan4bart_exportBARTState / createStoredBARTSampler / predictBART, init.cpp:354-446):
  // with a store of `capacity` samples every runSamplerWithResults call appends the trees and the response scale of its
  // kept draw; stored draws can be predicted from and flattened like the live sampler
  void set_keep_trees(long long capacity);
  long long num_stored() const { return store_len_; }
  // setControl(keepTrees = ...) as the Gibbs loop uses it (init.cpp:737-744): warm-up runs do not append to the store
  void set_keep_trees_active(bool on) { keep_trees_active_ = on; }
  void predict_stored(const double* x_test, long long rows, const double* test_offset, long long first, long long count, double* out);
  long long num_stored_nodes(long long sample);
  // (min, range) of the response scale of stored draws first .. first + count - 1 (what their leaf values are expressed in)
  void get_stored_scales(long long first, long long count, double* out2);
  void get_stored_trees(long long sample, int32_t* tree_no, long long* n_obs, int32_t* var, double* value);
  // exportBARTState (init.cpp:409-446): the stored draws, the cut points and what prediction needs, as one host blob
  long long stored_export_size();
  void stored_export(void* out, long long bytes);
  BartParams params();
  void varcount_device(unsigned int* d_out);

  // parity / measurement instrumentation
  void node_assignment(int tree, long long* out);
  int leaf_stats(int tree, int max_leaves, long long* heap, long long* count, double* sum, double* sumsq);
  void launch_leaf_stats(int tree, int num_leaves = -1);      // num_leaves < 0: asked from the device (one small synchronous copy)
  int tree_num_leaves(int tree);
  void get_residual(double* out);
  void set_trace(size_t cap_records);
  size_t get_trace(double* out, size_t cap_records);
  void set_tape(const double* tape, size_t len);
  void set_record(size_t cap);
  size_t get_record(double* out, size_t cap);
  unsigned long long rng_counter();
  void set_use_graph(bool g) { use_graph_ = g; if (!g) sweep_mode_ = 0; else if (sweep_mode_ == 0) sweep_mode_ = 1; }
  // 0: one launch per tree step; 1: the same kernels captured in a CUDA graph; 2: persistent on-chip sweep kernel
  // (sweep_kernel.cuh), the default whenever the chain fits
  void set_sweep_mode(int m);
  int sweep_mode() const { return sweep_mode_; }
  bool persistent_fits() const { return persistent_nq_ > 0; }
  // dbarts returns training fits with the offset added (init.cpp:828-829 subtracts it again); the Gibbs
  // loop asks for the tree-only fit directly
  void set_add_offset(bool a) { if (a != add_offset_) { add_offset_ = a; invalidate_graph(); } }
  void check_error_flag();

  long long n() const { return n_; }
  long long n_test() const { return nt_; }
  int p() const { return p_; }
  int num_trees() const { return T_; }
  bool is_binary() const { return cfg_.is_binary != 0; }
  cudaStream_t stream() const { return stream_; }
  double* d_train_out() const { return d_train_out_; }     // BART fit + offset, original units
  double* d_test_out() const { return d_test_out_; }
  double* d_latent_out() const { return d_latent_out_; }   // full latents (binary)
  double* d_offset() const { return d_offset_; }
  int grid() const { return grid_; }
  bool sharded() const { return shard_ != nullptr && shard_->world() > 1; }
  long long num_tree_steps() const { return num_tree_steps_; }
  // device time (CUDA events on the launching stream) spent in the sweep graphs since the last reset
  double tree_step_ms(bool reset);
  // cycles spent by the last block in: [0] its own pass, [1] partial reduction, [2] tree load, [3] decision + leaf draws,
  // [4] write-back + next tree load, [5] proposal, [6] descriptor publish, [7] number of steps
  void get_profile(unsigned long long* out24, bool reset);
  void set_profile(bool on) { if (on != profile_on_) { profile_on_ = on; invalidate_graph(); } }

 private:
  BartDev dev() const;
  ShardDev shard_dev() const;
  void launch_sweep_kernels(bool last_thin);
  void launch_persistent_sweep(bool last_thin);
  void setup_persistent();
  void test_fits_device(const uint8_t* d_xt, long long rows, long long rows_pad, const double* d_off, double* d_out, const DTree* trees = nullptr,
                        const double* scale = nullptr);
  void snapshot_trees();
  std::vector<DTree> download_stored(long long sample);
  void flatten_trees(std::vector<DTree>& trees, int32_t* tree_no, long long* n_obs, int32_t* var, double* value) const;
  void bin_matrix(const double* x, long long rows, long long rows_pad, std::vector<uint8_t>& out) const;
  std::vector<DTree> download_trees();
  void invalidate_graph();

  s4b_bart_config cfg_;
  cudaStream_t stream_;
  ShardContext* shard_ = nullptr;
  long long n_ = 0, nt_ = 0, npad_ = 0, npad_t_ = 0;
  int p_ = 0, T_ = 0, num_sms_ = 0, blocks_per_sm_ = 3, grid_ = 1, grid_ew_ = 1;
  std::vector<double> cuts_;
  bool scale_initialised_ = false;
  bool use_graph_ = true;
  bool add_offset_ = true;
  bool test_aliases_train_ = false;
  int sweep_mode_ = 1;
  int persistent_nq_ = 0, persistent_grid_ = 0;       // nq = kStreamNq: residuals streamed from global memory (L2)
  uint2* d_packs_ = nullptr;
  void gather_distinct_over_shards(std::vector<double>& sorted_distinct);   // use_quantiles on sharded rows (setup only)
  std::vector<int> cutless_;                                  // use_quantiles: constant predictors (no cut point; split weight 0)
  int* d_ncuts_var_ = nullptr; std::vector<int> ncuts_var_;   // bart_args n.cuts per predictor (empty = n_cuts everywhere)
  double* d_wt_ = nullptr;       // observation weights (zero padded), nullptr = unweighted
  uint32_t* d_split_w_ = nullptr;
  std::vector<double> split_probs_;       // normalised bart_args split.probs (empty = uniform)
  DTree* d_store_ = nullptr; double* d_store_scale_ = nullptr; long long store_cap_ = 0, store_len_ = 0;
  size_t persistent_smem_ = 0;
  // pipelined sweep kernel (sweep_pipe.cuh): ring of partial rows, barrier counters, per-step cell tables, per-sweep flag
  void* d_batch_args_ = nullptr; int batch_cap_ = 0;        // argument block of a batched launch (owned by the first fit of the batch)
  bool pipe_enabled_ = false;
  int pipe_count_words_ = 0;
  size_t pipe_smem_ = 0;
  unsigned long long* d_pipe_ring_ = nullptr; unsigned int* d_pipe_flag_ = nullptr; void* d_pipe_infos_ = nullptr;
  long long pipe_sweeps_ = 0, pipe_launches_ = 0;
  unsigned long long* d_pipe_ran_ = nullptr;
  int* d_pipe_pos_ = nullptr;          // first step still to do, one entry per launch of a sweep's segment sequence
 public:
  // tree steps that ran in the pipelined kernel (counted on the device)
  long long pipe_sweeps_done();
  // steps that did not fit the pipelined kernel so far: [0] all, [1] tree too large, [2] more than 8 statistic slots, [3] too many cells
  void pipe_reasons(unsigned int* out4);
  // sweeps launched through the pipelined kernel path so far (whether a given sweep ran pipelined is decided on the device)
  long long pipe_sweeps() const { return pipe_sweeps_; }
  bool pipe_enabled() const { return pipe_enabled_; }
  void set_pipe_enabled(bool on) { pipe_enabled_ = on && pipe_smem_ > 0; }
 private:
  unsigned int* d_barrier_ = nullptr;
  double* d_partials2_ = nullptr;
  double* d_tables_ = nullptr;
  StepDesc* d_descs_ = nullptr;
  double2* d_draws_ = nullptr;
  bool tape_set_ = false, rec_set_ = false, sequential_rng_ = false;
  int partial_stride_ = 0;
  int overlap_walk_ = 1;
  bool profile_on_ = false;
  bool keep_trees_active_ = true;
  // stand-alone leaf-statistics kernel (leaf_stats.cuh)
  double* d_leaf_partials_ = nullptr; unsigned int* d_leaf_ticket_ = nullptr; static constexpr int kLeafVariants = 4; int leaf_grid_[kLeafVariants] = {}; size_t leaf_smem_ = 0; bool leaf_generic_ = false;
  long long num_tree_steps_ = 0;

  uint8_t* d_xt_ = nullptr; uint8_t* d_xt_test_ = nullptr;
  double *d_R_ = nullptr, *d_yresc_ = nullptr, *d_y_ = nullptr, *d_offset_ = nullptr, *d_offset_in_ = nullptr;
  double *d_train_out_ = nullptr, *d_test_out_ = nullptr, *d_latent_out_ = nullptr;
  double *d_partials_ = nullptr, *d_minmax_ = nullptr, *d_stats_out_ = nullptr, *d_pgrow_ = nullptr, *d_scale_factor_ = nullptr;
  StepDesc* d_desc_ = nullptr; DTree* d_trees_ = nullptr; BartParams* d_params_ = nullptr; RngState* d_rng_ = nullptr;
  unsigned long long* d_prof_ = nullptr;
  unsigned int* d_ticket_ = nullptr; unsigned long long* d_trace_len_ = nullptr; unsigned int* d_varcount_ = nullptr;
  double* d_trace_ = nullptr; size_t trace_cap_ = 0;
  double* d_tape_ = nullptr; double* d_rec_ = nullptr;
  cudaGraphExec_t graph_exec_ = nullptr, graph_exec_thin_ = nullptr;
  cudaEvent_t ev_start_ = nullptr, ev_end_ = nullptr;
  bool ev_pending_ = false;
  double sweep_ms_ = 0.0;
};

// createStoredBARTSampler (init.cpp:409-446): prediction from exported draws without the training data or a live sampler
class StoredBart {
 public:
  StoredBart(const void* blob, long long bytes, cudaStream_t stream);
  ~StoredBart();
  StoredBart(const StoredBart&) = delete;
  StoredBart& operator=(const StoredBart&) = delete;
  long long count() const { return count_; }
  int p() const { return p_; }
  void get_scales(long long first, long long count, double* out2) const;
  void predict(const double* x_test, long long rows, const double* test_offset, long long first, long long count, double* out);

 private:
  cudaStream_t stream_;
  int p_ = 0, T_ = 0, n_cuts_ = 0, is_binary_ = 0;
  long long count_ = 0;
  std::vector<double> cuts_, scales_;
  DTree* d_store_ = nullptr; double* d_scale_ = nullptr; BartParams* d_params_ = nullptr;
};

}  // namespace s4b

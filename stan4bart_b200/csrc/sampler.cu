// stan4bart_b200/csrc/sampler.cu
// The alternating BART <-> Stan Gibbs loop and the extern "C" layer.
//   createSampler  /root/reference/src/init.cpp:190-310
//   run            /root/reference/src/init.cpp:678-965 (loop body :752-917)
// All N-length vectors the reference keeps on the host (bartOffset, stanOffset, bartLatents,
// init.cpp:143-145) live in HBM here; per iteration only the 58-odd Stan row and the tiny
// gradient reductions cross PCIe unless the caller asked for the fits (keep_fits).
#include "bart.hpp"
#include "glmm.hpp"
#include "nuts.hpp"

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>

namespace s4b {

// dst = a + b (b may be NULL: plain copy).  dst may alias a or b (in-place sums of the offset vectors): no __restrict__
__global__ void k_sum2(long long n, const double* a, const double* b, double* dst)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) dst[i] = b != nullptr ? a[i] + b[i] : a[i];
}

enum { kOffsetDefault = 0, kOffsetFixef = 1, kOffsetRanef = 2, kOffsetBart = 3, kOffsetParametric = 4 };   // init.cpp:84-88

// running sums of the kept draws (training fit, parametric mean, test fit) in one launch
__global__ void k_accumulate3(long long n, const double* __restrict__ s0, double* __restrict__ a0, const double* __restrict__ s1, double* __restrict__ a1,
                              long long n2, const double* __restrict__ s2, double* __restrict__ a2)
{
  const long long m = n > n2 ? n : n2;
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (long long) gridDim.x * blockDim.x) {
    if (i < n) { a0[i] += s0[i]; a1[i] += s1[i]; }
    if (i < n2) a2[i] += s2[i];
  }
}

// Several chains of one GPU whose BART sweeps are batched into ONE launch per Gibbs iteration (grid.y = chain).
// Every chain keeps its own host thread (its NUTS runs there) and stream; at its BART block the thread arrives here instead of
// launching its own sweep kernels.  The last chain to arrive orders the group's stream after every chain's stream, enqueues the
// batched step for all of them and records an event every chain's stream then waits for.  The chains advance in lock-step, one
// sweep at a time; a chain that never arrives (an error in its thread) times the others out instead of hanging them.
class BatchGroup {
 public:
  explicit BatchGroup(int count) : count_(count), fits_((size_t) count, nullptr), ready_((size_t) count, nullptr)
  {
    if (count < 1) throw std::invalid_argument("batch group: count < 1");
    S4B_CUDA(cudaStreamCreateWithFlags(&st_, cudaStreamNonBlocking));
    for (auto& e : ready_) S4B_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    S4B_CUDA(cudaEventCreateWithFlags(&done_, cudaEventDisableTiming));
  }
  ~BatchGroup() { for (auto e : ready_) cudaEventDestroy(e); cudaEventDestroy(done_); cudaStreamDestroy(st_); }
  int count() const { return count_; }
  long long launches() const { return launches_; }
  void run(BartFit* fit, cudaStream_t chain_stream)
  {
    std::unique_lock<std::mutex> lk(m_);
    const int idx = arrived_++;
    fits_[(size_t) idx] = fit;
    S4B_CUDA(cudaEventRecord(ready_[(size_t) idx], chain_stream));
    const unsigned long long gen = gen_;
    if (arrived_ == count_) {
      failed_.clear();
      try {
        for (int i = 0; i < count_; ++i) S4B_CUDA(cudaStreamWaitEvent(st_, ready_[(size_t) i], 0));
        BartFit::batched_sweeps_on(fits_.data(), count_, st_);
        S4B_CUDA(cudaEventRecord(done_, st_));
        ++launches_;
      } catch (const std::exception& e) { failed_ = e.what(); }
      arrived_ = 0; ++gen_;
      cv_.notify_all();
    } else if (!cv_.wait_for(lk, std::chrono::seconds(120), [&] { return gen_ != gen; })) {
      --arrived_;
      throw std::runtime_error("batch group: the other chains did not reach their BART block within 120 s");
    }
    if (!failed_.empty()) throw std::runtime_error("batch group: " + failed_);
    S4B_CUDA(cudaStreamWaitEvent(chain_stream, done_, 0));
    lk.unlock();
    fit->after_batched_sweeps();
  }
 private:
  int count_, arrived_ = 0; unsigned long long gen_ = 0; long long launches_ = 0;
  std::mutex m_; std::condition_variable cv_;
  std::vector<BartFit*> fits_; std::vector<cudaEvent_t> ready_; cudaEvent_t done_ = nullptr; cudaStream_t st_ = nullptr;
  std::string failed_;
};

class GibbsSampler {
 public:
  GibbsSampler(const s4b_bart_config& bcfg, const double* y_bart, const double* x_bart, const double* x_test, const s4b_glmm_data& gdata,
               const s4b_stan_control& sctl, const s4b_common_control& cctl, const double* bart_offset_init, cudaStream_t stream,
               ShardContext* shard = nullptr)
      : cc_(cctl), stream_(stream), glmm_(gdata, stream, shard), nuts_(glmm_, sctl, 1, cctl.warmup), bart_(bcfg, y_bart, x_bart, x_test, stream, shard)
  {
    n_ = bcfg.n; nt_ = bcfg.n_test; p_ = (int) bcfg.p;
    if (gdata.N != n_) throw std::invalid_argument("sampler: BART and Stan data disagree on N");
    num_pars_ = nuts_.num_pars();
    stan_curr_.assign((size_t) num_pars_, 0.0);
    auto dalloc = [&](double** p, size_t count) { S4B_CUDA(cudaMalloc(p, sizeof(double) * std::max<size_t>(count, 1))); zero_device_sync(*p, sizeof(double) * std::max<size_t>(count, 1), stream_); };
    dalloc(&d_bart_offset_, (size_t) n_); dalloc(&d_mean_train_, (size_t) n_); dalloc(&d_mean_param_, (size_t) n_); dalloc(&d_mean_test_, (size_t) std::max<long long>(nt_, 1));
    S4B_CUDA(cudaMalloc(&d_varcount_, sizeof(unsigned int) * (size_t) p_));
    S4B_CUDA(cudaEventCreateWithFlags(&ev_a_, cudaEventDisableTiming)); S4B_CUDA(cudaEventCreateWithFlags(&ev_b_, cudaEventDisableTiming));
    S4B_CUDA(cudaEventCreateWithFlags(&ev_c_, cudaEventDisableTiming));
    S4B_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    bart_.set_add_offset(false);
    if (cc_.offset_type < kOffsetDefault || cc_.offset_type > kOffsetParametric) throw std::invalid_argument("sampler: offset_type out of range");
    num_sms_ = device_sm_count();
    const int ew_grid = elementwise_grid(n_, 256, num_sms_);
    if (cc_.user_offset != nullptr) {
      dalloc(&d_user_offset_, (size_t) n_); dalloc(&d_stan_offset_, (size_t) n_);
      S4B_CUDA(cudaMemcpyAsync(d_user_offset_, cc_.user_offset, sizeof(double) * (size_t) n_, cudaMemcpyHostToDevice, stream_));
    }
    cc_.user_offset = nullptr;                                             // the caller's array is not kept
    if (bart_offset_init) S4B_CUDA(cudaMemcpyAsync(d_bart_offset_, bart_offset_init, sizeof(double) * (size_t) n_, cudaMemcpyHostToDevice, stream_));
    if (d_user_offset_ != nullptr && cc_.offset_type != kOffsetBart)       // init.cpp:236-247
      k_sum2<<<ew_grid, 256, 0, stream_>>>(n_, d_user_offset_, (bart_offset_init && cc_.offset_type == kOffsetDefault) ? d_bart_offset_ : nullptr, d_bart_offset_);
    bart_.set_offset_device(d_bart_offset_, true);                         // init.cpp:255
    if (!cc_.is_binary) bart_.set_sigma(cc_.sigma_init);                   // :256-257
    bart_.sample_trees_from_prior();                                       // :261
    bart_.run_sweeps();                                                    // :273  (first draw)
    glmm_.set_inputs_device(stan_offset_source(), cc_.is_binary ? bart_.d_latent_out() : nullptr);    // :275-291 (fit without the offset)
    S4B_CUDA(cudaStreamSynchronize(stream_));
    bart_.check_error_flag();
  }
  ~GibbsSampler()
  {
    if (h_plumb_) cudaFreeHost(h_plumb_);
    if (h_cb_) cudaFreeHost(h_cb_);
    if (h_k_) cudaFreeHost(h_k_);
    cudaFree(d_user_offset_); cudaFree(d_stan_offset_);
    cudaFree(d_bart_offset_); cudaFree(d_mean_train_); cudaFree(d_mean_param_); cudaFree(d_mean_test_); cudaFree(d_varcount_);
    cudaEventDestroy(ev_a_); cudaEventDestroy(ev_b_); cudaEventDestroy(ev_c_);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
  }

  int num_pars() const { return num_pars_; }
  BartFit& bart() { return bart_; }
  GlmmModel& glmm() { return glmm_; }
  NutsSampler& nuts() { return nuts_; }
  void set_batch_group(BatchGroup* g) { batch_group_ = g; }

  void run(int num_iter, bool is_warmup, double* stan, double* train, double* test, uint32_t* varcount, double* sigma)
  {
    if (num_iter < 1) throw std::invalid_argument("num_iter must be >= 1");
    const size_t n = (size_t) n_, nt = (size_t) nt_;
    ms_stan_ = ms_bart_ = 0.0;
    const long long grad0 = glmm_.num_grad_evals(), steps0 = bart_.num_tree_steps();
    const int acc_grid = elementwise_grid(n_, 256, num_sms_);
    const bool k_modeled = bart_.k_modeled();
    if (k_modeled && (size_t) num_iter > h_k_cap_) {
      if (h_k_) cudaFreeHost(h_k_);
      S4B_CUDA(cudaMallocHost(&h_k_, sizeof(double) * (size_t) num_iter)); h_k_cap_ = (size_t) num_iter;
    }
    last_k_.clear();
    bart_.set_keep_trees_active(!is_warmup);                                             // init.cpp:737-744
    // whatever ends the loop (an exception from a kernel, the callback asking to stop), no copy into the caller's buffers may
    // still be in flight when run() returns
    struct DrainCopies {
      GibbsSampler* s;
      ~DrainCopies() { cudaStreamSynchronize(s->copy_stream_); cudaStreamSynchronize(s->stream_); s->copy_pending_ = false; s->bart_.set_keep_trees_active(true); }
    } drain{this};
    for (int iter = 0; iter < num_iter; ++iter) {
      const size_t slot = cc_.keep_fits ? (size_t) iter : 0;
      auto t0 = std::chrono::steady_clock::now();
      // ---- A. Stan block (init.cpp:758-819) ----
      nuts_.run(is_warmup, stan_curr_.data());
      if (stan) std::memcpy(stan + slot * (size_t) num_pars_, stan_curr_.data(), sizeof(double) * (size_t) num_pars_);
      const double* constrained = stan_curr_.data() + 7;
      const double* beta = constrained + glmm_.num_params() + (glmm_.has_aux() ? 1 : 0);
      const double* b = beta + glmm_.K();
      if (d_user_offset_ == nullptr) glmm_.parametric_mean_device(beta, b, d_bart_offset_, true, true);               // :764
      else {                                                                             // :766-794
        const int ot = cc_.offset_type;
        if (ot == kOffsetParametric) S4B_CUDA(cudaMemcpyAsync(d_bart_offset_, d_user_offset_, sizeof(double) * n, cudaMemcpyDeviceToDevice, stream_));
        else {
          glmm_.parametric_mean_device(beta, b, d_bart_offset_, ot != kOffsetFixef, ot != kOffsetRanef);
          if (ot != kOffsetBart) k_sum2<<<acc_grid, 256, 0, stream_>>>(n_, d_bart_offset_, d_user_offset_, d_bart_offset_);
        }
      }
      if (host_plumbing_) {   // the reference's host vector `bartOffset` (init.cpp:143): D2H, then H2D into BART
        S4B_CUDA(cudaMemcpyAsync(h_plumb_, d_bart_offset_, sizeof(double) * n, cudaMemcpyDeviceToHost, stream_));
        S4B_CUDA(cudaMemcpyAsync(d_bart_offset_, h_plumb_, sizeof(double) * n, cudaMemcpyHostToDevice, stream_));
      }
      double aux = 1.0;
      if (!cc_.is_binary) { aux = constrained[glmm_.num_params()]; bart_.set_sigma(aux); }  // :796-800
      const int update_scale_mod = 1 << (8 * iter / num_iter);                           // :816
      bart_.set_offset_device(d_bart_offset_, is_warmup && (iter % update_scale_mod == 0));
      auto t1 = std::chrono::steady_clock::now();
      // ---- B. BART block (init.cpp:821-916) ----
      // the previous iteration's result copies (second stream) must be out of the fit buffers before they are rewritten
      if (copy_pending_) { S4B_CUDA(cudaStreamWaitEvent(stream_, ev_c_, 0)); copy_pending_ = false; }
      if (batch_group_ != nullptr) batch_group_->run(&bart_, stream_);                   // several chains, one launch
      else bart_.run_sweeps();                                                           // :824
      if (host_plumbing_) {   // `stanOffset` and `bartLatents` as host vectors (init.cpp:144-145, :835, :845-846)
        // the latents leave on the copy stream while the fit makes its round trip on the main stream: PCIe is full duplex
        if (cc_.is_binary) {
          S4B_CUDA(cudaEventRecord(ev_a_, stream_));
          S4B_CUDA(cudaStreamWaitEvent(copy_stream_, ev_a_, 0));
          S4B_CUDA(cudaMemcpyAsync(h_plumb_ + 2 * n, bart_.d_latent_out(), sizeof(double) * n, cudaMemcpyDeviceToHost, copy_stream_));
          S4B_CUDA(cudaEventRecord(ev_b_, copy_stream_));
        }
        // the reference computes `stanOffset` on the host FROM the results buffer dbarts filled (init.cpp:828-835): when the caller
        // collects the training fits, that buffer is the host vector of the round trip -- one device-to-host copy, not two
        double* h_train = train != nullptr ? train + slot * n : h_plumb_ + n;
        S4B_CUDA(cudaMemcpyAsync(h_train, bart_.d_train_out(), sizeof(double) * n, cudaMemcpyDeviceToHost, stream_));
        S4B_CUDA(cudaMemcpyAsync(bart_.d_train_out(), h_train, sizeof(double) * n, cudaMemcpyHostToDevice, stream_));
        if (cc_.is_binary) {
          S4B_CUDA(cudaStreamWaitEvent(stream_, ev_b_, 0));
          S4B_CUDA(cudaMemcpyAsync(bart_.d_latent_out(), h_plumb_ + 2 * n, sizeof(double) * n, cudaMemcpyHostToDevice, stream_));
        }
      }
      glmm_.set_inputs_device(stan_offset_source(), cc_.is_binary ? bart_.d_latent_out() : nullptr);   // :828-847
      if (!is_warmup) {
        k_accumulate3<<<acc_grid, 256, 0, stream_>>>(n_, bart_.d_train_out(), d_mean_train_, d_bart_offset_, d_mean_param_,
                                                     nt_, nt_ > 0 ? bart_.d_test_out() : nullptr, d_mean_test_);
        ++num_mean_draws_;
      }
      const bool train_copied = host_plumbing_ && train != nullptr;      // (already in the caller's buffer: the round trip above)
      if ((train && !train_copied) || (test && nt_ > 0)) {
        // the N-length results leave on the copy stream and overlap the next iteration's Stan block
        S4B_CUDA(cudaEventRecord(ev_a_, stream_));
        S4B_CUDA(cudaStreamWaitEvent(copy_stream_, ev_a_, 0));
        if (train && !train_copied) S4B_CUDA(cudaMemcpyAsync(train + slot * n, bart_.d_train_out(), sizeof(double) * n, cudaMemcpyDeviceToHost, copy_stream_));
        if (test && nt_ > 0) S4B_CUDA(cudaMemcpyAsync(test + slot * nt, bart_.d_test_out(), sizeof(double) * nt, cudaMemcpyDeviceToHost, copy_stream_));
        S4B_CUDA(cudaEventRecord(ev_c_, copy_stream_));
        copy_pending_ = true;
      }
      if (varcount) {
        bart_.varcount_device(d_varcount_);
        S4B_CUDA(cudaMemcpyAsync(varcount + slot * (size_t) p_, d_varcount_, sizeof(unsigned int) * (size_t) p_, cudaMemcpyDeviceToHost, stream_));
      }
      if (sigma) sigma[slot] = aux;
      if (k_modeled) S4B_CUDA(cudaMemcpyAsync(h_k_ + iter, bart_.d_k_ptr(), sizeof(double), cudaMemcpyDeviceToHost, stream_));
      S4B_CUDA(cudaStreamSynchronize(stream_));
      if (callback_ != nullptr) {                                                        // init.cpp:849-911
        if (!h_cb_) S4B_CUDA(cudaMallocHost(&h_cb_, sizeof(double) * (n + std::max<size_t>(nt, 1))));
        S4B_CUDA(cudaMemcpyAsync(h_cb_, bart_.d_train_out(), sizeof(double) * n, cudaMemcpyDeviceToHost, stream_));
        if (nt_ > 0) S4B_CUDA(cudaMemcpyAsync(h_cb_ + n, bart_.d_test_out(), sizeof(double) * nt, cudaMemcpyDeviceToHost, stream_));
        S4B_CUDA(cudaStreamSynchronize(stream_));
        if (callback_(callback_user_, iter, stan_curr_.data(), h_cb_, nt_ > 0 ? h_cb_ + n : nullptr) != 0)
          throw std::runtime_error("the iteration callback asked to stop");
      }
      auto t2 = std::chrono::steady_clock::now();
      ms_stan_ += std::chrono::duration<double, std::milli>(t1 - t0).count();
      ms_bart_ += std::chrono::duration<double, std::milli>(t2 - t1).count();
    }
    S4B_CUDA(cudaStreamSynchronize(copy_stream_));        // results are in the caller's buffers when run() returns
    copy_pending_ = false;
    if (k_modeled) last_k_.assign(h_k_, h_k_ + num_iter); else last_k_.assign((size_t) num_iter, bart_.config_k());
    bart_.check_error_flag();
    last_grad_evals_ = glmm_.num_grad_evals() - grad0;
    last_tree_steps_ = bart_.num_tree_steps() - steps0;
  }

  // what Stan sees as its offset (init.cpp:277-285, :831-839): the tree-only fit, plus / replaced by the user offset
  const double* stan_offset_source()
  {
    if (d_user_offset_ == nullptr) return bart_.d_train_out();
    if (cc_.offset_type == kOffsetBart) return d_user_offset_;
    if (cc_.offset_type != kOffsetDefault) return bart_.d_train_out();
    const int grid = elementwise_grid(n_, 256, num_sms_);
    k_sum2<<<grid, 256, 0, stream_>>>(n_, bart_.d_train_out(), d_user_offset_, d_stan_offset_);
    return d_stan_offset_;
  }
  void set_callback(s4b_iteration_callback fn, void* user) { callback_ = fn; callback_user_ = user; }
  const std::vector<double>& last_k() const { return last_k_; }
  void disengage_adaptation() { nuts_.disengage_adaptation(); }
  // route the N-length vectors of every iteration through pinned host memory, as a drop-in at the reference's
  // own host boundary would (bench.py's `e2e` leg); returns bytes moved per iteration in each direction
  void set_host_plumbing(bool on, long long* h2d_bytes, long long* d2h_bytes)
  {
    host_plumbing_ = on;
    if (on && !h_plumb_) S4B_CUDA(cudaMallocHost(&h_plumb_, sizeof(double) * 3 * (size_t) n_));
    long long per = (long long) sizeof(double) * n_ * (cc_.is_binary ? 3 : 2);
    if (h2d_bytes) *h2d_bytes = on ? per : 0;
    if (d2h_bytes) *d2h_bytes = on ? per : 0;
  }
  void parametric_mean(double* out) { glmm_.parametric_mean_host(stan_curr_.data() + 7, out, true, true); }
  void means(double* mean_train, double* mean_test, double* mean_param, long long* num_draws)
  {
    S4B_CUDA(cudaStreamSynchronize(stream_));
    const double inv = num_mean_draws_ > 0 ? 1.0 / (double) num_mean_draws_ : 0.0;
    auto fetch = [&](double* dst, const double* src, size_t count) {
      if (!dst || !count) return;
      S4B_CUDA(cudaMemcpy(dst, src, sizeof(double) * count, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < count; ++i) dst[i] *= inv;
    };
    fetch(mean_train, d_mean_train_, (size_t) n_); fetch(mean_test, d_mean_test_, (size_t) nt_); fetch(mean_param, d_mean_param_, (size_t) n_);
    if (num_draws) *num_draws = num_mean_draws_;
  }
  void last_run_stats(double* ms_stan, double* ms_bart, long long* n_grad, long long* n_steps) const
  {
    if (ms_stan) *ms_stan = ms_stan_; if (ms_bart) *ms_bart = ms_bart_;
    if (n_grad) *n_grad = last_grad_evals_; if (n_steps) *n_steps = last_tree_steps_;
  }

 private:
  s4b_common_control cc_;
  cudaStream_t stream_;
  GlmmModel glmm_;
  NutsSampler nuts_;
  BartFit bart_;
  long long n_ = 0, nt_ = 0; int p_ = 0, num_pars_ = 0, num_sms_ = 1;
  std::vector<double> stan_curr_, last_k_;
  double* h_k_ = nullptr; size_t h_k_cap_ = 0;
  double *d_user_offset_ = nullptr, *d_stan_offset_ = nullptr;
  s4b_iteration_callback callback_ = nullptr; void* callback_user_ = nullptr; double* h_cb_ = nullptr;
  double *d_bart_offset_ = nullptr, *d_mean_train_ = nullptr, *d_mean_param_ = nullptr, *d_mean_test_ = nullptr;
  unsigned int* d_varcount_ = nullptr;
  cudaEvent_t ev_a_ = nullptr, ev_b_ = nullptr, ev_c_ = nullptr;
  long long num_mean_draws_ = 0, last_grad_evals_ = 0, last_tree_steps_ = 0;
  double ms_stan_ = 0.0, ms_bart_ = 0.0;
  bool host_plumbing_ = false, copy_pending_ = false;
  BatchGroup* batch_group_ = nullptr;
  cudaStream_t copy_stream_ = nullptr;
  double* h_plumb_ = nullptr;
};

}  // namespace s4b

// =========================================================================================
// C ABI
// =========================================================================================
using namespace s4b;

struct gpubart_fit { std::unique_ptr<BartFit> owned; BartFit* fit; };
struct glmm_model { std::unique_ptr<GlmmModel> owned; GlmmModel* m; };
struct glmm_nuts { std::unique_ptr<NutsSampler> s; };
struct s4b_batch_group { std::unique_ptr<BatchGroup> g; };
struct s4b_shard { std::unique_ptr<ShardContext> ctx; };
struct gpubart_stored { std::unique_ptr<StoredBart> st; };
struct s4b_sampler { std::unique_ptr<GibbsSampler> s; gpubart_fit bart_view; glmm_model glmm_view; };

static thread_local std::string g_last_error;
static thread_local cudaStream_t g_stream = nullptr;
static thread_local bool g_stream_set = false;

static cudaStream_t current_stream()
{
  if (!g_stream_set) { S4B_CUDA(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking)); g_stream_set = true; }
  return g_stream;
}

#define S4B_API_BEGIN try {
#define S4B_API_END \
  return 0; } catch (const std::exception& e) { g_last_error = e.what(); return 1; } catch (...) { g_last_error = "unknown error"; return 1; }
#define S4B_REQUIRE(x) if (!(x)) throw std::invalid_argument("null or invalid argument: " #x)

extern "C" {

const char* s4b_last_error(void) { return g_last_error.c_str(); }

int s4b_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int s4b_set_device(int device)
{
  S4B_API_BEGIN
  S4B_CUDA(cudaSetDevice(device));
  S4B_API_END
}

int s4b_set_stream(void* cuda_stream)
{
  S4B_API_BEGIN
  if (cuda_stream == nullptr) { g_stream_set = false; g_stream = nullptr; }
  else { g_stream = (cudaStream_t) cuda_stream; g_stream_set = true; }
  S4B_API_END
}

// ---- gpubart ----
int gpubart_create(const s4b_bart_config* cfg, const double* y, const double* x, const double* x_test, gpubart_fit** out)
{
  S4B_API_BEGIN
  S4B_REQUIRE(cfg && y && x && out);
  if (cfg->n_test > 0) S4B_REQUIRE(x_test);
  if (s4b_device_count() < 1) throw std::runtime_error("no CUDA device: stan4bart_b200 has no CPU fallback");
  auto* h = new gpubart_fit;
  h->owned.reset(new BartFit(*cfg, y, x, x_test, current_stream()));
  h->fit = h->owned.get();
  *out = h;
  S4B_API_END
}
int gpubart_free(gpubart_fit* f) { S4B_API_BEGIN if (f && f->owned) delete f; S4B_API_END }
int gpubart_set_offset(gpubart_fit* f, const double* offset, int update_scale) { S4B_API_BEGIN S4B_REQUIRE(f); f->fit->set_offset_host(offset, update_scale != 0); S4B_API_END }
int gpubart_get_k(gpubart_fit* f, double* k) { S4B_API_BEGIN S4B_REQUIRE(f && k); *k = f->fit->current_k(); S4B_API_END }
int gpubart_set_sigma(gpubart_fit* f, double sigma) { S4B_API_BEGIN S4B_REQUIRE(f && sigma > 0); f->fit->set_sigma(sigma); S4B_API_END }
int gpubart_sample_trees_from_prior(gpubart_fit* f) { S4B_API_BEGIN S4B_REQUIRE(f); f->fit->sample_trees_from_prior(); S4B_API_END }
int gpubart_run_sampler_with_results(gpubart_fit* f, double* train, double* test, uint32_t* varcount, double* sigma)
{ S4B_API_BEGIN S4B_REQUIRE(f); f->fit->run(train, test, varcount, sigma); S4B_API_END }
int gpubart_run_batched(gpubart_fit* const* fits, int count)
{
  S4B_API_BEGIN
  S4B_REQUIRE(fits && count >= 1);
  std::vector<BartFit*> v((size_t) count);
  for (int c = 0; c < count; ++c) { S4B_REQUIRE(fits[c]); v[(size_t) c] = fits[c]->fit; }
  BartFit::run_sweeps_batched(v.data(), count);
  S4B_API_END
}
int gpubart_collect_results(gpubart_fit* f, double* train, double* test, uint32_t* varcount, double* sigma)
{ S4B_API_BEGIN S4B_REQUIRE(f); f->fit->collect_results(train, test, varcount, sigma); S4B_API_END }
int gpubart_store_latents(gpubart_fit* f, double* out) { S4B_API_BEGIN S4B_REQUIRE(f && out); f->fit->store_latents(out); S4B_API_END }
int gpubart_get_data_range(gpubart_fit* f, double* o) { S4B_API_BEGIN S4B_REQUIRE(f && o); BartParams P = f->fit->params(); o[0] = P.smin; o[1] = P.smax; o[2] = P.srange; S4B_API_END }
int gpubart_predict(gpubart_fit* f, const double* x_test, int64_t n, const double* test_offset, double* out)
{ S4B_API_BEGIN S4B_REQUIRE(f && x_test && out && n >= 0); f->fit->predict(x_test, n, test_offset, out); S4B_API_END }
int gpubart_num_nodes(gpubart_fit* f, int64_t* out) { S4B_API_BEGIN S4B_REQUIRE(f && out); *out = f->fit->num_nodes(); S4B_API_END }
int gpubart_get_trees(gpubart_fit* f, int32_t* tree_no, int64_t* n_obs, int32_t* var, double* value)
{ S4B_API_BEGIN S4B_REQUIRE(f && tree_no && n_obs && var && value); f->fit->get_trees(tree_no, (long long*) n_obs, var, value); S4B_API_END }
int gpubart_node_assignment(gpubart_fit* f, int tree, int64_t* heap) { S4B_API_BEGIN S4B_REQUIRE(f && heap); f->fit->node_assignment(tree, (long long*) heap); S4B_API_END }
int gpubart_leaf_stats(gpubart_fit* f, int tree, int max_leaves, int64_t* heap, int64_t* count, double* sum, double* sumsq, int* num_leaves)
{ S4B_API_BEGIN S4B_REQUIRE(f && heap && count && sum && sumsq && num_leaves); *num_leaves = f->fit->leaf_stats(tree, max_leaves, (long long*) heap, (long long*) count, sum, sumsq); S4B_API_END }
int gpubart_get_residual(gpubart_fit* f, double* out) { S4B_API_BEGIN S4B_REQUIRE(f && out); f->fit->get_residual(out); S4B_API_END }
int gpubart_set_trace(gpubart_fit* f, size_t cap) { S4B_API_BEGIN S4B_REQUIRE(f); f->fit->set_trace(cap); S4B_API_END }
int gpubart_get_trace(gpubart_fit* f, double* out, size_t cap, size_t* num) { S4B_API_BEGIN S4B_REQUIRE(f && num); *num = f->fit->get_trace(out, cap); S4B_API_END }
int gpubart_set_tape(gpubart_fit* f, const double* tape, size_t len) { S4B_API_BEGIN S4B_REQUIRE(f); f->fit->set_tape(tape, len); S4B_API_END }
int gpubart_set_record(gpubart_fit* f, size_t cap) { S4B_API_BEGIN S4B_REQUIRE(f); f->fit->set_record(cap); S4B_API_END }
int gpubart_get_record(gpubart_fit* f, double* out, size_t cap, size_t* len) { S4B_API_BEGIN S4B_REQUIRE(f && len); *len = f->fit->get_record(out, cap); S4B_API_END }
int gpubart_rng_counter(gpubart_fit* f, uint64_t* out) { S4B_API_BEGIN S4B_REQUIRE(f && out); *out = f->fit->rng_counter(); S4B_API_END }
int gpubart_set_use_graph(gpubart_fit* f, int g) { S4B_API_BEGIN S4B_REQUIRE(f); f->fit->set_use_graph(g != 0); S4B_API_END }
int gpubart_set_sweep_mode(gpubart_fit* f, int mode) { S4B_API_BEGIN S4B_REQUIRE(f); f->fit->set_sweep_mode(mode); S4B_API_END }
int gpubart_get_sweep_mode(gpubart_fit* f, int* mode) { S4B_API_BEGIN S4B_REQUIRE(f && mode); *mode = f->fit->sweep_mode(); S4B_API_END }
int gpubart_num_tree_steps(gpubart_fit* f, int64_t* out) { S4B_API_BEGIN S4B_REQUIRE(f && out); *out = f->fit->num_tree_steps(); S4B_API_END }
int gpubart_time_leaf_stats(gpubart_fit* f, int tree, int reps, double* ms)
{
  S4B_API_BEGIN
  S4B_REQUIRE(f && ms && reps > 0);
  cudaStream_t st = f->fit->stream();
  cudaEvent_t a, b; S4B_CUDA(cudaEventCreate(&a)); S4B_CUDA(cudaEventCreate(&b));
  const int leaves = f->fit->tree_num_leaves(tree);       // picks the kernel; asked once, outside the timed launches
  f->fit->launch_leaf_stats(tree, leaves);
  S4B_CUDA(cudaStreamSynchronize(st));
  S4B_CUDA(cudaEventRecord(a, st));
  for (int r = 0; r < reps; ++r) f->fit->launch_leaf_stats(tree, leaves);
  S4B_CUDA(cudaEventRecord(b, st));
  S4B_CUDA(cudaEventSynchronize(b));
  float t = 0.f; S4B_CUDA(cudaEventElapsedTime(&t, a, b));
  *ms = (double) t / reps;
  cudaEventDestroy(a); cudaEventDestroy(b);
  S4B_API_END
}

// ---- glmm ----
int glmm_create(const s4b_glmm_data* d, glmm_model** out)
{
  S4B_API_BEGIN
  S4B_REQUIRE(d && out);
  if (s4b_device_count() < 1) throw std::runtime_error("no CUDA device: stan4bart_b200 has no CPU fallback");
  auto* h = new glmm_model;
  h->owned.reset(new GlmmModel(*d, current_stream()));
  h->m = h->owned.get();
  *out = h;
  S4B_API_END
}
int glmm_free(glmm_model* m) { S4B_API_BEGIN if (m && m->owned) delete m; S4B_API_END }
int glmm_num_params(glmm_model* m, int* d, int* nc) { S4B_API_BEGIN S4B_REQUIRE(m); if (d) *d = m->m->num_params(); if (nc) *nc = m->m->num_constrained(); S4B_API_END }
int glmm_set_offset(glmm_model* m, const double* o) { S4B_API_BEGIN S4B_REQUIRE(m && o); m->m->set_offset_host(o); S4B_API_END }
int glmm_set_response(glmm_model* m, const double* y) { S4B_API_BEGIN S4B_REQUIRE(m && y); m->m->set_response_host(y); S4B_API_END }
int glmm_log_prob_grad(glmm_model* m, const double* q, double* lp, double* grad, int* status)
{ S4B_API_BEGIN S4B_REQUIRE(m && q && lp && grad && status); *status = m->m->log_prob_grad(q, lp, grad); S4B_API_END }
int glmm_write_array(glmm_model* m, const double* q, double* out) { S4B_API_BEGIN S4B_REQUIRE(m && q && out); m->m->write_array(q, out); S4B_API_END }
int glmm_parametric_mean(glmm_model* m, const double* c, double* out, int f, int r) { S4B_API_BEGIN S4B_REQUIRE(m && c && out); m->m->parametric_mean_host(c, out, f != 0, r != 0); S4B_API_END }
int glmm_data_terms(glmm_model* m, const double* beta, const double* b, double* S, double* gbeta, double* gb)
{ S4B_API_BEGIN S4B_REQUIRE(m && S && gbeta && gb); m->m->data_terms(beta, b, S, gbeta, gb); S4B_API_END }
// names of the stored Stan rows, '\n' separated: the 7 sampler diagnostics (mcmc/sample.hpp:43-44, base_nuts get_sampler_param_names)
// followed by the constrained parameter names -- what the reference puts into the dimnames of its `stan` result
// (src/stan_sampler.cpp:478-489, :577-596) and R splits with the regexes of R/stan4bart.R:241-246
int glmm_stan_row_names(glmm_model* m, char* out, size_t cap, size_t* needed)
{
  S4B_API_BEGIN
  S4B_REQUIRE(m && needed);
  std::string all = "lp__\naccept_stat__\nstepsize__\ntreedepth__\nn_leapfrog__\ndivergent__\nenergy__";
  for (const std::string& nm : m->m->param_names()) { all += "\n"; all += nm; }
  *needed = all.size() + 1;
  if (out != nullptr && cap > 0) { const size_t k = std::min(cap - 1, all.size()); std::memcpy(out, all.data(), k); out[k] = 0; }
  S4B_API_END
}
int glmm_set_mode(glmm_model* m, int mode) { S4B_API_BEGIN S4B_REQUIRE(m); m->m->set_mode(mode); S4B_API_END }
int glmm_get_mode(glmm_model* m, int* mode) { S4B_API_BEGIN S4B_REQUIRE(m && mode); *mode = m->m->mode(); S4B_API_END }
int glmm_num_device_passes(glmm_model* m, int64_t* out) { S4B_API_BEGIN S4B_REQUIRE(m && out); *out = m->m->num_device_passes(); S4B_API_END }
int glmm_time_data_pass(glmm_model* m, int reps, int flush_l2, double* ms, int* bulk)
{
  S4B_API_BEGIN
  S4B_REQUIRE(m && ms && reps > 0);
  *ms = m->m->time_data_pass(reps, flush_l2);
  if (bulk) *bulk = m->m->bulk_pass() ? 1 : 0;
  S4B_API_END
}
int glmm_num_grad_evals(glmm_model* m, int64_t* out) { S4B_API_BEGIN S4B_REQUIRE(m && out); *out = m->m->num_grad_evals(); S4B_API_END }

// ---- the Stan half alone: StanSampler over a model (src/stan_sampler.cpp:382-476, run(isWarmup) :478-489) ----
int glmm_nuts_create(glmm_model* m, const s4b_stan_control* ctl, int chain_id, int num_warmup, glmm_nuts** out)
{
  S4B_API_BEGIN
  S4B_REQUIRE(m && ctl && out && num_warmup >= 0);
  auto* h = new glmm_nuts;
  h->s.reset(new NutsSampler(*m->m, *ctl, chain_id, num_warmup));
  *out = h;
  S4B_API_END
}
int glmm_nuts_free(glmm_nuts* s) { S4B_API_BEGIN delete s; S4B_API_END }
int glmm_nuts_num_pars(glmm_nuts* s, int* out) { S4B_API_BEGIN S4B_REQUIRE(s && out); *out = s->s->num_pars(); S4B_API_END }
int glmm_nuts_run(glmm_nuts* s, int is_warmup, double* out) { S4B_API_BEGIN S4B_REQUIRE(s); s->s->run(is_warmup != 0, out); S4B_API_END }
int glmm_nuts_disengage_adaptation(glmm_nuts* s) { S4B_API_BEGIN S4B_REQUIRE(s); s->s->disengage_adaptation(); S4B_API_END }
int glmm_nuts_stepsize(glmm_nuts* s, double* out) { S4B_API_BEGIN S4B_REQUIRE(s && out); *out = s->s->stepsize(); S4B_API_END }

// ---- sampler ----
int s4b_sampler_create(const s4b_bart_config* bcfg, const double* y_bart, const double* x_bart, const double* x_test,
                       const s4b_glmm_data* gdata, const s4b_stan_control* sctl, const s4b_common_control* cctl,
                       const double* bart_offset_init, s4b_sampler** out)
{
  S4B_API_BEGIN
  S4B_REQUIRE(bcfg && y_bart && x_bart && gdata && sctl && cctl && out);
  if (bcfg->n_test > 0) S4B_REQUIRE(x_test);
  if (s4b_device_count() < 1) throw std::runtime_error("no CUDA device: stan4bart_b200 has no CPU fallback");
  auto* h = new s4b_sampler;
  h->s.reset(new GibbsSampler(*bcfg, y_bart, x_bart, x_test, *gdata, *sctl, *cctl, bart_offset_init, current_stream()));
  h->bart_view.fit = &h->s->bart();
  h->glmm_view.m = &h->s->glmm();
  *out = h;
  S4B_API_END
}
int s4b_sampler_free(s4b_sampler* s) { S4B_API_BEGIN delete s; S4B_API_END }
int s4b_sampler_num_stan_pars(s4b_sampler* s, int* out) { S4B_API_BEGIN S4B_REQUIRE(s && out); *out = s->s->num_pars(); S4B_API_END }
int s4b_batch_group_create(int count, s4b_batch_group** out)
{
  S4B_API_BEGIN
  S4B_REQUIRE(out && count >= 1);
  auto* h = new s4b_batch_group; h->g.reset(new BatchGroup(count)); *out = h;
  S4B_API_END
}
int s4b_batch_group_free(s4b_batch_group* g) { S4B_API_BEGIN delete g; S4B_API_END }
int s4b_batch_group_launches(s4b_batch_group* g, int64_t* out) { S4B_API_BEGIN S4B_REQUIRE(g && out); *out = g->g->launches(); S4B_API_END }
int s4b_sampler_set_batch_group(s4b_sampler* s, s4b_batch_group* g) { S4B_API_BEGIN S4B_REQUIRE(s); s->s->set_batch_group(g ? g->g.get() : nullptr); S4B_API_END }
int s4b_sampler_run(s4b_sampler* s, int num_iter, int is_warmup, double* stan, double* train, double* test, uint32_t* varcount, double* sigma)
{ S4B_API_BEGIN S4B_REQUIRE(s); s->s->run(num_iter, is_warmup != 0, stan, train, test, varcount, sigma); S4B_API_END }
int s4b_sampler_last_k(s4b_sampler* s, double* out, int capacity, int* count)
{
  S4B_API_BEGIN
  S4B_REQUIRE(s && out && capacity >= 0);
  const std::vector<double>& k = s->s->last_k();
  const int m = std::min<int>(capacity, (int) k.size());
  for (int i = 0; i < m; ++i) out[i] = k[(size_t) i];
  if (count) *count = m;
  S4B_API_END
}
int s4b_sampler_set_callback(s4b_sampler* s, s4b_iteration_callback fn, void* user) { S4B_API_BEGIN S4B_REQUIRE(s); s->s->set_callback(fn, user); S4B_API_END }
int s4b_sampler_disengage_adaptation(s4b_sampler* s) { S4B_API_BEGIN S4B_REQUIRE(s); s->s->disengage_adaptation(); S4B_API_END }
int s4b_sampler_get_bart_data_range(s4b_sampler* s, double* o) { S4B_API_BEGIN S4B_REQUIRE(s && o); BartParams P = s->s->bart().params(); o[0] = P.smin; o[1] = P.smax; S4B_API_END }
int s4b_sampler_get_parametric_mean(s4b_sampler* s, double* out) { S4B_API_BEGIN S4B_REQUIRE(s && out); s->s->parametric_mean(out); S4B_API_END }
int s4b_sampler_predict_bart(s4b_sampler* s, const double* x_test, int64_t n, const double* off, double* out)
{ S4B_API_BEGIN S4B_REQUIRE(s && x_test && out); s->s->bart().predict(x_test, n, off, out); S4B_API_END }
gpubart_fit* s4b_sampler_bart(s4b_sampler* s) { return s ? &s->bart_view : nullptr; }
glmm_model* s4b_sampler_glmm(s4b_sampler* s) { return s ? &s->glmm_view : nullptr; }
int s4b_sampler_get_means(s4b_sampler* s, double* mt, double* mte, double* mp, int64_t* nd)
{ S4B_API_BEGIN S4B_REQUIRE(s); long long k = 0; s->s->means(mt, mte, mp, &k); if (nd) *nd = k; S4B_API_END }
int s4b_sampler_set_host_plumbing(s4b_sampler* s, int on, int64_t* h2d, int64_t* d2h)
{ S4B_API_BEGIN S4B_REQUIRE(s); long long a = 0, b = 0; s->s->set_host_plumbing(on != 0, &a, &b); if (h2d) *h2d = a; if (d2h) *d2h = b; S4B_API_END }
int gpubart_get_profile(gpubart_fit* f, uint64_t* out24, int reset) { S4B_API_BEGIN S4B_REQUIRE(f && out24); f->fit->get_profile((unsigned long long*) out24, reset != 0); S4B_API_END }
int gpubart_set_keep_trees(gpubart_fit* f, int64_t capacity) { S4B_API_BEGIN S4B_REQUIRE(f && capacity >= 0); f->fit->set_keep_trees(capacity); S4B_API_END }
int gpubart_set_pipeline(gpubart_fit* f, int on) { S4B_API_BEGIN S4B_REQUIRE(f); f->fit->set_pipe_enabled(on != 0); S4B_API_END }
int gpubart_get_pipeline(gpubart_fit* f, int* enabled, int64_t* sweeps_launched, int64_t* steps_pipelined)
{
  S4B_API_BEGIN
  S4B_REQUIRE(f);
  if (enabled) *enabled = f->fit->pipe_enabled() ? 1 : 0;
  if (sweeps_launched) *sweeps_launched = f->fit->pipe_sweeps();
  if (steps_pipelined) *steps_pipelined = f->fit->pipe_sweeps_done();
  S4B_API_END
}
int gpubart_pipeline_misfits(gpubart_fit* f, uint32_t* out4) { S4B_API_BEGIN S4B_REQUIRE(f && out4); f->fit->pipe_reasons(out4); S4B_API_END }
int gpubart_set_keep_trees_active(gpubart_fit* f, int on) { S4B_API_BEGIN S4B_REQUIRE(f); f->fit->set_keep_trees_active(on != 0); S4B_API_END }
int gpubart_set_response(gpubart_fit* f, const double* y) { S4B_API_BEGIN S4B_REQUIRE(f && y); f->fit->set_response_host(y); S4B_API_END }
int gpubart_get_stored_scales(gpubart_fit* f, int64_t first, int64_t count, double* out2)
{ S4B_API_BEGIN S4B_REQUIRE(f && out2); f->fit->get_stored_scales(first, count, out2); S4B_API_END }
int gpubart_stored_get_scales(gpubart_stored* st, int64_t first, int64_t count, double* out2)
{ S4B_API_BEGIN S4B_REQUIRE(st && out2); st->st->get_scales(first, count, out2); S4B_API_END }
int gpubart_num_stored(gpubart_fit* f, int64_t* out) { S4B_API_BEGIN S4B_REQUIRE(f && out); *out = f->fit->num_stored(); S4B_API_END }
int gpubart_predict_stored(gpubart_fit* f, const double* x_test, int64_t n, const double* test_offset, int64_t first, int64_t count, double* out)
{ S4B_API_BEGIN S4B_REQUIRE(f && x_test && out && n >= 0); f->fit->predict_stored(x_test, n, test_offset, first, count, out); S4B_API_END }
int gpubart_num_stored_nodes(gpubart_fit* f, int64_t sample, int64_t* out) { S4B_API_BEGIN S4B_REQUIRE(f && out); *out = f->fit->num_stored_nodes(sample); S4B_API_END }
int gpubart_get_stored_trees(gpubart_fit* f, int64_t sample, int32_t* tree_no, int64_t* n_obs, int32_t* var, double* value)
{ S4B_API_BEGIN S4B_REQUIRE(f && tree_no && n_obs && var && value); f->fit->get_stored_trees(sample, tree_no, (long long*) n_obs, var, value); S4B_API_END }
int gpubart_stored_export_size(gpubart_fit* f, int64_t* bytes) { S4B_API_BEGIN S4B_REQUIRE(f && bytes); *bytes = f->fit->stored_export_size(); S4B_API_END }
int gpubart_stored_export(gpubart_fit* f, void* out, int64_t bytes) { S4B_API_BEGIN S4B_REQUIRE(f && out); f->fit->stored_export(out, bytes); S4B_API_END }
int gpubart_stored_import(const void* blob, int64_t bytes, gpubart_stored** out)
{
  S4B_API_BEGIN
  S4B_REQUIRE(blob && out);
  if (s4b_device_count() < 1) throw std::runtime_error("no CUDA device: stan4bart_b200 has no CPU fallback");
  auto* h = new gpubart_stored;
  h->st.reset(new StoredBart(blob, bytes, current_stream()));
  *out = h;
  S4B_API_END
}
int gpubart_stored_free(gpubart_stored* st) { S4B_API_BEGIN delete st; S4B_API_END }
int gpubart_stored_count(gpubart_stored* st, int64_t* out) { S4B_API_BEGIN S4B_REQUIRE(st && out); *out = st->st->count(); S4B_API_END }
int gpubart_stored_predict(gpubart_stored* st, const double* x_test, int64_t n, const double* test_offset, int64_t first, int64_t count, double* out)
{ S4B_API_BEGIN S4B_REQUIRE(st && x_test && out && n >= 0); st->st->predict(x_test, n, test_offset, first, count, out); S4B_API_END }
int gpubart_summary(gpubart_fit* f, char* out, size_t cap, size_t* needed)
{
  S4B_API_BEGIN
  S4B_REQUIRE(f && needed);
  const std::string sm = f->fit->summary();
  *needed = sm.size() + 1;
  if (out != nullptr && cap > 0) { const size_t k = std::min(cap - 1, sm.size()); std::memcpy(out, sm.data(), k); out[k] = 0; }
  S4B_API_END
}
int gpubart_set_profile(gpubart_fit* f, int on) { S4B_API_BEGIN S4B_REQUIRE(f); f->fit->set_profile(on != 0); S4B_API_END }
int gpubart_tree_step_ms(gpubart_fit* f, int reset, double* ms) { S4B_API_BEGIN S4B_REQUIRE(f && ms); *ms = f->fit->tree_step_ms(reset != 0); S4B_API_END }
int s4b_sampler_last_run_stats(s4b_sampler* s, double* ms_stan, double* ms_bart, int64_t* ng, int64_t* ns)
{ S4B_API_BEGIN S4B_REQUIRE(s); long long a = 0, b = 0; s->s->last_run_stats(ms_stan, ms_bart, &a, &b); if (ng) *ng = a; if (ns) *ns = b; S4B_API_END }

// ---- observation-sharded chains ----
int s4b_shard_create(int rank, int world, s4b_shard** out)
{
  S4B_API_BEGIN
  S4B_REQUIRE(out);
  if (s4b_device_count() < 1) throw std::runtime_error("no CUDA device: stan4bart_b200 has no CPU fallback");
  auto* h = new s4b_shard;
  h->ctx.reset(new ShardContext(rank, world));
  *out = h;
  S4B_API_END
}
int s4b_shard_free(s4b_shard* sh) { S4B_API_BEGIN delete sh; S4B_API_END }
int s4b_shard_ipc_handle(s4b_shard* sh, unsigned char* out64) { S4B_API_BEGIN S4B_REQUIRE(sh && out64); sh->ctx->ipc_handle(out64); S4B_API_END }
int s4b_shard_attach(s4b_shard* sh, const unsigned char* handles) { S4B_API_BEGIN S4B_REQUIRE(sh && handles); sh->ctx->attach(handles); S4B_API_END }
int s4b_shard_set_obs_range(s4b_shard* sh, int64_t first_obs, int64_t total_obs)
{ S4B_API_BEGIN S4B_REQUIRE(sh && first_obs >= 0 && total_obs >= first_obs); sh->ctx->set_obs_range(first_obs, total_obs); S4B_API_END }
int s4b_shard_allreduce(s4b_shard* sh, double* vec, int64_t n, int op)
{ S4B_API_BEGIN S4B_REQUIRE(sh && vec && n >= 0 && (op == 0 || op == 1)); sh->ctx->allreduce_host(vec, n, op == 0 ? kOpSum : kOpMax, current_stream()); S4B_API_END }
int s4b_shard_nccl_unique_id(unsigned char* out128) { S4B_API_BEGIN S4B_REQUIRE(out128); ShardContext::nccl_unique_id(out128); S4B_API_END }
int s4b_shard_nccl_init(s4b_shard* sh, const unsigned char* id128) { S4B_API_BEGIN S4B_REQUIRE(sh && id128); sh->ctx->nccl_init(id128); S4B_API_END }
int s4b_shard_use_nccl(s4b_shard* sh, int on) { S4B_API_BEGIN S4B_REQUIRE(sh); sh->ctx->use_nccl(on != 0); S4B_API_END }
int gpubart_create_sharded(const s4b_bart_config* cfg, const double* y, const double* x, const double* x_test, s4b_shard* sh, gpubart_fit** out)
{
  S4B_API_BEGIN
  S4B_REQUIRE(cfg && y && x && sh && out);
  if (cfg->n_test > 0) S4B_REQUIRE(x_test);
  auto* h = new gpubart_fit;
  h->owned.reset(new BartFit(*cfg, y, x, x_test, current_stream(), sh->ctx.get()));
  h->fit = h->owned.get();
  *out = h;
  S4B_API_END
}
int glmm_create_sharded(const s4b_glmm_data* d, s4b_shard* sh, glmm_model** out)
{
  S4B_API_BEGIN
  S4B_REQUIRE(d && sh && out);
  auto* h = new glmm_model;
  h->owned.reset(new GlmmModel(*d, current_stream(), sh->ctx.get()));
  h->m = h->owned.get();
  *out = h;
  S4B_API_END
}
int s4b_sampler_create_sharded(const s4b_bart_config* bcfg, const double* y_bart, const double* x_bart, const double* x_test,
                               const s4b_glmm_data* gdata, const s4b_stan_control* sctl, const s4b_common_control* cctl,
                               const double* bart_offset_init, s4b_shard* sh, s4b_sampler** out)
{
  S4B_API_BEGIN
  S4B_REQUIRE(bcfg && y_bart && x_bart && gdata && sctl && cctl && sh && out);
  if (bcfg->n_test > 0) S4B_REQUIRE(x_test);
  auto* h = new s4b_sampler;
  h->s.reset(new GibbsSampler(*bcfg, y_bart, x_bart, x_test, *gdata, *sctl, *cctl, bart_offset_init, current_stream(), sh->ctx.get()));
  h->bart_view.fit = &h->s->bart();
  h->glmm_view.m = &h->s->glmm();
  *out = h;
  S4B_API_END
}

}  // extern "C"

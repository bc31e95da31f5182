// stan4bart_b200/csrc/s4b_common.cuh
// Shared definitions for the sm_100a BART / GLMM kernels: device tree layout, the
// per-tree step descriptor exchanged between the controller and the N-length pass,
// and the s4b-rng v1 generator (Philox4x32-10 + AS241 inversion), implemented here
// independently of the CPU oracle (oracle/s4b_rng.h) from the same written spec.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#define S4B_MAX_LEAVES 64
#define S4B_MAX_NODES 127
#define S4B_NODE_CAP 128
#define S4B_MAX_DEPTH 30
#define S4B_TRACE_LEN 32
#define S4B_MAX_SLOTS 128
#define S4B_SLOT_CHUNK 8

#define S4B_STREAM_BART 0u
#define S4B_STREAM_STAN 1u
#define S4B_STREAM_LATENT 2u

namespace s4b {

struct CudaError : std::runtime_error {
  explicit CudaError(const std::string& m) : std::runtime_error(m) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line)
{
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %s (%d) at %s:%d: %s", cudaGetErrorName(e), (int) e, file, line, what);
    throw CudaError(buf);
  }
}
#define S4B_CUDA(x) ::s4b::cuda_check((x), #x, __FILE__, __LINE__)

// Zero a fresh allocation ON THE OBJECT'S STREAM and wait for it.  A plain cudaMemset runs asynchronously on the legacy
// default stream, which the library's non-blocking streams do not synchronise with: the memset could land after a
// stream-ordered copy into the same buffer (seen as a lost offset vector when two ranks time-share one GPU).
inline void zero_device_sync(void* p, size_t bytes, cudaStream_t stream)
{
  S4B_CUDA(cudaMemsetAsync(p, 0, bytes, stream));
  S4B_CUDA(cudaStreamSynchronize(stream));
}

// SM count of the current device (grids of the element-wise kernels are sized in multiples of it)
inline int device_sm_count()
{
  int dev = 0, sms = 0;
  S4B_CUDA(cudaGetDevice(&dev));
  S4B_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return sms > 0 ? sms : 1;
}
// grid of an element-wise kernel over `items` with `block` threads: at most 8 resident blocks per SM
inline int elementwise_grid(long long items, int block, int num_sms)
{
  const long long want = (items + block - 1) / block;
  const long long cap = (long long) num_sms * 8;
  return (int) (want < 1 ? 1 : (want > cap ? cap : want));
}

// ---------------------------------------------------------------------------------------
// RNG (spec "s4b-rng v1")
// ---------------------------------------------------------------------------------------
__host__ __device__ inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t& o0, uint32_t& o1)
{
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = (uint64_t) 0xD2511F53u * c0, p1 = (uint64_t) 0xCD9E8D57u * c2;
    uint32_t hi0 = (uint32_t) (p0 >> 32), lo0 = (uint32_t) p0, hi1 = (uint32_t) (p1 >> 32), lo1 = (uint32_t) p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o0 = c0; o1 = c1;
}

__host__ __device__ inline double bits_to_uniform(uint32_t o0, uint32_t o1)
{
  uint64_t k = ((uint64_t) (o0 >> 6) << 26) | (uint64_t) (o1 >> 6);
  return ((double) k + 0.5) * (1.0 / 4503599627370496.0);
}

// Wichura AS241 (PPND16)
__host__ __device__ inline double qnorm_as241(double p)
{
  double q = p - 0.5, r, val;
  if (fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    val = q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r + 45921.953931549871457) * r +
                  13731.693765509461125) * r + 1971.5909503065514427) * r + 133.14166789178437745) * r + 3.387132872796366608) /
          (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r + 21213.794301586595867) * r +
              5394.1960214247511077) * r + 687.1870074920579083) * r + 42.313330701600911252) * r + 1.0);
    return val;
  }
  r = q < 0.0 ? p : 1.0 - p;
  r = sqrt(-log(r));
  if (r <= 5.0) {
    r -= 1.6;
    val = (((((((r * 7.7454501427834140764e-4 + 0.0227238449892691845833) * r + 0.24178072517745061177) * r + 1.27045825245236838258) * r +
             3.64784832476320460504) * r + 5.7694972214606914055) * r + 4.6303378461565452959) * r + 1.42343711074968357734) /
          (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + 0.0151986665636164571966) * r + 0.14810397642748007459) * r +
              0.68976733498510000455) * r + 1.6763848301838038494) * r + 2.05319162663775882187) * r + 1.0);
  } else {
    r -= 5.0;
    val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + 0.0012426609473880784386) * r + 0.026532189526576123093) * r +
             0.29656057182850489123) * r + 1.7848265399172913358) * r + 5.4637849111641143699) * r + 6.6579046435011037772) /
          (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r + 7.868691311456132591e-4) * r +
              0.0148753612908506148525) * r + 0.13692988092273580531) * r + 0.59983220655588793769) * r + 1.0);
  }
  return q < 0.0 ? -val : val;
}

// sequential stream state (lives in device memory for the BART stream, host memory for NUTS)
// Counter layout (spec "s4b-rng v1"): c0 = draw index inside the substream, c1 = low 32 bits of the step,
// c2 = (sub << 30) | high 30 bits of the step, c3 = stream.  BART stream: sub 0 = proposal draws of a tree
// step, sub 1 = its decision draws (accept uniform, leaf normals), sub 2 = sampleTreesFromPrior.
struct RngState {
  uint32_t key0, key1;
  uint32_t stream;
  uint32_t tape_underrun;
  uint32_t sub, idx;
  unsigned long long step;
  unsigned long long counter;   // total draws consumed
  const double* tape;           // replay: interleaved uniforms / normals in consumption order
  unsigned long long tape_len, tape_pos;
  double* rec;                  // optional recording of every draw
  unsigned long long rec_cap, rec_len;
};

__host__ __device__ inline void rng_enter(RngState& g, unsigned long long step, uint32_t sub)
{
  if (g.step != step || g.sub != sub) { g.step = step; g.sub = sub; g.idx = 0; }
}
__host__ __device__ inline double keyed_stream_uniform(uint32_t k0, uint32_t k1, uint32_t stream, unsigned long long step, uint32_t sub, uint32_t idx)
{
  uint32_t o0, o1;
  philox4x32_10(idx, (uint32_t) step, (sub << 30) | (uint32_t) ((step >> 32) & 0x3FFFFFFFu), stream, k0, k1, o0, o1);
  return bits_to_uniform(o0, o1);
}
__host__ __device__ inline double rng_raw_uniform(RngState& g)
{
  double u = keyed_stream_uniform(g.key0, g.key1, g.stream, g.step, g.sub, g.idx);
  g.idx++;
  if (g.idx == 0) g.step++;
  return u;
}
__host__ __device__ inline double rng_note(RngState& g, double v)
{
  g.counter++;
  if (g.rec != nullptr) { if (g.rec_len < g.rec_cap) g.rec[g.rec_len] = v; g.rec_len++; }
  return v;
}
__host__ __device__ inline double rng_uniform(RngState& g)
{
  if (g.tape != nullptr) {
    if (g.tape_pos >= g.tape_len) { g.tape_underrun = 1; return 0.5; }
    return rng_note(g, g.tape[g.tape_pos++]);
  }
  return rng_note(g, rng_raw_uniform(g));
}
__host__ __device__ inline double rng_normal(RngState& g)
{
  if (g.tape != nullptr) {
    if (g.tape_pos >= g.tape_len) { g.tape_underrun = 1; return 0.0; }
    return rng_note(g, g.tape[g.tape_pos++]);
  }
  return rng_note(g, qnorm_as241(rng_raw_uniform(g)));
}
// Gamma(shape, 1) by Marsaglia & Tsang (2000) on the stream's normal and uniform draws (same recipe as oracle/s4b_rng.h)
__host__ __device__ inline double rng_gamma(RngState& g, double shape)
{
  double boost = 1.0;
  if (shape < 1.0) { boost = pow(rng_uniform(g), 1.0 / shape); shape += 1.0; }
  const double d = shape - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (;;) {
    const double x = rng_normal(g);
    double v = 1.0 + c * x;
    if (v <= 0.0) continue;
    v = v * v * v;
    const double u = rng_uniform(g);
    if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) return boost * d * v;
    if (g.tape != nullptr && g.tape_underrun) return boost * d;
  }
}
__host__ __device__ inline int rng_index(RngState& g, int n)
{
  int k = (int) (rng_uniform(g) * (double) n);
  return k >= n ? n - 1 : k;
}

__host__ __device__ inline double keyed_uniform(uint32_t k0, uint32_t k1, uint32_t obs, uint32_t epoch, uint32_t sub)
{
  uint32_t o0, o1;
  philox4x32_10(sub, epoch, obs, S4B_STREAM_LATENT, k0, k1, o0, o1);
  return bits_to_uniform(o0, o1);
}

// z ~ N(mean, 1) truncated to (0, inf) if positive else (-inf, 0), by inversion from one keyed uniform:
// x = m - Phi^-1(u Phi(m)) with m = +-mean (oracle/s4b_rng.h s4b_keyed_truncnorm); no data-dependent loop
__host__ __device__ inline double keyed_truncnorm(uint32_t k0, uint32_t k1, uint32_t obs, uint32_t epoch, double mean, bool positive)
{
  const double m = positive ? mean : -mean;
  const double u = keyed_uniform(k0, k1, obs, epoch, 0u);
  const double pm = 0.5 * erfc(-m * 0.70710678118654752440);
  double arg = u * pm;
  if (arg < 1e-300) arg = 1e-300;
  double x = m - qnorm_as241(arg);
  if (!(x > 0.0)) x = 0.0;
  return positive ? x : -x;
}

// ---------------------------------------------------------------------------------------
// device-resident tree: pre-order array, left child of i is i + 1
// ---------------------------------------------------------------------------------------
struct DNode {
  int16_t var;      // < 0 => bottom node
  int16_t cut;
  int16_t right;    // index of right child
  int16_t parent;   // -1 for the root
  int32_t n;        // observations in node at the last visit (bottom nodes)
  int32_t depth;
  double mu;        // leaf value, scaled units
};

struct DTree {
  int32_t num_nodes;
  int32_t pad;
  DNode nodes[S4B_NODE_CAP];
};

// packed traversal record: var (16 bits, 0xFFFF = leaf) | cut (8) | right (8)
__host__ __device__ inline uint32_t pack_trav(int var, int cut, int right)
{
  return ((uint32_t) (var < 0 ? 0xFFFF : var) << 16) | ((uint32_t) (cut & 0xFF) << 8) | (uint32_t) (right & 0xFF);
}

// Trees with at most S4B_BITMAP_INT internal nodes (nearly all of them) are also described for the "bitmap walk" of the
// persistent sweep kernel: every internal node's rule is evaluated for an observation (bit i = 1 iff x[var_i] <= cut_i,
// i = rank of the node among the internal nodes in index order) and the bit pattern indexes a table of bottom nodes.
#define S4B_BITMAP_INT 8
struct TravTree {
  int32_t n;
  int32_t pad;
  uint32_t trav[S4B_NODE_CAP];
  double val[S4B_NODE_CAP];
  uint8_t slot[S4B_NODE_CAP];
  int32_t n_int;                               // internal nodes, or 255 when the tree is too large for the bitmap walk
  uint32_t irec[S4B_BITMAP_INT];               // var << 8 | cut of internal node i
  uint8_t table[1 << S4B_BITMAP_INT];          // rule pattern -> index of the bottom node reached
};

// what the N-length pass of tree step t needs; written by the controller of step t - 1
struct StepDesc {
  // (A) fit/residual update of the previous tree: R += old_val[leaf_old] - new_val[leaf_new]
  int32_t a_valid;
  int32_t a_same;          // structure unchanged: a_old.val already holds mu_old - mu_new
  TravTree a_old, a_new;
  // (B) sufficient statistics for the current tree under its proposal
  int32_t b_tree;
  int32_t b_kind;          // 0 birth 1 death 2 change 3 swap, +10 aborted, -1 none
  int32_t b_node;          // pre-order index of the node the proposal acts on
  int32_t b_var, b_cut;    // birth: proposed rule
  int32_t b_num_leaves;    // L: slots [0, L) are the current leaves in index order
  int32_t b_nslots;
  int32_t b_child;         // swap: child index or -1 (both)
  TravTree b_cur;          // current tree, val = mu, slot = leaf slot
  TravTree b_prop;         // proposed tree (change / swap): slot = L + rank for leaves below b_node, 255 elsewhere
  double log_prior_trans;  // data-independent part of the MH ratio (plain ratio for birth / death, log for change / swap)
  double accept_thr;       // persistent sweep: log(u) - log(data-independent part): the step is accepted iff this is < delta log-lik
  int32_t new_var, new_cut;// change: proposed rule
  int32_t b_end;           // change / swap: one past the last pre-order index of the branch below b_node
  int32_t pad1;
};

struct BartParams {
  long long n, n_test;
  int p, num_trees, n_cuts, min_obs, is_binary, thin;
  double birth_death_prob, swap_prob, change_prob, birth_prob, base, power, leaf_prec;
  double sigma;             // scaled units
  double smin, smax, srange;
  uint32_t key0, key1;
  uint32_t latent_epoch;
  uint32_t error_flag;      // set by kernels on capacity / tape problems
  unsigned long long step_id;     // tree steps taken so far (keys the BART RNG substreams)
  unsigned long long prior_calls; // sampleTreesFromPrior calls so far
  // bart_args split.probs as integer weights (round(2^30 p_j / sum p), >= 1 when p_j > 0), nullptr = uniform
  const uint32_t* split_w;
  unsigned long long split_total; // sum of the weights
  int p_pos;                      // predictors with a positive weight
  int weighted;                   // observation weights present: leaf statistics are (count, sum w r, sum w)
  const int* ncuts_var;           // bart_args n.cuts given per predictor (each <= n_cuts), nullptr = n_cuts everywhere
  // leaf prior mu ~ N(0, (node_scale / (k sqrt(T)))^2); k_df > 0: k is sampled after every sweep under k ~ chi(k_df, scale)
  double k, k_df, k_inv_scale2, node_scale;
  // change rule without its proposal (Hastings) term: see s4b_bart_config::change_symmetric
  int change_symmetric, pad_cs;
};

// The opt-in limit of dynamic shared memory is an attribute of the kernel FUNCTION, not of a launch: every kernel that needs more than
// the default gets the most its static shared memory leaves, once and for all, so that two fits of different shapes in one process
// (other predictors, other tree counts, other models) cannot lower each other's limit between their launches.
inline cudaError_t s4b_allow_max_dynamic_smem(const void* fn)
{
  int dev = 0, optin = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  cudaFuncAttributes attr;
  e = cudaFuncGetAttributes(&attr, fn);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int) attr.sharedSizeBytes);
}

}  // namespace s4b

// stan4bart_b200/csrc/bart.cu
// BART half of the Gibbs sweep on one B200: kernels + the host object that owns device state.
// Mirrors the dbarts entry points stan4bart binds through BARTFunctionTable
// (/root/reference/src/init.cpp:54-81): initializeFit, setOffset, setSigma, sampleTreesFromPrior,
// runSamplerWithResults, storeLatents (getLatentVariables), predict, getTrees.
#include "bart.hpp"
#include "bart_kernels.cuh"
#include "sweep_kernel.cuh"
#include "sweep_pipe.cuh"
#include "leaf_stats.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace s4b {

constexpr int kBlock = 256;
constexpr int kStreamNq = 99;      // persistent_nq_ value of the streamed sweep variant

enum PassMode { kModeStep = 0, kModeStatsOnly = 1, kModeUpdateOnly = 2 };

// ---------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_words(void* dst, const void* src, int nbytes, int tid, int nthreads)
{
  uint32_t* d = (uint32_t*) dst; const uint32_t* s = (const uint32_t*) src;
  for (int i = tid; i < nbytes / 4; i += nthreads) d[i] = s[i];
}

__device__ __forceinline__ void copy_trav(TravTree& dst, const TravTree& src, int tid, int nthreads, bool with_val, bool with_slot)
{
  int n = src.n;
  if (tid == 0) dst.n = n;
  for (int i = tid; i < n; i += nthreads) {
    dst.trav[i] = src.trav[i];
    if (with_val) dst.val[i] = src.val[i];
    if (with_slot) dst.slot[i] = src.slot[i];
  }
}

__device__ __forceinline__ int traverse(const uint32_t* __restrict__ trav, const uint8_t* __restrict__ xt, long long npad, long long i)
{
  int node = 0;
  uint32_t tr = trav[0];
  while ((tr >> 16) != 0xFFFFu) {
    uint32_t x = __ldg(xt + (long long) (tr >> 16) * npad + i);
    node = (x <= ((tr >> 8) & 0xFFu)) ? node + 1 : (int) (tr & 0xFFu);
    tr = trav[node];
  }
  return node;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// the per-tree pass (see bart_kernels.cuh header comment)
// algorithmic bytes per observation: R read 8 + R write 8 (when an update is pending) + one
// u8 per tree level actually visited (typically 1-3)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock, 3) k_tree_step(BartDev dv, int mode, int propose_next, const __grid_constant__ ShardDev sh_param)
{
  __shared__ ShardDev sh;
  __shared__ unsigned long long s_seq;
  __shared__ int s_dead;
  __shared__ StepDesc sd;
  __shared__ StepDesc sd_out;
  __shared__ DTree tree;
  __shared__ LeafStat st[S4B_MAX_SLOTS];
  __shared__ double red[kBlock / 32][3 * S4B_SLOT_CHUNK];
  __shared__ BartParams prm;
  __shared__ double pgrow[S4B_MAX_DEPTH + 2];
  __shared__ unsigned int s_ticket;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const long long clk_start = clock64();

  // ---- stage the step descriptor ----
  if (tid == 0) {
    const StepDesc& g = *dv.desc;
    sd.a_valid = g.a_valid; sd.a_same = g.a_same;
    sd.b_tree = g.b_tree; sd.b_kind = g.b_kind; sd.b_node = g.b_node; sd.b_var = g.b_var; sd.b_cut = g.b_cut;
    sd.b_num_leaves = g.b_num_leaves; sd.b_nslots = g.b_nslots; sd.b_child = g.b_child;
    sd.log_prior_trans = g.log_prior_trans; sd.new_var = g.new_var; sd.new_cut = g.new_cut;
  }
  __syncthreads();
  const bool a_valid = sd.a_valid != 0 && mode != kModeStatsOnly;
  const bool a_same = sd.a_same != 0;
  const int kind = sd.b_kind;
  if (a_valid) { copy_trav(sd.a_old, dv.desc->a_old, tid, kBlock, true, false); if (!a_same) copy_trav(sd.a_new, dv.desc->a_new, tid, kBlock, true, false); }
  if (mode != kModeUpdateOnly) {
    copy_trav(sd.b_cur, dv.desc->b_cur, tid, kBlock, true, true);
    if (kind == 2 || kind == 3) copy_trav(sd.b_prop, dv.desc->b_prop, tid, kBlock, false, true);
  }
  __syncthreads();

  const long long n = dv.n, npad = dv.npad;
  const long long nquad = (n + 3) >> 2;
  const uint8_t* __restrict__ xt = dv.xt;
  double* __restrict__ R = dv.R;
  const int L = sd.b_num_leaves;
  const int nslots = mode == kModeUpdateOnly ? 0 : sd.b_nslots;
  const int nchunks = mode == kModeUpdateOnly ? 1 : (nslots + S4B_SLOT_CHUNK - 1) / S4B_SLOT_CHUNK;
  const bool two_trees = (kind == 2 || kind == 3);
  const int birth_node = kind == 0 ? sd.b_node : -1;
  const long long birth_col = (long long) (kind == 0 ? sd.b_var : 0) * npad;
  const uint32_t birth_cut = (uint32_t) sd.b_cut;

  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int base = chunk * S4B_SLOT_CHUNK;
    int cnt[S4B_SLOT_CHUNK]; double s1[S4B_SLOT_CHUNK], s2[S4B_SLOT_CHUNK];
#pragma unroll
    for (int k = 0; k < S4B_SLOT_CHUNK; ++k) { cnt[k] = 0; s1[k] = 0.0; s2[k] = 0.0; }
    const bool do_update = a_valid && chunk == 0;

    for (long long qd = (long long) blockIdx.x * kBlock + tid; qd < nquad; qd += (long long) gridDim.x * kBlock) {
      const long long i0 = qd << 2;
      double2 ra = *reinterpret_cast<const double2*>(R + i0);
      double2 rb = *reinterpret_cast<const double2*>(R + i0 + 2);
      double r[4] = { ra.x, ra.y, rb.x, rb.y };
      if (do_update) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          int leaf = traverse(sd.a_old.trav, xt, npad, i0 + j);
          r[j] += sd.a_old.val[leaf];
          if (!a_same) { int l2 = traverse(sd.a_new.trav, xt, npad, i0 + j); r[j] -= sd.a_new.val[l2]; }
        }
        *reinterpret_cast<double2*>(R + i0) = make_double2(r[0], r[1]);
        *reinterpret_cast<double2*>(R + i0 + 2) = make_double2(r[2], r[3]);
      }
      if (mode != kModeUpdateOnly) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const long long i = i0 + j;
          int leaf = traverse(sd.b_cur.trav, xt, npad, i);
          double pr = r[j] + sd.b_cur.val[leaf];
          int sa = sd.b_cur.slot[leaf];
          if (leaf == birth_node) sa = L + ((uint32_t) __ldg(xt + birth_col + i) > birth_cut ? 1 : 0);
          int sb = 255;
          if (two_trees) sb = sd.b_prop.slot[traverse(sd.b_prop.trav, xt, npad, i)];
          if (i >= n) { sa = 255; sb = 255; }
          sa -= base; sb -= base;
          double pp = pr * pr;
#pragma unroll
          for (int k = 0; k < S4B_SLOT_CHUNK; ++k) {
            bool m = (sa == k) | (sb == k);
            cnt[k] += m ? 1 : 0;
            s1[k] += m ? pr : 0.0;
            s2[k] += m ? pp : 0.0;
          }
        }
      }
    }
    if (mode == kModeUpdateOnly) return;

    // ---- block reduction of this chunk, fixed order ----
    const int kmax = min(S4B_SLOT_CHUNK, nslots - base);
#pragma unroll
    for (int k = 0; k < S4B_SLOT_CHUNK; ++k) {
      if (k < kmax) {
        int c = __reduce_add_sync(0xffffffffu, cnt[k]);
        double a = warp_sum(s1[k]);
        double b = warp_sum(s2[k]);
        if (lane == 0) { red[warp][3 * k] = (double) c; red[warp][3 * k + 1] = a; red[warp][3 * k + 2] = b; }
      }
    }
    __syncthreads();
    if (tid < 3 * kmax) {
      double acc = 0.0;
#pragma unroll
      for (int w = 0; w < kBlock / 32; ++w) acc += red[w][tid];
      dv.partials[(long long) (3 * base + tid) * gridDim.x + blockIdx.x] = acc;
    }
    __syncthreads();
  }

  // ---- last block to finish becomes the controller ----
  __threadfence();
  if (tid == 0) s_ticket = atomicAdd(dv.ticket, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  long long clk[8];
  clk[0] = clock64();

  // phase 1: reduce per-block partials (value-major layout, lanes stride over blocks)
  const int G = gridDim.x;
  for (int v = warp; v < 3 * nslots; v += kBlock / 32) {
    const double* src = dv.partials + (long long) v * G;
    double acc = 0.0;
    for (int b = lane; b < G; b += 32) acc += __ldcg(src + b);
    acc = warp_sum(acc);
    if (lane == 0) { double* dst = reinterpret_cast<double*>(&st[v / 3]); dst[v % 3] = acc; }
  }
  if (tid == 0) { prm = *dv.params; *dv.ticket = 0u; }
  if (tid < S4B_MAX_DEPTH + 2) pgrow[tid] = dv.pgrow[tid];
  __syncthreads();
  if (sh_param.world > 1) {
    // observation-sharded chain: the controller blocks of all ranks exchange their statistics through the peer-mapped
    // mailboxes (flag-carrying words, see shard.hpp) and add them in rank order -> identical decisions everywhere
    if (tid == 0) {
      sh.rank = sh_param.rank; sh.world = sh_param.world;
      for (int r = 0; r < kMaxRanks; ++r) sh.mail[r] = sh_param.mail[r];
      s_seq = sh_param.mail[sh_param.rank]->kseq + 1ull; s_dead = 0;
    }
    __syncthreads();
    const unsigned int seq32 = (unsigned int) s_seq;
    const int par = (int) (s_seq & 1ull);
    const int cnt = 3 * nslots, world = sh.world;
    double* stv = reinterpret_cast<double*>(st);
    for (int i = tid; i < world * cnt; i += kBlock) { const int dst = i / cnt, k = i - dst * cnt; mailbox_send_ll(&sh.mail[dst]->step_ll[par][sh.rank][k], stv[k], seq32); }
    __syncthreads();
    const Mailbox* mine = sh.mail[sh.rank];
    for (int i = tid; i < cnt; i += kBlock) {
      double acc = 0.0;
      for (int src = 0; src < world; ++src) {
        double v = 0.0;
        if (!s_dead && !mailbox_recv_ll(&mine->step_ll[par][src][i], seq32, &v)) s_dead = 1;
        acc = src == 0 ? v : acc + v;
      }
      stv[i] = acc;
    }
    __syncthreads();
    if (tid == 0) { sh.mail[sh.rank]->kseq = s_seq; if (s_dead) dv.params->error_flag |= 4u; }
  }
  clk[1] = clock64();
  if (mode == kModeStatsOnly) {
    for (int v = tid; v < 3 * nslots; v += kBlock) dv.stats_out[v] = reinterpret_cast<double*>(st)[v];
    return;
  }
  const int t_cur = sd.b_tree;
  {
    const DTree& g = dv.trees[t_cur];
    int nn = g.num_nodes;
    if (tid == 0) { tree.num_nodes = nn; tree.pad = 0; }
    copy_words(tree.nodes, g.nodes, nn * (int) sizeof(DNode), tid, kBlock);
  }
  __syncthreads();

  clk[2] = clock64();
  // phase 2: Metropolis decision + leaf draws (serial)
  if (tid == 0) {
    RngState rng = *dv.rng;
    double* trec = nullptr;
    if (dv.trace != nullptr) {
      unsigned long long k = *dv.trace_len;
      if (k < dv.trace_cap) { trec = dv.trace + k * S4B_TRACE_LEN; for (int i = 0; i < S4B_TRACE_LEN; ++i) trec[i] = 0.0; }
      *dv.trace_len = k + 1;
    }
    decide_and_draw(tree, prm, rng, sd, st, sd_out, trec, prm.step_id);
    dv.params->step_id = prm.step_id + 1ull;
    *dv.rng = rng;
    if (rng.tape_underrun) dv.params->error_flag |= 2u;
  }
  __syncthreads();
  clk[3] = clock64();
  // phase 3: write tree back, fetch the next one
  {
    DTree& g = dv.trees[t_cur];
    int nn = tree.num_nodes;
    if (tid == 0) g.num_nodes = nn;
    copy_words(g.nodes, tree.nodes, nn * (int) sizeof(DNode), tid, kBlock);
  }
  __syncthreads();
  if (propose_next) {
    const int t_next = (t_cur + 1) % prm.num_trees;
    const DTree& g = dv.trees[t_next];
    int nn = g.num_nodes;
    if (tid == 0) { tree.num_nodes = nn; }
    copy_words(tree.nodes, g.nodes, nn * (int) sizeof(DNode), tid, kBlock);
    __syncthreads();
    clk[4] = clock64();
    if (tid == 0) {
      RngState rng = *dv.rng;
      propose_step(tree, prm, pgrow, rng, sd_out, t_next, prm.step_id + 1ull);
      *dv.rng = rng;
      if (rng.tape_underrun) dv.params->error_flag |= 2u;
    }
  } else {
    clk[4] = clock64();
  }
  if (!propose_next && tid == 0) {
    sd_out.b_kind = -1; sd_out.b_tree = -1; sd_out.b_num_leaves = 0; sd_out.b_nslots = 0; sd_out.b_cur.n = 0; sd_out.b_prop.n = 0;
  }
  __syncthreads();
  clk[5] = clock64();
  // phase 4: publish the descriptor for the next launch
  {
    StepDesc& g = *dv.desc;
    if (tid == 0) {
      g.a_valid = sd_out.a_valid; g.a_same = sd_out.a_same;
      g.b_tree = sd_out.b_tree; g.b_kind = sd_out.b_kind; g.b_node = sd_out.b_node; g.b_var = sd_out.b_var; g.b_cut = sd_out.b_cut;
      g.b_num_leaves = sd_out.b_num_leaves; g.b_nslots = sd_out.b_nslots; g.b_child = sd_out.b_child;
      g.log_prior_trans = sd_out.log_prior_trans; g.new_var = sd_out.new_var; g.new_cut = sd_out.new_cut;
    }
    copy_trav(g.a_old, sd_out.a_old, tid, kBlock, true, false);
    if (!sd_out.a_same) copy_trav(g.a_new, sd_out.a_new, tid, kBlock, true, false);
    if (sd_out.b_kind != -1 || propose_next) {
      copy_trav(g.b_cur, sd_out.b_cur, tid, kBlock, true, true);
      if (sd_out.b_kind == 2 || sd_out.b_kind == 3) copy_trav(g.b_prop, sd_out.b_prop, tid, kBlock, false, true);
    }
  }
  if (tid == 0 && dv.prof != nullptr) {
    clk[6] = clock64();
    dv.prof[0] += (unsigned long long) (clk[0] - clk_start);
    for (int i = 1; i <= 6; ++i) dv.prof[i] += (unsigned long long) (clk[i] - clk[i - 1]);
    dv.prof[7] += 1ull;
  }
}

// first proposal of a sweep (tree 0); single block
__global__ void __launch_bounds__(kBlock) k_propose_first(BartDev dv)
{
  __shared__ DTree tree;
  __shared__ StepDesc sd_out;
  __shared__ BartParams prm;
  __shared__ double pgrow[S4B_MAX_DEPTH + 2];
  const int tid = threadIdx.x;
  if (tid == 0) prm = *dv.params;
  if (tid < S4B_MAX_DEPTH + 2) pgrow[tid] = dv.pgrow[tid];
  {
    const DTree& g = dv.trees[0];
    int nn = g.num_nodes;
    if (tid == 0) tree.num_nodes = nn;
    copy_words(tree.nodes, g.nodes, nn * (int) sizeof(DNode), tid, kBlock);
  }
  __syncthreads();
  if (tid == 0) {
    RngState rng = *dv.rng;
    sd_out.a_valid = 0; sd_out.a_same = 1; sd_out.a_old.n = 0; sd_out.a_new.n = 0;
    propose_step(tree, prm, pgrow, rng, sd_out, 0, prm.step_id);
    *dv.rng = rng;
    if (rng.tape_underrun) dv.params->error_flag |= 2u;
  }
  __syncthreads();
  StepDesc& g = *dv.desc;
  if (tid == 0) {
    g.a_valid = 0; g.a_same = 1;
    g.b_tree = sd_out.b_tree; g.b_kind = sd_out.b_kind; g.b_node = sd_out.b_node; g.b_var = sd_out.b_var; g.b_cut = sd_out.b_cut;
    g.b_num_leaves = sd_out.b_num_leaves; g.b_nslots = sd_out.b_nslots; g.b_child = sd_out.b_child;
    g.log_prior_trans = sd_out.log_prior_trans; g.new_var = sd_out.new_var; g.new_cut = sd_out.new_cut;
  }
  copy_trav(g.b_cur, sd_out.b_cur, tid, kBlock, true, true);
  if (sd_out.b_kind == 2 || sd_out.b_kind == 3) copy_trav(g.b_prop, sd_out.b_prop, tid, kBlock, false, true);
}

// descriptor for a stand-alone statistics pass over the current leaves of one tree
__global__ void k_build_stats_desc(BartDev dv, int tree_index)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  StepDesc& g = *dv.desc;
  const DTree& t = dv.trees[tree_index];
  g.a_valid = 0; g.a_same = 1; g.a_old.n = 0; g.a_new.n = 0;
  g.b_tree = tree_index; g.b_kind = -1; g.b_node = -1; g.b_var = -1; g.b_cut = -1; g.b_child = -1;
  g.log_prior_trans = 0.0; g.new_var = -1; g.new_cut = -1; g.b_prop.n = 0;
  g.b_cur.n = t.num_nodes;
  int leaf = 0;
  for (int k = 0; k < t.num_nodes; ++k) {
    const DNode& nd = t.nodes[k];
    g.b_cur.trav[k] = pack_trav(nd.var, nd.cut, nd.right);
    g.b_cur.val[k] = nd.mu;
    g.b_cur.slot[k] = nd.var < 0 ? (uint8_t) leaf++ : (uint8_t) 255;
  }
  g.b_num_leaves = leaf; g.b_nslots = leaf;
}

// one tree drawn from the CGM prior + its leaf values; fills the (A) part for an update-only pass
__global__ void k_prior_tree(BartDev dv, int tree_index)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  BartParams P = *dv.params;
  RngState rng = *dv.rng;
  rng_enter(rng, P.prior_calls * (unsigned long long) P.num_trees + (unsigned long long) tree_index, 2u);
  if (tree_index == P.num_trees - 1) dv.params->prior_calls = P.prior_calls + 1ull;
  DTree& t = dv.trees[tree_index];
  StepDesc& g = *dv.desc;
  // old tree -> a_old
  g.a_valid = 1; g.a_same = 0; g.b_kind = -1; g.b_tree = tree_index;
  g.a_old.n = t.num_nodes;
  for (int k = 0; k < t.num_nodes; ++k) { g.a_old.trav[k] = pack_trav(t.nodes[k].var, t.nodes[k].cut, t.nodes[k].right); g.a_old.val[k] = t.nodes[k].mu; }
  // collapse to the root, then grow in pre-order
  t.num_nodes = 1;
  DNode& root = t.nodes[0];
  root.var = -1; root.cut = -1; root.right = -1; root.parent = -1; root.depth = 0; root.n = 0;
  int leaves = 1;
  for (int i = 0; i < t.num_nodes; ++i) {
    int navail = t_num_vars_available(t, P, i);
    bool birthable = navail > 0 && t.nodes[i].depth < S4B_MAX_DEPTH && leaves < S4B_MAX_LEAVES;
    double pg = birthable ? t_growth_prob_depth(dv.pgrow, navail, t.nodes[i].depth) : 0.0;
    double u = rng_uniform(rng);
    if (!(u < pg)) continue;
    int var = t_draw_var(t, P, i, navail, rng);
    int lo, hi; t_split_interval(t, s4b_ncuts(P, var), i, var, lo, hi);
    int cut = lo + rng_index(rng, hi - lo + 1);
    t_insert_children(t, i, var, cut);
    ++leaves;
  }
  for (int k = 0; k < t.num_nodes; ++k) if (t.nodes[k].var < 0) t.nodes[k].mu = rng_normal(rng) / sqrt(P.leaf_prec);
  g.a_new.n = t.num_nodes;
  for (int k = 0; k < t.num_nodes; ++k) { g.a_new.trav[k] = pack_trav(t.nodes[k].var, t.nodes[k].cut, t.nodes[k].right); g.a_new.val[k] = t.nodes[k].mu; }
  *dv.rng = rng;
  if (rng.tape_underrun) dv.params->error_flag |= 2u;
}

// ---------------------------------------------------------------------------------------
// sweep epilogue: pending update of the last tree, probit latents, results
//   train_out  = BART fit in original units (+ offset if add_offset)   [SURVEY a2]
//   latent_out = full latent z (binary)                                [SURVEY a9]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_finish_sweep(BartDev dv, double* __restrict__ train_out, double* __restrict__ latent_out, int add_offset,
                                                         double* __restrict__ test_alias_out)
{
  __shared__ TravTree a_old, a_new;
  __shared__ BartParams prm;
  const int tid = threadIdx.x;
  const StepDesc& g = *dv.desc;
  const bool a_valid = g.a_valid != 0;
  const bool a_same = g.a_same != 0;
  if (tid == 0) prm = *dv.params;
  if (a_valid) { copy_trav(a_old, g.a_old, tid, kBlock, true, false); if (!a_same) copy_trav(a_new, g.a_new, tid, kBlock, true, false); }
  __syncthreads();
  const long long n = dv.n, npad = dv.npad;
  const bool binary = prm.is_binary != 0;
  for (long long i = (long long) blockIdx.x * kBlock + tid; i < n; i += (long long) gridDim.x * kBlock) {
    double r = dv.R[i];
    if (a_valid) {
      r += a_old.val[traverse(a_old.trav, dv.xt, npad, i)];
      if (!a_same) r -= a_new.val[traverse(a_new.trav, dv.xt, npad, i)];
    }
    double yr = dv.yresc[i];
    double tf = yr - r;                      // total fit, scaled units
    double off = dv.offset[i];
    if (binary) {
      double z = keyed_truncnorm(prm.key0, prm.key1, (uint32_t) (i + dv.obs_offset), prm.latent_epoch, tf + off, dv.y[i] > 0.0);
      yr = z - off;
      dv.yresc[i] = yr;
      r = yr - tf;
      if (latent_out != nullptr) latent_out[i] = z;
    }
    dv.R[i] = r;
    if (train_out != nullptr) {
      double f = binary ? tf : prm.smin + (tf + 0.5) * prm.srange;
      train_out[i] = add_offset ? f + off : f;
      // test design identical to the training design (same binned rows): the test fit is the training fit
      if (test_alias_out != nullptr) test_alias_out[i] = f;
    }
  }
}

// k ~ chi(df, scale) hyperprior of the leaf prior (bart_args k = chi(1.25, Inf); the reference's `!kPrior->isFixed`,
// src/init.cpp:731): given the L leaf values of all trees, k^2 ~ Gamma((L + df) / 2, rate = T sum mu^2 / (2 node_scale^2)
// + 1 / (2 scale^2)).  One block; the per-thread partial sums are added in thread order (deterministic).
__global__ void __launch_bounds__(256) k_draw_k(BartDev dv)
{
  __shared__ double s_sum[256];
  __shared__ int s_cnt[256];
  const int tid = threadIdx.x;
  BartParams& P = *dv.params;
  double sum = 0.0; int cnt = 0;
  for (int t = tid; t < P.num_trees; t += 256) {
    const DTree& tr = dv.trees[t];
    for (int k = 0; k < tr.num_nodes; ++k) if (tr.nodes[k].var < 0) { const double mu = tr.nodes[k].mu; sum += mu * mu; ++cnt; }
  }
  s_sum[tid] = sum; s_cnt[tid] = cnt;
  __syncthreads();
  if (tid != 0) return;
  double S = 0.0; int L = 0;
  for (int i = 0; i < 256; ++i) { S += s_sum[i]; L += s_cnt[i]; }
  RngState rng = *dv.rng;
  rng_enter(rng, P.step_id, 3);
  const double shape = 0.5 * ((double) L + P.k_df);
  const double rate = 0.5 * (S * (double) P.num_trees / (P.node_scale * P.node_scale) + P.k_inv_scale2);
  const double k = sqrt(rng_gamma(rng, shape) / rate);
  const double sd_leaf = P.node_scale / (k * sqrt((double) P.num_trees));
  P.k = k;
  P.leaf_prec = 1.0 / (sd_leaf * sd_leaf);
  if (rng.tape_underrun) P.error_flag |= 2u;
  *dv.rng = rng;
}

__global__ void k_bump_epoch_clear_update(BartDev dv, int bump)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) { if (bump) dv.params->latent_epoch += 1u; dv.desc->a_valid = 0; }
}

// test-sample fits: sum over all trees of the leaf value reached by each test row.  The block's rows are staged in
// shared memory (p bytes per row), the trees are streamed through shared memory in batches; all walks are on chip.
// `trees` / `scale` (smin, srange) default to the live sampler state; stored samples (keepTrees) pass their own
__global__ void __launch_bounds__(kBlock) k_test_fits(BartDev dv, const uint8_t* __restrict__ xt_test, long long n_test, long long npad_test,
                                                      const double* __restrict__ test_offset, double* __restrict__ out, int unscale, int p,
                                                      const DTree* __restrict__ trees_in, const double* __restrict__ scale_in)
{
  const DTree* __restrict__ trees = trees_in != nullptr ? trees_in : dv.trees;
  extern __shared__ unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const int T = dv.params->num_trees;
  const int cap_nodes = 2048;
  uint32_t* trav = reinterpret_cast<uint32_t*>(smem_raw);
  double* val = reinterpret_cast<double*>(smem_raw + sizeof(uint32_t) * cap_nodes);
  uint8_t* rows = smem_raw + (sizeof(uint32_t) + sizeof(double)) * cap_nodes;      // [p][kBlock]
  __shared__ int tree_start[257];
  const long long i = (long long) blockIdx.x * kBlock + tid;
  for (int v = 0; v < p; ++v) rows[v * kBlock + tid] = i < n_test ? xt_test[(long long) v * npad_test + i] : (uint8_t) 0;
  double acc = 0.0;
  int t0 = 0;
  while (t0 < T) {
    __syncthreads();
    if (tid == 0) {
      int used = 0, t = t0, k = 0;
      while (t < T && k < 256 && used + trees[t].num_nodes <= cap_nodes) { tree_start[k++] = used; used += trees[t].num_nodes; ++t; }
      tree_start[k] = used;
      tree_start[256] = k;
    }
    __syncthreads();
    const int nt = tree_start[256];
    for (int k = tid; k < nt * 8; k += kBlock) {
      // 8 threads per tree copy its nodes
      const int tr = k >> 3, sub = k & 7;
      const DTree& g = trees[t0 + tr];
      const int base = tree_start[tr];
      for (int j = sub; j < g.num_nodes; j += 8) { const DNode& nd = g.nodes[j]; trav[base + j] = pack_trav(nd.var, nd.cut, nd.right); val[base + j] = nd.mu; }
    }
    __syncthreads();
    if (i < n_test) {
      for (int k = 0; k < nt; ++k) {
        const uint32_t* tv = trav + tree_start[k];
        int node = 0;
        uint32_t tr = tv[0];
        while ((tr >> 16) != 0xFFFFu) {
          const uint32_t x = rows[(tr >> 16) * kBlock + tid];
          node = (x <= ((tr >> 8) & 0xFFu)) ? node + 1 : (int) (tr & 0xFFu);
          tr = tv[node];
        }
        acc += val[tree_start[k] + node];
      }
    }
    t0 += nt;
  }
  if (i < n_test) {
    const double smin = scale_in != nullptr ? scale_in[0] : dv.params->smin, srange = scale_in != nullptr ? scale_in[1] : dv.params->srange;
    double f = unscale ? smin + (acc + 0.5) * srange : acc;
    out[i] = f + (test_offset != nullptr ? test_offset[i] : 0.0);
  }
}

// ---------------------------------------------------------------------------------------
// setOffset / setSigma / rescale
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_minmax(BartDev dv, const double* __restrict__ new_offset, double* __restrict__ part /* [2][G] */)
{
  __shared__ double smin_s[kBlock / 32], smax_s[kBlock / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double mn = INFINITY, mx = -INFINITY;
  for (long long i = (long long) blockIdx.x * kBlock + tid; i < dv.n; i += (long long) gridDim.x * kBlock) {
    double v = dv.y[i] - (new_offset != nullptr ? new_offset[i] : 0.0);
    mn = fmin(mn, v); mx = fmax(mx, v);
  }
  for (int o = 16; o > 0; o >>= 1) { mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if (lane == 0) { smin_s[warp] = mn; smax_s[warp] = mx; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < kBlock / 32; ++w) { mn = fmin(mn, smin_s[w]); mx = fmax(mx, smax_s[w]); }
    part[blockIdx.x] = mn; part[gridDim.x + blockIdx.x] = mx;
  }
}

// sharded chains: fold the block partials into (-min, max) so that one max-reduction over the ranks finishes both, then
// unpack into the [2][1] layout k_update_scale reads
__global__ void k_minmax_pack(const double* __restrict__ part, int G, double* __restrict__ packed)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double mn = INFINITY, mx = -INFINITY;
  for (int b = 0; b < G; ++b) { mn = fmin(mn, part[b]); mx = fmax(mx, part[G + b]); }
  packed[0] = -mn; packed[1] = mx;
}
__global__ void k_minmax_unpack(const double* __restrict__ packed, double* __restrict__ part1)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) { part1[0] = -packed[0]; part1[1] = packed[1]; }
}

// single thread: finish min/max, update the scale, sigma and (optionally) leaf values
__global__ void k_update_scale(BartDev dv, const double* __restrict__ part, int G, double* __restrict__ scale_factor_out)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  BartParams& P = *dv.params;
  double mn = INFINITY, mx = -INFINITY;
  for (int b = 0; b < G; ++b) { mn = fmin(mn, part[b]); mx = fmax(mx, part[G + b]); }
  double old_range = P.srange;
  double sigma_unscaled = old_range > 0.0 ? P.sigma * old_range : P.sigma;
  double range = mx - mn;
  if (!(range > 0.0)) range = 1.0;
  P.smin = mn; P.smax = mx; P.srange = range;
  double s = 1.0;
  if (old_range > 0.0) { s = old_range / range; P.sigma = sigma_unscaled / range; }
  *scale_factor_out = s;
}

__global__ void k_scale_leaves(BartDev dv, const double* __restrict__ scale_factor)
{
  const double s = *scale_factor;
  const int t = blockIdx.x;
  DTree& tr = dv.trees[t];
  for (int k = threadIdx.x; k < tr.num_nodes; k += blockDim.x) if (tr.nodes[k].var < 0) tr.nodes[k].mu *= s;
}

// yresc / R refresh for a new offset.  scale_factor == nullptr => scale unchanged.
__global__ void __launch_bounds__(kBlock) k_apply_offset(BartDev dv, const double* __restrict__ new_offset, const double* __restrict__ scale_factor)
{
  const double s = scale_factor != nullptr ? *scale_factor : 1.0;
  const double smin = dv.params->smin, srange = dv.params->srange;
  for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < dv.n; i += (long long) gridDim.x * kBlock) {
    double off = new_offset != nullptr ? new_offset[i] : 0.0;
    double tf = (dv.yresc[i] - dv.R[i]) * s;
    double yr = (dv.y[i] - off - smin) / srange - 0.5;
    dv.yresc[i] = yr;
    dv.R[i] = yr - tf;
    dv.offset[i] = off;
  }
}

// binary: new offset => redraw the latents around (fit + offset)  (oracle_bart.c sample_latents)
__global__ void __launch_bounds__(kBlock) k_apply_offset_binary(BartDev dv, const double* __restrict__ new_offset, double* __restrict__ latent_out)
{
  const BartParams& P = *dv.params;
  const uint32_t k0 = P.key0, k1 = P.key1, epoch = P.latent_epoch;
  for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < dv.n; i += (long long) gridDim.x * kBlock) {
    double off = new_offset != nullptr ? new_offset[i] : 0.0;
    double tf = dv.yresc[i] - dv.R[i];
    double z = keyed_truncnorm(k0, k1, (uint32_t) (i + dv.obs_offset), epoch, tf + off, dv.y[i] > 0.0);
    double yr = z - off;
    dv.yresc[i] = yr;
    dv.R[i] = yr - tf;
    dv.offset[i] = off;
    if (latent_out != nullptr) latent_out[i] = z;
  }
}

__global__ void k_set_sigma(BartDev dv, double sigma)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) dv.params->sigma = dv.params->is_binary ? 1.0 : sigma / dv.params->srange;
}

__global__ void k_store_latents(BartDev dv, double* __restrict__ out)
{
  for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < dv.n; i += (long long) gridDim.x * blockDim.x) out[i] = dv.yresc[i] + dv.offset[i];
}

// keepTrees: copy the current trees and the current response scale into slot `slot` of the store
__global__ void k_snapshot_trees(BartDev dv, DTree* __restrict__ store, double* __restrict__ scales)
{
  const int t = blockIdx.x;
  const DTree& g = dv.trees[t];
  DTree& d = store[t];
  const int nn = g.num_nodes;
  if (threadIdx.x == 0) { d.num_nodes = nn; d.pad = 0; }
  for (int i = threadIdx.x; i < nn * (int) (sizeof(DNode) / 4); i += blockDim.x) reinterpret_cast<uint32_t*>(d.nodes)[i] = reinterpret_cast<const uint32_t*>(g.nodes)[i];
  if (t == 0 && threadIdx.x == 0) { scales[0] = dv.params->smin; scales[1] = dv.params->srange; }
}

__global__ void k_varcount(BartDev dv, unsigned int* __restrict__ out)
{
  const int T = dv.params->num_trees;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < T; t += gridDim.x * blockDim.x) {
    const DTree& tr = dv.trees[t];
    for (int k = 0; k < tr.num_nodes; ++k) if (tr.nodes[k].var >= 0) atomicAdd(out + tr.nodes[k].var, 1u);
  }
}

// observation -> node partition of one tree as heap indices (integer, bit-exact contract)
__global__ void __launch_bounds__(kBlock) k_node_assignment(BartDev dv, int tree_index, long long* __restrict__ out)
{
  __shared__ uint32_t trav[S4B_NODE_CAP];
  const DTree& g = dv.trees[tree_index];
  for (int k = threadIdx.x; k < g.num_nodes; k += kBlock) trav[k] = pack_trav(g.nodes[k].var, g.nodes[k].cut, g.nodes[k].right);
  __syncthreads();
  for (long long i = (long long) blockIdx.x * kBlock + threadIdx.x; i < dv.n; i += (long long) gridDim.x * kBlock) {
    int node = 0; long long h = 1;
    uint32_t tr = trav[0];
    while ((tr >> 16) != 0xFFFFu) {
      uint32_t x = __ldg(dv.xt + (long long) (tr >> 16) * dv.npad + i);
      bool left = x <= ((tr >> 8) & 0xFFu);
      node = left ? node + 1 : (int) (tr & 0xFFu);
      h = 2 * h + (left ? 0 : 1);
      tr = trav[node];
    }
    out[i] = h;
  }
}

// ---------------------------------------------------------------------------------------
// host object
// ---------------------------------------------------------------------------------------
static int env_int(const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; }

// use_quantiles on an observation-sharded chain: every rank holds the sorted distinct values of its own rows; the cut points are a
// function of the sorted distinct values of all rows, so the ranks exchange theirs (setup only).  The exchange is an all-gather
// written as a rank-ordered sum over the peer mailboxes: every rank contributes its values at its own offset of a zero vector
// (x + 0 is exact), first the counts, then the values.  Every rank ends with the same vector and therefore the same cuts.
void BartFit::gather_distinct_over_shards(std::vector<double>& u)
{
  const int world = shard_->world(), rank = shard_->rank();
  std::vector<double> counts((size_t) world, 0.0);
  counts[(size_t) rank] = (double) u.size();
  shard_->allreduce_host(counts.data(), world, kOpSum, stream_);
  size_t total = 0, first = 0;
  for (int r = 0; r < world; ++r) { if (r < rank) first += (size_t) counts[(size_t) r]; total += (size_t) counts[(size_t) r]; }
  std::vector<double> all(total, 0.0);
  std::copy(u.begin(), u.end(), all.begin() + (std::ptrdiff_t) first);
  shard_->allreduce_host(all.data(), (long long) total, kOpSum, stream_);
  // the segments are sorted: merge them pairwise in rank order, then drop the values that several shards hold
  size_t done = (size_t) counts[0];
  for (int r = 1; r < world; ++r) {
    const size_t next = done + (size_t) counts[(size_t) r];
    std::inplace_merge(all.begin(), all.begin() + (std::ptrdiff_t) done, all.begin() + (std::ptrdiff_t) next);
    done = next;
  }
  all.erase(std::unique(all.begin(), all.end()), all.end());
  u.swap(all);
}

BartFit::BartFit(const s4b_bart_config& cfg, const double* y, const double* x, const double* x_test, cudaStream_t stream, ShardContext* shard)
    : cfg_(cfg), stream_(stream), shard_(shard)
{
  if (shard_ != nullptr && !shard_->attached()) throw std::invalid_argument("sharded fit: attach the peer mailboxes first");
  if (cfg.n_cuts < 1 || cfg.n_cuts > 255) throw std::invalid_argument("n_cuts must be in [1, 255] (u8 bins)");
  if (cfg.p < 1 || cfg.p > 32767) throw std::invalid_argument("p out of range");
  if (cfg.num_trees < 1) throw std::invalid_argument("num_trees < 1");
  if (cfg.n < 1) throw std::invalid_argument("n < 1");
  n_ = cfg.n; p_ = (int) cfg.p; nt_ = cfg.n_test; T_ = cfg.num_trees;
  npad_ = (n_ + 15) / 16 * 16; npad_t_ = (nt_ + 15) / 16 * 16;
  int dev = 0; S4B_CUDA(cudaGetDevice(&dev));
  S4B_CUDA(cudaDeviceGetAttribute(&num_sms_, cudaDevAttrMultiProcessorCount, dev));
  blocks_per_sm_ = env_int("S4B_BLOCKS_PER_SM", 3);
  long long want = ((n_ + 3) / 4 + kBlock - 1) / kBlock;
  grid_ = (int) std::max<long long>(1, std::min<long long>(want, (long long) num_sms_ * blocks_per_sm_));
  long long want1 = (n_ + kBlock - 1) / kBlock;
  grid_ew_ = (int) std::max<long long>(1, std::min<long long>(want1, (long long) num_sms_ * 8));

  if (cfg.n_cuts_var != nullptr) {
    ncuts_var_.assign(cfg.n_cuts_var, cfg.n_cuts_var + p_);
    for (int j = 0; j < p_; ++j) if (ncuts_var_[(size_t) j] < 1 || ncuts_var_[(size_t) j] > cfg.n_cuts) throw std::invalid_argument("n_cuts_var entries must be in [1, n_cuts]");
  }
  cfg_.n_cuts_var = nullptr;
  if (cfg.use_quantiles != 0) {
    // bart_args use.quantiles: every predictor gets its own number of cuts (at most its n.cuts; fewer when it has few distinct values)
    if (ncuts_var_.empty()) ncuts_var_.assign((size_t) p_, cfg.n_cuts);
  }
  // ---- cut points (uniform over the training range, or between the distinct sorted values) and binning, host side (setup only) ----
  cuts_.resize((size_t) p_ * cfg.n_cuts);
  std::vector<uint8_t> xt((size_t) p_ * npad_, 0);
  std::vector<double> range((size_t) 2 * p_);          // (-min, max) per predictor: one max-reduction over the shards
  for (int j = 0; j < p_; ++j) {
    const double* col = x + (size_t) j * n_;
    double mn = col[0], mx = col[0];
    for (long long i = 1; i < n_; ++i) { mn = std::min(mn, col[i]); mx = std::max(mx, col[i]); }
    range[(size_t) 2 * j] = -mn; range[(size_t) 2 * j + 1] = mx;
  }
  if (sharded()) shard_->allreduce_host(range.data(), (long long) range.size(), kOpMax, stream_);
  for (int j = 0; j < p_; ++j) {
    const double* col = x + (size_t) j * n_;
    const double mn = -range[(size_t) 2 * j], mx = range[(size_t) 2 * j + 1];
    // a predictor with fewer cuts than n_cuts (bart_args n.cuts as a vector) pads its row with +inf: binning never goes past its last cut
    int mj = ncuts_var_.empty() ? cfg.n_cuts : ncuts_var_[(size_t) j];
    double* c = cuts_.data() + (size_t) j * cfg.n_cuts;
    if (cfg.use_quantiles != 0) {
      // dbarts' quantile rule (as recalled; dbarts is not vendored -- DESIGN.md section 6): distinct values sorted; few of them => a cut in
      // every gap, otherwise mj cuts every (distinct / mj) values starting half a step in; a cut = midpoint of two neighbouring values
      std::vector<double> u(col, col + n_);
      std::sort(u.begin(), u.end());
      u.erase(std::unique(u.begin(), u.end()), u.end());
      if (sharded()) gather_distinct_over_shards(u);      // the rule is stated on the distinct values of the WHOLE column
      const size_t nu = u.size();
      size_t num, step, offset;
      if (nu <= (size_t) mj + 1) { num = nu - 1; step = 1; offset = 0; }
      else { num = (size_t) mj; step = nu / num; offset = step / 2; }
      for (size_t k = 0; k < num; ++k) { const size_t idx = std::min(k * step + offset, nu - 2); c[k] = 0.5 * (u[idx] + u[idx + 1]); }
      for (int k = (int) num; k < cfg.n_cuts; ++k) c[k] = std::numeric_limits<double>::infinity();
      // a constant predictor has no cut: it is taken out of the variable selection (split weight 0, exactly like split.probs = 0)
      // and keeps one unreachable cut at +inf so that the interval bookkeeping stays well defined
      if (num == 0) { cutless_.push_back(j); num = 1; c[0] = std::numeric_limits<double>::infinity(); }
      mj = (int) num;
      ncuts_var_[(size_t) j] = mj;
    } else {
      const double inc = (mx - mn) / (double) (mj + 1);
      for (int k = 0; k < cfg.n_cuts; ++k) c[k] = k < mj ? mn + (double) (k + 1) * inc : std::numeric_limits<double>::infinity();
    }
    uint8_t* dst = xt.data() + (size_t) j * npad_;
    for (long long i = 0; i < n_; ++i) dst[i] = (uint8_t) (std::lower_bound(c, c + cfg.n_cuts, col[i]) - c);
  }
  if (!ncuts_var_.empty()) {
    S4B_CUDA(cudaMalloc(&d_ncuts_var_, sizeof(int) * (size_t) p_));
    S4B_CUDA(cudaMemcpy(d_ncuts_var_, ncuts_var_.data(), sizeof(int) * (size_t) p_, cudaMemcpyHostToDevice));
  }
  S4B_CUDA(cudaMalloc(&d_xt_, xt.size()));
  S4B_CUDA(cudaMemcpy(d_xt_, xt.data(), xt.size(), cudaMemcpyHostToDevice));
  if (nt_ > 0) {
    std::vector<uint8_t> xtt; bin_matrix(x_test, nt_, npad_t_, xtt);
    // counterfactual designs that differ from the training design only outside the BART covariates (README example,
    // treatment = z) bin to identical rows: their test fits equal the training fits and need no second traversal
    test_aliases_train_ = nt_ == n_ && xtt.size() == xt.size() && std::memcmp(xtt.data(), xt.data(), xt.size()) == 0 && getenv("S4B_NO_TEST_ALIAS") == nullptr;
    S4B_CUDA(cudaMalloc(&d_xt_test_, xtt.size()));
    S4B_CUDA(cudaMemcpy(d_xt_test_, xtt.data(), xtt.size(), cudaMemcpyHostToDevice));
    S4B_CUDA(cudaMalloc(&d_test_out_, sizeof(double) * (size_t) npad_t_));
  }
  auto dalloc = [&](double** p, size_t count) { S4B_CUDA(cudaMalloc(p, sizeof(double) * count)); zero_device_sync(*p, sizeof(double) * count, stream_); };
  dalloc(&d_R_, (size_t) npad_); dalloc(&d_yresc_, (size_t) npad_); dalloc(&d_y_, (size_t) npad_); dalloc(&d_offset_, (size_t) npad_);
  dalloc(&d_train_out_, (size_t) npad_); dalloc(&d_latent_out_, (size_t) npad_); dalloc(&d_offset_in_, (size_t) npad_);
  S4B_CUDA(cudaMemcpy(d_y_, y, sizeof(double) * (size_t) n_, cudaMemcpyHostToDevice));
  dalloc(&d_partials_, (size_t) 3 * S4B_MAX_SLOTS * grid_);
  dalloc(&d_minmax_, (size_t) 2 * grid_ew_ + 8);
  dalloc(&d_stats_out_, (size_t) 3 * S4B_MAX_SLOTS);
  S4B_CUDA(cudaMalloc(&d_desc_, sizeof(StepDesc))); zero_device_sync(d_desc_, sizeof(StepDesc), stream_);
  S4B_CUDA(cudaMalloc(&d_ticket_, sizeof(unsigned int))); zero_device_sync(d_ticket_, sizeof(unsigned int), stream_);
  S4B_CUDA(cudaMalloc(&d_prof_, sizeof(unsigned long long) * 24)); zero_device_sync(d_prof_, sizeof(unsigned long long) * 24, stream_);
  S4B_CUDA(cudaMalloc(&d_trace_len_, sizeof(unsigned long long))); zero_device_sync(d_trace_len_, sizeof(unsigned long long), stream_);
  S4B_CUDA(cudaMalloc(&d_varcount_, sizeof(unsigned int) * (size_t) p_));
  // trees: single root each
  std::vector<DTree> trees((size_t) T_);
  std::memset(trees.data(), 0, sizeof(DTree) * trees.size());
  for (auto& t : trees) { t.num_nodes = 1; t.nodes[0].var = -1; t.nodes[0].cut = -1; t.nodes[0].right = -1; t.nodes[0].parent = -1; t.nodes[0].n = (int32_t) std::min<long long>(sharded() ? shard_->total_obs() : n_, 2147483647LL); }
  S4B_CUDA(cudaMalloc(&d_trees_, sizeof(DTree) * trees.size()));
  S4B_CUDA(cudaMemcpy(d_trees_, trees.data(), sizeof(DTree) * trees.size(), cudaMemcpyHostToDevice));
  // params
  BartParams P; std::memset(&P, 0, sizeof P);
  P.n = n_; P.n_test = nt_; P.p = p_; P.num_trees = T_; P.n_cuts = cfg.n_cuts; P.min_obs = cfg.min_obs; P.is_binary = cfg.is_binary; P.thin = cfg.thin;
  P.birth_death_prob = cfg.birth_death_prob; P.swap_prob = cfg.swap_prob; P.change_prob = cfg.change_prob; P.birth_prob = cfg.birth_prob;
  P.base = cfg.base; P.power = cfg.power;
  double sd_leaf = cfg.node_scale / (cfg.k * std::sqrt((double) cfg.num_trees));
  P.leaf_prec = 1.0 / (sd_leaf * sd_leaf);
  P.k = cfg.k; P.node_scale = cfg.node_scale; P.ncuts_var = d_ncuts_var_;
  if (cfg.k_df < 0.0 || !std::isfinite(cfg.k_df)) throw std::invalid_argument("k_df must be >= 0");
  P.k_df = cfg.k_df;
  P.change_symmetric = cfg.change_symmetric != 0 ? 1 : 0; P.pad_cs = 0;
  P.k_inv_scale2 = (cfg.k_scale > 0.0 && std::isfinite(cfg.k_scale)) ? 1.0 / (cfg.k_scale * cfg.k_scale) : 0.0;
  P.sigma = 1.0; P.smin = -0.5; P.smax = 0.5; P.srange = cfg.is_binary ? 1.0 : 0.0;
  P.key0 = (uint32_t) cfg.seed; P.key1 = (uint32_t) (cfg.seed >> 32);
  std::vector<double> sp_eff;
  if (cfg.split_probs != nullptr) sp_eff.assign(cfg.split_probs, cfg.split_probs + p_);
  if (!cutless_.empty()) {
    if (sp_eff.empty()) sp_eff.assign((size_t) p_, 1.0);
    for (int j : cutless_) sp_eff[(size_t) j] = 0.0;
  }
  if (!sp_eff.empty()) {
    // bart_args split.probs -> integer weights round(2^30 p_j / sum p), at least 1 for a positive probability (same recipe
    // as the oracle: selection and rule priors are then exact integer arithmetic on both sides)
    const double* split_probs = sp_eff.data();
    double sum = 0.0;
    for (int j = 0; j < p_; ++j) { if (!(split_probs[j] >= 0.0)) throw std::invalid_argument("split_probs must be non-negative"); sum += split_probs[j]; }
    if (!(sum > 0.0)) throw std::invalid_argument("split_probs must not all be zero (and with use_quantiles at least one predictor must take two values)");
    std::vector<uint32_t> w((size_t) p_);
    unsigned long long total = 0; int pos = 0;
    for (int j = 0; j < p_; ++j) {
      const double tj = split_probs[j] / sum;
      const double wj = std::floor(std::ldexp(tj, 30) + 0.5);
      w[(size_t) j] = split_probs[j] > 0.0 ? (wj < 1.0 ? 1u : (uint32_t) wj) : 0u;
      total += w[(size_t) j]; pos += w[(size_t) j] != 0u;
    }
    S4B_CUDA(cudaMalloc(&d_split_w_, sizeof(uint32_t) * (size_t) p_));
    S4B_CUDA(cudaMemcpy(d_split_w_, w.data(), sizeof(uint32_t) * (size_t) p_, cudaMemcpyHostToDevice));
    P.split_w = d_split_w_; P.split_total = total; P.p_pos = pos;
    split_probs_.resize((size_t) p_);
    for (int j = 0; j < p_; ++j) split_probs_[(size_t) j] = split_probs[j] / sum;
  }
  cfg_.split_probs = nullptr;        // the caller's array is not kept
  if (cfg.weights != nullptr) {
    // dbarts data weights (R/stan4bart_fit.R:449): y_i ~ N(f(x_i), sigma^2 / w_i)
    for (long long i = 0; i < n_; ++i) if (!(cfg.weights[i] >= 0.0) || !std::isfinite(cfg.weights[i])) throw std::invalid_argument("weights must be finite and non-negative");
    dalloc(&d_wt_, (size_t) npad_);
    S4B_CUDA(cudaMemcpy(d_wt_, cfg.weights, sizeof(double) * (size_t) n_, cudaMemcpyHostToDevice));
    P.weighted = 1;
  }
  cfg_.weights = nullptr;
  S4B_CUDA(cudaMalloc(&d_params_, sizeof(BartParams)));
  S4B_CUDA(cudaMemcpy(d_params_, &P, sizeof P, cudaMemcpyHostToDevice));
  std::vector<double> pg(S4B_MAX_DEPTH + 2);
  for (int d = 0; d < S4B_MAX_DEPTH + 2; ++d) pg[d] = cfg.base / std::pow(1.0 + (double) d, cfg.power);
  dalloc(&d_pgrow_, pg.size());
  S4B_CUDA(cudaMemcpy(d_pgrow_, pg.data(), sizeof(double) * pg.size(), cudaMemcpyHostToDevice));
  RngState rs; std::memset(&rs, 0, sizeof rs);
  rs.key0 = P.key0; rs.key1 = P.key1; rs.stream = S4B_STREAM_BART;
  S4B_CUDA(cudaMalloc(&d_rng_, sizeof(RngState)));
  S4B_CUDA(cudaMemcpy(d_rng_, &rs, sizeof rs, cudaMemcpyHostToDevice));
  S4B_CUDA(cudaMalloc(&d_scale_factor_, sizeof(double)));
  S4B_CUDA(cudaEventCreate(&ev_start_)); S4B_CUDA(cudaEventCreate(&ev_end_));

  setup_persistent();
  if (cfg.is_binary) {
    // latents start at +-1 with zero offset (oracle_bart.c or_bart_create)
    std::vector<double> init((size_t) n_);
    for (long long i = 0; i < n_; ++i) init[(size_t) i] = y[i] > 0.0 ? 1.0 : -1.0;
    S4B_CUDA(cudaMemcpy(d_yresc_, init.data(), sizeof(double) * (size_t) n_, cudaMemcpyHostToDevice));
    S4B_CUDA(cudaMemcpy(d_R_, init.data(), sizeof(double) * (size_t) n_, cudaMemcpyHostToDevice));
  } else {
    set_offset_device(nullptr, true);
  }
  S4B_CUDA(cudaStreamSynchronize(stream_));
}

BartFit::~BartFit()
{
  if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
  if (graph_exec_thin_) cudaGraphExecDestroy(graph_exec_thin_);
  if (ev_start_) cudaEventDestroy(ev_start_);
  if (ev_end_) cudaEventDestroy(ev_end_);
  cudaFree(d_xt_); cudaFree(d_xt_test_); cudaFree(d_R_); cudaFree(d_yresc_); cudaFree(d_y_); cudaFree(d_offset_);
  cudaFree(d_train_out_); cudaFree(d_latent_out_); cudaFree(d_offset_in_); cudaFree(d_test_out_);
  cudaFree(d_partials_); cudaFree(d_minmax_); cudaFree(d_stats_out_); cudaFree(d_desc_); cudaFree(d_ticket_);
  cudaFree(d_trace_len_); cudaFree(d_trees_); cudaFree(d_params_); cudaFree(d_pgrow_); cudaFree(d_rng_); cudaFree(d_scale_factor_);
  cudaFree(d_split_w_); cudaFree(d_wt_); cudaFree(d_ncuts_var_); cudaFree(d_store_); cudaFree(d_store_scale_); cudaFree(d_packs_); cudaFree(d_barrier_); cudaFree(d_partials2_); cudaFree(d_pipe_ring_); cudaFree(d_pipe_flag_); cudaFree(d_pipe_infos_); cudaFree(d_batch_args_); cudaFree(d_pipe_ran_); cudaFree(d_pipe_pos_); cudaFree(d_leaf_partials_); cudaFree(d_leaf_ticket_); cudaFree(d_tables_); cudaFree(d_descs_); cudaFree(d_draws_); cudaFree(d_prof_); cudaFree(d_trace_); cudaFree(d_tape_); cudaFree(d_rec_); cudaFree(d_varcount_);
}

template <int NQ>
static size_t sweep_smem_bytes(int p)
{
  size_t base = ((sizeof(SweepSmem) + 15) / 16) * 16;
  size_t bins = (size_t) (kBinSlots + 1) * kWorkers * sizeof(double2) + (size_t) kWorkers * sizeof(unsigned long long);
  size_t tile = (size_t) p * NQ * kWorkers * sizeof(uint32_t);
  return base + bins + tile;
}

void BartFit::setup_persistent()
{
  // the persistent sweep needs every CTA co-resident and the chain's residuals / predictors on chip
  persistent_nq_ = 0;
  const char* env = getenv("S4B_SWEEP_MODE");
  int dev = 0; S4B_CUDA(cudaGetDevice(&dev));
  int coop = 0; S4B_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  int max_smem = 0; S4B_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (coop) {
    const long long nquad = (n_ + 3) / 4;
    // several chains per GPU (config D): every chain can be confined to a share of the SMs (s4b_bart_config::max_ctas, or
    // S4B_MAX_CTAS), so that the sweep kernels of different chains, launched from their own host threads and streams, are
    // resident side by side and hide each other's barrier / decision latency
    long long cta_cap = num_sms_;
    if (cfg_.max_ctas > 0) cta_cap = std::min<long long>(cta_cap, cfg_.max_ctas);
    if (getenv("S4B_MAX_CTAS") && atoi(getenv("S4B_MAX_CTAS")) > 0) cta_cap = std::min<long long>(cta_cap, atoi(getenv("S4B_MAX_CTAS")));
    auto try_nq = [&](int nq, size_t smem, const void* fn, const void* fn_seq, const void* fn_nosq) -> bool {
      if (smem > (size_t) max_smem || p_ > 511) return false;      // traversal records carry 9 bits of variable index
      for (const void* f : { fn, fn_seq, fn_nosq }) {
        if (s4b_allow_max_dynamic_smem((const void*) f) != cudaSuccess) { cudaGetLastError(); return false; }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, kSweepBlock, smem) != cudaSuccess) { cudaGetLastError(); return false; }
        if (per_sm < 1) return false;
      }
      long long grid = cta_cap;                        // one CTA per SM (or the share of the SMs this chain was given)
      long long need = (nquad + (long long) nq * kWorkers - 1) / ((long long) nq * kWorkers);
      if (need > grid) return false;
      persistent_nq_ = nq; persistent_grid_ = (int) std::max<long long>(1, std::min<long long>(grid, (nquad + kWorkers - 1) / kWorkers)); persistent_smem_ = smem;
      return true;
    };
    const int force_nq = getenv("S4B_FORCE_NQ") ? atoi(getenv("S4B_FORCE_NQ")) : 0;      // tests: a given register variant at any size
#define S4B_SWEEP_FNS(NQ) (const void*) k_sweep<NQ, false>, (const void*) k_sweep<NQ, true>, (const void*) k_sweep<NQ, false, false, false>
    if ((force_nq != 0 && force_nq != 1) || !try_nq(1, sweep_smem_bytes<1>(p_), S4B_SWEEP_FNS(1)))
      if ((force_nq != 0 && force_nq != 2) || !try_nq(2, sweep_smem_bytes<2>(p_), S4B_SWEEP_FNS(2)))
        if ((force_nq != 0 && force_nq != 4) || !try_nq(4, sweep_smem_bytes<4>(p_), S4B_SWEEP_FNS(4)))
          if (force_nq == 0 || force_nq == 6) try_nq(6, sweep_smem_bytes<6>(p_), S4B_SWEEP_FNS(6));
#undef S4B_SWEEP_FNS
    // shards beyond the register file (or when forced, for the tests): residuals and node indices streamed from global memory
    const bool force_stream = getenv("S4B_FORCE_STREAM") != nullptr && atoi(getenv("S4B_FORCE_STREAM")) != 0;
    if (persistent_nq_ == 0 || force_stream) {
      const size_t smem = sweep_smem_bytes<1>(0);
      const long long rounds = (nquad / cta_cap + 1 + kWorkers - 1) / kWorkers;
      bool ok = smem <= (size_t) max_smem && p_ <= 511 && rounds <= 63;      // 8-bit count fields: at most 63 rounds of 4 observations
      for (const void* f : { (const void*) k_sweep<1, false, true>, (const void*) k_sweep<1, true, true>, (const void*) k_sweep<1, false, true, false> }) {
        if (!ok) break;
        if (s4b_allow_max_dynamic_smem((const void*) f) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, f, kSweepBlock, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); ok = false; }
      }
      if (ok) {
        persistent_nq_ = kStreamNq; persistent_smem_ = smem;
        persistent_grid_ = (int) std::max<long long>(1, std::min<long long>(cta_cap, (nquad + kWorkers - 1) / kWorkers));
        S4B_CUDA(cudaMalloc(&d_packs_, sizeof(uint2) * 2 * (size_t) nquad));
        zero_device_sync(d_packs_, sizeof(uint2) * 2 * (size_t) nquad, stream_);
      }
    }
  }
  if (persistent_nq_ > 0) {
    partial_stride_ = 3 * S4B_MAX_SLOTS * persistent_grid_;
    S4B_CUDA(cudaMalloc(&d_partials2_, sizeof(double) * 2 * (size_t) partial_stride_));
    zero_device_sync(d_partials2_, sizeof(double) * 2 * (size_t) partial_stride_, stream_);
    S4B_CUDA(cudaMalloc(&d_barrier_, sizeof(unsigned int)));
    zero_device_sync(d_barrier_, sizeof(unsigned int), stream_);
    // host-side tables (glibc): growth probabilities by depth, their logs, log of small integers
    std::vector<double> tab((size_t) kTabSize, 0.0);
    for (int d = 0; d < 32; ++d) {
      double pg = cfg_.base / std::pow(1.0 + (double) d, cfg_.power);
      tab[(size_t) kTabPg + d] = pg; tab[(size_t) kTabLogPg + d] = std::log(pg); tab[(size_t) kTabLog1mPg + d] = std::log(1.0 - pg);
    }
    for (int i = 1; i < kLogTab; ++i) tab[(size_t) kTabLogInt + i] = std::log((double) i);
    S4B_CUDA(cudaMalloc(&d_descs_, sizeof(StepDesc) * (size_t) T_));
    zero_device_sync(d_descs_, sizeof(StepDesc) * (size_t) T_, stream_);
    S4B_CUDA(cudaMalloc(&d_draws_, sizeof(double2) * 32 * (size_t) T_));
    {
      size_t psmem = ((sizeof(double) * kTabSize + sizeof(BartParams) + sizeof(RngState) + 15) / 16) * 16 + sizeof(PrepSmemWarp) * kPrepWarps;
      S4B_CUDA(s4b_allow_max_dynamic_smem((const void*) k_prepare_sweep));
    }
    S4B_CUDA(cudaMalloc(&d_tables_, sizeof(double) * tab.size()));
    S4B_CUDA(cudaMemcpy(d_tables_, tab.data(), sizeof(double) * tab.size(), cudaMemcpyHostToDevice));
    sweep_mode_ = 2;
    // ---- pipelined kernel: register-resident, unweighted, unsharded chains; the cross table gets what shared memory is left ----
    if (persistent_nq_ != kStreamNq && d_wt_ == nullptr && !sharded() && !(getenv("S4B_PIPE") && atoi(getenv("S4B_PIPE")) == 0)) {
      const void* fn = persistent_nq_ == 1 ? (const void*) k_sweep_pipe<1> : persistent_nq_ == 2 ? (const void*) k_sweep_pipe<2>
                     : persistent_nq_ == 4 ? (const void*) k_sweep_pipe<4> : (const void*) k_sweep_pipe<6>;
      const size_t fixed = ((sizeof(PipeSmem) + 15) / 16) * 16 + (size_t) (kPipeSlots + 1) * kWorkers * sizeof(double) + 2 * (size_t) kWorkers * sizeof(unsigned long long)
                         + (size_t) p_ * persistent_nq_ * kWorkers * sizeof(uint32_t);
      // the cross table (slot of this step) x (cell of the previous step) gets what shared memory is left: one byte counter per entry and thread
      int entries = 0;
      if (fixed < (size_t) max_smem) entries = (int) std::min<size_t>((size_t) kPipeCross, ((size_t) max_smem - fixed) / kWorkers - 1);
      entries &= ~3;
      if (entries >= 4 * kPipeSlots) {
        const int words = entries;            // (member name kept: capacity of the cross table in entries)
        const size_t smem = fixed + (size_t) (entries + 1) * kWorkers;
        int per_sm = 0;
        if (persistent_nq_ == 4) s4b_allow_max_dynamic_smem((const void*) k_sweep_pipe<4, true>);
        if (s4b_allow_max_dynamic_smem((const void*) fn) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kSweepBlock, smem) == cudaSuccess && per_sm >= 1) {
          pipe_count_words_ = words; pipe_smem_ = smem;
          S4B_CUDA(cudaMalloc(&d_pipe_ring_, sizeof(unsigned long long) * 2 * kPipeRing * kPipeAcc));
          zero_device_sync(d_pipe_ring_, sizeof(unsigned long long) * 2 * kPipeRing * kPipeAcc, stream_);
          S4B_CUDA(cudaMalloc(&d_pipe_flag_, sizeof(unsigned int) * 4));
          zero_device_sync(d_pipe_flag_, sizeof(unsigned int) * 4, stream_);
          S4B_CUDA(cudaMalloc(&d_pipe_infos_, sizeof(PipeInfo) * (size_t) T_));
          zero_device_sync(d_pipe_infos_, sizeof(PipeInfo) * (size_t) T_, stream_);
          S4B_CUDA(cudaMalloc(&d_pipe_pos_, sizeof(int) * (2 * kPipeSegments + 1)));
          S4B_CUDA(cudaMalloc(&d_pipe_ran_, sizeof(unsigned long long)));
          zero_device_sync(d_pipe_ran_, sizeof(unsigned long long), stream_);
          pipe_enabled_ = true;
        } else cudaGetLastError();
      }
    }
  }
  if (d_wt_ != nullptr && persistent_nq_ == 0)
    throw std::invalid_argument("weighted fit: the sweep kernel does not fit this many rows on one GPU (shard the chain by rows)");
  if (env) set_sweep_mode(atoi(env));
  if (getenv("S4B_OVERLAP_WALK")) overlap_walk_ = atoi(getenv("S4B_OVERLAP_WALK"));
}

void BartFit::set_sweep_mode(int m)
{
  if (m == 2 && persistent_nq_ == 0) throw std::invalid_argument("persistent sweep kernel does not fit this problem (n, p) on this GPU");
  if (m < 0 || m > 2) throw std::invalid_argument("sweep mode must be 0, 1 or 2");
  if (m != 2 && d_wt_ != nullptr) throw std::invalid_argument("weighted fits run the persistent sweep kernel only (sweep mode 2)");
  sweep_mode_ = m; use_graph_ = m != 0;
}

void BartFit::launch_persistent_sweep(bool last_thin)
{
  BartDev dv = dev();
  dv.partials = d_partials2_;
  S4B_CUDA(cudaMemsetAsync(d_barrier_, 0, sizeof(unsigned int), stream_));
  unsigned int* bar = d_barrier_;
  int stride = partial_stride_;
  const double* tabs = d_tables_;
  // replay / record keep strict program order inside the kernel; otherwise proposals and decision draws are produced up front
  const StepDesc* descs = sequential_rng_ ? nullptr : d_descs_;
  const double2* draws = sequential_rng_ ? nullptr : d_draws_;
  // Production runs (no parity trace, no replay, no cycle counters) of chains the pipelined kernel supports split the sweep
  // into segments, decided on the device from k_prepare_sweep's per-step flags: the pipelined kernel takes every run of
  // consecutive steps that fit it, the synchronous kernel the single steps in between (and whatever is left after
  // kPipeSegments rounds).  Launches with nothing to do return at once.
  const bool pipe = pipe_enabled_ && !sequential_rng_ && trace_cap_ == 0 && !profile_on_;
  PipeInfo* infos = pipe ? static_cast<PipeInfo*>(d_pipe_infos_) : nullptr;
  if (!sequential_rng_) {
    size_t psmem = ((sizeof(double) * kTabSize + sizeof(BartParams) + sizeof(RngState) + 15) / 16) * 16 + sizeof(PrepSmemWarp) * kPrepWarps;
    if (pipe) S4B_CUDA(cudaMemsetAsync(d_pipe_pos_, 0, sizeof(int) * (2 * kPipeSegments + 1), stream_));
    k_prepare_sweep<<<(T_ + kPrepWarps - 1) / kPrepWarps, kPrepWarps * 32, psmem, stream_>>>(dv, d_descs_, d_draws_, d_tables_, infos, d_pipe_flag_,
                                                                                         kPipeCells);
  }
  int overlap = overlap_walk_;
  ShardDev sh = shard_dev();
  // the sums of squares are accumulated only when the parity trace (which reports the individual log-likelihoods) is on
  // (weighted fits use the two-value bins for sum w r and sum w)
  const bool sq = trace_cap_ > 0 || sequential_rng_ || d_wt_ != nullptr;
  const void* fn;
#define S4B_PICK(NQ) (sequential_rng_ ? (const void*) k_sweep<NQ, true> : (sq ? (const void*) k_sweep<NQ, false> : (const void*) k_sweep<NQ, false, false, false>))
  if (persistent_nq_ == kStreamNq) fn = sequential_rng_ ? (const void*) k_sweep<1, true, true> : (sq ? (const void*) k_sweep<1, false, true> : (const void*) k_sweep<1, false, true, false>);
  else if (persistent_nq_ == 1) fn = S4B_PICK(1);
  else if (persistent_nq_ == 2) fn = S4B_PICK(2);
  else if (persistent_nq_ == 4) fn = S4B_PICK(4);
  else fn = S4B_PICK(6);
#undef S4B_PICK
  if (!pipe) {
    const int* pos_in = nullptr; int* pos_out = nullptr; int max_steps = T_;
    void* args[] = { &dv, &bar, &stride, &tabs, &descs, &draws, &overlap, &sh, &pos_in, &pos_out, &max_steps };
    S4B_CUDA(cudaLaunchCooperativeKernel(fn, dim3(persistent_grid_), dim3(kSweepBlock), args, persistent_smem_, stream_));
  } else {
    static const bool pipe_prof = getenv("S4B_PIPE_PROF") != nullptr;
    const void* pfn = persistent_nq_ == 1 ? (const void*) k_sweep_pipe<1> : persistent_nq_ == 2 ? (const void*) k_sweep_pipe<2>
                    : persistent_nq_ == 4 ? (pipe_prof ? (const void*) k_sweep_pipe<4, true> : (const void*) k_sweep_pipe<4>) : (const void*) k_sweep_pipe<6>;
    unsigned long long* ring = d_pipe_ring_; int words = pipe_count_words_;
    const StepDesc* pdescs = d_descs_; const PipeInfo* pinfos = infos; const double2* pdraws = d_draws_; unsigned long long* ran = d_pipe_ran_;
    unsigned long long* pprof = pipe_prof ? d_prof_ : nullptr;
    for (int k = 0; k < kPipeSegments; ++k) {
      const int* pin = d_pipe_pos_ + 2 * k; int* pmid = d_pipe_pos_ + 2 * k + 1; int* pout = d_pipe_pos_ + 2 * k + 2;
      int parity = (int) (pipe_launches_++ & 1);
      void* pargs[] = { &dv, &ring, &parity, &pdescs, &pinfos, &pdraws, &pin, &pmid, &words, &ran, &pprof };
      S4B_CUDA(cudaLaunchCooperativeKernel(pfn, dim3(persistent_grid_), dim3(kSweepBlock), pargs, pipe_smem_, stream_));
      if (k > 0) S4B_CUDA(cudaMemsetAsync(d_barrier_, 0, sizeof(unsigned int), stream_));
      const int* sin = pmid; int max_steps = k + 1 < kPipeSegments ? 1 : T_;
      void* args[] = { &dv, &bar, &stride, &tabs, &descs, &draws, &overlap, &sh, &sin, &pout, &max_steps };
      S4B_CUDA(cudaLaunchCooperativeKernel(fn, dim3(persistent_grid_), dim3(kSweepBlock), args, persistent_smem_, stream_));
    }
    ++pipe_sweeps_;
  }
  k_finish_sweep<<<grid_ew_, kBlock, 0, stream_>>>(dv, last_thin ? d_train_out_ : nullptr, d_latent_out_, add_offset_ ? 1 : 0,
                                                   (last_thin && test_aliases_train_) ? d_test_out_ : nullptr);
  k_bump_epoch_clear_update<<<1, 32, 0, stream_>>>(dv, cfg_.is_binary ? 1 : 0);
  S4B_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// Several chains, one launch (config D, grid.y = chain).  Every fit of the batch was created with the same shape
// class (same kernel instantiation and shared-memory size) and a share of the SMs (s4b_bart_config::max_ctas) such that all the
// grids are co-resident; they live on one stream.  Per sweep: every chain's k_prepare_sweep, ONE cooperative k_sweep_batch over
// (CTAs per chain) x (chains), every chain's epilogue.  The chains' results are those of running them one by one (same kernels'
// code, same draws): tests/test_bart_gpu.py::test_chains_batched_in_one_launch.
// ---------------------------------------------------------------------------------------
void BartFit::run_sweeps_batched(BartFit* const* fits, int count)
{
  if (count < 1) return;
  for (int c = 0; c < count; ++c) if (fits[c]->stream_ != fits[0]->stream_) throw std::invalid_argument("batched sweep: the fits must share the stream");
  for (int c = 0; c < count; ++c) fits[c]->tree_step_ms(false);
  S4B_CUDA(cudaEventRecord(fits[0]->ev_start_, fits[0]->stream_));
  batched_sweeps_on(fits, count, fits[0]->stream_);
  S4B_CUDA(cudaEventRecord(fits[0]->ev_end_, fits[0]->stream_));
  fits[0]->ev_pending_ = true;
  for (int c = 0; c < count; ++c) fits[c]->after_batched_sweeps();
}

// what follows the sweep kernels of one runSamplerWithResults step, on the fit's own stream (run_sweeps() does the same)
void BartFit::after_batched_sweeps()
{
  draw_k();
  num_tree_steps_ += (long long) cfg_.thin * T_;
  if (nt_ > 0 && !test_aliases_train_) test_fits_device(d_xt_test_, nt_, npad_t_, nullptr, d_test_out_);
  if (keep_trees_active_) snapshot_trees();
}

// the kernels of the batched step on stream `st` (the fits' own streams are not touched: the caller orders them against `st`)
void BartFit::batched_sweeps_on(BartFit* const* fits, int count, cudaStream_t st)
{
  BartFit& f0 = *fits[0];
  long long ctas = 0;
  for (int c = 0; c < count; ++c) {
    BartFit& f = *fits[c];
    if (f.sweep_mode_ != 2 || f.persistent_nq_ == 0) throw std::invalid_argument("batched sweep: every fit needs the persistent sweep kernel");
    if (f.sequential_rng_ || f.trace_cap_ > 0 || f.profile_on_ || f.d_wt_ != nullptr || f.sharded())
      throw std::invalid_argument("batched sweep: traced, replayed, profiled, weighted and sharded fits run on their own");
    if (f.persistent_nq_ != f0.persistent_nq_ || f.persistent_smem_ != f0.persistent_smem_ || f.persistent_grid_ != f0.persistent_grid_ ||
        f.cfg_.thin != f0.cfg_.thin)
      throw std::invalid_argument("batched sweep: the fits must share the kernel layout (rows per thread, predictors), the CTA count and `thin`");
    ctas += f.persistent_grid_;
  }
  if (ctas > f0.num_sms_) throw std::invalid_argument("batched sweep: the chains' CTAs exceed the SMs (create the fits with max_ctas = SMs / chains)");
  if (f0.d_batch_args_ == nullptr || f0.batch_cap_ < count) {
    cudaFree(f0.d_batch_args_);
    S4B_CUDA(cudaMalloc(&f0.d_batch_args_, sizeof(SweepBatchArgs) * (size_t) count));
    f0.batch_cap_ = count;
  }
  const void* fn = f0.persistent_nq_ == kStreamNq ? (const void*) k_sweep_batch<1, true> : f0.persistent_nq_ == 1 ? (const void*) k_sweep_batch<1>
                 : f0.persistent_nq_ == 2 ? (const void*) k_sweep_batch<2> : f0.persistent_nq_ == 4 ? (const void*) k_sweep_batch<4> : (const void*) k_sweep_batch<6>;
  S4B_CUDA(s4b_allow_max_dynamic_smem(fn));
  std::vector<SweepBatchArgs> args((size_t) count);
  const size_t psmem = ((sizeof(double) * kTabSize + sizeof(BartParams) + sizeof(RngState) + 15) / 16) * 16 + sizeof(PrepSmemWarp) * kPrepWarps;
  for (int k = 0; k < f0.cfg_.thin; ++k) {
    const bool last = (k + 1) == f0.cfg_.thin;
    for (int c = 0; c < count; ++c) {
      BartFit& f = *fits[c];
      if (k > 0) f.draw_k_on(st);
      BartDev dv = f.dev();
      dv.partials = f.d_partials2_;
      S4B_CUDA(cudaMemsetAsync(f.d_barrier_, 0, sizeof(unsigned int), st));
      k_prepare_sweep<<<(f.T_ + kPrepWarps - 1) / kPrepWarps, kPrepWarps * 32, psmem, st>>>(dv, f.d_descs_, f.d_draws_, f.d_tables_, nullptr, f.d_pipe_flag_, kPipeCells);
      SweepBatchArgs& a = args[(size_t) c];
      std::memset(&a, 0, sizeof a);
      a.dv = dv; a.barrier_counter = f.d_barrier_; a.partial_stride = f.partial_stride_; a.tables = f.d_tables_; a.descs = f.d_descs_; a.draws = f.d_draws_;
      a.overlap_walk = f.overlap_walk_; a.sh = f.shard_dev(); a.max_steps = f.T_;
    }
    S4B_CUDA(cudaMemcpyAsync(f0.d_batch_args_, args.data(), sizeof(SweepBatchArgs) * (size_t) count, cudaMemcpyHostToDevice, st));
    S4B_CUDA(cudaStreamSynchronize(st));      // (the pageable staging copy has left `args` before it is rewritten)
    const SweepBatchArgs* d_args = static_cast<const SweepBatchArgs*>(f0.d_batch_args_);
    void* kargs[] = { &d_args };
    S4B_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned) f0.persistent_grid_, (unsigned) count), dim3(kSweepBlock), kargs, f0.persistent_smem_, st));
    for (int c = 0; c < count; ++c) {
      BartFit& f = *fits[c];
      BartDev dv = f.dev();
      dv.partials = f.d_partials2_;
      k_finish_sweep<<<f.grid_ew_, kBlock, 0, st>>>(dv, last ? f.d_train_out_ : nullptr, f.d_latent_out_, f.add_offset_ ? 1 : 0,
                                                    (last && f.test_aliases_train_) ? f.d_test_out_ : nullptr);
      k_bump_epoch_clear_update<<<1, 32, 0, st>>>(dv, f.cfg_.is_binary ? 1 : 0);
    }
    S4B_CUDA(cudaGetLastError());
  }
}

void BartFit::pipe_reasons(unsigned int* out4)
{
  for (int i = 0; i < 4; ++i) out4[i] = 0;
  if (d_pipe_flag_ == nullptr) return;
  S4B_CUDA(cudaStreamSynchronize(stream_));
  S4B_CUDA(cudaMemcpy(out4, d_pipe_flag_, sizeof(unsigned int) * 4, cudaMemcpyDeviceToHost));
}

long long BartFit::pipe_sweeps_done()
{
  if (d_pipe_ran_ == nullptr) return 0;
  unsigned long long k = 0;
  S4B_CUDA(cudaStreamSynchronize(stream_));
  S4B_CUDA(cudaMemcpy(&k, d_pipe_ran_, sizeof k, cudaMemcpyDeviceToHost));
  return (long long) k;
}

void BartFit::bin_matrix(const double* x, long long rows, long long rows_pad, std::vector<uint8_t>& out) const
{
  out.assign((size_t) p_ * rows_pad, 0);
  for (int j = 0; j < p_; ++j) {
    const double* c = cuts_.data() + (size_t) j * cfg_.n_cuts;
    const double* col = x + (size_t) j * rows;
    uint8_t* dst = out.data() + (size_t) j * rows_pad;
    for (long long i = 0; i < rows; ++i) dst[i] = (uint8_t) (std::lower_bound(c, c + cfg_.n_cuts, col[i]) - c);
  }
}

ShardDev BartFit::shard_dev() const
{
  ShardDev sh; std::memset(&sh, 0, sizeof sh); sh.world = 1;
  if (sharded()) sh = shard_->dev();
  return sh;
}

BartDev BartFit::dev() const
{
  BartDev d;
  d.n = n_; d.npad = npad_; d.obs_offset = shard_ != nullptr ? shard_->obs_offset() : 0; d.xt = d_xt_; d.R = d_R_; d.yresc = d_yresc_; d.y = d_y_; d.offset = d_offset_;
  d.desc = d_desc_; d.trees = d_trees_; d.params = d_params_; d.pgrow = d_pgrow_; d.rng = d_rng_;
  d.partials = d_partials_; d.ticket = d_ticket_; d.packs = d_packs_; d.wt = d_wt_; d.trace = d_trace_; d.trace_cap = trace_cap_; d.trace_len = d_trace_len_;
  d.stats_out = d_stats_out_; d.prof = profile_on_ ? d_prof_ : nullptr;
  return d;
}

void BartFit::invalidate_graph()
{
  if (graph_exec_) { cudaGraphExecDestroy(graph_exec_); graph_exec_ = nullptr; }
  if (graph_exec_thin_) { cudaGraphExecDestroy(graph_exec_thin_); graph_exec_thin_ = nullptr; }
}

void BartFit::set_trace(size_t cap_records)
{
  S4B_CUDA(cudaStreamSynchronize(stream_));
  cudaFree(d_trace_); d_trace_ = nullptr; trace_cap_ = cap_records;
  if (cap_records) { S4B_CUDA(cudaMalloc(&d_trace_, sizeof(double) * S4B_TRACE_LEN * cap_records)); zero_device_sync(d_trace_, sizeof(double) * S4B_TRACE_LEN * cap_records, stream_); }
  zero_device_sync(d_trace_len_, sizeof(unsigned long long), stream_);
  invalidate_graph();
}

size_t BartFit::get_trace(double* out, size_t cap_records)
{
  S4B_CUDA(cudaStreamSynchronize(stream_));
  unsigned long long k = 0;
  S4B_CUDA(cudaMemcpy(&k, d_trace_len_, sizeof k, cudaMemcpyDeviceToHost));
  size_t m = std::min<size_t>({ (size_t) k, trace_cap_, cap_records });
  if (m && out) S4B_CUDA(cudaMemcpy(out, d_trace_, sizeof(double) * S4B_TRACE_LEN * m, cudaMemcpyDeviceToHost));
  return (size_t) k;
}

void BartFit::set_tape(const double* tape, size_t len)
{
  S4B_CUDA(cudaStreamSynchronize(stream_));
  cudaFree(d_tape_); d_tape_ = nullptr;
  RngState rs; S4B_CUDA(cudaMemcpy(&rs, d_rng_, sizeof rs, cudaMemcpyDeviceToHost));
  if (tape && len) {
    S4B_CUDA(cudaMalloc(&d_tape_, sizeof(double) * len));
    S4B_CUDA(cudaMemcpy(d_tape_, tape, sizeof(double) * len, cudaMemcpyHostToDevice));
  }
  rs.tape = d_tape_; rs.tape_len = d_tape_ ? len : 0; rs.tape_pos = 0; rs.tape_underrun = 0;
  tape_set_ = d_tape_ != nullptr; sequential_rng_ = tape_set_ || rec_set_;
  S4B_CUDA(cudaMemcpy(d_rng_, &rs, sizeof rs, cudaMemcpyHostToDevice));
}

void BartFit::set_record(size_t cap)
{
  S4B_CUDA(cudaStreamSynchronize(stream_));
  cudaFree(d_rec_); d_rec_ = nullptr;
  RngState rs; S4B_CUDA(cudaMemcpy(&rs, d_rng_, sizeof rs, cudaMemcpyDeviceToHost));
  if (cap) S4B_CUDA(cudaMalloc(&d_rec_, sizeof(double) * cap));
  rs.rec = d_rec_; rs.rec_cap = cap; rs.rec_len = 0;
  rec_set_ = d_rec_ != nullptr; sequential_rng_ = tape_set_ || rec_set_;
  S4B_CUDA(cudaMemcpy(d_rng_, &rs, sizeof rs, cudaMemcpyHostToDevice));
}

size_t BartFit::get_record(double* out, size_t cap)
{
  S4B_CUDA(cudaStreamSynchronize(stream_));
  RngState rs; S4B_CUDA(cudaMemcpy(&rs, d_rng_, sizeof rs, cudaMemcpyDeviceToHost));
  size_t m = std::min<size_t>({ (size_t) rs.rec_len, (size_t) rs.rec_cap, cap });
  if (m && out) S4B_CUDA(cudaMemcpy(out, d_rec_, sizeof(double) * m, cudaMemcpyDeviceToHost));
  return (size_t) rs.rec_len;
}

unsigned long long BartFit::rng_counter()
{
  S4B_CUDA(cudaStreamSynchronize(stream_));
  RngState rs; S4B_CUDA(cudaMemcpy(&rs, d_rng_, sizeof rs, cudaMemcpyDeviceToHost));
  return rs.counter;
}

void BartFit::check_error_flag()
{
  BartParams P = params();
  if (P.error_flag & 2u) throw std::runtime_error("s4b: RNG tape underrun in replay mode");
  if (P.error_flag & 4u) throw std::runtime_error("s4b: a peer rank stopped answering during a sharded sweep");
  if (P.error_flag & 8u) throw std::runtime_error("s4b: a CTA never reached the grid barrier of the sweep kernel");
  if (P.error_flag) throw std::runtime_error("s4b: device error flag set");
}

BartParams BartFit::params()
{
  S4B_CUDA(cudaStreamSynchronize(stream_));
  BartParams P; S4B_CUDA(cudaMemcpy(&P, d_params_, sizeof P, cudaMemcpyDeviceToHost));
  return P;
}

void BartFit::set_offset_device(const double* d_offset, bool update_scale)
{
  BartDev dv = dev();
  if (cfg_.is_binary) {
    k_apply_offset_binary<<<grid_ew_, kBlock, 0, stream_>>>(dv, d_offset, d_latent_out_);
    k_bump_epoch_clear_update<<<1, 32, 0, stream_>>>(dv, 1);
    S4B_CUDA(cudaGetLastError());
    return;
  }
  bool first = !scale_initialised_;
  if (update_scale || first) {
    k_minmax<<<grid_ew_, kBlock, 0, stream_>>>(dv, d_offset, d_minmax_);
    if (sharded()) {
      // the response range is a property of the whole data set: max-reduce (-min, max) over the ranks
      double* packed = d_minmax_ + 2 * (size_t) grid_ew_;
      k_minmax_pack<<<1, 32, 0, stream_>>>(d_minmax_, grid_ew_, packed);
      shard_->allreduce(packed, 2, kOpMax, stream_);
      k_minmax_unpack<<<1, 32, 0, stream_>>>(packed, packed + 2);
      k_update_scale<<<1, 32, 0, stream_>>>(dv, packed + 2, 1, d_scale_factor_);
    } else
      k_update_scale<<<1, 32, 0, stream_>>>(dv, d_minmax_, grid_ew_, d_scale_factor_);
    if (!first) k_scale_leaves<<<T_, 128, 0, stream_>>>(dv, d_scale_factor_);
    k_apply_offset<<<grid_ew_, kBlock, 0, stream_>>>(dv, d_offset, d_scale_factor_);
    scale_initialised_ = true;
  } else {
    k_apply_offset<<<grid_ew_, kBlock, 0, stream_>>>(dv, d_offset, nullptr);
  }
  S4B_CUDA(cudaGetLastError());
}

void BartFit::set_offset_host(const double* offset, bool update_scale)
{
  if (offset) S4B_CUDA(cudaMemcpyAsync(d_offset_in_, offset, sizeof(double) * (size_t) n_, cudaMemcpyHostToDevice, stream_));
  set_offset_device(offset ? d_offset_in_ : nullptr, update_scale);
}

void BartFit::set_response_host(const double* y)
{
  // the new response replaces y; re-applying the current offset refreshes the rescaled response and the residuals
  // (binary: the latents are redrawn for the new signs), exactly what a setOffset with unchanged values does
  S4B_CUDA(cudaMemcpyAsync(d_y_, y, sizeof(double) * (size_t) n_, cudaMemcpyHostToDevice, stream_));
  S4B_CUDA(cudaMemcpyAsync(d_offset_in_, d_offset_, sizeof(double) * (size_t) n_, cudaMemcpyDeviceToDevice, stream_));
  set_offset_device(d_offset_in_, false);
  S4B_CUDA(cudaStreamSynchronize(stream_));
}

void BartFit::set_sigma(double sigma)
{
  k_set_sigma<<<1, 32, 0, stream_>>>(dev(), sigma);
  S4B_CUDA(cudaGetLastError());
}

void BartFit::sample_trees_from_prior()
{
  BartDev dv = dev();
  for (int t = 0; t < T_; ++t) {
    k_prior_tree<<<1, 32, 0, stream_>>>(dv, t);
    k_tree_step<<<grid_, kBlock, 0, stream_>>>(dv, kModeUpdateOnly, 0, shard_dev());
  }
  k_bump_epoch_clear_update<<<1, 32, 0, stream_>>>(dv, 0);
  S4B_CUDA(cudaGetLastError());
}

void BartFit::launch_sweep_kernels(bool last_thin)
{
  BartDev dv = dev();
  k_propose_first<<<1, kBlock, 0, stream_>>>(dv);
  for (int t = 0; t < T_; ++t) k_tree_step<<<grid_, kBlock, 0, stream_>>>(dv, kModeStep, t + 1 < T_ ? 1 : 0, shard_dev());
  k_finish_sweep<<<grid_ew_, kBlock, 0, stream_>>>(dv, last_thin ? d_train_out_ : nullptr, d_latent_out_, add_offset_ ? 1 : 0,
                                                   (last_thin && test_aliases_train_) ? d_test_out_ : nullptr);
  k_bump_epoch_clear_update<<<1, 32, 0, stream_>>>(dv, cfg_.is_binary ? 1 : 0);
}

void BartFit::run_sweeps()
{
  // one runSamplerWithResults(fit, 0, results[numSamples = 1]): `thin` sweeps, the last one kept
  const bool use_graph = use_graph_;
  tree_step_ms(false);                       // fold in the previous sweep's event pair
  S4B_CUDA(cudaEventRecord(ev_start_, stream_));
  for (int k = 0; k < cfg_.thin; ++k) {
    bool last = (k + 1) == cfg_.thin;
    if (k > 0) draw_k();                   // k of the previous (thinned) sweep; the last one is drawn after the loop
    if (sweep_mode_ == 2) { launch_persistent_sweep(last); continue; }
    if (!use_graph) { launch_sweep_kernels(last); S4B_CUDA(cudaGetLastError()); continue; }
    cudaGraphExec_t& ge = last ? graph_exec_ : graph_exec_thin_;
    if (!ge) {
      cudaGraph_t graph;
      S4B_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
      launch_sweep_kernels(last);
      S4B_CUDA(cudaStreamEndCapture(stream_, &graph));
      S4B_CUDA(cudaGraphInstantiate(&ge, graph, 0));
      S4B_CUDA(cudaGraphDestroy(graph));
    }
    S4B_CUDA(cudaGraphLaunch(ge, stream_));
  }
  draw_k();
  S4B_CUDA(cudaEventRecord(ev_end_, stream_));
  ev_pending_ = true;
  num_tree_steps_ += (long long) cfg_.thin * T_;
  if (nt_ > 0 && !test_aliases_train_) test_fits_device(d_xt_test_, nt_, npad_t_, nullptr, d_test_out_);
  // keepTrees (no-op unless a store was requested); the reference switches dbarts' keepTrees off for warm-up runs
  // (init.cpp:737-744), so warm-up sweeps never take slots of the store
  if (keep_trees_active_) snapshot_trees();
}

void BartFit::draw_k() { draw_k_on(stream_); }
void BartFit::draw_k_on(cudaStream_t st)
{
  if (!(cfg_.k_df > 0.0)) return;
  k_draw_k<<<1, 256, 0, st>>>(dev());
  S4B_CUDA(cudaGetLastError());
}

double BartFit::current_k()
{
  double k = 0.0;
  S4B_CUDA(cudaMemcpyAsync(&k, d_k_ptr(), sizeof(double), cudaMemcpyDeviceToHost, stream_));
  S4B_CUDA(cudaStreamSynchronize(stream_));
  return k;
}

void BartFit::set_keep_trees(long long capacity)
{
  S4B_CUDA(cudaStreamSynchronize(stream_));
  cudaFree(d_store_); cudaFree(d_store_scale_); d_store_ = nullptr; d_store_scale_ = nullptr;
  store_cap_ = capacity > 0 ? capacity : 0; store_len_ = 0;
  if (store_cap_ > 0) {
    S4B_CUDA(cudaMalloc(&d_store_, sizeof(DTree) * (size_t) T_ * (size_t) store_cap_));
    S4B_CUDA(cudaMalloc(&d_store_scale_, sizeof(double) * 2 * (size_t) store_cap_));
  }
}

void BartFit::snapshot_trees()
{
  if (store_cap_ == 0) return;
  if (store_len_ >= store_cap_) throw std::runtime_error("keepTrees: the tree store is full (gpubart_set_keep_trees capacity)");
  k_snapshot_trees<<<T_, 128, 0, stream_>>>(dev(), d_store_ + (size_t) store_len_ * (size_t) T_, d_store_scale_ + 2 * (size_t) store_len_);
  S4B_CUDA(cudaGetLastError());
  ++store_len_;
}

void BartFit::predict_stored(const double* x_test, long long rows, const double* test_offset, long long first, long long count, double* out)
{
  if (first < 0 || count < 0 || first + count > store_len_) throw std::invalid_argument("stored sample range out of bounds");
  if (rows <= 0 || count == 0) return;
  long long rows_pad = (rows + 15) / 16 * 16;
  std::vector<uint8_t> xtt; bin_matrix(x_test, rows, rows_pad, xtt);
  uint8_t* d_x; double* d_o; double* d_off = nullptr;
  S4B_CUDA(cudaMalloc(&d_x, xtt.size())); S4B_CUDA(cudaMalloc(&d_o, sizeof(double) * (size_t) rows));
  S4B_CUDA(cudaMemcpyAsync(d_x, xtt.data(), xtt.size(), cudaMemcpyHostToDevice, stream_));
  if (test_offset) { S4B_CUDA(cudaMalloc(&d_off, sizeof(double) * (size_t) rows)); S4B_CUDA(cudaMemcpyAsync(d_off, test_offset, sizeof(double) * (size_t) rows, cudaMemcpyHostToDevice, stream_)); }
  for (long long s = 0; s < count; ++s) {
    test_fits_device(d_x, rows, rows_pad, d_off, d_o, d_store_ + (size_t) (first + s) * (size_t) T_, d_store_scale_ + 2 * (size_t) (first + s));
    S4B_CUDA(cudaMemcpyAsync(out + (size_t) s * (size_t) rows, d_o, sizeof(double) * (size_t) rows, cudaMemcpyDeviceToHost, stream_));
  }
  S4B_CUDA(cudaStreamSynchronize(stream_));
  cudaFree(d_x); cudaFree(d_o); cudaFree(d_off);
}

void BartFit::get_stored_scales(long long first, long long count, double* out2)
{
  if (first < 0 || count < 0 || first + count > store_len_) throw std::invalid_argument("stored sample range out of bounds");
  if (count == 0) return;
  S4B_CUDA(cudaStreamSynchronize(stream_));
  S4B_CUDA(cudaMemcpy(out2, d_store_scale_ + 2 * (size_t) first, sizeof(double) * 2 * (size_t) count, cudaMemcpyDeviceToHost));
}

std::vector<DTree> BartFit::download_stored(long long sample)
{
  if (sample < 0 || sample >= store_len_) throw std::invalid_argument("stored sample index out of bounds");
  S4B_CUDA(cudaStreamSynchronize(stream_));
  std::vector<DTree> trees((size_t) T_);
  S4B_CUDA(cudaMemcpy(trees.data(), d_store_ + (size_t) sample * (size_t) T_, sizeof(DTree) * trees.size(), cudaMemcpyDeviceToHost));
  return trees;
}

long long BartFit::num_stored_nodes(long long sample)
{
  long long c = 0;
  for (const auto& t : download_stored(sample)) c += t.num_nodes;
  return c;
}

void BartFit::get_stored_trees(long long sample, int32_t* tree_no, long long* n_obs, int32_t* var, double* value)
{
  auto trees = download_stored(sample);
  flatten_trees(trees, tree_no, n_obs, var, value);
}

// ---- exported draws: header, cut points, then per draw the response scale and the trees (used nodes only) ----
struct StoredHeader { unsigned long long magic; int p, T, n_cuts, is_binary; long long count; };
static const unsigned long long kStoredMagic = 0x5334425354524545ull;      // "S4BSTREE"

long long BartFit::stored_export_size()
{
  long long bytes = (long long) sizeof(StoredHeader) + (long long) sizeof(double) * (long long) cuts_.size();
  for (long long s = 0; s < store_len_; ++s) {
    bytes += 2 * (long long) sizeof(double);
    for (const auto& t : download_stored(s)) bytes += (long long) sizeof(int32_t) * 2 + (long long) sizeof(DNode) * t.num_nodes;
  }
  return bytes;
}

void BartFit::stored_export(void* out, long long bytes)
{
  if (bytes < stored_export_size()) throw std::invalid_argument("stored_export: buffer too small");
  unsigned char* w = static_cast<unsigned char*>(out);
  StoredHeader h; std::memset(&h, 0, sizeof h);
  h.magic = kStoredMagic; h.p = p_; h.T = T_; h.n_cuts = cfg_.n_cuts; h.is_binary = cfg_.is_binary; h.count = store_len_;
  std::memcpy(w, &h, sizeof h); w += sizeof h;
  std::memcpy(w, cuts_.data(), sizeof(double) * cuts_.size()); w += sizeof(double) * cuts_.size();
  std::vector<double> scales((size_t) 2 * (size_t) std::max<long long>(store_len_, 1));
  if (store_len_ > 0) S4B_CUDA(cudaMemcpy(scales.data(), d_store_scale_, sizeof(double) * 2 * (size_t) store_len_, cudaMemcpyDeviceToHost));
  for (long long s = 0; s < store_len_; ++s) {
    std::memcpy(w, scales.data() + 2 * s, 2 * sizeof(double)); w += 2 * sizeof(double);
    for (const auto& t : download_stored(s)) {
      int32_t hdr[2] = { t.num_nodes, 0 };
      std::memcpy(w, hdr, sizeof hdr); w += sizeof hdr;
      std::memcpy(w, t.nodes, sizeof(DNode) * (size_t) t.num_nodes); w += sizeof(DNode) * (size_t) t.num_nodes;
    }
  }
}

StoredBart::StoredBart(const void* blob, long long bytes, cudaStream_t stream) : stream_(stream)
{
  const unsigned char* r = static_cast<const unsigned char*>(blob);
  const unsigned char* end = r + bytes;
  StoredHeader h;
  if (bytes < (long long) sizeof h) throw std::invalid_argument("stored sampler: blob too short");
  std::memcpy(&h, r, sizeof h); r += sizeof h;
  if (h.magic != kStoredMagic || h.p < 1 || h.T < 1 || h.n_cuts < 1 || h.n_cuts > 255 || h.count < 0) throw std::invalid_argument("stored sampler: not an exported BART state");
  p_ = h.p; T_ = h.T; n_cuts_ = h.n_cuts; is_binary_ = h.is_binary; count_ = h.count;
  const size_t ncut = (size_t) p_ * (size_t) n_cuts_;
  if (r + sizeof(double) * ncut > end) throw std::invalid_argument("stored sampler: truncated blob");
  cuts_.resize(ncut); std::memcpy(cuts_.data(), r, sizeof(double) * ncut); r += sizeof(double) * ncut;
  std::vector<DTree> trees((size_t) T_ * (size_t) std::max<long long>(count_, 1));
  std::memset(trees.data(), 0, sizeof(DTree) * trees.size());
  std::vector<double> scales((size_t) 2 * (size_t) std::max<long long>(count_, 1), 0.0);
  for (long long s = 0; s < count_; ++s) {
    if (r + 2 * sizeof(double) > end) throw std::invalid_argument("stored sampler: truncated blob");
    std::memcpy(scales.data() + 2 * s, r, 2 * sizeof(double)); r += 2 * sizeof(double);
    for (int t = 0; t < T_; ++t) {
      int32_t hdr[2];
      if (r + sizeof hdr > end) throw std::invalid_argument("stored sampler: truncated blob");
      std::memcpy(hdr, r, sizeof hdr); r += sizeof hdr;
      if (hdr[0] < 1 || hdr[0] > S4B_NODE_CAP || r + sizeof(DNode) * (size_t) hdr[0] > end) throw std::invalid_argument("stored sampler: corrupt tree record");
      DTree& d = trees[(size_t) s * (size_t) T_ + (size_t) t];
      d.num_nodes = hdr[0];
      std::memcpy(d.nodes, r, sizeof(DNode) * (size_t) hdr[0]); r += sizeof(DNode) * (size_t) hdr[0];
    }
  }
  S4B_CUDA(cudaMalloc(&d_store_, sizeof(DTree) * trees.size()));
  S4B_CUDA(cudaMemcpy(d_store_, trees.data(), sizeof(DTree) * trees.size(), cudaMemcpyHostToDevice));
  S4B_CUDA(cudaMalloc(&d_scale_, sizeof(double) * scales.size()));
  S4B_CUDA(cudaMemcpy(d_scale_, scales.data(), sizeof(double) * scales.size(), cudaMemcpyHostToDevice));
  scales_ = scales;
  BartParams P; std::memset(&P, 0, sizeof P);
  P.p = p_; P.num_trees = T_; P.n_cuts = n_cuts_; P.is_binary = is_binary_; P.smin = -0.5; P.smax = 0.5; P.srange = 1.0;
  S4B_CUDA(cudaMalloc(&d_params_, sizeof(BartParams)));
  S4B_CUDA(cudaMemcpy(d_params_, &P, sizeof P, cudaMemcpyHostToDevice));
}

StoredBart::~StoredBart() { cudaFree(d_store_); cudaFree(d_scale_); cudaFree(d_params_); }

void StoredBart::get_scales(long long first, long long count, double* out2) const
{
  if (first < 0 || count < 0 || first + count > count_) throw std::invalid_argument("stored sample range out of bounds");
  for (long long i = 0; i < 2 * count; ++i) out2[i] = scales_[(size_t) (2 * first + i)];
}

void StoredBart::predict(const double* x_test, long long rows, const double* test_offset, long long first, long long count, double* out)
{
  if (first < 0 || count < 0 || first + count > count_) throw std::invalid_argument("stored sample range out of bounds");
  if (rows <= 0 || count == 0) return;
  const long long rows_pad = (rows + 15) / 16 * 16;
  std::vector<uint8_t> xtt((size_t) p_ * (size_t) rows_pad, 0);
  for (int j = 0; j < p_; ++j) {
    const double* c = cuts_.data() + (size_t) j * n_cuts_;
    const double* col = x_test + (size_t) j * rows;
    uint8_t* dst = xtt.data() + (size_t) j * rows_pad;
    for (long long i = 0; i < rows; ++i) dst[i] = (uint8_t) (std::lower_bound(c, c + n_cuts_, col[i]) - c);
  }
  uint8_t* d_x; double* d_o; double* d_off = nullptr;
  S4B_CUDA(cudaMalloc(&d_x, xtt.size())); S4B_CUDA(cudaMalloc(&d_o, sizeof(double) * (size_t) rows));
  S4B_CUDA(cudaMemcpyAsync(d_x, xtt.data(), xtt.size(), cudaMemcpyHostToDevice, stream_));
  if (test_offset) { S4B_CUDA(cudaMalloc(&d_off, sizeof(double) * (size_t) rows)); S4B_CUDA(cudaMemcpyAsync(d_off, test_offset, sizeof(double) * (size_t) rows, cudaMemcpyHostToDevice, stream_)); }
  BartDev dv; std::memset(&dv, 0, sizeof dv);
  dv.params = d_params_;
  const size_t smem = (sizeof(uint32_t) + sizeof(double)) * 2048 + (size_t) p_ * kBlock;
  if (smem > 48 * 1024) S4B_CUDA(s4b_allow_max_dynamic_smem((const void*) k_test_fits));
  const int grid = (int) ((rows + kBlock - 1) / kBlock);
  for (long long s = 0; s < count; ++s) {
    k_test_fits<<<grid, kBlock, smem, stream_>>>(dv, d_x, rows, rows_pad, d_off, d_o, is_binary_ ? 0 : 1, p_, d_store_ + (size_t) (first + s) * (size_t) T_,
                                                 d_scale_ + 2 * (size_t) (first + s));
    S4B_CUDA(cudaGetLastError());
    S4B_CUDA(cudaMemcpyAsync(out + (size_t) s * (size_t) rows, d_o, sizeof(double) * (size_t) rows, cudaMemcpyDeviceToHost, stream_));
  }
  S4B_CUDA(cudaStreamSynchronize(stream_));
  cudaFree(d_x); cudaFree(d_o); cudaFree(d_off);
}

void BartFit::get_profile(unsigned long long* out8, bool reset)
{
  S4B_CUDA(cudaStreamSynchronize(stream_));
  S4B_CUDA(cudaMemcpy(out8, d_prof_, sizeof(unsigned long long) * 24, cudaMemcpyDeviceToHost));
  if (reset) zero_device_sync(d_prof_, sizeof(unsigned long long) * 24, stream_);
}

double BartFit::tree_step_ms(bool reset)
{
  if (ev_pending_) {
    S4B_CUDA(cudaEventSynchronize(ev_end_));
    float ms = 0.f; S4B_CUDA(cudaEventElapsedTime(&ms, ev_start_, ev_end_));
    sweep_ms_ += (double) ms;
    ev_pending_ = false;
  }
  double r = sweep_ms_;
  if (reset) sweep_ms_ = 0.0;
  return r;
}

void BartFit::test_fits_device(const uint8_t* d_xt, long long rows, long long rows_pad, const double* d_off, double* d_out, const DTree* trees,
                               const double* scale)
{
  size_t smem = (sizeof(uint32_t) + sizeof(double)) * 2048 + (size_t) p_ * kBlock;
  if (smem > 48 * 1024) S4B_CUDA(s4b_allow_max_dynamic_smem((const void*) k_test_fits));
  int grid = (int) ((rows + kBlock - 1) / kBlock);
  k_test_fits<<<grid, kBlock, smem, stream_>>>(dev(), d_xt, rows, rows_pad, d_off, d_out, cfg_.is_binary ? 0 : 1, p_, trees, scale);
  S4B_CUDA(cudaGetLastError());
}

void BartFit::run(double* train, double* test, uint32_t* varcount, double* sigma)
{
  run_sweeps();
  collect_results(train, test, varcount, sigma);
}

// the results of the last run_sweeps() / run_sweeps_batched(): what runSamplerWithResults hands back (any pointer may be NULL)
void BartFit::collect_results(double* train, double* test, uint32_t* varcount, double* sigma)
{
  if (train) S4B_CUDA(cudaMemcpyAsync(train, d_train_out_, sizeof(double) * (size_t) n_, cudaMemcpyDeviceToHost, stream_));
  if (test && nt_ > 0) S4B_CUDA(cudaMemcpyAsync(test, d_test_out_, sizeof(double) * (size_t) nt_, cudaMemcpyDeviceToHost, stream_));
  if (varcount) {
    S4B_CUDA(cudaMemsetAsync(d_varcount_, 0, sizeof(unsigned int) * (size_t) p_, stream_));
    k_varcount<<<1, 256, 0, stream_>>>(dev(), d_varcount_);
    S4B_CUDA(cudaMemcpyAsync(varcount, d_varcount_, sizeof(unsigned int) * (size_t) p_, cudaMemcpyDeviceToHost, stream_));
  }
  BartParams P = params();   // synchronises
  if (P.error_flag) check_error_flag();
  if (sigma) *sigma = cfg_.is_binary ? 1.0 : P.sigma * P.srange;
}

void BartFit::varcount_device(unsigned int* d_out)
{
  S4B_CUDA(cudaMemsetAsync(d_out, 0, sizeof(unsigned int) * (size_t) p_, stream_));
  k_varcount<<<1, 256, 0, stream_>>>(dev(), d_out);
}

void BartFit::store_latents(double* out)
{
  k_store_latents<<<grid_ew_, kBlock, 0, stream_>>>(dev(), d_latent_out_);
  S4B_CUDA(cudaMemcpyAsync(out, d_latent_out_, sizeof(double) * (size_t) n_, cudaMemcpyDeviceToHost, stream_));
  S4B_CUDA(cudaStreamSynchronize(stream_));
}

void BartFit::get_residual(double* out)
{
  S4B_CUDA(cudaMemcpyAsync(out, d_R_, sizeof(double) * (size_t) n_, cudaMemcpyDeviceToHost, stream_));
  S4B_CUDA(cudaStreamSynchronize(stream_));
}

void BartFit::node_assignment(int tree, long long* out)
{
  if (tree < 0 || tree >= T_) throw std::invalid_argument("tree index out of range");
  long long* d; S4B_CUDA(cudaMalloc(&d, sizeof(long long) * (size_t) n_));
  k_node_assignment<<<grid_ew_, kBlock, 0, stream_>>>(dev(), tree, d);
  S4B_CUDA(cudaMemcpyAsync(out, d, sizeof(long long) * (size_t) n_, cudaMemcpyDeviceToHost, stream_));
  S4B_CUDA(cudaStreamSynchronize(stream_));
  cudaFree(d);
}

int BartFit::leaf_stats(int tree, int max_leaves, long long* heap, long long* count, double* sum, double* sumsq)
{
  if (tree < 0 || tree >= T_) throw std::invalid_argument("tree index out of range");
  launch_leaf_stats(tree);
  if (d_leaf_ticket_ != nullptr && !leaf_generic_) {
    int fits = 1;
    S4B_CUDA(cudaMemcpyAsync(&fits, reinterpret_cast<int*>(d_leaf_ticket_ + 1), sizeof fits, cudaMemcpyDeviceToHost, stream_));
    S4B_CUDA(cudaStreamSynchronize(stream_));
    if (!fits) { leaf_generic_ = true; launch_leaf_stats(tree); leaf_generic_ = false; }      // a tree with more than kLeafSlots bottom nodes
  }
  std::vector<double> st((size_t) 3 * S4B_MAX_SLOTS);
  S4B_CUDA(cudaMemcpyAsync(st.data(), d_stats_out_, sizeof(double) * st.size(), cudaMemcpyDeviceToHost, stream_));
  std::vector<DTree> trees = download_trees();       // (sharded chains: the kernel already exchanged the statistics)
  const DTree& t = trees[(size_t) tree];
  int leaf = 0;
  for (int k = 0; k < t.num_nodes; ++k) if (t.nodes[k].var < 0) {
    if (leaf < max_leaves) {
      // heap index by walking parents
      long long path[64]; int len = 0; int child = k, par = t.nodes[k].parent;
      while (par >= 0) { path[len++] = (child == par + 1) ? 0 : 1; child = par; par = t.nodes[par].parent; }
      long long h = 1; for (int q = len - 1; q >= 0; --q) h = 2 * h + path[q];
      heap[leaf] = h; count[leaf] = (long long) st[(size_t) 3 * leaf]; sum[leaf] = st[(size_t) 3 * leaf + 1]; sumsq[leaf] = st[(size_t) 3 * leaf + 2];
    }
    ++leaf;
  }
  return leaf;
}

int BartFit::tree_num_leaves(int tree)
{
  if (tree < 0 || tree >= T_) throw std::invalid_argument("tree index out of range");
  int nn = 1;
  S4B_CUDA(cudaMemcpyAsync(&nn, &d_trees_[tree].num_nodes, sizeof nn, cudaMemcpyDeviceToHost, stream_));
  S4B_CUDA(cudaStreamSynchronize(stream_));
  return (nn + 1) / 2;                      // a binary tree
}

void BartFit::launch_leaf_stats(int tree, int num_leaves)
{
  // the dedicated one-launch kernels (leaf_stats.cuh) for unweighted, unsharded fits: register bins + software pipelining for trees with
  // at most 4 bottom nodes, shared-memory bins up to kLeafSlots; the latter reports through d_leaf_fits_ when a tree has more bottom
  // nodes than it handles, and leaf_stats() then repeats the pass with the generic per-tree kernels below
  if (d_wt_ == nullptr && !sharded() && !leaf_generic_) {
    using LeafKernel = void (*)(long long, long long, const uint8_t*, const double*, const DTree*, int, double*, unsigned int*, double*, int*);
    static const LeafKernel kernels[kLeafVariants] = { k_leaf_stats, k_leaf_stats_small<2, 4, 1>, k_leaf_stats_small<4, 3, 2>, k_leaf_stats_small<8, 2, 2> };
    if (d_leaf_partials_ == nullptr) {
      leaf_smem_ = ((sizeof(LeafSmem) + 15) / 16) * 16 + (size_t) (kLeafSlots + 1) * kLeafBlock * (sizeof(double2) + sizeof(int));
      int max_grid = 1;
      for (int v = 0; v < kLeafVariants; ++v) {
        S4B_CUDA(s4b_allow_max_dynamic_smem((const void*) kernels[v]));
        // one wave: as many CTAs as are resident at once (registers and the 46 KB of bins decide), every thread >= 4 quads
        int per_sm = 0;
        S4B_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernels[v], kLeafBlock, leaf_smem_));
        per_sm = std::max(1, per_sm);
        leaf_grid_[v] = (int) std::max<long long>(1, std::min<long long>(((n_ + 3) / 4 + 4 * kLeafBlock - 1) / (4 * kLeafBlock), (long long) num_sms_ * per_sm));
        max_grid = std::max(max_grid, leaf_grid_[v]);
      }
      S4B_CUDA(cudaMalloc(&d_leaf_partials_, sizeof(double) * 3 * kLeafSlots * (size_t) max_grid));
      S4B_CUDA(cudaMalloc(&d_leaf_ticket_, sizeof(unsigned int) + sizeof(int)));
      zero_device_sync(d_leaf_ticket_, sizeof(unsigned int) + sizeof(int), stream_);
    }
    int* fits = reinterpret_cast<int*>(d_leaf_ticket_ + 1);
    if (num_leaves < 0) num_leaves = tree_num_leaves(tree);
    // S4B_LEAF_REG_BINS=0: the shared-memory bins for every tree
    const char* rb = getenv("S4B_LEAF_REG_BINS");
    int v = (rb != nullptr && atoi(rb) == 0) ? 0 : num_leaves <= 2 ? 1 : num_leaves <= 4 ? 2 : num_leaves <= kLeafSlots ? 3 : 0;
    if (v == 2 && n_ / ((long long) leaf_grid_[2] * kLeafBlock) + 8 > 65535) v = 0;      // its row counts are 16-bit fields per thread
    kernels[v]<<<leaf_grid_[v], kLeafBlock, leaf_smem_, stream_>>>(n_, npad_, d_xt_, d_R_, d_trees_, tree, d_leaf_partials_, d_leaf_ticket_, d_stats_out_, fits);
    S4B_CUDA(cudaGetLastError());
    return;
  }
  BartDev dv = dev();
  k_build_stats_desc<<<1, 32, 0, stream_>>>(dv, tree);
  k_tree_step<<<grid_, kBlock, 0, stream_>>>(dv, kModeStatsOnly, 0, shard_dev());
  k_bump_epoch_clear_update<<<1, 32, 0, stream_>>>(dv, 0);
  S4B_CUDA(cudaGetLastError());
}

std::vector<DTree> BartFit::download_trees()
{
  S4B_CUDA(cudaStreamSynchronize(stream_));
  std::vector<DTree> trees((size_t) T_);
  S4B_CUDA(cudaMemcpy(trees.data(), d_trees_, sizeof(DTree) * trees.size(), cudaMemcpyDeviceToHost));
  return trees;
}

std::string BartFit::summary() const
{
  char line[256];
  std::string out = "Running BART with ";
  out += cfg_.is_binary ? "binary" : "numeric"; out += " y\n\n";
  snprintf(line, sizeof line, "number of trees: %d\nnumber of training observations: %lld\nnumber of test observations: %lld\nnumber of explanatory variables: %d\n",
           T_, n_, nt_, p_);
  out += line;
  snprintf(line, sizeof line, "Prior:\n\tk prior fixed to %g\n\tpower and base for tree prior: %g %g\n", cfg_.k, cfg_.power, cfg_.base);
  out += line;
  out += "\ttree split probabilities: ";
  for (int j = 0; j < p_; ++j) {
    snprintf(line, sizeof line, "%g%s", split_probs_.empty() ? 1.0 / (double) p_ : split_probs_[(size_t) j], j + 1 < p_ ? ", " : "\n");
    out += line;
  }
  snprintf(line, sizeof line, "\tuse quantiles for rule cut points: %s\n\tproposal probabilities: birth/death %.2f, swap %.2f, change %.2f; birth %.2f\n",
           cfg_.use_quantiles != 0 ? "true" : "false", cfg_.birth_death_prob, cfg_.swap_prob, cfg_.change_prob, cfg_.birth_prob);
  out += line;
  snprintf(line, sizeof line, "Cutoff rules c in x<=c vs x>c\nnumber of cuts: %d per predictor (%s)\n", cfg_.n_cuts,
           cfg_.use_quantiles != 0 ? "at most; between the distinct sorted training values" : "uniform over the training range");
  out += line;
  return out;
}

long long BartFit::num_nodes()
{
  long long c = 0;
  for (const auto& t : download_trees()) c += t.num_nodes;
  return c;
}

void BartFit::flatten_trees(std::vector<DTree>& trees, int32_t* tree_no, long long* n_obs, int32_t* var, double* value) const
{
  long long pos = 0;
  for (int t = 0; t < T_; ++t) {
    DTree& tr = trees[(size_t) t];
    // the device keeps observation counts for bottom nodes only; internal nodes are the sum of their children
    for (int k = tr.num_nodes - 1; k >= 0; --k) if (tr.nodes[k].var >= 0) tr.nodes[k].n = tr.nodes[k + 1].n + tr.nodes[tr.nodes[k].right].n;
    for (int k = 0; k < tr.num_nodes; ++k, ++pos) {
      const DNode& nd = tr.nodes[k];
      tree_no[pos] = t; n_obs[pos] = nd.n;
      if (nd.var < 0) { var[pos] = -1; value[pos] = nd.mu; }
      else { var[pos] = nd.var; value[pos] = cuts_[(size_t) nd.var * cfg_.n_cuts + nd.cut]; }
    }
  }
}

void BartFit::get_trees(int32_t* tree_no, long long* n_obs, int32_t* var, double* value)
{
  auto trees = download_trees();
  flatten_trees(trees, tree_no, n_obs, var, value);
}

void BartFit::predict(const double* x_test, long long rows, const double* test_offset, double* out)
{
  if (rows <= 0) return;
  long long rows_pad = (rows + 15) / 16 * 16;
  std::vector<uint8_t> xtt; bin_matrix(x_test, rows, rows_pad, xtt);
  uint8_t* d_x; double* d_o; double* d_off = nullptr;
  S4B_CUDA(cudaMalloc(&d_x, xtt.size())); S4B_CUDA(cudaMalloc(&d_o, sizeof(double) * (size_t) rows));
  S4B_CUDA(cudaMemcpyAsync(d_x, xtt.data(), xtt.size(), cudaMemcpyHostToDevice, stream_));
  if (test_offset) { S4B_CUDA(cudaMalloc(&d_off, sizeof(double) * (size_t) rows)); S4B_CUDA(cudaMemcpyAsync(d_off, test_offset, sizeof(double) * (size_t) rows, cudaMemcpyHostToDevice, stream_)); }
  test_fits_device(d_x, rows, rows_pad, d_off, d_o);
  S4B_CUDA(cudaMemcpyAsync(out, d_o, sizeof(double) * (size_t) rows, cudaMemcpyDeviceToHost, stream_));
  S4B_CUDA(cudaStreamSynchronize(stream_));
  cudaFree(d_x); cudaFree(d_o); cudaFree(d_off);
}

}  // namespace s4b

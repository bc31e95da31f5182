// stan4bart_b200/csrc/sweep_kernel.cuh
// Persistent, on-chip BART sweep for sm_100a: the primary path whenever one chain's working set fits the
// register file + shared memory of the GPU (n <= 148 x 480 x 16 ~ 1.1 M observations per GPU).
//
//   * ONE cooperative launch per sweep, one CTA per SM: 15 worker warps + 1 helper warp;
//   * every worker thread owns NQ x 4 observations for the whole sweep: their residuals R live in REGISTERS,
//     their binned predictors in a shared-memory tile (filled once per sweep with coalesced 32-bit loads);
//     inside the 200-tree loop nothing is read from global memory except the tiny per-CTA partials;
//   * per tree step
//       workers : accumulate (n, sum, sum^2) per leaf slot in lane-private shared-memory bins keyed by the cached leaf
//                 index (no atomics, branch-free), fixed-order row reduction, one row of partials, grid barrier; the
//                 walk of the NEXT tree (shared memory only, 4-way ILP) overlaps the controller's decision;
//       helper  : meanwhile fetches the next tree, pre-computes the decision draws of this step and draws
//                 the PROPOSAL OF THE NEXT TREE (keyed RNG substreams make it independent of this step);
//       all CTAs: reduce all partial rows in the same fixed order and run the same warp-parallel Metropolis
//                 decision + leaf draws: bitwise identical everywhere, so no broadcast / second barrier;
//       workers : R += mu_old[leaf] - mu_new[leaf'] from the leaf indices cached during the walk.
//
// Arithmetic follows the reference path restated in oracle/oracle_bart.c (dbarts).
#pragma once

#include "bart_kernels.cuh"
#include "shard.hpp"

namespace s4b {

constexpr int kWorkers = 480;               // 15 worker warps + 1 helper warp = 512 threads => 128 registers per thread
constexpr int kWorkerWarps = kWorkers / 32;
constexpr int kSweepBlock = kWorkers + 32;     // + helper warp
constexpr int kSweepWarps = kSweepBlock / 32;
constexpr int kBinSlots = 8;                   // lane-private shared-memory bin rows per pass (+ 1 trash row)
constexpr int kLogTab = 1024;

// host-computed tables (glibc log, so they are bit-identical to the CPU oracle's calls):
// [0,32) growth probability by depth, [32,64) its log, [64,96) log(1 - p), [96, 96 + kLogTab) log(i)
constexpr int kTabPg = 0, kTabLogPg = 32, kTabLog1mPg = 64, kTabLogInt = 96, kTabSize = 96 + kLogTab;

struct UpdateDesc {                 // how to apply the accepted / rejected step to the residuals
  int32_t mode;                     // 0: structure unchanged; 1: birth accepted; 2: death accepted; 3: change / swap accepted
  int32_t node;
  double delta[S4B_NODE_CAP];       // mode 0: mu_old - mu_new by node index
  double val_old[S4B_NODE_CAP];     // mu before the step, by old node index
  double val_new[S4B_NODE_CAP];     // mu after the step, by new node index
  uint8_t remap[S4B_NODE_CAP];      // old leaf index -> new leaf index (modes 0-2)
};

struct CtlScratch {
  int16_t navail[S4B_NODE_CAP];
  uint8_t flag[S4B_NODE_CAP];
  double ubuf[32], zbuf[32];        // pre-generated draws of the current substream
  double ll[S4B_MAX_SLOTS];         // per statistic slot: integrated log-likelihood,
  double pmean[S4B_MAX_SLOTS + 1];  //   posterior mean of the leaf value,
  double psd[S4B_MAX_SLOTS + 1];    //   posterior sd (entry [nslots] describes the merged parent of a birth / death)
  DNode tmp[S4B_NODE_CAP];
  int32_t pad0;
  int32_t draws_total;              // draws consumed through this scratch since kernel start
  int32_t prof_on, pad1;            // cycle counters below are kept only when profiling is enabled (gpubart_set_profile)
  long long dbg[8];                 // cycle counters (diagnostics)
  long long fine[4];                // decision, first part: slot summaries, staging, ratio, accept
};

// ---------------------------------------------------------------------------------------
// warp-cooperative RNG over one substream: 32 draws are generated at once (lane i -> draw base + i)
// and consumed in order.  Replay (tape) and recording force strictly sequential program order.
// ---------------------------------------------------------------------------------------
// one out-of-line copy: Philox + AS241 are ~500 instructions and the refill sits on many rarely taken paths of the
// controller, whose instruction footprint is what its latency is made of
__device__ __noinline__ void warp_rng_fill(const RngState* g, CtlScratch* cs, unsigned long long step, uint32_t sub, uint32_t base, int lane)
{
  const RngState& r = *g;
  double u, z;
  if (r.tape != nullptr) {
    unsigned long long idx = r.tape_pos + (unsigned long long) lane;
    u = idx < r.tape_len ? r.tape[idx] : 0.5;
    z = idx < r.tape_len ? u : 0.0;
  } else {
    u = keyed_stream_uniform(r.key0, r.key1, r.stream, step, sub, base + (uint32_t) lane);
    z = qnorm_as241(u);
  }
  __syncwarp();
  cs->ubuf[lane] = u; cs->zbuf[lane] = z;
  __syncwarp();
}

struct WarpRng {
  RngState* g;        // shared-memory copy (tape / record bookkeeping), identical in every CTA
  CtlScratch* cs;
  unsigned long long step;
  uint32_t sub, base; // buffer holds draws [base, base + 32) of (step, sub)
  int pos;            // consumed from the buffer (warp-uniform register)
  int lane;
  bool writer;        // CTA 0 appends to the record buffer
  bool recording;

  __device__ void fill()
  {
    warp_rng_fill(g, cs, step, sub, base, lane);
    pos = 0;
  }
  // the buffer was filled by someone else (pre-computed draws): start consuming at its beginning
  __device__ void adopt() { pos = 0; __syncwarp(); }
  __device__ void enter(unsigned long long s, uint32_t sb) { step = s; sub = sb; base = 0; pos = 0; recording = g->rec != nullptr; }
  // account for the consumed prefix (all lanes call)
  __device__ void commit()
  {
    const int k = pos;
    base += (uint32_t) k;
    if (lane == 0) {
      RngState& r = *g;
      if (r.tape != nullptr) { if (r.tape_pos + (unsigned long long) k > r.tape_len) r.tape_underrun = 1; r.tape_pos += (unsigned long long) k; }
      cs->draws_total += k;
    }
    pos = 0;
    __syncwarp();
  }
  __device__ void note(double v)
  {
    RngState& r = *g;
    if (lane == 0) { if (writer && r.rec_len < r.rec_cap) r.rec[r.rec_len] = v; r.rec_len++; }
  }
  __device__ double uniform()
  {
    if (pos >= 32) { commit(); fill(); }
    const double v = cs->ubuf[pos++];
    if (recording) note(v);
    return v;
  }
  __device__ int index(int n) { int k = (int) (uniform() * (double) n); return k >= n ? n - 1 : k; }
  // reserve `count` (<= 32) consecutive normal draws; returns the buffer position of the first one
  __device__ int reserve_normals(int count)
  {
    if (pos + count > 32) { commit(); fill(); }
    const int p = pos;
    pos += count;
    if (recording) for (int i = 0; i < count; ++i) note(cs->zbuf[p + i]);
    return p;
  }
};

// ---------------------------------------------------------------------------------------
// warp helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int w_count_flag(const CtlScratch& cs, int nn, uint8_t bit, int lane)
{
  int c = 0;
  for (int base = 0; base < nn; base += 32) {
    int k = base + lane;
    c += __popc(__ballot_sync(0xffffffffu, k < nn && (cs.flag[k] & bit)));
  }
  return c;
}
__device__ __forceinline__ int nth_set_bit(unsigned m, int n)
{
  for (int i = 0; i < n; ++i) m &= m - 1;
  return __ffs(m) - 1;
}
// index of the pick-th node (in index order) whose flag has `bit`
__device__ __forceinline__ int w_select_flag(const CtlScratch& cs, int nn, uint8_t bit, int pick, int lane)
{
  int found = -1;
  for (int base = 0; base < nn; base += 32) {
    int k = base + lane;
    unsigned m = __ballot_sync(0xffffffffu, k < nn && (cs.flag[k] & bit));
    int c = __popc(m);
    if (found < 0) { if (pick < c) found = base + nth_set_bit(m, pick); else pick -= c; }
  }
  return found;
}
__device__ __forceinline__ double w_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double w_min(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int w_maxi(int v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int w_mini(int v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// branch-free traversal record: var (9 bits) | cut (8) | left (7) | pad (1) | right (7); a leaf points at itself with cut 255
__host__ __device__ inline uint32_t pack_trav2(int var, int cut, int left, int right)
{
  return ((uint32_t) (var & 0x1FF) << 23) | ((uint32_t) (cut & 0xFF) << 15) | ((uint32_t) (left & 0x7F) << 8) | (uint32_t) (right & 0x7F);
}
__device__ __forceinline__ int trav2_var(uint32_t tr) { return (int) (tr >> 23); }
__device__ __forceinline__ int trav2_cut(uint32_t tr) { return (int) ((tr >> 15) & 0xFFu); }

enum : uint8_t { kFLeaf = 1, kFBirthable = 2, kFNog = 4, kFSwappable = 8, kFInternal = 16 };

// per-node attributes, one lane per node
__device__ inline void w_node_attrs(const DTree& t, const BartParams& P, CtlScratch& cs, int num_leaves_for_cap, int lane)
{
  const int nn = t.num_nodes;
  for (int k = lane; k < nn; k += 32) {
    int na = t_num_vars_available(t, P, k);
    cs.navail[k] = (int16_t) na;
    uint8_t f = 0;
    if (t.nodes[k].var < 0) {
      f |= kFLeaf;
      if (na > 0 && t.nodes[k].depth < S4B_MAX_DEPTH && num_leaves_for_cap < S4B_MAX_LEAVES) f |= kFBirthable;
    } else {
      f |= kFInternal;
      bool ll = t.nodes[k + 1].var < 0, rl = t.nodes[t.nodes[k].right].var < 0;
      if (ll && rl) f |= kFNog;
      if (!ll || !rl) f |= kFSwappable;
    }
    cs.flag[k] = f;
  }
  __syncwarp();
}

// fills trav / val / slot of a TravTree (one lane per node); tv.pad receives the depth of the deepest leaf;
// returns the number of leaves
__device__ inline int w_fill_trav(const DTree& t, TravTree& tv, bool with_mu, int lane)
{
  const int nn = t.num_nodes;
  int leaves = 0, maxd = 0;
  for (int base = 0; base < nn; base += 32) {
    int k = base + lane;
    bool leaf = k < nn && t.nodes[k].var < 0;
    unsigned m = __ballot_sync(0xffffffffu, leaf);
    if (k < nn) {
      const DNode& nd = t.nodes[k];
      tv.trav[k] = leaf ? pack_trav2(0, 255, k, k) : pack_trav2(nd.var, nd.cut, k + 1, nd.right);
      if (leaf) { tv.slot[k] = (uint8_t) (leaves + __popc(m & ((1u << lane) - 1u))); if (with_mu) tv.val[k] = nd.mu; maxd = max(maxd, nd.depth); }
      else tv.slot[k] = 255;
    }
    leaves += __popc(m);
  }
  maxd = w_maxi(maxd);
  const int n_int = nn - leaves;
  const bool bitmap = nn <= 32 && n_int <= S4B_BITMAP_INT;
  if (lane == 0) { tv.n = nn; tv.pad = maxd; tv.n_int = bitmap ? n_int : 255; }
  if (bitmap) {
    // rule records in internal-node order and the pattern -> bottom-node table of the bitmap walk
    const unsigned imask = __ballot_sync(0xffffffffu, lane < nn && t.nodes[lane].var >= 0);
    if ((imask >> lane) & 1u) tv.irec[__popc(imask & ((1u << lane) - 1u))] = ((uint32_t) t.nodes[lane].var << 8) | (uint32_t) (t.nodes[lane].cut & 0xFF);
    for (int e = lane; e < (1 << n_int); e += 32) {
      int node = 0;
      while ((imask >> node) & 1u) {
        const int id = __popc(imask & ((1u << node) - 1u));
        node = ((e >> id) & 1) ? node + 1 : (int) t.nodes[node].right;
      }
      tv.table[e] = (uint8_t) node;
    }
  }
  __syncwarp();
  return leaves;
}

// i-th available variable at node `node` (lanes scan variables)
__device__ inline int w_ith_available_var(const DTree& t, const BartParams& P, int node, int ith, int lane)
{
  int found = -1;
  for (int base = 0; base < P.p; base += 32) {
    int j = base + lane;
    const bool avail = j < P.p && t_var_available(t, P, node, j);
    unsigned m = __ballot_sync(0xffffffffu, avail);
    int c = __popc(m);
    if (found < 0) { if (ith < c) found = base + nth_set_bit(m, ith); else ith -= c; }
  }
  return found;
}

// split weights: the available variable whose cumulative weight (index order) first exceeds r; integer arithmetic, so the
// warp scan gives exactly the oracle's sequential result
__device__ inline int w_weighted_var(const DTree& t, const BartParams& P, int node, unsigned long long r, int lane)
{
  int found = -1, last = -1;
  unsigned long long carried = 0ull;
  for (int base = 0; base < P.p; base += 32) {
    int j = base + lane;
    const bool avail = j < P.p && t_var_available(t, P, node, j);
    unsigned long long cum = avail ? (unsigned long long) P.split_w[j] : 0ull;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long up = __shfl_up_sync(0xffffffffu, cum, o); if (lane >= o) cum += up; }
    const unsigned hit = __ballot_sync(0xffffffffu, avail && carried + cum > r);
    const unsigned am = __ballot_sync(0xffffffffu, avail);
    if (found < 0 && hit) found = base + __ffs(hit) - 1;
    if (am) last = base + 31 - __clz(am);
    carried += __shfl_sync(0xffffffffu, cum, 31);
  }
  return found >= 0 ? found : last;
}

// one draw of a splitting variable at `node` (one uniform either way)
__device__ inline int w_draw_var(const DTree& t, const BartParams& P, int node, int navail, WarpRng& rng, int lane)
{
  if (P.split_w == nullptr) return w_ith_available_var(t, P, node, rng.index(navail), lane);
  const unsigned long long W = t_avail_weight(t, P, node);
  return w_weighted_var(t, P, node, weighted_position(rng.uniform(), W), lane);
}

__device__ __forceinline__ double tab_log_int(const double* tab, int i) { return i < kLogTab ? tab[kTabLogInt + i] : log((double) i); }

// log prior of the branch [node, end): lanes over nodes, table look-ups, fixed-order warp sum
__device__ inline double w_branch_log_prior(const DTree& t, const BartParams& P, const double* tab, int node, int end, int lane)
{
  double acc = 0.0;
  for (int base = node; base < end; base += 32) {
    int k = base + lane;
    double term = 0.0;
    if (k < end) {
      int navail = t_num_vars_available(t, P, k);
      int d = t.nodes[k].depth;
      if (t.nodes[k].var < 0) term = navail > 0 ? tab[kTabLog1mPg + d] : 0.0;
      else {
        int lo, hi; t_split_interval(t, s4b_ncuts(P, t.nodes[k].var), k, t.nodes[k].var, lo, hi);
        const double lvar = P.split_w == nullptr ? -tab_log_int(tab, navail) : log((double) P.split_w[t.nodes[k].var] / (double) t_avail_weight(t, P, k));
        term = tab[kTabLogPg + d] + lvar - tab_log_int(tab, hi - lo + 1);
      }
    }
    acc += w_sum(term);
  }
  return acc;
}

// slots of the proposed tree: L + rank for leaves below `node`, 255 elsewhere; returns the slot count
__device__ inline int w_assign_prop_slots(const DTree& t, TravTree& prop, int node, int end, int L, int lane)
{
  const int nn = t.num_nodes;
  int s = L;
  for (int base = 0; base < nn; base += 32) {
    int k = base + lane;
    bool in = k < nn && k >= node && k < end && t.nodes[k].var < 0;
    unsigned m = __ballot_sync(0xffffffffu, in);
    if (k < nn) prop.slot[k] = in ? (uint8_t) (s + __popc(m & ((1u << lane) - 1u))) : (uint8_t) 255;
    s += __popc(m);
  }
  __syncwarp();
  return s;
}

// ---------------------------------------------------------------------------------------
// proposal (warp-cooperative; same draw order as oracle_bart.c metropolis_jump)
// ---------------------------------------------------------------------------------------
__device__ inline void w_propose(DTree& t, const BartParams& P, const double* tab, WarpRng& rng, StepDesc& d, CtlScratch& cs, int tree_index, int lane)
{
  const int nn = t.num_nodes;
  const int L = w_fill_trav(t, d.b_cur, true, lane);
  w_node_attrs(t, P, cs, L, lane);
  int kind = -1, node = -1, b_var = -1, b_cut = -1, child = -1, new_var = -1, new_cut = -1, nslots = L, b_end = -1;
  double lpt = 0.0;

  const double u = rng.uniform();
  if (u < P.birth_death_prob) {
    const int n_birthable = w_count_flag(cs, nn, kFBirthable, lane);
    const double p_birth = n_birthable == 0 ? 0.0 : (nn == 1 ? 1.0 : P.birth_prob);
    const int n_nog = w_count_flag(cs, nn, kFNog, lane);
    if (rng.uniform() < p_birth) {
      node = w_select_flag(cs, nn, kFBirthable, rng.index(n_birthable), lane);
      const int navail = cs.navail[node];
      const int depth = t.nodes[node].depth;
      const double pg_parent = t_growth_prob_depth(tab + kTabPg, navail, depth);
      b_var = w_draw_var(t, P, node, navail, rng, lane);
      int lo, hi; t_split_interval(t, s4b_ncuts(P, b_var), node, b_var, lo, hi);
      b_cut = lo + rng.index(hi - lo + 1);
      const int navail_l = navail - ((b_cut - 1 < lo) ? 1 : 0);
      const int navail_r = navail - ((b_cut + 1 > hi) ? 1 : 0);
      const double pg_l = t_growth_prob_depth(tab + kTabPg, navail_l, depth + 1), pg_r = t_growth_prob_depth(tab + kTabPg, navail_r, depth + 1);
      const int Lnew = L + 1;
      bool any_new = false;
      for (int base = 0; base < nn; base += 32) {
        int k = base + lane;
        bool b = k < nn && k != node && (cs.flag[k] & kFLeaf) && cs.navail[k] > 0 && t.nodes[k].depth < S4B_MAX_DEPTH && Lnew < S4B_MAX_LEAVES;
        if (__ballot_sync(0xffffffffu, b)) any_new = true;
      }
      if (!any_new && depth + 1 < S4B_MAX_DEPTH && Lnew < S4B_MAX_LEAVES && (navail_l > 0 || navail_r > 0)) any_new = true;
      const double p_death_new = 1.0 - (any_new ? P.birth_prob : 0.0);
      const int par = t.nodes[node].parent;
      const bool parent_was_nog = par >= 0 && (cs.flag[par] & kFNog);
      const int n_nog_new = n_nog + 1 - (parent_was_nog ? 1 : 0);
      const double prior_ratio = pg_parent * (1.0 - pg_l) * (1.0 - pg_r) / (1.0 - pg_parent);
      const double trans_ratio = (p_death_new * (1.0 / (double) n_nog_new)) / (p_birth * (1.0 / (double) n_birthable));
      kind = 0; nslots = L + 2; lpt = prior_ratio * trans_ratio;
    } else {
      node = w_select_flag(cs, nn, kFNog, rng.index(n_nog), lane);
      const int left = node + 1, right = t.nodes[node].right;
      const double pg_parent = t_growth_prob_depth(tab + kTabPg, cs.navail[node], t.nodes[node].depth);
      const double pg_l = t_growth_prob_depth(tab + kTabPg, cs.navail[left], t.nodes[left].depth);
      const double pg_r = t_growth_prob_depth(tab + kTabPg, cs.navail[right], t.nodes[right].depth);
      const int Lnew = L - 1;
      int n_birthable_new = 0;
      for (int base = 0; base < nn; base += 32) {
        int k = base + lane;
        bool b = k < nn && k != left && k != right && (cs.flag[k] & kFLeaf) && cs.navail[k] > 0 && t.nodes[k].depth < S4B_MAX_DEPTH && Lnew < S4B_MAX_LEAVES;
        n_birthable_new += __popc(__ballot_sync(0xffffffffu, b));
      }
      if (t.nodes[node].depth < S4B_MAX_DEPTH && Lnew < S4B_MAX_LEAVES) ++n_birthable_new;
      double p_birth_new = node == 0 ? 1.0 : P.birth_prob;
      if (n_birthable_new == 0) p_birth_new = 0.0;
      const double p_select_birth = n_birthable_new > 0 ? 1.0 / (double) n_birthable_new : 0.0;
      const double p_death = 1.0 - p_birth;
      const double prior_ratio = (1.0 - pg_parent) / (pg_parent * (1.0 - pg_l) * (1.0 - pg_r));
      const double trans_ratio = (p_birth_new * p_select_birth) / (p_death * (1.0 / (double) n_nog));
      kind = 1; b_var = t.nodes[node].var; b_cut = t.nodes[node].cut; lpt = prior_ratio * trans_ratio;
    }
  } else if (u < P.birth_death_prob + P.swap_prob) {
    // ---- swap ----
    kind = 13;
    const int n_sw = w_count_flag(cs, nn, kFSwappable, lane);
    if (n_sw > 0) {
      node = w_select_flag(cs, nn, kFSwappable, rng.index(n_sw), lane);
      const int left = node + 1, right = t.nodes[node].right;
      const bool li = t.nodes[left].var >= 0, ri = t.nodes[right].var >= 0;
      const bool both_same = li && ri && t.nodes[left].var == t.nodes[right].var && t.nodes[left].cut == t.nodes[right].cut;
      if (!both_same) {
        if (li && ri) child = rng.uniform() < 0.5 ? left : right;
        else child = li ? left : right;
      }
      const int end = t_subtree_end(t, node);
      const int pv = t.nodes[node].var, pc = t.nodes[node].cut;
      const int cv = both_same ? t.nodes[left].var : t.nodes[child].var, cc = both_same ? t.nodes[left].cut : t.nodes[child].cut;
      const double old_lp = w_branch_log_prior(t, P, tab, node, end, lane);
      __syncwarp();
      if (lane == 0) {
        t.nodes[node].var = (int16_t) cv; t.nodes[node].cut = (int16_t) cc;
        if (both_same) { t.nodes[left].var = (int16_t) pv; t.nodes[left].cut = (int16_t) pc; t.nodes[right].var = (int16_t) pv; t.nodes[right].cut = (int16_t) pc; }
        else { t.nodes[child].var = (int16_t) pv; t.nodes[child].cut = (int16_t) pc; }
      }
      __syncwarp();
      bool bad = false;
      for (int base = node; base < end; base += 32) {
        int k = base + lane;
        bool b = false;
        if (k < end && t.nodes[k].var >= 0) { int lo, hi; t_split_interval(t, s4b_ncuts(P, t.nodes[k].var), k, t.nodes[k].var, lo, hi); b = t.nodes[k].cut < lo || t.nodes[k].cut > hi; }
        if (__ballot_sync(0xffffffffu, b)) bad = true;
      }
      double new_lp = 0.0;
      if (!bad) {
        new_lp = w_branch_log_prior(t, P, tab, node, end, lane);
        w_fill_trav(t, d.b_prop, false, lane);
      }
      __syncwarp();
      if (lane == 0) {
        t.nodes[node].var = (int16_t) pv; t.nodes[node].cut = (int16_t) pc;
        if (both_same) { t.nodes[left].var = (int16_t) cv; t.nodes[left].cut = (int16_t) cc; t.nodes[right].var = (int16_t) cv; t.nodes[right].cut = (int16_t) cc; }
        else { t.nodes[child].var = (int16_t) cv; t.nodes[child].cut = (int16_t) cc; }
      }
      __syncwarp();
      if (!bad) { nslots = w_assign_prop_slots(t, d.b_prop, node, end, L, lane); kind = 3; lpt = new_lp - old_lp; b_end = end; }
    }
  } else {
    // ---- change ----
    kind = 12;
    const int n_nb = w_count_flag(cs, nn, kFInternal, lane);
    if (n_nb > 0) {
      node = w_select_flag(cs, nn, kFInternal, rng.index(n_nb), lane);
      new_var = w_draw_var(t, P, node, cs.navail[node], rng, lane);
      int lo, hi; t_split_interval(t, s4b_ncuts(P, new_var), node, new_var, lo, hi);
      const int end = t_subtree_end(t, node), rstart = t.nodes[node].right;
      int lo_c = lo, hi_c = hi;
      for (int base = node + 1; base < end; base += 32) {
        int k = base + lane;
        if (k < end && t.nodes[k].var == new_var) {
          int c = t.nodes[k].cut;
          if (k < rstart) lo_c = max(lo_c, c + 1); else hi_c = min(hi_c, c - 1);
        }
      }
      lo = w_maxi(lo_c); hi = w_mini(hi_c);
      if (lo <= hi) {
        new_cut = lo + rng.index(hi - lo + 1);
        const double old_lp = w_branch_log_prior(t, P, tab, node, end, lane);
        const int ov = t.nodes[node].var, oc = t.nodes[node].cut;
        __syncwarp();
        if (lane == 0) { t.nodes[node].var = (int16_t) new_var; t.nodes[node].cut = (int16_t) new_cut; }
        __syncwarp();
        const double new_lp = w_branch_log_prior(t, P, tab, node, end, lane);
        w_fill_trav(t, d.b_prop, false, lane);
        if (lane == 0) { t.nodes[node].var = (int16_t) ov; t.nodes[node].cut = (int16_t) oc; }
        __syncwarp();
        nslots = w_assign_prop_slots(t, d.b_prop, node, end, L, lane);
        // proposal (Hastings) term: |I_new| P(old var) / (|I_old| P(new var)), I = interval left by ancestors and descendants
        double log_hastings = 0.0;
        if (!P.change_symmetric && new_var != ov) {
          int olo, ohi; t_split_interval(t, s4b_ncuts(P, ov), node, ov, olo, ohi);
          int olo_c = olo, ohi_c = ohi;
          for (int base = node + 1; base < end; base += 32) {
            int k = base + lane;
            if (k < end && t.nodes[k].var == ov) {
              int c = t.nodes[k].cut;
              if (k < rstart) olo_c = max(olo_c, c + 1); else ohi_c = min(ohi_c, c - 1);
            }
          }
          olo = w_maxi(olo_c); ohi = w_mini(ohi_c);
          log_hastings = tab_log_int(tab, hi - lo + 1) - tab_log_int(tab, ohi - olo + 1);
          if (P.split_w != nullptr) log_hastings += log((double) P.split_w[ov] / (double) P.split_w[new_var]);
        }
        kind = 2; lpt = (new_lp - old_lp) + log_hastings; b_end = end;
      }
    }
  }
  if (lane == 0) {
    d.a_valid = 0; d.a_same = 1;
    d.b_tree = tree_index; d.b_kind = kind; d.b_node = node; d.b_var = b_var; d.b_cut = b_cut; d.b_child = child;
    d.b_num_leaves = L; d.b_nslots = nslots; d.log_prior_trans = lpt; d.new_var = new_var; d.new_cut = new_cut; d.b_end = b_end;
    if (kind != 2 && kind != 3) d.b_prop.n = 0;
  }
  __syncwarp();
}

// the data-independent side of the Metropolis test, on the log scale (no exp on the decision's critical path)
__device__ __forceinline__ double accept_threshold(double u, int kind, double prior_trans)
{
  if (kind == 0 || kind == 1) return prior_trans > 0.0 ? log(u) - log(prior_trans) : INFINITY;
  return log(u) - prior_trans;
}

// posterior mean / sd of the leaf value and integrated log-likelihood of one leaf (two divisions, one log, one sqrt).
// SQ = false: the sum of squares is not available and not needed -- every Metropolis ratio compares two partitions of the
// SAME observations, so the -sum(r^2) / (2 sigma^2) terms of the two sides cancel exactly; `ll` is then the integrated
// log-likelihood without that common term: 1/2 log(a / (a + n/s2)) + 1/2 (n/s2)^2 avg^2 / (a + n/s2)
// weighted: the slot holds (count, sum w r, sum w) -- y_i ~ N(mu, sigma^2 / w_i), so sum w takes the place of the count in the
// data precision and the mean is the weighted one; the (cancelling) sum w r^2 is never formed
template <bool SQ>
__device__ __forceinline__ void slot_summary(const LeafStat& s, double inv_sigsq, double leaf_prec, bool weighted, double& pmean, double& psd, double& ll)
{
  const double neff = (SQ && weighted) ? s.sumsq : s.n;
  const double dp = neff * inv_sigsq;
  const double rinv = 1.0 / (leaf_prec + dp);
  psd = sqrt(rinv);
  if (neff <= 0.0) { pmean = 0.0; ll = 0.0; return; }
  const double avg = s.sum / neff;
  pmean = dp * avg * rinv;
  if (SQ && !weighted) {
    double ss = s.sumsq - s.n * avg * avg;
    if (ss < 0.0) ss = 0.0;
    ll = 0.5 * log(leaf_prec * rinv) - 0.5 * ss * inv_sigsq - 0.5 * ((leaf_prec * avg) * (dp * avg)) * rinv;
  } else {
    ll = 0.5 * log(leaf_prec * rinv) + 0.5 * ((dp * avg) * (dp * avg)) * rinv;
  }
}

// ---------------------------------------------------------------------------------------
// decision + leaf draws (warp-cooperative)
// ---------------------------------------------------------------------------------------
template <bool SQ>
__device__ inline void w_decide(DTree& t, const BartParams& P, WarpRng& rng, const StepDesc& in, const LeafStat* stats, UpdateDesc& upd,
                                CtlScratch& cs, double* trace_rec, int lane, double inv_sigsq, double accept_thr)
{
  const int L = in.b_num_leaves, kind = in.b_kind, node = in.b_node, nslots = in.b_nslots;
  const int nn_old = t.num_nodes;
  const long long d0 = clock64();
  // slot nslots = the merged parent of a birth / death step, summarised in the same parallel pass
  int sl_bd = 0, sr_bd = 0;
  if (kind == 0 || kind == 1) { sl_bd = kind == 0 ? L : in.b_cur.slot[node + 1]; sr_bd = kind == 0 ? L + 1 : in.b_cur.slot[t.nodes[node].right]; }
  const int nsum = nslots + ((kind == 0 || kind == 1) ? 1 : 0);
  // lane s summarises slot s; with at most 32 slots (nearly always) the summaries also stay in registers and the
  // decision below reads them with shuffles instead of a shared-memory round trip
  const bool small = nsum <= 32;
  double my_ll = 0.0, my_n = 0.0;
  for (int s = lane; s < nsum; s += 32) {
    LeafStat st = s < nslots ? stats[s] : LeafStat{ stats[sl_bd].n + stats[sr_bd].n, stats[sl_bd].sum + stats[sr_bd].sum, stats[sl_bd].sumsq + stats[sr_bd].sumsq };
    double pm, ps, ll;
    slot_summary<SQ>(st, inv_sigsq, P.leaf_prec, P.weighted != 0, pm, ps, ll);
    cs.pmean[s] = pm; cs.psd[s] = ps; cs.ll[s] = ll;
    my_ll = ll; my_n = st.n;
  }
  const long long f1 = clock64();
  for (int k = lane; k < nn_old; k += 32) upd.val_old[k] = in.b_cur.val[k];
  if (!small) __syncwarp();
  const long long f2 = clock64();
  bool accept = false;
  double ratio = -1.0, old_ll = 0.0, new_ll = 0.0, n_first = 0.0, n_second = 0.0;
  if (kind == 0 || kind == 1) {
    const int sl = sl_bd, sr = sr_bd;
    double ll_par, ll_l, ll_r, n_l, n_r;
    if (small) {
      ll_par = __shfl_sync(0xffffffffu, my_ll, nslots); ll_l = __shfl_sync(0xffffffffu, my_ll, sl); ll_r = __shfl_sync(0xffffffffu, my_ll, sr);
      n_l = __shfl_sync(0xffffffffu, my_n, sl); n_r = __shfl_sync(0xffffffffu, my_n, sr);
    } else { ll_par = cs.ll[nslots]; ll_l = cs.ll[sl]; ll_r = cs.ll[sr]; n_l = stats[sl].n; n_r = stats[sr].n; }
    const double ll_ch = ll_l + ll_r;
    if (kind == 0) { old_ll = ll_par; new_ll = ll_ch; } else { old_ll = ll_ch; new_ll = ll_par; }
    ratio = in.log_prior_trans * exp(new_ll - old_ll);
    const bool too_small = kind == 0 && (n_l < (double) P.min_obs || n_r < (double) P.min_obs);
    if (too_small) ratio = 0.0;
    (void) rng.uniform();
    accept = !too_small && accept_thr < new_ll - old_ll;
    n_first = n_l; n_second = n_r;
  } else if (kind == 2 || kind == 3) {
    if (small) __syncwarp();
    const int end = in.b_end;
    double a_old = 0.0, a_new = 0.0, mn = 1e300;
    int first_leaf = -1, second_leaf = -1, seen = 0;
    for (int base = node; base < end; base += 32) {
      int k = base + lane;
      bool leaf = k < end && t.nodes[k].var < 0;
      double to = 0.0, tn = 0.0, nk = 1e300;
      if (leaf) { to = cs.ll[in.b_cur.slot[k]]; tn = cs.ll[in.b_prop.slot[k]]; nk = stats[in.b_prop.slot[k]].n; }
      a_old += w_sum(to); a_new += w_sum(tn); mn = fmin(mn, w_min(nk));
      unsigned m = __ballot_sync(0xffffffffu, leaf);
      while (m && seen < 2) { int b = __ffs(m) - 1; if (seen == 0) first_leaf = base + b; else second_leaf = base + b; ++seen; m &= m - 1; }
    }
    old_ll = a_old; new_ll = a_new;
    ratio = exp(in.log_prior_trans + (new_ll - old_ll));
    const bool too_small = mn < (double) P.min_obs;
    if (too_small) ratio = 0.0;
    (void) rng.uniform();
    accept = !too_small && accept_thr < new_ll - old_ll;
    if (first_leaf >= 0) n_first = stats[in.b_prop.slot[first_leaf]].n;
    if (second_leaf >= 0) n_second = stats[in.b_prop.slot[second_leaf]].n;
  }
  __syncwarp();        // summaries and staged leaf values visible to all lanes

  // ---- structural change, all lanes ----
  const long long d1 = clock64();
  if (lane == 0 && cs.prof_on) { cs.fine[0] += f1 - d0; cs.fine[1] += f2 - f1; cs.fine[2] += d1 - f2; }
  const int amode = !accept ? 0 : (kind == 0 ? 1 : (kind == 1 ? 2 : 3));
  if (amode == 1 || amode == 2) {
    for (int k = lane; k < nn_old; k += 32) cs.tmp[k] = t.nodes[k];
    __syncwarp();
    const int shift = amode == 1 ? 2 : -2;
    const int pivot = amode == 1 ? node : node + 2;     // indices > pivot move
    for (int k = lane; k < nn_old; k += 32) {
      if (amode == 2 && (k == node + 1 || k == node + 2)) continue;
      DNode nd = cs.tmp[k];
      if (nd.var >= 0 && nd.right > pivot) nd.right = (int16_t) (nd.right + shift);
      if (nd.parent > pivot) nd.parent = (int16_t) (nd.parent + shift);
      t.nodes[k > pivot ? k + shift : k] = nd;
    }
    __syncwarp();
    if (lane == 0) {
      DNode& nd = t.nodes[node];
      if (amode == 1) {
        nd.var = (int16_t) in.b_var; nd.cut = (int16_t) in.b_cut; nd.right = (int16_t) (node + 2);
        for (int c = 1; c <= 2; ++c) { DNode& ch = t.nodes[node + c]; ch.var = -1; ch.cut = -1; ch.right = -1; ch.parent = (int16_t) node; ch.n = 0; ch.depth = nd.depth + 1; ch.mu = 0.0; }
      } else { nd.var = -1; nd.cut = -1; nd.right = -1; }
      t.num_nodes = nn_old + shift;
    }
    __syncwarp();
  } else if (amode == 3) {
    const int end = in.b_end;
    for (int k = node + lane; k < end; k += 32) if (t.nodes[k].var >= 0) { uint32_t tv = in.b_prop.trav[k]; t.nodes[k].var = (int16_t) trav2_var(tv); t.nodes[k].cut = (int16_t) trav2_cut(tv); }
    __syncwarp();
  }

  // ---- leaf draws: one lane per node of the final tree ----
  const long long d2 = clock64();
  const int nn = t.num_nodes;
  int leaves_before = 0;
  for (int base = 0; base < nn; base += 32) {
    const int k = base + lane;
    const bool leaf = k < nn && t.nodes[k].var < 0;
    const unsigned m = __ballot_sync(0xffffffffu, leaf);
    const int cnt = __popc(m);
    const int p0 = cnt > 0 ? rng.reserve_normals(cnt) : 0;
    if (leaf) {
      const int j = __popc(m & ((1u << lane) - 1u));
      int old_k = k;
      if (amode == 1) old_k = (k <= node) ? k : (k <= node + 2 ? node : k - 2);
      else if (amode == 2) old_k = (k <= node) ? k : k + 2;
      // which statistic slot describes this leaf (-1: the merged parent of a rejected birth / accepted death)
      int sl;
      if (kind == 0) {
        if (accept && k == node + 1) sl = L;
        else if (accept && k == node + 2) sl = L + 1;
        else if (!accept && k == node) sl = -1;
        else sl = in.b_cur.slot[old_k];
      } else if (kind == 1 && accept && k == node) sl = -1;
      else if ((kind == 2 || kind == 3) && accept && in.b_prop.slot[k] != 255) sl = in.b_prop.slot[k];
      else sl = in.b_cur.slot[old_k];
      const int ss = sl >= 0 ? sl : nslots;
      const double nobs = sl >= 0 ? stats[sl].n : stats[sl_bd].n + stats[sr_bd].n;
      const double mu = cs.pmean[ss] + cs.psd[ss] * cs.zbuf[p0 + j];
      t.nodes[k].mu = mu;
      t.nodes[k].n = (int32_t) nobs;
      upd.val_new[k] = mu;
      if (trace_rec != nullptr && 11 + leaves_before + j < S4B_TRACE_LEN) trace_rec[11 + leaves_before + j] = mu;
    }
    leaves_before += cnt;
  }
  // ---- residual update descriptor ----
  const long long d3 = clock64();
  for (int k = lane; k < nn_old; k += 32) {
    int nk = k;
    if (amode == 1) nk = k > node ? k + 2 : k;                                   // the split leaf itself is resolved by the caller
    else if (amode == 2) nk = (k == node + 1 || k == node + 2) ? node : (k > node + 2 ? k - 2 : k);
    upd.remap[k] = (uint8_t) nk;
    if (amode == 0) upd.delta[k] = t.nodes[k].var < 0 ? in.b_cur.val[k] - t.nodes[k].mu : 0.0;
  }
  if (lane == 0) { upd.mode = amode; upd.node = node; }
  if (trace_rec != nullptr && lane == 0) {
    trace_rec[0] = (double) kind;
    trace_rec[1] = (kind >= 0 && node >= 0) ? (double) t_heap_index(t, node) : 0.0;
    if (kind == 0 || kind == 1) { trace_rec[2] = in.b_var; trace_rec[3] = in.b_cut; }
    else if (kind == 2 || kind == 12) { trace_rec[2] = node >= 0 ? in.new_var : 0; trace_rec[3] = kind == 2 ? in.new_cut : 0; }
    else if ((kind == 3 || kind == 13) && node >= 0) { trace_rec[2] = in.b_child >= 0 ? (double) t_heap_index(t, in.b_child) : -1.0; }
    trace_rec[4] = accept ? 1.0 : 0.0; trace_rec[5] = ratio; trace_rec[6] = old_ll; trace_rec[7] = new_ll;
    trace_rec[8] = (double) leaves_before; trace_rec[9] = n_first; trace_rec[10] = n_second;
  }
  __syncwarp();
  if (lane == 0 && cs.prof_on) { const long long d4 = clock64(); cs.dbg[0] += d1 - d0; cs.dbg[1] += d2 - d1; cs.dbg[2] += d3 - d2; cs.dbg[3] += d4 - d3; }
}

// ---------------------------------------------------------------------------------------
// Fast path of the decision for trees of at most 30 nodes (practically all of them): everything that depends only on
// the tree and its proposal -- the node array of the accepted outcome, which statistic slot and which normal draw
// belong to every bottom node under either outcome, the index remap of the residual update -- is prepared by the
// controller BEFORE the grid barrier, while it would otherwise idle.  After the barrier what is left is the arithmetic
// on the statistics: lane s summarises slot s in registers, the ratio is assembled with shuffles, and lane k finishes
// node k of the final tree.  Results are identical to w_decide (same formulas, same draw positions).
// ---------------------------------------------------------------------------------------
struct FastPlan {
  bool valid;               // warp-uniform
  int nn_acc;               // nodes of the tree if the proposal is accepted
  int sl_bd, sr_bd;         // birth / death: slots of the two children
  unsigned mask_rej, mask_acc;   // bottom nodes of the final tree under either outcome (bit k = node k)
  unsigned mask_under;      // change / swap: bottom nodes below the proposal node
  int slot_rej, slot_acc;   // lane k: statistic slot of node k as a bottom node of the final tree (nslots = merged parent)
  int slot_cur, slot_prop;  // lane k: slots under the current / proposed rule (change / swap ratio)
};

// the plan crosses the grid barrier in shared memory: held in registers it would be live across the workers' code and
// cost every thread of the block a dozen registers
struct FastPlanSmem {
  int valid, nn_acc, sl_bd, sr_bd;
  unsigned mask_rej, mask_acc, mask_under;
  int pad;
  uint8_t slot_rej[32], slot_acc[32], slot_cur[32], slot_prop[32];
};
__device__ inline void plan_store(FastPlanSmem& d, const FastPlan& pl, int lane)
{
  if (lane == 0) { d.valid = pl.valid ? 1 : 0; d.nn_acc = pl.nn_acc; d.sl_bd = pl.sl_bd; d.sr_bd = pl.sr_bd; d.mask_rej = pl.mask_rej; d.mask_acc = pl.mask_acc; d.mask_under = pl.mask_under; }
  d.slot_rej[lane] = (uint8_t) pl.slot_rej; d.slot_acc[lane] = (uint8_t) pl.slot_acc; d.slot_cur[lane] = (uint8_t) pl.slot_cur; d.slot_prop[lane] = (uint8_t) pl.slot_prop;
  __syncwarp();
}
__device__ inline FastPlan plan_load(const FastPlanSmem& d, int lane)
{
  FastPlan pl;
  pl.valid = d.valid != 0; pl.nn_acc = d.nn_acc; pl.sl_bd = d.sl_bd; pl.sr_bd = d.sr_bd; pl.mask_rej = d.mask_rej; pl.mask_acc = d.mask_acc; pl.mask_under = d.mask_under;
  pl.slot_rej = d.slot_rej[lane]; pl.slot_acc = d.slot_acc[lane]; pl.slot_cur = d.slot_cur[lane]; pl.slot_prop = d.slot_prop[lane];
  return pl;
}

__device__ inline FastPlan w_plan(const DTree& t, const StepDesc& in, UpdateDesc& upd, CtlScratch& cs, int lane)
{
  FastPlan pl;
  const int L = in.b_num_leaves, kind = in.b_kind, node = in.b_node, nslots = in.b_nslots;
  const int nn_old = t.num_nodes;
  const bool bd = kind == 0 || kind == 1;
  const int nsum = nslots + (bd ? 1 : 0);
  pl.valid = nn_old + 2 <= 32 && nsum <= 32;
  pl.nn_acc = nn_old; pl.sl_bd = 0; pl.sr_bd = 0; pl.mask_rej = 0; pl.mask_acc = 0; pl.mask_under = 0;
  pl.slot_rej = 0; pl.slot_acc = 0; pl.slot_cur = 0; pl.slot_prop = 0;
  if (!pl.valid) return pl;
  const int k = lane;
  DNode me; me.var = -1; me.cut = -1; me.right = -1; me.parent = -1; me.n = 0; me.depth = 0; me.mu = 0.0;
  if (k < nn_old) me = t.nodes[k];
  const bool leaf_old = k < nn_old && me.var < 0;
  const int my_slot = k < nn_old ? (int) in.b_cur.slot[k] : 0;
  pl.mask_rej = __ballot_sync(0xffffffffu, leaf_old);
  pl.slot_cur = my_slot;
  // staged values the residual update reads whatever the outcome
  if (k < nn_old) upd.val_old[k] = in.b_cur.val[k];
  if (kind == 0) {
    pl.sl_bd = L; pl.sr_bd = L + 1;
    pl.nn_acc = nn_old + 2;
    // accepted tree: nodes after `node` move up by two, `node` becomes internal with two fresh bottom children
    const int depth_node = __shfl_sync(0xffffffffu, me.depth, node);
    if (k < nn_old) {
      DNode nd = me;
      if (nd.var >= 0 && nd.right > node) nd.right = (int16_t) (nd.right + 2);
      if (nd.parent > node) nd.parent = (int16_t) (nd.parent + 2);
      if (k == node) { nd.var = (int16_t) in.b_var; nd.cut = (int16_t) in.b_cut; nd.right = (int16_t) (node + 2); }
      cs.tmp[k > node ? k + 2 : k] = nd;
      upd.remap[k] = (uint8_t) (k > node ? k + 2 : k);
    }
    if (k == 0) {
      for (int c = 1; c <= 2; ++c) { DNode& ch = cs.tmp[node + c]; ch.var = -1; ch.cut = -1; ch.right = -1; ch.parent = (int16_t) node; ch.n = 0; ch.depth = depth_node + 1; ch.mu = 0.0; }
    }
    pl.slot_rej = k == node ? nslots : my_slot;
    const int old_k = (k <= node) ? k : (k <= node + 2 ? node : k - 2);
    const int old_slot = __shfl_sync(0xffffffffu, my_slot, old_k & 31);
    const bool old_leaf = (pl.mask_rej >> (old_k & 31)) & 1u;
    const bool leaf_acc = k < pl.nn_acc && (k == node + 1 || k == node + 2 || (k != node && old_leaf));
    pl.slot_acc = k == node + 1 ? L : (k == node + 2 ? L + 1 : old_slot);
    pl.mask_acc = __ballot_sync(0xffffffffu, leaf_acc);
  } else if (kind == 1) {
    const int right = __shfl_sync(0xffffffffu, (int) me.right, node);
    pl.sl_bd = __shfl_sync(0xffffffffu, my_slot, node + 1); pl.sr_bd = __shfl_sync(0xffffffffu, my_slot, right & 31);
    pl.nn_acc = nn_old - 2;
    if (k < nn_old && k != node + 1 && k != node + 2) {
      DNode nd = me;
      if (nd.var >= 0 && nd.right > node + 2) nd.right = (int16_t) (nd.right - 2);
      if (nd.parent > node + 2) nd.parent = (int16_t) (nd.parent - 2);
      if (k == node) { nd.var = -1; nd.cut = -1; nd.right = -1; }
      cs.tmp[k > node + 2 ? k - 2 : k] = nd;
    }
    if (k < nn_old) upd.remap[k] = (uint8_t) ((k == node + 1 || k == node + 2) ? node : (k > node + 2 ? k - 2 : k));
    pl.slot_rej = my_slot;
    const int old_k = (k <= node) ? k : k + 2;
    const int old_slot = __shfl_sync(0xffffffffu, my_slot, old_k & 31);
    const bool old_leaf = old_k < 32 && ((pl.mask_rej >> (old_k & 31)) & 1u);
    const bool leaf_acc = k < pl.nn_acc && (k == node || old_leaf);
    pl.slot_acc = k == node ? nslots : old_slot;
    pl.mask_acc = __ballot_sync(0xffffffffu, leaf_acc);
  } else if (kind == 2 || kind == 3) {
    const int end = in.b_end;
    if (k < nn_old) {
      DNode nd = me;
      if (k >= node && k < end && nd.var >= 0) { const uint32_t tv = in.b_prop.trav[k]; nd.var = (int16_t) trav2_var(tv); nd.cut = (int16_t) trav2_cut(tv); }
      cs.tmp[k] = nd;
      upd.remap[k] = (uint8_t) k;
    }
    const int ps = k < nn_old ? (int) in.b_prop.slot[k] : 255;
    pl.slot_prop = ps == 255 ? 0 : ps;
    pl.slot_rej = my_slot;
    pl.slot_acc = ps != 255 ? ps : my_slot;
    pl.mask_acc = pl.mask_rej;
    pl.mask_under = __ballot_sync(0xffffffffu, leaf_old && k >= node && k < end);
  } else {
    if (k < nn_old) upd.remap[k] = (uint8_t) k;
    pl.slot_rej = my_slot; pl.slot_acc = my_slot; pl.mask_acc = pl.mask_rej;
  }
  __syncwarp();
  return pl;
}

template <bool SQ>
__device__ inline void w_decide_fast(const FastPlan& pl, DTree& t, const BartParams& P, WarpRng& rng, const StepDesc& in, const LeafStat* stats,
                                     UpdateDesc& upd, CtlScratch& cs, double* trace_rec, int lane, double inv_sigsq, double accept_thr)
{
  const int L = in.b_num_leaves, kind = in.b_kind, node = in.b_node, nslots = in.b_nslots;
  const bool bd = kind == 0 || kind == 1;
  const int nsum = nslots + (bd ? 1 : 0);
  const int nn_old = t.num_nodes;          // read by every lane before lane 0 rewrites it below
  __syncwarp();
  const long long d0 = clock64();
  // ---- lane s: summary of slot s (slot nslots = the merged parent of a birth / death step) ----
  double my_ll = 0.0, my_n = 0.0, my_pm = 0.0, my_ps = 0.0;
  if (lane < nsum) {
    LeafStat st;
    if (lane < nslots) st = stats[lane];
    else { const LeafStat a = stats[pl.sl_bd], b = stats[pl.sr_bd]; st.n = a.n + b.n; st.sum = a.sum + b.sum; st.sumsq = a.sumsq + b.sumsq; }
    slot_summary<SQ>(st, inv_sigsq, P.leaf_prec, P.weighted != 0, my_pm, my_ps, my_ll);
    my_n = st.n;
  }
  const long long f1 = clock64();
  // ---- leaf values of node `lane` under BOTH outcomes, issued before the ratio chain so that their shuffles and loads
  // complete in its shadow.  The accept uniform (if any) is draw 0, the normals follow at the same positions either way ----
  const bool has_u = kind >= 0 && kind <= 3;
  const int z0 = has_u ? 1 : 0;
  const int j_rej = __popc(pl.mask_rej & ((1u << lane) - 1u)), j_acc = __popc(pl.mask_acc & ((1u << lane) - 1u));
  const double pm_r = __shfl_sync(0xffffffffu, my_pm, pl.slot_rej & 31), ps_r = __shfl_sync(0xffffffffu, my_ps, pl.slot_rej & 31);
  const double n_r_ = __shfl_sync(0xffffffffu, my_n, pl.slot_rej & 31);
  const double pm_a = __shfl_sync(0xffffffffu, my_pm, pl.slot_acc & 31), ps_a = __shfl_sync(0xffffffffu, my_ps, pl.slot_acc & 31);
  const double n_a_ = __shfl_sync(0xffffffffu, my_n, pl.slot_acc & 31);
  const double mu_rej = pm_r + ps_r * cs.zbuf[(z0 + j_rej) & 31], mu_acc = pm_a + ps_a * cs.zbuf[(z0 + j_acc) & 31];
  // ---- Metropolis test on the log scale: accept iff log u - log(prior x transition) < delta log-likelihood ----
  bool accept = false;
  double ratio = -1.0, old_ll = 0.0, new_ll = 0.0, n_first = 0.0, n_second = 0.0;
  if (bd) {
    const double ll_par = __shfl_sync(0xffffffffu, my_ll, nslots), ll_l = __shfl_sync(0xffffffffu, my_ll, pl.sl_bd), ll_r = __shfl_sync(0xffffffffu, my_ll, pl.sr_bd);
    const double n_l = __shfl_sync(0xffffffffu, my_n, pl.sl_bd), n_r = __shfl_sync(0xffffffffu, my_n, pl.sr_bd);
    const double ll_ch = ll_l + ll_r;
    if (kind == 0) { old_ll = ll_par; new_ll = ll_ch; } else { old_ll = ll_ch; new_ll = ll_par; }
    const bool too_small = kind == 0 && (n_l < (double) P.min_obs || n_r < (double) P.min_obs);
    (void) rng.uniform();
    accept = !too_small && accept_thr < new_ll - old_ll;
    if (trace_rec != nullptr) ratio = too_small ? 0.0 : in.log_prior_trans * exp(new_ll - old_ll);
    n_first = n_l; n_second = n_r;
  } else if (kind == 2 || kind == 3) {
    const bool under = (pl.mask_under >> lane) & 1u;
    const double to_ = __shfl_sync(0xffffffffu, my_ll, pl.slot_cur & 31), tn_ = __shfl_sync(0xffffffffu, my_ll, pl.slot_prop & 31);
    const double nk_ = __shfl_sync(0xffffffffu, my_n, pl.slot_prop & 31);
    old_ll = w_sum(under ? to_ : 0.0); new_ll = w_sum(under ? tn_ : 0.0);
    const double mn = w_min(under ? nk_ : 1e300);
    const bool too_small = mn < (double) P.min_obs;
    (void) rng.uniform();
    accept = !too_small && accept_thr < new_ll - old_ll;
    if (trace_rec != nullptr) {
      ratio = too_small ? 0.0 : exp(in.log_prior_trans + (new_ll - old_ll));
      unsigned m = pl.mask_under;
      if (m) { const int b = __ffs(m) - 1; n_first = __shfl_sync(0xffffffffu, nk_, b); m &= m - 1; }
      if (m) { const int b = __ffs(m) - 1; n_second = __shfl_sync(0xffffffffu, nk_, b); }
    }
  }
  const long long d1 = clock64();
  // ---- final tree: lane k finishes node k ----
  const int amode = !accept ? 0 : (kind == 0 ? 1 : (kind == 1 ? 2 : 3));
  const unsigned mask = accept ? pl.mask_acc : pl.mask_rej;
  const int cnt = __popc(mask);
  if (cnt > 0) (void) rng.reserve_normals(cnt);          // bookkeeping: the positions were fixed above
  const bool leaf = (mask >> lane) & 1u;
  const int j = accept ? j_acc : j_rej;
  const double nobs = accept ? n_a_ : n_r_;
  const double mu = leaf ? (accept ? mu_acc : mu_rej) : 0.0;
  if (amode != 0) {
    const int nn = pl.nn_acc;
    if (lane < nn) { DNode nd = cs.tmp[lane]; if (leaf) { nd.mu = mu; nd.n = (int32_t) nobs; } t.nodes[lane] = nd; }
    if (lane == 0) t.num_nodes = nn;
    if (leaf) upd.val_new[lane] = mu;
  } else if (leaf) {
    t.nodes[lane].mu = mu; t.nodes[lane].n = (int32_t) nobs;
    upd.val_new[lane] = mu;
  }
  if (amode == 0 && lane < nn_old) upd.delta[lane] = leaf ? in.b_cur.val[lane] - mu : 0.0;
  if (lane == 0) { upd.mode = amode; upd.node = node; }
  if (trace_rec != nullptr) {
    if (leaf && 11 + j < S4B_TRACE_LEN) trace_rec[11 + j] = mu;
    __syncwarp();
    if (lane == 0) {
      trace_rec[0] = (double) kind;
      trace_rec[1] = (kind >= 0 && node >= 0) ? (double) t_heap_index(t, node) : 0.0;
      if (kind == 0 || kind == 1) { trace_rec[2] = in.b_var; trace_rec[3] = in.b_cut; }
      else if (kind == 2 || kind == 12) { trace_rec[2] = node >= 0 ? in.new_var : 0; trace_rec[3] = kind == 2 ? in.new_cut : 0; }
      else if ((kind == 3 || kind == 13) && node >= 0) { trace_rec[2] = in.b_child >= 0 ? (double) t_heap_index(t, in.b_child) : -1.0; }
      trace_rec[4] = accept ? 1.0 : 0.0; trace_rec[5] = ratio; trace_rec[6] = old_ll; trace_rec[7] = new_ll;
      trace_rec[8] = (double) cnt; trace_rec[9] = n_first; trace_rec[10] = n_second;
    }
  }
  __syncwarp();
  if (lane == 0 && cs.prof_on) { const long long d4 = clock64(); cs.dbg[0] += d1 - d0; cs.dbg[2] += d4 - d1; cs.fine[0] += f1 - d0; cs.fine[2] += d1 - f1; }
}

// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

struct SweepSmem {
  StepDesc sd[2];
  DTree tree[2];
  UpdateDesc upd[2];   // by step parity: the streamed variant applies step t's update while it accumulates step t + 1
  CtlScratch csd;      // decision scratch
  CtlScratch csp;      // proposal scratch
  LeafStat st[S4B_MAX_SLOTS];
  double inv_sigsq;    // 1 / sigma^2, fixed for the whole sweep
  long long pc[6], wk[4];   // phase cycle counters (profiling runs)
  FastPlanSmem plan;   // this step's decision plan (controller, written before the grid barrier)
  int peer_dead;       // a peer rank stopped answering (sharded mode): stop waiting, flag the error
  unsigned long long seq_base;   // sharded chains: exchanges completed before this launch (Mailbox::kseq)
  ShardDev sh;         // copy of the kernel parameter (indexed dynamically; keeps it out of local memory)
  RngState rng;
  BartParams prm;
  double tab[kTabSize];
};

// walk one tree for the four observations of a quad (branch-free records, 4-way ILP); returns packed node indices
__device__ __forceinline__ uint32_t walk_quad(const uint32_t* __restrict__ trav, const uint32_t* __restrict__ tile, int tile_stride, int qslot, int depth)
{
  int n0 = 0, n1 = 0, n2 = 0, n3 = 0;
#pragma unroll 1
  for (int lvl = 0; lvl < depth; ++lvl) {
    const uint32_t t0 = trav[n0], t1 = trav[n1], t2 = trav[n2], t3 = trav[n3];
    const uint32_t w0 = tile[(t0 >> 23) * tile_stride + qslot], w1 = tile[(t1 >> 23) * tile_stride + qslot];
    const uint32_t w2 = tile[(t2 >> 23) * tile_stride + qslot], w3 = tile[(t3 >> 23) * tile_stride + qslot];
    n0 = ((w0 & 0xFFu) <= ((t0 >> 15) & 0xFFu)) ? (int) ((t0 >> 8) & 0x7Fu) : (int) (t0 & 0x7Fu);
    n1 = (((w1 >> 8) & 0xFFu) <= ((t1 >> 15) & 0xFFu)) ? (int) ((t1 >> 8) & 0x7Fu) : (int) (t1 & 0x7Fu);
    n2 = (((w2 >> 16) & 0xFFu) <= ((t2 >> 15) & 0xFFu)) ? (int) ((t2 >> 8) & 0x7Fu) : (int) (t2 & 0x7Fu);
    n3 = ((w3 >> 24) <= ((t3 >> 15) & 0xFFu)) ? (int) ((t3 >> 8) & 0x7Fu) : (int) (t3 & 0x7Fu);
  }
  return (uint32_t) n0 | ((uint32_t) n1 << 8) | ((uint32_t) n2 << 16) | ((uint32_t) n3 << 24);
}

// leaf indices (and the auxiliary byte: proposed-tree leaf for change / swap, split side for a birth) of every owned quad
// bitmap walk: evaluate every internal node's rule for the 4 observations of each quad (one shared-memory word per rule
// and quad), collect the outcomes as one bit pattern per observation, look the bottom node up
template <int NQ>
__device__ __forceinline__ void bitmap_walk(const TravTree& tv, const uint8_t* __restrict__ table, const uint32_t* __restrict__ tile, int tile_stride, int tid,
                                            uint32_t (&pack)[NQ])
{
  uint32_t pat[NQ];
#pragma unroll
  for (int j = 0; j < NQ; ++j) pat[j] = 0u;
  const int n_int = tv.n_int;
#pragma unroll 1
  for (int i = 0; i < n_int; ++i) {
    const uint32_t rec = tv.irec[i];
    const uint32_t* col = tile + (rec >> 8) * tile_stride + tid;
    const uint32_t cut4 = (rec & 0xFFu) * 0x01010101u;
#pragma unroll
    for (int j = 0; j < NQ; ++j) pat[j] |= __vsetleu4(col[j * kWorkers], cut4) << i;
  }
#pragma unroll
  for (int j = 0; j < NQ; ++j)
    pack[j] = (uint32_t) table[pat[j] & 0xFFu] | ((uint32_t) table[(pat[j] >> 8) & 0xFFu] << 8) | ((uint32_t) table[(pat[j] >> 16) & 0xFFu] << 16) |
              ((uint32_t) table[pat[j] >> 24] << 24);
}

template <int NQ>
__device__ __forceinline__ void walk_step(const StepDesc& sd, const uint32_t* __restrict__ tile, int tile_stride, int tid, unsigned valid_mask,
                                          uint32_t (&leaf_pack)[NQ], uint32_t (&aux_pack)[NQ])
{
  const int kind = sd.b_kind;
  const int depth = sd.b_cur.pad;
  if (sd.b_cur.n_int <= S4B_BITMAP_INT) {
    // (quads beyond the data hold zeros in the tile: walking them is harmless, nothing of theirs is ever used)
    bitmap_walk<NQ>(sd.b_cur, sd.b_cur.table, tile, tile_stride, tid, leaf_pack);
    if (kind == 2 || kind == 3) bitmap_walk<NQ>(sd.b_prop, sd.b_cur.table, tile, tile_stride, tid, aux_pack);
    else if (kind == 0) {
      const uint32_t* col = tile + sd.b_var * tile_stride + tid;
      const uint32_t cut4 = (uint32_t) sd.b_cut * 0x01010101u;
#pragma unroll
      for (int j = 0; j < NQ; ++j) aux_pack[j] = __vcmpgtu4(col[j * kWorkers], cut4) & 0x01010101u;
    } else {
#pragma unroll
      for (int j = 0; j < NQ; ++j) aux_pack[j] = 0u;
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    uint32_t lp = 0, ap = 0;
    if ((valid_mask >> j) & 1u) {
      const int qslot = j * kWorkers + tid;
      lp = walk_quad(sd.b_cur.trav, tile, tile_stride, qslot, depth);
      if (kind == 2 || kind == 3) ap = walk_quad(sd.b_prop.trav, tile, tile_stride, qslot, depth);
      else if (kind == 0) {
        const uint32_t w = tile[sd.b_var * tile_stride + qslot];
        const uint32_t c = (uint32_t) sd.b_cut;
        ap = ((w & 0xFFu) > c ? 1u : 0u) | (((w >> 8) & 0xFFu) > c ? 0x100u : 0u) | (((w >> 16) & 0xFFu) > c ? 0x10000u : 0u) | ((w >> 24) > c ? 0x1000000u : 0u);
      }
    }
    leaf_pack[j] = lp; aux_pack[j] = ap;
  }
}

// warp-cooperative copy of the used part of a step descriptor
__device__ inline void w_copy_desc(StepDesc& dst, const StepDesc& src, int lane)
{
  const int kind = src.b_kind;
  const int n = src.b_cur.n;
  if (lane == 0) {
    dst.a_valid = 0; dst.a_same = 1;
    dst.b_tree = src.b_tree; dst.b_kind = kind; dst.b_node = src.b_node; dst.b_var = src.b_var; dst.b_cut = src.b_cut;
    dst.b_num_leaves = src.b_num_leaves; dst.b_nslots = src.b_nslots; dst.b_child = src.b_child;
    dst.log_prior_trans = src.log_prior_trans; dst.accept_thr = src.accept_thr; dst.new_var = src.new_var; dst.new_cut = src.new_cut; dst.b_end = src.b_end;
    dst.b_cur.n = n; dst.b_cur.pad = src.b_cur.pad;
    dst.b_prop.n = (kind == 2 || kind == 3) ? n : 0; dst.b_prop.pad = src.b_cur.pad;
    dst.b_cur.n_int = src.b_cur.n_int; dst.b_prop.n_int = src.b_cur.n_int;
  }
  {
    const int n_int = src.b_cur.n_int;
    if (n_int <= S4B_BITMAP_INT) {
      if (lane < n_int) { dst.b_cur.irec[lane] = src.b_cur.irec[lane]; if (kind == 2 || kind == 3) dst.b_prop.irec[lane] = src.b_prop.irec[lane]; }
      // the structure of the proposed tree of a change / swap step is that of the current one: one table serves both
      for (int e = lane * 4; e < (1 << n_int); e += 128) *reinterpret_cast<uint32_t*>(&dst.b_cur.table[e]) = *reinterpret_cast<const uint32_t*>(&src.b_cur.table[e]);
    }
  }
  for (int k = lane; k < n; k += 32) {
    dst.b_cur.trav[k] = src.b_cur.trav[k]; dst.b_cur.val[k] = src.b_cur.val[k]; dst.b_cur.slot[k] = src.b_cur.slot[k];
    if (kind == 2 || kind == 3) { dst.b_prop.trav[k] = src.b_prop.trav[k]; dst.b_prop.slot[k] = src.b_prop.slot[k]; }
  }
  __syncwarp();
}

// Everything of a sweep that does not depend on the data: with keyed RNG substreams the proposal of every tree (it needs
// only that tree's structure) and the decision draws of every step can be produced before the sweep, one warp per tree.
constexpr int kPrepWarps = 8;            // trees per block: the kernel is bound by instruction fetch (large, cold code): many warps per SM share the fetched lines
struct PrepSmemWarp { DTree tree; CtlScratch cs; StepDesc sd; };

// ---- pipelined sweep (sweep_pipe.cuh): data-independent description of a step's cells ----
constexpr int kPipeSegments = 2;          // (pipelined run, synchronous step) rounds per sweep; the last synchronous launch takes the rest
constexpr int kPipeSlots = 12;            // statistic slots a step of the pipelined kernel may have (+ 1 trash row of bins)
constexpr int kPipeCells = 48;            // capacity of the per-step cell tables (change / swap on a branch of k leaves: k^2 cells; slots <= 12 => k <= 6)
constexpr int kPipeCross = 288;           // capacity of a step's cross table (its slots) x (the previous step's cells); a pair of steps beyond it drains the pipeline
constexpr int kPipeRing = 4;
constexpr int kPipeDescs = 3;

// data-independent description of step t for the pipelined kernel (k_prepare_sweep): everything keyed by STATISTIC SLOT, so
// that the workers go from a row's rule pattern straight to its slot (one table look-up) and from the slot to the current leaf
// value; and the step's cells: which node a cell's rows sit in now (cell_a) and in the accepted tree (cell_f; the rejected
// tree keeps cell_a)
struct PipeInfo {
  int32_t ncells;
  int32_t ok;                              // this step fits the pipelined kernel
  int32_t slot_b;                          // birth: slot of the node to split (its rows take slot L + side); 255 otherwise
  int32_t nslots;                          // statistic slots of the step (the cross table of step t has nslots_t x ncells_{t-1} entries)
  double vs[kPipeSlots + 2];               // leaf value of the current tree by slot (a birth's two new slots repeat the parent's)
  uint8_t cellbase[16];                    // change / swap: first cell of slot s (a slot outside the branch: its only cell)
  uint8_t cell_a[kPipeCells], cell_f[kPipeCells];
  uint8_t stab[1 << S4B_BITMAP_INT];       // rule pattern -> slot under the current rules
  uint8_t ptab[1 << S4B_BITMAP_INT];       // change / swap: pattern under the proposed rules -> proposed slot (255 outside the branch)
};

// one warp per tree, after w_propose: fills infos[t]; returns whether the step fits
__device__ inline bool w_pipe_info(const DTree& t, const StepDesc& d, PipeInfo& pi, int /* max_cells: the pair limit is checked by the kernel */, int lane)
{
  const int nn = t.num_nodes, kind = d.b_kind, node = d.b_node, L = d.b_num_leaves, nslots = d.b_nslots;
  const bool bd = kind == 0 || kind == 1;
  const int n_int = d.b_cur.n_int;
  bool ok = nn + 2 <= 32 && nslots + (bd ? 1 : 0) <= 32 && nslots <= kPipeSlots && n_int <= S4B_BITMAP_INT;
  int ncells = nslots;
  for (int c = lane; c < kPipeCells; c += 32) { pi.cell_a[c] = 0; pi.cell_f[c] = 0; }
  if (lane < 16) pi.cellbase[lane] = 0;
  if (lane < kPipeSlots + 2) pi.vs[lane] = 0.0;
  __syncwarp();
  if (ok) {
    const bool leaf = lane < nn && t.nodes[lane].var < 0;
    const int slot = lane < nn ? (int) d.b_cur.slot[lane] : 255;
    if (leaf) pi.vs[slot] = d.b_cur.val[lane];
    if (kind == 0 && lane < 2) pi.vs[L + lane] = d.b_cur.val[node];
    if (kind == 2 || kind == 3) {
      const bool inside = leaf && d.b_prop.slot[lane] != 255;
      const unsigned m_in = __ballot_sync(0xffffffffu, inside), m_out = __ballot_sync(0xffffffffu, leaf && !inside);
      const int k_in = __popc(m_in), n_out = __popc(m_out);
      ncells = n_out + k_in * k_in;
      ok = ncells <= kPipeCells;
      if (ok) {
        const unsigned below = (1u << lane) - 1u;
        if (leaf && !inside) { const int c = __popc(m_out & below); pi.cellbase[slot] = (uint8_t) c; pi.cell_a[c] = (uint8_t) lane; pi.cell_f[c] = (uint8_t) lane; }
        if (inside) {
          const int r = __popc(m_in & below);
          pi.cellbase[slot] = (uint8_t) (n_out + r * k_in);
          // pair (this node now, j-th node of the branch after an accepted change / swap): prop slot L + j belongs to the j-th inside node
          for (int j = 0; j < k_in; ++j) { const int f = nth_set_bit(m_in, j); pi.cell_a[n_out + r * k_in + j] = (uint8_t) lane; pi.cell_f[n_out + r * k_in + j] = (uint8_t) f; }
        }
      }
    } else {
      ok = ncells <= kPipeCells;
      if (ok && leaf && slot < kPipeCells) {
        int f = lane;
        if (kind == 0) f = lane > node ? lane + 2 : lane;
        else if (kind == 1) f = (lane == node + 1 || lane == node + 2) ? node : (lane > node + 2 ? lane - 2 : lane);
        pi.cell_a[slot] = (uint8_t) lane; pi.cell_f[slot] = (uint8_t) f;
      }
      if (ok && kind == 0 && lane < 2) { pi.cell_a[L + lane] = (uint8_t) node; pi.cell_f[L + lane] = (uint8_t) (node + 1 + lane); }
    }
    if (ok) {
      const bool two = kind == 2 || kind == 3;
      for (int e = lane; e < (1 << n_int); e += 32) {
        const int nd = d.b_cur.table[e];
        pi.stab[e] = d.b_cur.slot[nd];
        pi.ptab[e] = two ? d.b_prop.slot[nd] : (uint8_t) 255;
      }
    }
  }
  if (lane == 0) { pi.ncells = ncells; pi.ok = ok ? 1 : 0; pi.slot_b = (ok && kind == 0) ? (int) d.b_cur.slot[node] : 255; pi.nslots = nslots; }
  __syncwarp();
  return ok;
}

__global__ void __launch_bounds__(kPrepWarps * 32) k_prepare_sweep(BartDev dv, StepDesc* __restrict__ descs, double2* __restrict__ draws,
                                                                   const double* __restrict__ tables, PipeInfo* __restrict__ infos,
                                                                   unsigned int* __restrict__ pipe_not_ok, int pipe_max_cells)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* tab = reinterpret_cast<double*>(smem_raw);
  BartParams* prm = reinterpret_cast<BartParams*>(tab + kTabSize);
  RngState* rng = reinterpret_cast<RngState*>(prm + 1);
  PrepSmemWarp* pw = reinterpret_cast<PrepSmemWarp*>(smem_raw + ((sizeof(double) * kTabSize + sizeof(BartParams) + sizeof(RngState) + 15) / 16) * 16);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kTabSize; i += kPrepWarps * 32) tab[i] = tables[i];
  if (tid == 0) { *prm = *dv.params; *rng = *dv.rng; }
  __syncthreads();
  const int T = prm->num_trees;
  const int t = blockIdx.x * kPrepWarps + warp;
  if (t >= T) return;
  PrepSmemWarp& W = pw[warp];
  {
    const DTree& g = dv.trees[t];
    int nn = g.num_nodes;
    if (lane == 0) { W.tree.num_nodes = nn; W.tree.pad = 0; W.cs.draws_total = 0; }
    for (int i = lane; i < nn * (int) (sizeof(DNode) / 4); i += 32) reinterpret_cast<uint32_t*>(W.tree.nodes)[i] = reinterpret_cast<const uint32_t*>(g.nodes)[i];
    __syncwarp();
  }
  const unsigned long long step = prm->step_id + (unsigned long long) t;
  WarpRng rngp; rngp.g = rng; rngp.cs = &W.cs; rngp.lane = lane; rngp.writer = false;
  rngp.enter(step, 0u); rngp.fill();
  w_propose(W.tree, *prm, tab, rngp, W.sd, W.cs, t, lane);
  rngp.commit();
  w_copy_desc(descs[t], W.sd, lane);
  // does this step fit the pipelined kernel?  one step that does not sends the whole sweep to the synchronous kernel
  if (infos != nullptr) {
    const bool ok = w_pipe_info(W.tree, W.sd, infos[t], pipe_max_cells, lane);
    if (!ok && lane == 0) {                                    // (diagnostics: steps that took the synchronous kernel, by reason)
      atomicAdd(pipe_not_ok, 1u);
      const bool bd = W.sd.b_kind == 0 || W.sd.b_kind == 1;
      if (W.tree.num_nodes + 2 > 32 || W.sd.b_nslots + (bd ? 1 : 0) > 32) atomicAdd(pipe_not_ok + 1, 1u);
      else if (W.sd.b_nslots > kPipeSlots) atomicAdd(pipe_not_ok + 2, 1u);
      else atomicAdd(pipe_not_ok + 3, 1u);
    }
  }
  // decision draws of this step: (uniform, normal) pairs for draw indices 0..31
  {
    const double u = keyed_stream_uniform(rng->key0, rng->key1, rng->stream, step, 1u, (uint32_t) lane);
    draws[t * 32 + lane] = make_double2(u, qnorm_as241(u));
    // draw 0 is the accept uniform of a birth / death / change / swap step: u < prior x exp(delta) <=> log u - log prior < delta
    if (lane == 0) descs[t].accept_thr = accept_threshold(u, W.sd.b_kind, W.sd.log_prior_trans);
  }
  if (lane == 0) atomicAdd(&dv.rng->counter, (unsigned long long) W.cs.draws_total);
}

// ---------------------------------------------------------------------------------------
// Streamed variant (STREAM = true): shards that do not fit the register file keep the residuals and the cached node
// indices in global memory (L2 resident up to a few million rows) and stream them through the same per-step phases in
// rounds of kWorkers quads; the binned predictors are read from global memory by the walks.  Every quad is owned by one
// thread for the whole sweep, so no synchronisation is needed for R or the node-index buffers.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void stream_walk(const StepDesc& sd, const uint32_t* __restrict__ xt32, int col_words, long long q_lo, long long q_hi, int tid,
                                            uint2* packs_out)
{
  for (long long q0 = q_lo; q0 < q_hi; q0 += kWorkers) {
    const long long q = q0 + tid;
    if (q < q_hi) {
      uint32_t lp[1], ap[1];
      walk_step<1>(sd, xt32 + q0, col_words, tid, 1u, lp, ap);
      packs_out[q] = make_uint2(lp[0], ap[0]);
    }
  }
}

// one observation into its lane-private bin: (sum, sum^2) as one 16-byte read-modify-write, or the sum alone (SQ = false:
// half the shared-memory traffic of the accumulation)
template <bool SQ>
__device__ __forceinline__ void bin_add(double2* __restrict__ bin_s, int idx, double pr)
{
  if (SQ) { double2 v = bin_s[idx]; v.x += pr; v.y = fma(pr, pr, v.y); bin_s[idx] = v; }
  else { double* b1 = reinterpret_cast<double*>(bin_s); b1[idx] += pr; }
}

// weighted observation (SQ layout only): (sum w r, sum w)
__device__ __forceinline__ void bin_add_w(double2* __restrict__ bin_s, int idx, double pr, double w)
{
  double2 v = bin_s[idx]; v.x = fma(w, pr, v.x); v.y += w; bin_s[idx] = v;
}

// register-resident variants: the residuals live in registers, the weights of a quad are re-read (L2) at every tree step
__device__ __forceinline__ void reg_quad_weights(const double* __restrict__ wt, long long q, unsigned int live, double (&w)[4])
{
  if (live == 0u) return;
  const double2 a = __ldg(reinterpret_cast<const double2*>(wt + 4 * q)), b = __ldg(reinterpret_cast<const double2*>(wt + 4 * q + 2));
  w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
}

// (no __restrict__ / read-only qualifiers on R and the node-index buffers: they are rewritten inside the same kernel, and a
// non-coherent load would return stale values)
template <bool SQ>
__device__ __forceinline__ void stream_acc_quad(const StepDesc& sd, const double2 ra, const double2 rb, const uint2 pk, long long q, long long n, int tid, int base,
                                                int kmax, bool two_trees, int birth_node, int L, double2* __restrict__ bin_s, unsigned long long& cpk,
                                                const bool weighted, const double2 wa, const double2 wb)
{
  const double r[4] = { ra.x, ra.y, rb.x, rb.y };
  const double w[4] = { wa.x, wa.y, wb.x, wb.y };
  double pr[4]; int row[4], row2[4];
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int leaf = (pk.x >> (8 * o)) & 0xFF;
    const int aux = (pk.y >> (8 * o)) & 0xFF;
    pr[o] = r[o] + sd.b_cur.val[leaf];
    const bool ok = 4 * q + o < n;
    const int sa = (two_trees ? (int) sd.b_cur.slot[leaf] : (leaf == birth_node ? L + aux : (int) sd.b_cur.slot[leaf])) - base;
    row[o] = ((unsigned) sa < (unsigned) kmax && ok) ? sa : kBinSlots;
    row2[o] = kBinSlots;
    if (two_trees) { const int sb = (int) sd.b_prop.slot[aux] - base; row2[o] = ((unsigned) sb < (unsigned) kmax && ok) ? sb : kBinSlots; }
  }
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    if (SQ && weighted) bin_add_w(bin_s, row[o] * kWorkers + tid, pr[o], w[o]); else bin_add<SQ>(bin_s, row[o] * kWorkers + tid, pr[o]);
    if (row[o] < kBinSlots) cpk += 1ull << (8 * row[o]);
    if (two_trees) {
      if (SQ && weighted) bin_add_w(bin_s, row2[o] * kWorkers + tid, pr[o], w[o]); else bin_add<SQ>(bin_s, row2[o] * kWorkers + tid, pr[o]);
      if (row2[o] < kBinSlots) cpk += 1ull << (8 * row2[o]);
    }
  }
}

__device__ __forceinline__ void stream_upd_quad(const UpdateDesc& upd, double2& ra, double2& rb, const uint2 pk, int amode, int unode);

// two rounds per iteration: all global loads of both rounds are in flight before the first is used.  `upd_prev` != nullptr:
// the previous step's residual update (node indices in packs_prev) is applied on the way, R is read and written once
template <bool SQ>
__device__ __forceinline__ void stream_accumulate(const StepDesc& sd, const UpdateDesc* upd_prev, double* Rg, const uint2* packs_prev, const uint2* packs,
                                                  long long q_lo, long long q_hi, long long n, int tid, int base, int kmax, double2* __restrict__ bin_s,
                                                  unsigned long long& cpk, const double* __restrict__ wt)
{
  const int kind = sd.b_kind, L = sd.b_num_leaves;
  const bool two_trees = (kind == 2 || kind == 3);
  const int birth_node = kind == 0 ? sd.b_node : -1;
  const bool fused = upd_prev != nullptr;
  const bool weighted = wt != nullptr;
  const int pmode = fused ? upd_prev->mode : 0, pnode = fused ? upd_prev->node : 0;
  for (long long q0 = q_lo; q0 < q_hi; q0 += 2 * kWorkers) {
    const long long qa = q0 + tid, qb = qa + kWorkers;
    const bool la = qa < q_hi, lb = qb < q_hi;
    double2 a0 = make_double2(0.0, 0.0), a1 = a0, b0 = a0, b1 = a0; uint2 pa = make_uint2(0u, 0u), pb = pa, ua = pa, ub = pa;
    double2 wa0 = make_double2(1.0, 1.0), wa1 = wa0, wb0 = wa0, wb1 = wa0;
    if (la) { a0 = *reinterpret_cast<const double2*>(Rg + 4 * qa); a1 = *reinterpret_cast<const double2*>(Rg + 4 * qa + 2); pa = packs[qa]; if (fused) ua = packs_prev[qa]; }
    if (lb) { b0 = *reinterpret_cast<const double2*>(Rg + 4 * qb); b1 = *reinterpret_cast<const double2*>(Rg + 4 * qb + 2); pb = packs[qb]; if (fused) ub = packs_prev[qb]; }
    if (SQ && weighted) {
      if (la) { wa0 = __ldg(reinterpret_cast<const double2*>(wt + 4 * qa)); wa1 = __ldg(reinterpret_cast<const double2*>(wt + 4 * qa + 2)); }
      if (lb) { wb0 = __ldg(reinterpret_cast<const double2*>(wt + 4 * qb)); wb1 = __ldg(reinterpret_cast<const double2*>(wt + 4 * qb + 2)); }
    }
    if (la) {
      if (fused) { stream_upd_quad(*upd_prev, a0, a1, ua, pmode, pnode); *reinterpret_cast<double2*>(Rg + 4 * qa) = a0; *reinterpret_cast<double2*>(Rg + 4 * qa + 2) = a1; }
      stream_acc_quad<SQ>(sd, a0, a1, pa, qa, n, tid, base, kmax, two_trees, birth_node, L, bin_s, cpk, weighted, wa0, wa1);
    }
    if (lb) {
      if (fused) { stream_upd_quad(*upd_prev, b0, b1, ub, pmode, pnode); *reinterpret_cast<double2*>(Rg + 4 * qb) = b0; *reinterpret_cast<double2*>(Rg + 4 * qb + 2) = b1; }
      stream_acc_quad<SQ>(sd, b0, b1, pb, qb, n, tid, base, kmax, two_trees, birth_node, L, bin_s, cpk, weighted, wb0, wb1);
    }
  }
}

__device__ __forceinline__ void stream_upd_quad(const UpdateDesc& upd, double2& ra, double2& rb, const uint2 pk, int amode, int unode)
{
  double r[4] = { ra.x, ra.y, rb.x, rb.y };
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    const int leaf = (pk.x >> (8 * o)) & 0xFF;
    const int aux = (pk.y >> (8 * o)) & 0xFF;
    if (amode == 0) r[o] += upd.delta[leaf];
    else {
      int nl;
      if (amode == 3) nl = aux;
      else if (amode == 1 && leaf == unode) nl = unode + 1 + aux;
      else nl = upd.remap[leaf];
      r[o] += upd.val_old[leaf] - upd.val_new[nl];
    }
  }
  ra = make_double2(r[0], r[1]); rb = make_double2(r[2], r[3]);
}

__device__ __forceinline__ void stream_update(const UpdateDesc& upd, double* Rg, const uint2* packs, long long q_lo, long long q_hi, int tid)
{
  const int amode = upd.mode, unode = upd.node;
  for (long long q0 = q_lo; q0 < q_hi; q0 += 2 * kWorkers) {
    const long long qa = q0 + tid, qb = qa + kWorkers;
    const bool la = qa < q_hi, lb = qb < q_hi;
    double2 a0 = make_double2(0.0, 0.0), a1 = a0, b0 = a0, b1 = a0; uint2 pa = make_uint2(0u, 0u), pb = pa;
    if (la) { a0 = *reinterpret_cast<const double2*>(Rg + 4 * qa); a1 = *reinterpret_cast<const double2*>(Rg + 4 * qa + 2); pa = packs[qa]; }
    if (lb) { b0 = *reinterpret_cast<const double2*>(Rg + 4 * qb); b1 = *reinterpret_cast<const double2*>(Rg + 4 * qb + 2); pb = packs[qb]; }
    if (la) { stream_upd_quad(upd, a0, a1, pa, amode, unode); *reinterpret_cast<double2*>(Rg + 4 * qa) = a0; *reinterpret_cast<double2*>(Rg + 4 * qa + 2) = a1; }
    if (lb) { stream_upd_quad(upd, b0, b1, pb, amode, unode); *reinterpret_cast<double2*>(Rg + 4 * qb) = b0; *reinterpret_cast<double2*>(Rg + 4 * qb + 2) = b1; }
  }
}

// SEQ: replay / record keep the strict program order of the draws (proposals and draws produced inside the loop);
// the production instantiation (SEQ = false) carries none of that code
// SQ: also accumulate the per-slot sums of squares (needed only to report the individual log-likelihoods in the parity
// trace; the Metropolis ratios do not depend on them, see slot_summary)
// The body of the sweep: everything of a chain comes in through its arguments, and the grid is used through blockIdx.x / gridDim.x only,
// so the same code serves one chain per launch (k_sweep) and several chains in one launch, one per blockIdx.y (k_sweep_batch)
template <int NQ, bool SEQ, bool STREAM = false, bool SQ = true>
__device__ __forceinline__ void sweep_body(const BartDev& dv, unsigned int* barrier_counter, int partial_stride, const double* __restrict__ tables,
                                           const StepDesc* __restrict__ descs, const double2* __restrict__ draws, int overlap_walk,
                                           const ShardDev& sh_param, const int* __restrict__ pos_in,
                                           int* __restrict__ pos_out, int max_steps)
{
  // A sweep can be split into segments of consecutive tree steps handled by alternating launches of the pipelined kernel
  // (sweep_pipe.cuh: runs of steps that fit it) and of this one (the steps that do not): *pos_in = first step still to do,
  // max_steps = how many this launch may take.  pos_in == nullptr: the whole sweep.
  if (pos_in != nullptr && *pos_in >= dv.params->num_trees) { if (blockIdx.x == 0 && threadIdx.x == 0) *pos_out = *pos_in; return; }
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SweepSmem& S = *reinterpret_cast<SweepSmem*>(smem_raw);
  // lane-private statistic bins: [slot][thread] -> (sum, sum^2) and count; no atomics, no bank conflicts
  double2* bin_s = reinterpret_cast<double2*>(smem_raw + ((sizeof(SweepSmem) + 15) / 16) * 16);
  unsigned long long* bin_c = reinterpret_cast<unsigned long long*>(bin_s + (kBinSlots + 1) * kWorkers);   // packed per-thread counts
  uint32_t* tile = reinterpret_cast<uint32_t*>(bin_c + kWorkers);                                      // [p][NQ * kWorkers]
  constexpr int tile_stride = NQ * kWorkers;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_worker = tid < kWorkers;            // warps 0..14: observations; warp 15: controller (proposals + decisions)
  const int G = gridDim.x, cta = blockIdx.x;
  const long long n = dv.n, npad = dv.npad;
  const long long nquad = (n + 3) >> 2;
  // balanced partition: CTA c owns the contiguous quads [c nquad / G, (c + 1) nquad / G), dealt to its threads in rounds of
  // kWorkers (coalesced); every SM then moves the same number of observations through its shared-memory bins
  const long long q_lo = nquad * cta / G, q_hi = nquad * (cta + 1) / G;

  // ---- one-time loads: residuals -> registers, binned predictors -> shared tile, controller state ----
  double R[NQ][4];
  unsigned valid_mask = 0;       // bit j: quad j exists
  unsigned obs_mask = 0;         // bit 4j + o: observation is a real one (not padding)
#pragma unroll
  for (int j = 0; j < NQ; ++j) {
    const long long q = q_lo + (long long) j * kWorkers + tid;
    if (!STREAM && is_worker && q < q_hi) {
      valid_mask |= 1u << j;
      for (int o = 0; o < 4; ++o) if (4 * q + o < n) obs_mask |= 1u << (4 * j + o);
      double2 a = *reinterpret_cast<const double2*>(dv.R + 4 * q), b = *reinterpret_cast<const double2*>(dv.R + 4 * q + 2);
      R[j][0] = a.x; R[j][1] = a.y; R[j][2] = b.x; R[j][3] = b.y;
    } else { R[j][0] = R[j][1] = R[j][2] = R[j][3] = 0.0; }
  }
  if (tid == 0) {
    S.sh.rank = sh_param.rank; S.sh.world = sh_param.world; S.sh.obs_offset = sh_param.obs_offset;
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r) S.sh.mail[r] = sh_param.mail[r];
    S.seq_base = sh_param.world > 1 ? sh_param.mail[sh_param.rank]->kseq : 0ull;
  }
  const int world = sh_param.world;
  if (tid == 0) { const double sg = dv.params->sigma; S.inv_sigsq = 1.0 / (sg * sg); for (int i = 0; i < 6; ++i) S.pc[i] = 0; for (int i = 0; i < 4; ++i) S.wk[i] = 0; }
  if (tid == 0) { S.peer_dead = 0; S.prm = *dv.params; S.rng = *dv.rng; S.csd.draws_total = 0; S.csp.draws_total = 0; S.csd.prof_on = dv.prof != nullptr ? 1 : 0; S.csp.prof_on = 0; for (int i = 0; i < 8; ++i) S.csd.dbg[i] = 0; for (int i = 0; i < 4; ++i) S.csd.fine[i] = 0; }
  for (int i = tid; i < kTabSize; i += kSweepBlock) S.tab[i] = tables[i];
  for (int i = tid; i < 3 * S4B_MAX_SLOTS; i += kSweepBlock) reinterpret_cast<double*>(S.st)[i] = 0.0;
  __syncthreads();
  const int p = S.prm.p, T = S.prm.num_trees;
  const int t_begin = pos_in != nullptr ? *pos_in : 0;
  const int t_end = pos_in != nullptr ? min(T, t_begin + max_steps) : T;
  const unsigned long long step0 = S.prm.step_id;
  // replay / record need strict program order: proposals and draws are then produced inside the loop
  constexpr bool sequential_rng = SEQ;
  const uint32_t* xt32 = reinterpret_cast<const uint32_t*>(dv.xt);
  const int col_words = (int) (npad >> 2);
  if (!STREAM && is_worker) {
    for (int v = 0; v < p; ++v)
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        const long long q = q_lo + (long long) j * kWorkers + tid;
        tile[v * tile_stride + j * kWorkers + tid] = ((valid_mask >> j) & 1u) ? __ldg(xt32 + (long long) v * col_words + q) : 0u;
      }
  }
  {
    const DTree& g = dv.trees[t_begin];
    int nn = g.num_nodes;
    if (tid == 0) { S.tree[t_begin & 1].num_nodes = nn; S.tree[t_begin & 1].pad = 0; }
    for (int i = tid; i < nn * (int) (sizeof(DNode) / 4); i += kSweepBlock) reinterpret_cast<uint32_t*>(S.tree[t_begin & 1].nodes)[i] = reinterpret_cast<const uint32_t*>(g.nodes)[i];
  }
  __syncthreads();
  WarpRng rngp; rngp.g = &S.rng; rngp.cs = &S.csp; rngp.lane = lane; rngp.writer = cta == 0;     // proposal draws
  WarpRng rngd; rngd.g = &S.rng; rngd.cs = &S.csd; rngd.lane = lane; rngd.writer = cta == 0;     // decision draws
  if (!is_worker) {
    if (sequential_rng) {
      rngp.enter(step0 + (unsigned long long) t_begin, 0u); rngp.fill();
      w_propose(S.tree[t_begin & 1], S.prm, S.tab, rngp, S.sd[t_begin & 1], S.csp, t_begin, lane);
      rngp.commit();
    } else w_copy_desc(S.sd[t_begin & 1], descs[t_begin], lane);
  }
  __syncthreads();
  uint32_t leaf_pack[NQ], aux_pack[NQ];
  if (is_worker) {
    if (STREAM) stream_walk(S.sd[t_begin & 1], xt32, col_words, q_lo, q_hi, tid, dv.packs + (size_t) (t_begin & 1) * (size_t) nquad);
    else walk_step<NQ>(S.sd[t_begin & 1], tile, tile_stride, tid, valid_mask, leaf_pack, aux_pack);
  }

  // phase counters live in shared memory (thread 0, profiling runs only): as registers they would be carried through the
  // whole loop by every thread.  S.wk[0..3]: worker sub-phases (zero bins, accumulate, wait + row reduce, second barrier)
  const bool prof_on = dv.prof != nullptr;
  for (int t = t_begin; t < t_end; ++t) {
    const long long c0 = clock64();
    StepDesc& sd = S.sd[t & 1];
    StepDesc& sd_next = S.sd[(t + 1) & 1];
    DTree& tree = S.tree[t & 1];
    DTree& tree_next = S.tree[(t + 1) & 1];
    const int kind = sd.b_kind;
    const int L = sd.b_num_leaves;
    const int nslots = sd.b_nslots;
    const bool two_trees = (kind == 2 || kind == 3);
    const int birth_node = kind == 0 ? sd.b_node : -1;
    double* partials = dv.partials + (size_t) (t & 1) * partial_stride;

    if (is_worker) {
      // ---- accumulate (n, sum, sum^2) per statistic slot from the cached leaf indices ----
      const int nchunks = (nslots + kBinSlots - 1) / kBinSlots;
      for (int chunk = 0; chunk < nchunks; ++chunk) {
        const int base = chunk * kBinSlots;
        const int kmax = min(kBinSlots, nslots - base);
        const long long w0 = clock64();
        if (SQ) { for (int k = 0; k < kmax; ++k) bin_s[k * kWorkers + tid] = make_double2(0.0, 0.0); }
        else { for (int k = 0; k < kmax; ++k) reinterpret_cast<double*>(bin_s)[k * kWorkers + tid] = 0.0; }
        // observation counts stay in a register: one 6-bit field per bin row (a thread adds at most 32 to a row)
        unsigned long long cpk = 0ull;
        const bool weighted = dv.wt != nullptr;
        const long long w1 = clock64();
        // branch-free: every observation adds into exactly one bin row (row kBinSlots is a trash row for padding and for
        // slots outside this pass); loads first (independent), then the read-modify-write chain
        if (STREAM) {
          stream_accumulate<SQ>(sd, (t > t_begin && chunk == 0) ? &S.upd[(t - 1) & 1] : nullptr, dv.R, dv.packs + (size_t) ((t + 1) & 1) * (size_t) nquad,
                            dv.packs + (size_t) (t & 1) * (size_t) nquad, q_lo, q_hi, n, tid, base, kmax, bin_s, cpk, dv.wt);
        } else if (!two_trees) {
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            double pr[4]; int row[4];
            double wq[4] = { 1.0, 1.0, 1.0, 1.0 };
            if (SQ && weighted) reg_quad_weights(dv.wt, q_lo + (long long) j * kWorkers + tid, (obs_mask >> (4 * j)) & 0xFu, wq);
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              const int leaf = (leaf_pack[j] >> (8 * o)) & 0xFF;
              const int aux = (aux_pack[j] >> (8 * o)) & 0xFF;
              pr[o] = R[j][o] + sd.b_cur.val[leaf];
              const int sa = (leaf == birth_node ? L + aux : (int) sd.b_cur.slot[leaf]) - base;
              row[o] = ((unsigned) sa < (unsigned) kmax && ((obs_mask >> (4 * j + o)) & 1u)) ? sa : kBinSlots;
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              if (SQ && weighted) bin_add_w(bin_s, row[o] * kWorkers + tid, pr[o], wq[o]); else bin_add<SQ>(bin_s, row[o] * kWorkers + tid, pr[o]);
              cpk += 1ull << (6 * row[o]);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            double pr[4]; int row[4], row2[4];
            double wq[4] = { 1.0, 1.0, 1.0, 1.0 };
            if (SQ && weighted) reg_quad_weights(dv.wt, q_lo + (long long) j * kWorkers + tid, (obs_mask >> (4 * j)) & 0xFu, wq);
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              const int leaf = (leaf_pack[j] >> (8 * o)) & 0xFF;
              const int aux = (aux_pack[j] >> (8 * o)) & 0xFF;
              pr[o] = R[j][o] + sd.b_cur.val[leaf];
              const bool ok = (obs_mask >> (4 * j + o)) & 1u;
              const int sa = (int) sd.b_cur.slot[leaf] - base;
              const int sb = (int) sd.b_prop.slot[aux] - base;
              row[o] = ((unsigned) sa < (unsigned) kmax && ok) ? sa : kBinSlots;
              row2[o] = ((unsigned) sb < (unsigned) kmax && ok) ? sb : kBinSlots;
            }
#pragma unroll
            for (int o = 0; o < 4; ++o) {
              if (SQ && weighted) { bin_add_w(bin_s, row[o] * kWorkers + tid, pr[o], wq[o]); bin_add_w(bin_s, row2[o] * kWorkers + tid, pr[o], wq[o]); }
              else { bin_add<SQ>(bin_s, row[o] * kWorkers + tid, pr[o]); bin_add<SQ>(bin_s, row2[o] * kWorkers + tid, pr[o]); }
              cpk += (1ull << (6 * row[o])) + (1ull << (6 * row2[o]));
            }
          }
        }
        bin_c[tid] = cpk;
        const long long w2 = clock64();
        named_bar_sync(1, kWorkers);
        // row-wise reduction: task r < kmax sums (sum, sum^2) of slot r over the CTA's threads, task kmax + r its counts
        for (int task = warp; task < 2 * kmax; task += kWorkerWarps) {
          if (task < kmax) {
            if (SQ) {
              double a = 0.0, b = 0.0;
#pragma unroll
              for (int i = 0; i < kWorkerWarps; ++i) { const double2 v = bin_s[task * kWorkers + i * 32 + lane]; a += v.x; b += v.y; }
              a = w_sum(a); b = w_sum(b);
              if (lane == 0) { partials[(size_t) (3 * (base + task) + 1) * G + cta] = a; partials[(size_t) (3 * (base + task) + 2) * G + cta] = b; }
            } else {
              double a = 0.0;
#pragma unroll
              for (int i = 0; i < kWorkerWarps; ++i) a += reinterpret_cast<const double*>(bin_s)[task * kWorkers + i * 32 + lane];
              a = w_sum(a);
              if (lane == 0) partials[(size_t) (3 * (base + task) + 1) * G + cta] = a;
            }
          } else {
            const int r = task - kmax;
            int c = 0;
#pragma unroll
            for (int i = 0; i < kWorkerWarps; ++i) c += STREAM ? (int) ((bin_c[i * 32 + lane] >> (8 * r)) & 255ull) : (int) ((bin_c[i * 32 + lane] >> (6 * r)) & 63ull);
            c = __reduce_add_sync(0xffffffffu, c);
            if (lane == 0) partials[(size_t) (3 * (base + r)) * G + cta] = (double) c;
          }
        }
        const long long w3 = clock64();
        if (chunk + 1 < nchunks) named_bar_sync(1, kWorkers);        // the bins are reused by the next pass
        if (tid == 0 && prof_on) { S.wk[0] += w1 - w0; S.wk[1] += w2 - w1; S.wk[2] += w3 - w2; S.wk[3] += clock64() - w3; }
      }
      // ---- arrive at the grid barrier and wait for every CTA's partial rows.  The rows are stored by lane 0 of
      // several warps: all of them must have issued their stores before thread 0 publishes them (fence + arrive) ----
      named_bar_sync(1, kWorkers);
      if (tid == 0) {
        if (prof_on) S.pc[0] += clock64() - c0;
        // release-arrive: orders this CTA's partial rows (made visible to thread 0 by the named barrier) before the count
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(barrier_counter) : "memory");
        const unsigned int target = (unsigned int) (t - t_begin + 1) * (unsigned int) G;
        unsigned int v;
        const long long w0 = clock64();
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(barrier_counter) : "memory");
          if (v < target && (S.peer_dead || clock64() - w0 > 4000000000LL)) { S.peer_dead = 2; break; }     // a CTA never arrived: fail, do not hang
        } while (v < target);
      }
    } else {
      // ---- controller warp: plan this step's decision, fetch tree t+1, pre-compute this step's decision draws, propose for t+1 ----
      { const FastPlan pl = w_plan(tree, sd, S.upd[t & 1], S.csd, lane); plan_store(S.plan, pl, lane); }
      const long long h0 = clock64();
      if (t + 1 < t_end) {
        const DTree& g = dv.trees[t + 1];
        int nn = g.num_nodes;
        if (lane == 0) { tree_next.num_nodes = nn; tree_next.pad = 0; }
        for (int i = lane; i < nn * (int) (sizeof(DNode) / 4); i += 32) reinterpret_cast<uint32_t*>(tree_next.nodes)[i] = reinterpret_cast<const uint32_t*>(g.nodes)[i];
        __syncwarp();
      }
      const long long h1 = clock64();
      rngd.enter(step0 + (unsigned long long) t, 1u);
      if (!sequential_rng) {
        // pre-computed by k_prepare_sweep: this step's decision draws and the next tree's proposal
        const double2 dz = __ldcg(draws + t * 32 + lane);
        S.csd.ubuf[lane] = dz.x; S.csd.zbuf[lane] = dz.y;
        rngd.adopt();
        const long long h2 = clock64();
        if (t + 1 < t_end) w_copy_desc(sd_next, descs[t + 1], lane);
        if (lane == 0 && S.csd.prof_on) { S.csd.dbg[4] += h1 - h0; S.csd.dbg[5] += h2 - h1; S.csd.dbg[7] += clock64() - h2; }
      }
    }
    __syncthreads();                                                        // [A] partials complete everywhere, next proposal ready
    const long long c2 = clock64();

    // ---- every CTA: reduce all partial rows in the same fixed order ----
    for (int w = warp; w < (SQ ? 3 : 2) * nslots; w += kSweepWarps) {
      const int v = SQ ? w : (w >> 1) * 3 + (w & 1);          // without sums of squares only the (n, sum) rows exist
      const double* src = partials + (size_t) v * G;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, a4 = 0.0;
      { int b = lane;       if (b < G) a0 = __ldcg(src + b); }
      { int b = lane + 32;  if (b < G) a1 = __ldcg(src + b); }
      { int b = lane + 64;  if (b < G) a2 = __ldcg(src + b); }
      { int b = lane + 96;  if (b < G) a3 = __ldcg(src + b); }
      { int b = lane + 128; if (b < G) a4 = __ldcg(src + b); }
      double acc = (((a0 + a1) + a2) + a3) + a4;
      for (int b = lane + 160; b < G; b += 32) acc += __ldcg(src + b);
      acc = w_sum(acc);
      if (lane == 0) reinterpret_cast<double*>(&S.st[v / 3])[v % 3] = acc;
    }
    __syncthreads();                                                        // [B] statistics in shared memory
    const long long cx = clock64();
    if (world > 1) {
      const ShardDev& sh = S.sh;
      // ---- observation-sharded chain: exchange the per-slot statistics with the peer GPUs (shard.hpp).  CTA 0 stores
      // this rank's sums straight into every peer's mailbox over NVLink, each double as two flag-carrying 8-byte words
      // (no fence, one hop); every CTA polls the local mailbox and adds the contributions in rank order, so all CTAs
      // of all ranks hold bitwise identical statistics and take the same decision ----
      const unsigned long long seq = S.seq_base + (unsigned long long) t + 1ull;
      const unsigned int seq32 = (unsigned int) seq;
      const int par = (int) (seq & 1ull);
      const int cnt = 3 * nslots;
      double* stv = reinterpret_cast<double*>(S.st);
      double mine_k = 0.0;
      if (tid < cnt) mine_k = stv[tid];
      if (cta == 0) {
        for (int i = tid; i < world * cnt; i += kSweepBlock) {
          const int dst = i / cnt, k = i - dst * cnt;
          mailbox_send_ll(&sh.mail[dst]->step_ll[par][sh.rank][k], stv[k], seq32);
        }
      }
      __syncthreads();                                   // everyone has read the local sums before they are overwritten
      const Mailbox* mine = sh.mail[sh.rank];
      for (int i = tid; i < cnt; i += kSweepBlock) {
        double acc = 0.0;
        for (int src = 0; src < world; ++src) {
          double v = 0.0;
          if (src == sh.rank && i == tid) v = mine_k;    // own contribution needs no round trip
          else if (!S.peer_dead && !mailbox_recv_ll(&mine->step_ll[par][src][i], seq32, &v)) S.peer_dead = 1;
          acc = src == 0 ? v : acc + v;
        }
        stv[i] = acc;
      }
      __syncthreads();
    }
    const long long c3 = clock64();
    uint32_t leaf_next[NQ], aux_next[NQ];
    if (!is_worker) {
      // ---- controller: Metropolis decision + leaf draws for tree t ----
      double* trec = nullptr;
      if (SQ && dv.trace != nullptr && cta == 0) {
        unsigned long long k = *dv.trace_len;
        if (k < dv.trace_cap) { trec = dv.trace + k * S4B_TRACE_LEN; for (int i = lane; i < S4B_TRACE_LEN; i += 32) trec[i] = 0.0; }
        __syncwarp();
        if (lane == 0) *dv.trace_len = k + 1;
      }
      if (sequential_rng) rngd.fill();
      // replay / record draw in program order: the threshold is formed here; otherwise k_prepare_sweep stored it with the proposal
      const double thr = sequential_rng ? accept_threshold(S.csd.ubuf[0], sd.b_kind, sd.log_prior_trans) : sd.accept_thr;
      if (S.plan.valid) { const FastPlan plan = plan_load(S.plan, lane); w_decide_fast<SQ>(plan, tree, S.prm, rngd, sd, S.st, S.upd[t & 1], S.csd, trec, lane, S.inv_sigsq, thr); }
      else w_decide<SQ>(tree, S.prm, rngd, sd, S.st, S.upd[t & 1], S.csd, trec, lane, S.inv_sigsq, thr);
      rngd.commit();
      if (sequential_rng && t + 1 < t_end) {
        rngp.enter(step0 + (unsigned long long) (t + 1), 0u); rngp.fill();
        w_propose(tree_next, S.prm, S.tab, rngp, sd_next, S.csp, t + 1, lane);
        rngp.commit();
      }
    } else if (!sequential_rng && overlap_walk && t + 1 < t_end) {
      // ---- workers, concurrently: walk tree t+1 (independent of this step's decision) ----
      if (STREAM) stream_walk(sd_next, xt32, col_words, q_lo, q_hi, tid, dv.packs + (size_t) ((t + 1) & 1) * (size_t) nquad);
      else walk_step<NQ>(sd_next, tile, tile_stride, tid, valid_mask, leaf_next, aux_next);
    }
    __syncthreads();                                                        // [C] decision known
    const long long c5 = clock64();
    if (is_worker) {
      // ---- fit / residual update from the cached leaf indices ----
      const UpdateDesc& upd = S.upd[t & 1];
      const int amode = upd.mode, unode = upd.node;
      if (STREAM) {
        // applied together with the next step's accumulation (one read and one write of R per step); the last step has no successor
        if (t + 1 == t_end) stream_update(upd, dv.R, dv.packs + (size_t) (t & 1) * (size_t) nquad, q_lo, q_hi, tid);
      } else if (amode == 0) {
#pragma unroll
        for (int j = 0; j < NQ; ++j)
#pragma unroll
          for (int o = 0; o < 4; ++o) R[j][o] += upd.delta[(leaf_pack[j] >> (8 * o)) & 0xFF];
      } else {
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            const int leaf = (leaf_pack[j] >> (8 * o)) & 0xFF;
            const int aux = (aux_pack[j] >> (8 * o)) & 0xFF;
            int nl;
            if (amode == 3) nl = aux;
            else if (amode == 1 && leaf == unode) nl = unode + 1 + aux;
            else nl = upd.remap[leaf];
            R[j][o] += upd.val_old[leaf] - upd.val_new[nl];
          }
        }
      }
      if (STREAM) {
        if (t + 1 < t_end && (sequential_rng || !overlap_walk)) stream_walk(sd_next, xt32, col_words, q_lo, q_hi, tid, dv.packs + (size_t) ((t + 1) & 1) * (size_t) nquad);
      } else if (t + 1 < t_end) {
        if (sequential_rng || !overlap_walk) walk_step<NQ>(sd_next, tile, tile_stride, tid, valid_mask, leaf_pack, aux_pack);
        else {
#pragma unroll
          for (int j = 0; j < NQ; ++j) { leaf_pack[j] = leaf_next[j]; aux_pack[j] = aux_next[j]; }
        }
      }
    } else if (cta == 0) {
      // controller of CTA 0 persists the tree that was just decided
      DTree& g = dv.trees[t];
      int nn = tree.num_nodes;
      if (lane == 0) g.num_nodes = nn;
      for (int i = lane; i < nn * (int) (sizeof(DNode) / 4); i += 32) reinterpret_cast<uint32_t*>(g.nodes)[i] = reinterpret_cast<const uint32_t*>(tree.nodes)[i];
    }
    // (no barrier here: the update tables are double buffered, and the tree / descriptor buffers of step t are touched by
    // the controller alone from now on, so workers run straight into the next accumulation)
    if (tid == 0 && prof_on) { S.pc[1] += c2 - c0; S.pc[2] += cx - c2; S.pc[4] += c3 - cx; S.pc[3] += c5 - c3; S.pc[5] += clock64() - c5; }
  }

  // ---- write back ----
#pragma unroll
  for (int j = 0; j < NQ; ++j) if ((valid_mask >> j) & 1u) {
    const long long q = q_lo + (long long) j * kWorkers + tid;
    *reinterpret_cast<double2*>(dv.R + 4 * q) = make_double2(R[j][0], R[j][1]);
    *reinterpret_cast<double2*>(dv.R + 4 * q + 2) = make_double2(R[j][2], R[j][3]);
  }
  if (cta == 0 && tid == 0) {
    RngState out = S.rng;
    out.counter += (unsigned long long) (S.csd.draws_total + S.csp.draws_total);
    *dv.rng = out;
    if (out.tape_underrun) dv.params->error_flag |= 2u;
    if (S.peer_dead) dv.params->error_flag |= (S.peer_dead == 2 ? 8u : 4u);
    if (t_end == T) dv.params->step_id = step0 + (unsigned long long) T;
    if (pos_out != nullptr) *pos_out = t_end;
    if (world > 1) S.sh.mail[S.sh.rank]->kseq = S.seq_base + (unsigned long long) T;
    dv.desc->a_valid = 0;
    if (dv.prof != nullptr) {
      // [0] accumulate + CTA reduction (CTA 0), [1] ... + grid barrier + controller wait, [2] statistics reduce,
      // [3] decision (overlapped with the next walk), [4] cross-rank exchange (sharded chains), [5] update, [7] steps
      for (int i = 0; i < 6; ++i) dv.prof[i] += (unsigned long long) S.pc[i];
      dv.prof[7] += (unsigned long long) (t_end - t_begin);
      // [8..11] decision: slot summaries + accept, structure, leaf draws, update descriptor; [12..15] controller before the
      // barrier: tree fetch, decision-draw prefill, proposal-draw prefill, proposal
      for (int i = 0; i < 8; ++i) dv.prof[8 + i] += (unsigned long long) S.csd.dbg[i];
      for (int i = 0; i < 4; ++i) dv.prof[16 + i] += (unsigned long long) S.wk[i];
      for (int i = 0; i < 4; ++i) dv.prof[20 + i] += (unsigned long long) S.csd.fine[i];
    }
  }
}

template <int NQ, bool SEQ, bool STREAM = false, bool SQ = true>
__global__ void __launch_bounds__(kSweepBlock, 1) k_sweep(const __grid_constant__ BartDev dv, unsigned int* barrier_counter, int partial_stride, const double* __restrict__ tables,
                                                               const StepDesc* __restrict__ descs, const double2* __restrict__ draws, int overlap_walk,
                                                               const __grid_constant__ ShardDev sh_param, const int* __restrict__ pos_in,
                                                               int* __restrict__ pos_out, int max_steps)
{
  sweep_body<NQ, SEQ, STREAM, SQ>(dv, barrier_counter, partial_stride, tables, descs, draws, overlap_walk, sh_param, pos_in, pos_out, max_steps);
}

// Several chains in ONE cooperative launch (grid.y = chain): chain c = blockIdx.y runs its sweep on the gridDim.x
// CTAs of its row with its own state, barrier counter and partial rows; the chains share nothing but the launch.  All chains of a
// batch use the same instantiation (same rows-per-thread layout and shared-memory size) and gridDim.x * gridDim.y <= the SM count.
struct SweepBatchArgs {
  BartDev dv; unsigned int* barrier_counter; int partial_stride; const double* tables; const StepDesc* descs; const double2* draws;
  int overlap_walk; ShardDev sh; int max_steps;
};
template <int NQ, bool STREAM = false>
__global__ void __launch_bounds__(kSweepBlock, 1) k_sweep_batch(const SweepBatchArgs* __restrict__ chains)
{
  __shared__ SweepBatchArgs a;
  for (int i = threadIdx.x; i < (int) (sizeof(SweepBatchArgs) / 4); i += kSweepBlock) reinterpret_cast<uint32_t*>(&a)[i] = reinterpret_cast<const uint32_t*>(chains + blockIdx.y)[i];
  __syncthreads();
  sweep_body<NQ, false, STREAM, false>(a.dv, a.barrier_counter, a.partial_stride, a.tables, a.descs, a.draws, a.overlap_walk, a.sh, nullptr, nullptr, a.max_steps);
}

}  // namespace s4b

// stan4bart_b200/csrc/bart_kernels.cuh
// Hand-written sm_100a kernels for the BART half of the Gibbs sweep (SURVEY.md 8a rows
// a2-a10; the reference reaches this code through dbarts' runSamplerWithResults,
// /root/reference/src/init.cpp:824).
//
// Data model: per chain we keep the FULL residual R = y* - sum_t fit_t
// (fp64[N]) and the binned predictors xt (u8 [P][Npad], column major).  Node membership is
// never stored: every pass re-derives it by walking the (tiny) tree held in shared memory,
// which makes the observation -> node partition a pure function of (tree, xt) -- bit exact by
// construction -- and removes dbarts' T x N treeFits array and index partitions.
//
// One kernel per tree step:
//   k_tree_step  (a) applies the fit/residual update of tree t-1 (R += mu_old - mu_new),
//                (b) accumulates per-leaf (n, sum, sum^2) of the partial residual of tree t
//                    under its *proposed* structure (warp-shuffle segmented fp64 reduction,
//                    fixed order => run-to-run deterministic),
//                (c) the last block to finish reduces the per-block partials, takes the
//                    Metropolis decision for tree t, draws its leaf values and draws the
//                    proposal for tree t+1 -- no host round trip inside a sweep.
#pragma once

#include "s4b_common.cuh"

namespace s4b {

// --------------------------------------------------------------------------------------
// controller-side tree helpers (single thread, tree in shared memory)
// --------------------------------------------------------------------------------------
__device__ __forceinline__ bool t_is_leaf(const DTree& t, int i) { return t.nodes[i].var < 0; }

// cut points of predictor `var` (bart_args n.cuts may name one count per predictor)
__device__ __forceinline__ int s4b_ncuts(const BartParams& P, int var) { return P.ncuts_var != nullptr ? P.ncuts_var[var] : P.n_cuts; }

__device__ inline void t_split_interval(const DTree& t, int n_cuts, int i, int var, int& lo, int& hi)
{
  lo = 0; hi = n_cuts - 1;
  int child = i, par = t.nodes[i].parent;
  while (par >= 0) {
    if (t.nodes[par].var == var) {
      int c = t.nodes[par].cut;
      if (child == par + 1) { if (c - 1 < hi) hi = c - 1; }
      else                  { if (c + 1 > lo) lo = c + 1; }
    }
    child = par; par = t.nodes[par].parent;
  }
}

// number of predictors that still have a cut available below node i (data independent); with split weights only
// predictors of positive weight count
__device__ inline int t_num_vars_available(const DTree& t, const BartParams& P, int i)
{
  int blocked = 0;
  int par = t.nodes[i].parent;
  // only variables used by an ancestor can be exhausted; visit each distinct one once
  while (par >= 0) {
    int v = t.nodes[par].var;
    bool seen = false;
    for (int a = t.nodes[i].parent; a != par; a = t.nodes[a].parent) if (t.nodes[a].var == v) { seen = true; break; }
    if (!seen && (P.split_w == nullptr || P.split_w[v] != 0u)) { int lo, hi; t_split_interval(t, s4b_ncuts(P, v), i, v, lo, hi); if (hi < lo) ++blocked; }
    par = t.nodes[par].parent;
  }
  return (P.split_w != nullptr ? P.p_pos : P.p) - blocked;
}

__device__ inline bool t_var_available(const DTree& t, const BartParams& P, int i, int j)
{
  if (P.split_w != nullptr && P.split_w[j] == 0u) return false;
  bool used = false;
  for (int a = t.nodes[i].parent; a >= 0; a = t.nodes[a].parent) if (t.nodes[a].var == j) { used = true; break; }
  if (!used) return true;
  int lo, hi; t_split_interval(t, s4b_ncuts(P, j), i, j, lo, hi);
  return hi >= lo;
}

__device__ inline int t_ith_available_var(const DTree& t, const BartParams& P, int i, int ith)
{
  for (int j = 0; j < P.p; ++j) if (t_var_available(t, P, i, j)) { if (ith == 0) return j; --ith; }
  return -1;
}

// split weights: total weight of the predictors available below node i
__device__ inline unsigned long long t_avail_weight(const DTree& t, const BartParams& P, int i)
{
  unsigned long long w = P.split_total;
  int par = t.nodes[i].parent;
  while (par >= 0) {
    int v = t.nodes[par].var;
    bool seen = false;
    for (int a = t.nodes[i].parent; a != par; a = t.nodes[a].parent) if (t.nodes[a].var == v) { seen = true; break; }
    if (!seen && P.split_w[v] != 0u) { int lo, hi; t_split_interval(t, s4b_ncuts(P, v), i, v, lo, hi); if (hi < lo) w -= P.split_w[v]; }
    par = t.nodes[par].parent;
  }
  return w;
}
// position r in [0, W) of the cumulative weights of the available predictors (index order)
__device__ inline unsigned long long weighted_position(double u, unsigned long long W)
{
  unsigned long long r = (unsigned long long) (u * (double) W);
  return r >= W ? W - 1ull : r;
}
__device__ inline int t_weighted_var(const DTree& t, const BartParams& P, int i, unsigned long long r)
{
  unsigned long long cum = 0ull;
  int last = -1;
  for (int j = 0; j < P.p; ++j) if (t_var_available(t, P, i, j)) { last = j; cum += P.split_w[j]; if (cum > r) return j; }
  return last;
}
// one draw from the rule prior's variable distribution at node i (one uniform either way)
__device__ inline int t_draw_var(const DTree& t, const BartParams& P, int i, int navail, RngState& rng)
{
  if (P.split_w == nullptr) return t_ith_available_var(t, P, i, rng_index(rng, navail));
  const unsigned long long W = t_avail_weight(t, P, i);
  return t_weighted_var(t, P, i, weighted_position(rng_uniform(rng), W));
}
// log prior probability of the splitting variable of internal node i
__device__ inline double t_log_var_prior(const DTree& t, const BartParams& P, int i, int navail)
{
  if (P.split_w == nullptr) return -log((double) navail);
  return log((double) P.split_w[t.nodes[i].var] / (double) t_avail_weight(t, P, i));
}

__device__ inline double t_growth_prob_depth(const double* pgrow, int navail, int depth) { return navail > 0 ? pgrow[depth] : 0.0; }
__device__ inline double t_growth_prob(const DTree& t, const BartParams& P, const double* pgrow, int i)
{
  return t_growth_prob_depth(pgrow, t_num_vars_available(t, P, i), t.nodes[i].depth);
}
__device__ inline bool t_is_birthable(const DTree& t, const BartParams& P, int i, int num_leaves)
{
  return t_num_vars_available(t, P, i) > 0 && t.nodes[i].depth < S4B_MAX_DEPTH && num_leaves < S4B_MAX_LEAVES;
}
__device__ inline int t_subtree_end(const DTree& t, int i)
{
  int d = t.nodes[i].depth, k = i + 1;
  while (k < t.num_nodes && t.nodes[k].depth > d) ++k;
  return k;
}
__device__ inline long long t_heap_index(const DTree& t, int i)
{
  // root = 1, children 2h, 2h + 1
  int path[S4B_MAX_DEPTH + 2]; int len = 0;
  int child = i, par = t.nodes[i].parent;
  while (par >= 0) { path[len++] = (child == par + 1) ? 0 : 1; child = par; par = t.nodes[par].parent; }
  long long h = 1;
  for (int k = len - 1; k >= 0; --k) h = 2 * h + path[k];
  return h;
}

__device__ inline void t_insert_children(DTree& t, int i, int var, int cut)
{
  int n = t.num_nodes;
  for (int k = n - 1; k > i; --k) t.nodes[k + 2] = t.nodes[k];
  for (int k = 0; k < n + 2; ++k) {
    if (k == i + 1 || k == i + 2) continue;
    if (t.nodes[k].var >= 0 && t.nodes[k].right > i) t.nodes[k].right += 2;
    if (t.nodes[k].parent > i) t.nodes[k].parent += 2;
  }
  DNode& nd = t.nodes[i];
  nd.var = (int16_t) var; nd.cut = (int16_t) cut; nd.right = (int16_t) (i + 2);
  for (int c = 1; c <= 2; ++c) {
    DNode& ch = t.nodes[i + c];
    ch.var = -1; ch.cut = -1; ch.right = -1; ch.parent = (int16_t) i; ch.n = 0; ch.depth = nd.depth + 1; ch.mu = 0.0;
  }
  t.num_nodes = n + 2;
}

__device__ inline void t_remove_children(DTree& t, int i)
{
  int n = t.num_nodes;
  for (int k = i + 3; k < n; ++k) t.nodes[k - 2] = t.nodes[k];
  t.num_nodes = n - 2;
  for (int k = 0; k < n - 2; ++k) {
    if (k == i) continue;
    if (t.nodes[k].var >= 0 && t.nodes[k].right > i + 2) t.nodes[k].right -= 2;
    if (t.nodes[k].parent > i + 2) t.nodes[k].parent -= 2;
  }
  t.nodes[i].var = -1; t.nodes[i].cut = -1; t.nodes[i].right = -1;
}

// log prior of the branch rooted at i: growth / no-growth and rule probabilities
__device__ inline double t_branch_log_prior(const DTree& t, const BartParams& P, const double* pgrow, int i)
{
  int end = t_subtree_end(t, i);
  double r = 0.0;
  for (int k = i; k < end; ++k) {
    int navail = t_num_vars_available(t, P, k);
    double pg = t_growth_prob_depth(pgrow, navail, t.nodes[k].depth);
    if (t_is_leaf(t, k)) r += log(1.0 - pg);
    else {
      int lo, hi; t_split_interval(t, s4b_ncuts(P, t.nodes[k].var), k, t.nodes[k].var, lo, hi);
      r += log(pg) + t_log_var_prior(t, P, k, navail) - log((double) (hi - lo + 1));
    }
  }
  return r;
}
__device__ inline bool t_rules_valid(const DTree& t, const BartParams& P, int i)
{
  int end = t_subtree_end(t, i);
  for (int k = i; k < end; ++k) if (!t_is_leaf(t, k)) {
    int lo, hi; t_split_interval(t, s4b_ncuts(P, t.nodes[k].var), k, t.nodes[k].var, lo, hi);
    if (t.nodes[k].cut < lo || t.nodes[k].cut > hi) return false;
  }
  return true;
}

__device__ inline void t_fill_trav(const DTree& t, TravTree& tv, bool with_mu)
{
  tv.n = t.num_nodes;
  int leaf = 0;
  for (int k = 0; k < t.num_nodes; ++k) {
    const DNode& nd = t.nodes[k];
    tv.trav[k] = pack_trav(nd.var, nd.cut, nd.right);
    if (nd.var < 0) { tv.slot[k] = (uint8_t) leaf++; if (with_mu) tv.val[k] = nd.mu; }
    else tv.slot[k] = 255;
  }
}

struct LeafStat { double n, sum, sumsq; };

__device__ inline double leaf_loglik(const LeafStat& s, double sigma, double leaf_prec)
{
  if (s.n <= 0.0) return 0.0;
  double avg = s.sum / s.n;
  double ss = s.sumsq - s.n * avg * avg;
  if (ss < 0.0) ss = 0.0;
  double sigsq = sigma * sigma;
  double dp = s.n / sigsq;
  double r = 0.5 * log(leaf_prec / (leaf_prec + dp));
  r -= 0.5 * ss / sigsq;
  r -= 0.5 * ((leaf_prec * avg) * (dp * avg)) / (leaf_prec + dp);
  return r;
}

// --------------------------------------------------------------------------------------
// proposal for the tree held in `t` (thread 0 of the controller block)
// RNG consumption order mirrors the CPU oracle / dbarts step functions exactly.
// --------------------------------------------------------------------------------------
__device__ inline void propose_step(DTree& t, const BartParams& P, const double* pgrow, RngState& rng, StepDesc& d, int tree_index, unsigned long long step_id)
{
  rng_enter(rng, step_id, 0u);
  d.b_tree = tree_index;
  d.b_kind = -1; d.b_node = -1; d.b_var = -1; d.b_cut = -1; d.b_child = -1; d.new_var = -1; d.new_cut = -1;
  d.log_prior_trans = 0.0;
  t_fill_trav(t, d.b_cur, true);
  int L = 0;
  for (int k = 0; k < t.num_nodes; ++k) if (t_is_leaf(t, k)) ++L;
  d.b_num_leaves = L; d.b_nslots = L;
  d.b_prop.n = 0;

  double u = rng_uniform(rng);
  if (u < P.birth_death_prob) {
    // ---- birth / death (dbarts birthDeathRule.cpp) ----
    int n_birthable = 0;
    for (int k = 0; k < t.num_nodes; ++k) if (t_is_leaf(t, k) && t_is_birthable(t, P, k, L)) ++n_birthable;
    double p_birth = n_birthable == 0 ? 0.0 : (t.num_nodes == 1 ? 1.0 : P.birth_prob);
    if (rng_uniform(rng) < p_birth) {
      int pick = rng_index(rng, n_birthable);
      int node = -1;
      for (int k = 0; k < t.num_nodes; ++k) if (t_is_leaf(t, k) && t_is_birthable(t, P, k, L)) { if (pick == 0) { node = k; break; } --pick; }
      int navail = t_num_vars_available(t, P, node);
      int depth = t.nodes[node].depth;
      double pg_parent = t_growth_prob_depth(pgrow, navail, depth);
      int var = t_draw_var(t, P, node, navail, rng);
      int lo, hi; t_split_interval(t, s4b_ncuts(P, var), node, var, lo, hi);
      int cut = lo + rng_index(rng, hi - lo + 1);
      // children: same availability except possibly `var`
      int navail_l = navail - ((cut - 1 < lo) ? 1 : 0);
      int navail_r = navail - ((cut + 1 > hi) ? 1 : 0);
      double pg_l = t_growth_prob_depth(pgrow, navail_l, depth + 1), pg_r = t_growth_prob_depth(pgrow, navail_r, depth + 1);
      // tree after the birth: L + 1 leaves
      int Lnew = L + 1;
      bool any_birthable_new = false;
      for (int k = 0; k < t.num_nodes && !any_birthable_new; ++k)
        if (k != node && t_is_leaf(t, k) && t_is_birthable(t, P, k, Lnew)) any_birthable_new = true;
      if (!any_birthable_new && depth + 1 < S4B_MAX_DEPTH && Lnew < S4B_MAX_LEAVES && (navail_l > 0 || navail_r > 0)) any_birthable_new = true;
      double p_death_new = 1.0 - (any_birthable_new ? P.birth_prob : 0.0);
      int n_nog = 0;
      for (int k = 0; k < t.num_nodes; ++k) if (!t_is_leaf(t, k) && t_is_leaf(t, k + 1) && t_is_leaf(t, t.nodes[k].right)) ++n_nog;
      int par = t.nodes[node].parent;
      bool parent_was_nog = par >= 0 && t_is_leaf(t, par + 1) && t_is_leaf(t, t.nodes[par].right);
      int n_nog_new = n_nog + 1 - (parent_was_nog ? 1 : 0);
      double prior_ratio = pg_parent * (1.0 - pg_l) * (1.0 - pg_r) / (1.0 - pg_parent);
      double trans_ratio = (p_death_new * (1.0 / (double) n_nog_new)) / (p_birth * (1.0 / (double) n_birthable));
      d.b_kind = 0; d.b_node = node; d.b_var = var; d.b_cut = cut; d.b_nslots = L + 2;
      d.log_prior_trans = prior_ratio * trans_ratio;     // BD steps carry the plain ratio
    } else {
      int n_nog = 0;
      for (int k = 0; k < t.num_nodes; ++k) if (!t_is_leaf(t, k) && t_is_leaf(t, k + 1) && t_is_leaf(t, t.nodes[k].right)) ++n_nog;
      int pick = rng_index(rng, n_nog);
      int node = -1;
      for (int k = 0; k < t.num_nodes; ++k) if (!t_is_leaf(t, k) && t_is_leaf(t, k + 1) && t_is_leaf(t, t.nodes[k].right)) { if (pick == 0) { node = k; break; } --pick; }
      int left = node + 1, right = t.nodes[node].right;
      double pg_parent = t_growth_prob(t, P, pgrow, node);
      double pg_l = t_growth_prob(t, P, pgrow, left), pg_r = t_growth_prob(t, P, pgrow, right);
      int Lnew = L - 1;
      int n_birthable_new = 0;
      for (int k = 0; k < t.num_nodes; ++k) if (k != left && k != right && t_is_leaf(t, k) && t_is_birthable(t, P, k, Lnew)) ++n_birthable_new;
      if (t.nodes[node].depth < S4B_MAX_DEPTH && Lnew < S4B_MAX_LEAVES) ++n_birthable_new;
      double p_birth_new = node == 0 ? 1.0 : P.birth_prob;
      if (n_birthable_new == 0) p_birth_new = 0.0;
      double p_select_birth = n_birthable_new > 0 ? 1.0 / (double) n_birthable_new : 0.0;
      double p_death = 1.0 - p_birth;
      double prior_ratio = (1.0 - pg_parent) / (pg_parent * (1.0 - pg_l) * (1.0 - pg_r));
      double trans_ratio = (p_birth_new * p_select_birth) / (p_death * (1.0 / (double) n_nog));
      d.b_kind = 1; d.b_node = node; d.b_var = t.nodes[node].var; d.b_cut = t.nodes[node].cut;
      d.log_prior_trans = prior_ratio * trans_ratio;
    }
    return;
  }

  bool is_swap = u < P.birth_death_prob + P.swap_prob;
  if (!is_swap) {
    // ---- change rule ----
    d.b_kind = 12;
    int n_nb = 0;
    for (int k = 0; k < t.num_nodes; ++k) if (!t_is_leaf(t, k)) ++n_nb;
    if (n_nb == 0) return;
    int pick = rng_index(rng, n_nb);
    int node = -1;
    for (int k = 0; k < t.num_nodes; ++k) if (!t_is_leaf(t, k)) { if (pick == 0) { node = k; break; } --pick; }
    int navail = t_num_vars_available(t, P, node);
    int new_var = t_draw_var(t, P, node, navail, rng);
    int lo, hi; t_split_interval(t, s4b_ncuts(P, new_var), node, new_var, lo, hi);
    int end = t_subtree_end(t, node), rstart = t.nodes[node].right;
    for (int k = node + 1; k < end; ++k) if (!t_is_leaf(t, k) && t.nodes[k].var == new_var) {
      int c = t.nodes[k].cut;
      if (k < rstart) { if (c + 1 > lo) lo = c + 1; }
      else            { if (c - 1 < hi) hi = c - 1; }
    }
    d.b_node = node; d.new_var = new_var;
    if (lo > hi) return;
    int new_cut = lo + rng_index(rng, hi - lo + 1);
    d.new_cut = new_cut;
    double old_lp = t_branch_log_prior(t, P, pgrow, node);
    int ov = t.nodes[node].var, oc = t.nodes[node].cut;
    t.nodes[node].var = (int16_t) new_var; t.nodes[node].cut = (int16_t) new_cut;
    double new_lp = t_branch_log_prior(t, P, pgrow, node);
    t_fill_trav(t, d.b_prop, false);
    t.nodes[node].var = (int16_t) ov; t.nodes[node].cut = (int16_t) oc;
    int s = L;
    for (int k = 0; k < t.num_nodes; ++k) d.b_prop.slot[k] = (k >= node && k < end && t_is_leaf(t, k)) ? (uint8_t) s++ : (uint8_t) 255;
    // proposal (Hastings) term of the change step: the cut of the new rule was drawn from the interval that ancestors and
    // descendants leave for new_var, the reverse move draws the old cut from the interval of the old variable
    // (oracle_bart.c change_rule; exactness: tests/test_exact_posterior.py)
    double log_hastings = 0.0;
    if (!P.change_symmetric && new_var != ov) {
      int olo, ohi; t_split_interval(t, s4b_ncuts(P, ov), node, ov, olo, ohi);
      for (int k = node + 1; k < end; ++k) if (!t_is_leaf(t, k) && t.nodes[k].var == ov) {
        int c = t.nodes[k].cut;
        if (k < rstart) { if (c + 1 > olo) olo = c + 1; }
        else            { if (c - 1 < ohi) ohi = c - 1; }
      }
      log_hastings = log((double) (hi - lo + 1)) - log((double) (ohi - olo + 1));
      if (P.split_w != nullptr) log_hastings += log((double) P.split_w[ov] / (double) P.split_w[new_var]);
    }
    d.b_kind = 2; d.b_nslots = s; d.log_prior_trans = (new_lp - old_lp) + log_hastings;
    return;
  }

  // ---- swap rule ----
  d.b_kind = 13;
  int n_sw = 0;
  for (int k = 0; k < t.num_nodes; ++k) if (!t_is_leaf(t, k) && (!t_is_leaf(t, k + 1) || !t_is_leaf(t, t.nodes[k].right))) ++n_sw;
  if (n_sw == 0) return;
  int pick = rng_index(rng, n_sw);
  int node = -1;
  for (int k = 0; k < t.num_nodes; ++k) if (!t_is_leaf(t, k) && (!t_is_leaf(t, k + 1) || !t_is_leaf(t, t.nodes[k].right))) { if (pick == 0) { node = k; break; } --pick; }
  d.b_node = node;
  int left = node + 1, right = t.nodes[node].right;
  bool li = !t_is_leaf(t, left), ri = !t_is_leaf(t, right);
  bool both_same = li && ri && t.nodes[left].var == t.nodes[right].var && t.nodes[left].cut == t.nodes[right].cut;
  int child = -1;
  if (!both_same) {
    if (li && ri) child = rng_uniform(rng) < 0.5 ? left : right;
    else child = li ? left : right;
  }
  d.b_child = child;
  int pv = t.nodes[node].var, pc = t.nodes[node].cut;
  int cv = both_same ? t.nodes[left].var : t.nodes[child].var, cc = both_same ? t.nodes[left].cut : t.nodes[child].cut;
  double old_lp = t_branch_log_prior(t, P, pgrow, node);
  t.nodes[node].var = (int16_t) cv; t.nodes[node].cut = (int16_t) cc;
  if (both_same) { t.nodes[left].var = (int16_t) pv; t.nodes[left].cut = (int16_t) pc; t.nodes[right].var = (int16_t) pv; t.nodes[right].cut = (int16_t) pc; }
  else { t.nodes[child].var = (int16_t) pv; t.nodes[child].cut = (int16_t) pc; }
  bool ok = t_rules_valid(t, P, node);
  double new_lp = 0.0;
  int end = t_subtree_end(t, node);
  if (ok) { new_lp = t_branch_log_prior(t, P, pgrow, node); t_fill_trav(t, d.b_prop, false); }
  t.nodes[node].var = (int16_t) pv; t.nodes[node].cut = (int16_t) pc;
  if (both_same) { t.nodes[left].var = (int16_t) cv; t.nodes[left].cut = (int16_t) cc; t.nodes[right].var = (int16_t) cv; t.nodes[right].cut = (int16_t) cc; }
  else { t.nodes[child].var = (int16_t) cv; t.nodes[child].cut = (int16_t) cc; }
  if (!ok) return;
  int s = L;
  for (int k = 0; k < t.num_nodes; ++k) d.b_prop.slot[k] = (k >= node && k < end && t_is_leaf(t, k)) ? (uint8_t) s++ : (uint8_t) 255;
  d.b_kind = 3; d.b_nslots = s; d.log_prior_trans = new_lp - old_lp;
}

// --------------------------------------------------------------------------------------
// Metropolis decision + leaf draws for the tree described by `in` (thread 0).
// `stats[slot]` are the reduced sufficient statistics.  Fills the (A) part of `out`.
// --------------------------------------------------------------------------------------
__device__ inline void decide_and_draw(DTree& t, const BartParams& P, RngState& rng, const StepDesc& in, const LeafStat* stats,
                                       StepDesc& out, double* trace_rec, unsigned long long step_id)
{
  rng_enter(rng, step_id, 1u);
  const int L = in.b_num_leaves;
  const int kind = in.b_kind;
  const int node = in.b_node;
  bool accept = false;
  double ratio = -1.0, old_ll = 0.0, new_ll = 0.0;
  double n_first = 0.0, n_second = 0.0;

  if (kind == 0) {
    LeafStat l = stats[L], r = stats[L + 1];
    LeafStat par = { l.n + r.n, l.sum + r.sum, l.sumsq + r.sumsq };
    old_ll = leaf_loglik(par, P.sigma, P.leaf_prec);
    new_ll = leaf_loglik(l, P.sigma, P.leaf_prec) + leaf_loglik(r, P.sigma, P.leaf_prec);
    ratio = in.log_prior_trans * exp(new_ll - old_ll);
    if (l.n < (double) P.min_obs || r.n < (double) P.min_obs) ratio = 0.0;
    accept = rng_uniform(rng) < ratio;
    n_first = l.n; n_second = r.n;
  } else if (kind == 1) {
    int li = node + 1, ri = t.nodes[node].right;
    LeafStat l = stats[in.b_cur.slot[li]], r = stats[in.b_cur.slot[ri]];
    LeafStat par = { l.n + r.n, l.sum + r.sum, l.sumsq + r.sumsq };
    old_ll = leaf_loglik(l, P.sigma, P.leaf_prec) + leaf_loglik(r, P.sigma, P.leaf_prec);
    new_ll = leaf_loglik(par, P.sigma, P.leaf_prec);
    ratio = in.log_prior_trans * exp(new_ll - old_ll);
    accept = rng_uniform(rng) < ratio;
    n_first = l.n; n_second = r.n;
  } else if (kind == 2 || kind == 3) {
    int end = t_subtree_end(t, node);
    double min_n = 1e300; int seen = 0;
    for (int k = node; k < end; ++k) if (t_is_leaf(t, k)) {
      old_ll += leaf_loglik(stats[in.b_cur.slot[k]], P.sigma, P.leaf_prec);
      const LeafStat& s = stats[in.b_prop.slot[k]];
      new_ll += leaf_loglik(s, P.sigma, P.leaf_prec);
      if (s.n < min_n) min_n = s.n;
      if (seen == 0) n_first = s.n; else if (seen == 1) n_second = s.n;
      ++seen;
    }
    ratio = exp(in.log_prior_trans + (new_ll - old_ll));
    if (min_n < (double) P.min_obs) ratio = 0.0;
    accept = rng_uniform(rng) < ratio;
  }

  // ---- leaf draws in bottom-node order of the final tree; stats per final leaf ----
  // a_old = tree before the step with mu_old; a_new = tree after with mu_new
  out.a_valid = 1;
  out.a_old.n = in.b_cur.n;
  for (int k = 0; k < in.b_cur.n; ++k) { out.a_old.trav[k] = in.b_cur.trav[k]; out.a_old.val[k] = in.b_cur.val[k]; }
  bool structure_changed = accept && (kind >= 0 && kind <= 3);
  if (accept) {
    if (kind == 0) t_insert_children(t, node, in.b_var, in.b_cut);
    else if (kind == 1) t_remove_children(t, node);
    else {
      int end = t_subtree_end(t, node);
      for (int k = node; k < end; ++k) if (!t_is_leaf(t, k)) { uint32_t tv = in.b_prop.trav[k]; t.nodes[k].var = (int16_t) (tv >> 16); t.nodes[k].cut = (int16_t) ((tv >> 8) & 0xFF); }
    }
  }
  double sigsq = P.sigma * P.sigma;
  int nleaf = 0;
  // walk the leaves of the final tree; map each to its statistics
  int old_k = 0;   // running index into the OLD pre-order array
  for (int k = 0; k < t.num_nodes; ++k) {
    // old index corresponding to new index k
    if (accept && kind == 0)      old_k = (k <= node) ? k : (k <= node + 2 ? node : k - 2);
    else if (accept && kind == 1) old_k = (k <= node) ? k : k + 2;
    else old_k = k;
    if (!t_is_leaf(t, k)) continue;
    LeafStat s;
    if (kind == 0) {
      if (accept && k == node + 1) s = stats[L];
      else if (accept && k == node + 2) s = stats[L + 1];
      else if (!accept && k == node) { s.n = stats[L].n + stats[L + 1].n; s.sum = stats[L].sum + stats[L + 1].sum; s.sumsq = stats[L].sumsq + stats[L + 1].sumsq; }
      else s = stats[in.b_cur.slot[old_k]];
    } else if (kind == 1 && accept && k == node) {
      const LeafStat& a = stats[in.b_cur.slot[node + 1]];
      const LeafStat& b = stats[in.b_cur.slot[(int) (in.b_cur.trav[node] & 0xFF)]];
      s.n = a.n + b.n; s.sum = a.sum + b.sum; s.sumsq = a.sumsq + b.sumsq;
    } else if ((kind == 2 || kind == 3) && accept && in.b_prop.slot[k] != 255) {
      s = stats[in.b_prop.slot[k]];
    } else {
      s = stats[in.b_cur.slot[old_k]];
    }
    double avg = s.n > 0.0 ? s.sum / s.n : 0.0;
    double dp = s.n / sigsq;
    double post_mean = dp * avg / (P.leaf_prec + dp);
    double post_sd = 1.0 / sqrt(P.leaf_prec + dp);
    double mu = post_mean + post_sd * rng_normal(rng);
    t.nodes[k].mu = mu;
    t.nodes[k].n = (int32_t) s.n;
    if (trace_rec != nullptr && 11 + nleaf < S4B_TRACE_LEN) trace_rec[11 + nleaf] = mu;
    ++nleaf;
  }
  // observation counts of internal nodes (FlattenedTrees numObservations, init.cpp:583-666)
  for (int k = t.num_nodes - 1; k >= 0; --k) if (!t_is_leaf(t, k)) t.nodes[k].n = t.nodes[k + 1].n + t.nodes[t.nodes[k].right].n;
  if (structure_changed) {
    out.a_same = 0;
    t_fill_trav(t, out.a_new, true);
  } else {
    out.a_same = 1;
    out.a_new.n = 0;
    for (int k = 0; k < t.num_nodes; ++k) if (t_is_leaf(t, k)) out.a_old.val[k] = in.b_cur.val[k] - t.nodes[k].mu;
  }
  if (trace_rec != nullptr) {
    trace_rec[0] = (double) kind;
    if (kind >= 0) {
      // heap index is taken on the tree in which `node` still exists with the same ancestors
      trace_rec[1] = node >= 0 ? (double) t_heap_index(t, node) : 0.0;
    }
    if (kind == 0) { trace_rec[2] = in.b_var; trace_rec[3] = in.b_cut; }
    else if (kind == 1) { trace_rec[2] = in.b_var; trace_rec[3] = in.b_cut; }
    else if (kind == 2 || kind == 12) { trace_rec[2] = node >= 0 ? in.new_var : 0; trace_rec[3] = kind == 2 ? in.new_cut : 0; }
    else if ((kind == 3 || kind == 13) && node >= 0) { trace_rec[2] = in.b_child >= 0 ? (double) t_heap_index(t, in.b_child) : -1.0; }
    trace_rec[4] = accept ? 1.0 : 0.0; trace_rec[5] = ratio;
    trace_rec[6] = old_ll; trace_rec[7] = new_ll; trace_rec[8] = (double) nleaf;
    trace_rec[9] = n_first; trace_rec[10] = n_second;
  }
}

}  // namespace s4b

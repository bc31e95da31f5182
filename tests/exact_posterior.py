"""Exact posterior of a ONE-tree BART model on a tiny data set, by brute-force enumeration from the model itself.

TEST INFRASTRUCTURE.  Nothing here is derived from dbarts, from oracle/ or from the CUDA path: it states the model
(Chipman, George & McCulloch 2010) and enumerates.  It is the external pin of the BART half of the
oracle: a Metropolis-Hastings sampler whose prior, transition or likelihood ratios are wrong for any of the four
moves (birth, death, change, swap) does not leave this distribution invariant, and the visit frequencies of a long
chain show it.

Model (all in the sampler's rescaled units):
    y_i = mu_{leaf(x_i)} + e_i,  e_i ~ N(0, sigma^2 / w_i)          (w_i = 1 without weights)
    mu_l ~ N(0, 1 / a)                                               a = (k sqrt(T) / node_scale)^2, T = 1
    tree ~ CGM prior: a node at depth d with at least one variable that still has a cut point splits with
           probability base (1 + d)^-power; the variable is drawn from the available ones (uniformly, or in
           proportion to split.probs), the cut uniformly from the cut points that ancestors left for it.
    support: every bottom node holds at least `min_obs` rows (the sampler never accepts a move that breaks this).
Rule (var j, cut c): left iff bin_j(x) <= c, with bin_j(x) = #{cut points of j that are < x}.
"""
import itertools
import math

import numpy as np


def uniform_cuts(col, ncuts):
    mn, mx = float(np.min(col)), float(np.max(col))
    return mn + (np.arange(ncuts) + 1.0) * (mx - mn) / (ncuts + 1.0)


def bin_matrix(x, ncuts):
    """x [n, p] -> integer bins [n, p] against `ncuts` uniform cut points per column."""
    x = np.asarray(x, dtype=np.float64)
    out = np.zeros(x.shape, dtype=np.int64)
    for j in range(x.shape[1]):
        cuts = uniform_cuts(x[:, j], ncuts)
        out[:, j] = (x[:, j][:, None] > cuts[None, :]).sum(axis=1)
    return out


def enumerate_trees(p, ncuts, base, power, split_probs=None, cell_ok=None):
    """All trees over p variables with `ncuts` cut points each, as (rules, log_prior, leaves):
    rules  = tuple of (heap, var, cut) sorted by heap index (root = 1, children 2h and 2h + 1),
    leaves = tuple of (heap, cell) in depth-first, left-first order; cell[j] = (lo, hi) cut points still available.
    cell_ok(cell) == False marks a region that may not appear as a bottom node (too few rows); nothing below such a
    region can be valid either, so the whole branch is pruned (the unrestricted count for 2 x 3 cut points is 2.6 million)."""
    sp = None if split_probs is None else np.asarray(split_probs, dtype=np.float64)
    memo = {}

    def heap_of(path):
        h = 1
        for bit in path:
            h = 2 * h + bit
        return h

    def rec(cell, depth):
        """Sub-trees below a node with `cell` at `depth`; positions are paths (tuples of 0 = left / 1 = right) from that node."""
        if (cell, depth) in memo:
            return memo[(cell, depth)]
        if cell_ok is not None and not cell_ok(cell):
            memo[(cell, depth)] = []
            return []
        avail = [j for j in range(p) if cell[j][1] >= cell[j][0] and (sp is None or sp[j] > 0)]
        pg = base / (1.0 + depth) ** power if avail else 0.0
        out = [((), math.log(1.0 - pg), (((), cell),))]
        for j in avail:
            pj = 1.0 / len(avail) if sp is None else sp[j] / sum(sp[a] for a in avail)
            lo, hi = cell[j]
            for c in range(lo, hi + 1):
                lp_rule = math.log(pg) + math.log(pj) - math.log(hi - lo + 1)
                lcell = tuple((lo, c - 1) if a == j else cell[a] for a in range(p))
                rcell = tuple((c + 1, hi) if a == j else cell[a] for a in range(p))
                for (lr, llp, ll), (rr, rlp, rl) in itertools.product(rec(lcell, depth + 1), rec(rcell, depth + 1)):
                    rules = (((), j, c),) + tuple(((0,) + q, v, k) for q, v, k in lr) + tuple(((1,) + q, v, k) for q, v, k in rr)
                    leaves = tuple(((0,) + q, cl) for q, cl in ll) + tuple(((1,) + q, cl) for q, cl in rl)
                    out.append((rules, lp_rule + llp + rlp, leaves))
        memo[(cell, depth)] = out
        return out

    root = tuple((0, ncuts - 1) for _ in range(p))
    return [(tuple(sorted((heap_of(q), v, k) for q, v, k in r)), lp, tuple((heap_of(q), cl) for q, cl in lv)) for r, lp, lv in rec(root, 0)]


def leaf_rows(bins, cell):
    """Rows that fall into the node whose remaining cut points are `cell`: bins lo .. hi + 1 of every variable."""
    ok = np.ones(bins.shape[0], dtype=bool)
    for j, (lo, hi) in enumerate(cell):
        ok &= (bins[:, j] >= lo) & (bins[:, j] <= hi + 1)
    return np.nonzero(ok)[0]


def leaf_log_marginal(y, w, sigma, a):
    """log of  integral prod_i N(y_i | mu, sigma^2 / w_i) N(mu | 0, 1 / a) dmu  without the tree-independent constants."""
    if y.size == 0:
        return 0.0, 0.0
    prec = float(np.sum(w)) / sigma ** 2
    lin = float(np.sum(w * y)) / sigma ** 2
    quad = float(np.sum(w * y * y)) / sigma ** 2
    return 0.5 * math.log(a / (a + prec)) - 0.5 * (quad - lin * lin / (a + prec)), lin / (a + prec)


class ExactPosterior:
    def __init__(self, x, y_scaled, sigma_scaled, leaf_prec, ncuts, base=0.95, power=2.0, min_obs=5, weights=None, split_probs=None):
        self.bins = bin_matrix(x, ncuts)
        n, p = self.bins.shape
        y = np.asarray(y_scaled, dtype=np.float64)
        w = np.ones(n) if weights is None else np.asarray(weights, dtype=np.float64)
        self.keys, logpost, self.fit_mean, self.num_leaves = [], [], [], []
        cells = {}

        def cell_stats(cell):
            if cell not in cells:
                rows = leaf_rows(self.bins, cell)
                cells[cell] = (rows,) + leaf_log_marginal(y[rows], w[rows], sigma_scaled, leaf_prec)
            return cells[cell]
        for rules, lp, leaves in enumerate_trees(p, ncuts, base, power, split_probs, lambda cell: cell_stats(cell)[0].size >= min_obs):
            ll, fit, ok = 0.0, np.zeros(n), True
            for _, cell in leaves:
                rows, l, m = cell_stats(cell)
                if rows.size < min_obs:
                    ok = False
                    break
                ll += l
                fit[rows] = m
            if ok:
                self.keys.append(rules); logpost.append(lp + ll); self.fit_mean.append(fit); self.num_leaves.append(len(leaves))
        logpost = np.array(logpost)
        pr = np.exp(logpost - logpost.max())
        self.prob = pr / pr.sum()
        self.index = {k: i for i, k in enumerate(self.keys)}
        self.fit_mean = np.array(self.fit_mean)
        self.expected_fit = self.prob @ self.fit_mean                 # E[f(x_i) | y], scaled units
        self.num_leaves = np.array(self.num_leaves)


# ---------------------------------------------------------------------------------------------------------------
# chain side: rebuild the sequence of tree structures (and leaf values) from a sampler's parity trace
# ---------------------------------------------------------------------------------------------------------------
def dfs_leaves(rules):
    """Bottom nodes (heap indices) in depth-first, left-first order for rules {heap: (var, cut)}."""
    out, stack = [], [1]
    while stack:
        h = stack.pop()
        if h in rules:
            stack.append(2 * h + 1); stack.append(2 * h)
        else:
            out.append(h)
    return out


def row_leaf_index(bins, rules, order):
    pos = {h: k for k, h in enumerate(order)}
    idx = np.zeros(bins.shape[0], dtype=np.int64)
    for i in range(bins.shape[0]):
        h = 1
        while h in rules:
            j, c = rules[h]
            h = 2 * h if bins[i, j] <= c else 2 * h + 1
        idx[i] = pos[h]
    return idx


def replay_trace(trace, bins, start_rules=None, skip=0):
    """Walk a one-tree chain's trace (tests/common.py layout: kind, heap, var | child heap, cut, accept, ..., leaf values
    from column 11).  Returns (visit counts by structure key, sum over steps of the per-row fit, number of steps counted);
    the first `skip` steps only move the state (burn-in)."""
    rules = dict(start_rules or {})
    visits, cache = {}, {}
    n = bins.shape[0]
    kinds, heaps, accepts = trace[:, 0].astype(np.int64), trace[:, 1].astype(np.int64), trace[:, 4] != 0.0
    nleaves = trace[:, 8].astype(np.int64)
    key = None
    # per structure: rows of the trace spent in it (the leaf values are summed per structure at the end, vectorised)
    steps_in = {}
    for step in range(len(trace)):
        if accepts[step] or key is None:
            kind, h = int(kinds[step]), int(heaps[step])
            if accepts[step]:
                rec = trace[step]
                if kind == 0 or kind == 2:
                    rules[h] = (int(rec[2]), int(rec[3]))
                elif kind == 1:
                    del rules[h]
                elif kind == 3:
                    c = int(rec[2])
                    if c < 0:
                        par = rules[h]
                        rules[h] = rules[2 * h]
                        rules[2 * h] = par; rules[2 * h + 1] = par
                    else:
                        rules[h], rules[c] = rules[c], rules[h]
            key = tuple(sorted((hh, v, c) for hh, (v, c) in rules.items()))
            if key not in cache:
                order = dfs_leaves(rules)
                cache[key] = (len(order), row_leaf_index(bins, rules, order))
                steps_in[key] = []
        if step < skip:
            continue
        assert nleaves[step] == cache[key][0], "trace and replayed tree disagree on the number of bottom nodes"
        steps_in[key].append(step)
    fit_sum = np.zeros(n)
    for key, steps in steps_in.items():
        if not steps:
            continue
        nl, idx = cache[key]
        visits[key] = len(steps)
        fit_sum += trace[np.array(steps), 11:11 + nl].sum(axis=0)[idx]
    return visits, fit_sum, len(trace) - skip

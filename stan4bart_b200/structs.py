"""ctypes mirrors of the plain-C configuration structs in include/stan4bart_b200.h.

The test-side CPU checker declares identical layouts so that tests can drive both sides
with the same inputs.  Field meaning follows the reference:
  BartConfig     dbarts Control/Model as set up by R/stan4bart_fit.R:437-479
  GlmmData       the 44-field `data.stan` list, R/stan4bart_fit.R:259-365,
                 parsed by src/stan_sampler.cpp:112-380
  StanControl    src/stan_sampler.hpp:28-42 (defaults src/stan_sampler.cpp:395-458)
  CommonControl  src/init.cpp:1015-1051
"""
import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint32_p = C.POINTER(C.c_uint32)


class BartConfig(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("p", C.c_int64), ("n_test", C.c_int64),
        ("num_trees", C.c_int32), ("n_cuts", C.c_int32), ("thin", C.c_int32), ("min_obs", C.c_int32),
        ("is_binary", C.c_int32), ("max_ctas", C.c_int32),
        ("birth_death_prob", C.c_double), ("swap_prob", C.c_double), ("change_prob", C.c_double),
        ("birth_prob", C.c_double), ("base", C.c_double), ("power", C.c_double), ("k", C.c_double),
        ("node_scale", C.c_double), ("seed", C.c_uint64), ("split_probs", C.POINTER(C.c_double)),
        ("weights", C.POINTER(C.c_double)), ("k_df", C.c_double), ("k_scale", C.c_double),
        ("n_cuts_var", C.POINTER(C.c_int32)), ("change_symmetric", C.c_int32), ("use_quantiles", C.c_int32),
    ]


class GlmmData(C.Structure):
    _fields_ = [
        ("N", C.c_int64), ("K", C.c_int32), ("is_binary", C.c_int32), ("prior_dist", C.c_int32),
        ("prior_dist_for_aux", C.c_int32), ("t", C.c_int32), ("q", C.c_int32), ("len_theta_L", C.c_int32),
        ("len_concentration", C.c_int32), ("len_regularization", C.c_int32), ("reserved", C.c_int32),
        ("num_non_zero", C.c_int64),
        ("X", c_double_p), ("y", c_double_p), ("prior_scale", c_double_p), ("prior_mean", c_double_p),
        ("prior_scale_for_aux", C.c_double), ("prior_mean_for_aux", C.c_double), ("prior_df_for_aux", C.c_double),
        ("p", c_int32_p), ("l", c_int32_p), ("shape", c_double_p), ("scale", c_double_p),
        ("concentration", c_double_p), ("regularization", c_double_p),
        ("w", c_double_p), ("v", c_int32_p), ("u", c_int32_p), ("weights", c_double_p),
        ("prior_df", c_double_p), ("num_normals", c_int32_p),
        ("global_prior_df", C.c_double), ("global_prior_scale", C.c_double), ("slab_df", C.c_double), ("slab_scale", C.c_double),
    ]


class StanControl(C.Structure):
    _fields_ = [
        ("seed", C.c_uint32), ("skip", C.c_int32), ("init_radius", C.c_double), ("adapt_gamma", C.c_double),
        ("adapt_delta", C.c_double), ("adapt_kappa", C.c_double), ("adapt_t0", C.c_double),
        ("adapt_init_buffer", C.c_uint32), ("adapt_term_buffer", C.c_uint32), ("adapt_window", C.c_uint32),
        ("max_treedepth", C.c_int32), ("stepsize", C.c_double), ("stepsize_jitter", C.c_double),
    ]


class CommonControl(C.Structure):
    _fields_ = [
        ("warmup", C.c_int32), ("iter", C.c_int32), ("is_binary", C.c_int32), ("keep_fits", C.c_int32),
        ("sigma_init", C.c_double), ("offset_type", C.c_int32), ("reserved", C.c_int32), ("user_offset", C.POINTER(C.c_double)),
    ]


def dptr(a):
    return a.ctypes.data_as(c_double_p) if a is not None else c_double_p()


def i32ptr(a):
    return a.ctypes.data_as(c_int32_p) if a is not None else c_int32_p()


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def bart_config(n, p, n_test=0, num_trees=75, n_cuts=100, thin=1, min_obs=5, is_binary=False,
                base=0.95, power=2.0, k=2.0, node_scale=None, seed=0,
                birth_death_prob=0.5, swap_prob=0.1, change_prob=0.4, birth_prob=0.5, split_probs=None, max_ctas=0, weights=None,
                k_df=0.0, k_scale=float("inf"), change_symmetric=False, use_quantiles=False):
    """dbarts defaults as used by stan4bart (R/stan4bart_fit.R:437-479).  split_probs: relative probabilities of the p
    predictors (bart_args split.probs), None = uniform."""
    if node_scale is None:
        node_scale = 3.0 if is_binary else 0.5
    n_cuts_var = None
    if np.ndim(n_cuts) > 0:          # bart_args n.cuts as a vector: recycled over the predictors (R/stan4bart_fit.R:451)
        n_cuts_var = np.ascontiguousarray(np.resize(np.asarray(n_cuts, dtype=np.int32), p))
        n_cuts = int(n_cuts_var.max())
    cfg = BartConfig(n=n, p=p, n_test=n_test, num_trees=num_trees, n_cuts=n_cuts, thin=thin, min_obs=min_obs,
                     is_binary=int(is_binary), max_ctas=int(max_ctas), birth_death_prob=birth_death_prob, swap_prob=swap_prob,
                     change_prob=change_prob, birth_prob=birth_prob, base=base, power=power, k=k,
                     node_scale=node_scale, seed=seed, k_df=float(k_df), k_scale=float(k_scale), change_symmetric=int(bool(change_symmetric)),
                     use_quantiles=int(bool(use_quantiles)))
    if split_probs is not None:
        sp = np.ascontiguousarray(split_probs, dtype=np.float64)
        if sp.shape != (p,) or np.any(sp < 0) or not np.any(sp > 0):
            raise ValueError("split_probs must be p non-negative numbers, not all zero")
        cfg._split_probs = sp                       # keeps the array alive with the struct
        cfg.split_probs = sp.ctypes.data_as(C.POINTER(C.c_double))
    if n_cuts_var is not None:
        cfg._n_cuts_var = n_cuts_var
        cfg.n_cuts_var = n_cuts_var.ctypes.data_as(C.POINTER(C.c_int32))
    if weights is not None:
        wt = np.ascontiguousarray(weights, dtype=np.float64)
        if wt.shape != (n,) or not np.all(np.isfinite(wt)) or np.any(wt < 0):
            raise ValueError("weights must be n finite non-negative numbers")
        cfg._weights = wt
        cfg.weights = wt.ctypes.data_as(C.POINTER(C.c_double))
    return cfg


def stan_control(seed=0, skip=1, init_radius=2.0, adapt_gamma=0.05, adapt_delta=0.8, adapt_kappa=0.75, adapt_t0=10.0,
                 adapt_init_buffer=75, adapt_term_buffer=50, adapt_window=25, max_treedepth=10, stepsize=1.0,
                 stepsize_jitter=0.0):
    """Defaults of src/stan_sampler.cpp:395-458."""
    return StanControl(seed=seed, skip=skip, init_radius=init_radius, adapt_gamma=adapt_gamma, adapt_delta=adapt_delta,
                       adapt_kappa=adapt_kappa, adapt_t0=adapt_t0, adapt_init_buffer=adapt_init_buffer,
                       adapt_term_buffer=adapt_term_buffer, adapt_window=adapt_window, max_treedepth=max_treedepth,
                       stepsize=stepsize, stepsize_jitter=stepsize_jitter)


class StanData:
    """Owns the numpy arrays behind a GlmmData struct (keeps them alive)."""

    def __init__(self, X, y, is_binary, prior_dist, prior_scale, prior_mean, prior_dist_for_aux, prior_scale_for_aux,
                 prior_mean_for_aux, prior_df_for_aux, p, l, shape, scale, concentration, regularization, w, v, u, q):
        self.X = np.asfortranarray(X, dtype=np.float64)
        self.N, self.K = self.X.shape
        self.y = f64(y)
        self.is_binary = bool(is_binary)
        self.prior_dist = int(prior_dist)
        self.prior_scale = f64(prior_scale)
        self.prior_mean = f64(prior_mean)
        self.prior_dist_for_aux = int(prior_dist_for_aux)
        self.prior_scale_for_aux = float(prior_scale_for_aux)
        self.prior_mean_for_aux = float(prior_mean_for_aux)
        self.prior_df_for_aux = float(prior_df_for_aux)
        self.p = i32(p)
        self.l = i32(l)
        self.t = len(self.p)
        self.q = int(q)
        self.len_theta_L = int(sum(pi * (pi - 1) // 2 + pi for pi in self.p))
        self.shape = f64(shape)
        self.scale = f64(scale)
        self.concentration = f64(concentration)
        self.regularization = f64(regularization)
        self.w = f64(w)
        self.v = i32(v)
        self.u = i32(u)
        self.len_rho = int(sum(self.p) - self.t)
        self.len_z_T = int(sum((pi - 2) * (pi - 1) for pi in self.p if pi > 2))        # continuous.stan:258
        # hyper-parameters of the non-default coefficient priors (continuous.stan:207-215; prior_dist 2 .. 7)
        self.prior_df = np.ones(self.K)
        self.global_prior_df = 1.0
        self.global_prior_scale = 0.01
        self.slab_df = 4.0
        self.slab_scale = 2.5
        self.num_normals = np.full(self.K, 2, dtype=np.int32)

    # parameter block of continuous.stan:262-279 ------------------------------------------------------------
    @property
    def hs(self):
        return {3: 2, 4: 4}.get(self.prior_dist, 0)

    @property
    def len_z_beta(self):
        return int(np.sum(self.num_normals)) if self.prior_dist == 7 else self.K

    @property
    def coef_extra_names(self):
        """global, local, caux, mix, one_over_lambda: the extra parameters of the shrinkage priors, in declaration order."""
        hs, K = self.hs, self.K
        names = [f"global.{i + 1}" for i in range(hs)]
        names += [f"local.{j + 1}.{k + 1}" for j in range(hs) for k in range(K)]
        names += ["caux.1"] if hs > 0 else []
        names += [f"mix.1.{k + 1}" for k in range(K)] if self.prior_dist in (5, 6) else []
        names += ["one_over_lambda.1"] if self.prior_dist == 6 else []
        return names

    @property
    def num_params(self):
        return (self.len_z_beta + len(self.coef_extra_names) + self.q + self.len_z_T + self.len_rho + len(self.concentration)
                + self.t + (0 if self.is_binary else 1))

    @property
    def num_constrained(self):
        return self.num_params + (0 if self.is_binary else 1) + self.K + self.q + self.len_theta_L

    def rows(self, lo, hi):
        """The same model restricted to observations [lo, hi): the shard of one rank of an observation-sharded chain.
        Priors, the ranef layout and the column space of Z (q) are those of the whole data set."""
        lo, hi = int(lo), int(hi)
        k0, k1 = int(self.u[lo]), int(self.u[hi])
        out = StanData(self.X[lo:hi], self.y[lo:hi], self.is_binary, self.prior_dist, self.prior_scale, self.prior_mean,
                       self.prior_dist_for_aux, self.prior_scale_for_aux, self.prior_mean_for_aux, self.prior_df_for_aux,
                       self.p, self.l, self.shape, self.scale, self.concentration, self.regularization,
                       self.w[k0:k1], self.v[k0:k1], self.u[lo:hi + 1] - k0, self.q)
        for name in ("xbar", "term_order", "prior_df", "global_prior_df", "global_prior_scale", "slab_df", "slab_scale", "num_normals"):
            if hasattr(self, name):
                setattr(out, name, getattr(self, name))
        if getattr(self, "weights", None) is not None:
            out.weights = np.ascontiguousarray(self.weights[lo:hi])
        return out

    def struct(self):
        return GlmmData(
            N=self.N, K=self.K, is_binary=int(self.is_binary), prior_dist=self.prior_dist,
            prior_dist_for_aux=self.prior_dist_for_aux, t=self.t, q=self.q, len_theta_L=self.len_theta_L,
            len_concentration=len(self.concentration), len_regularization=len(self.regularization), reserved=0,
            num_non_zero=len(self.w),
            X=dptr(self.X), y=dptr(self.y), prior_scale=dptr(self.prior_scale), prior_mean=dptr(self.prior_mean),
            prior_scale_for_aux=self.prior_scale_for_aux, prior_mean_for_aux=self.prior_mean_for_aux,
            prior_df_for_aux=self.prior_df_for_aux,
            p=i32ptr(self.p), l=i32ptr(self.l), shape=dptr(self.shape), scale=dptr(self.scale),
            concentration=dptr(self.concentration), regularization=dptr(self.regularization),
            w=dptr(self.w), v=i32ptr(self.v), u=i32ptr(self.u), weights=self._weights_ptr(),
            prior_df=self._keep("prior_df", np.float64, dptr), num_normals=self._keep("num_normals", np.int32, i32ptr),
            global_prior_df=self.global_prior_df, global_prior_scale=self.global_prior_scale, slab_df=self.slab_df,
            slab_scale=self.slab_scale)

    def _keep(self, name, dtype, to_ptr):
        arr = np.ascontiguousarray(getattr(self, name), dtype=dtype)
        if arr.shape != (self.K,):
            raise ValueError(f"{name} must have one entry per fixed-effect coefficient")
        setattr(self, name, arr)
        return to_ptr(arr)

    def _weights_ptr(self):
        """data.stan `weights` (R/stan4bart_fit.R:255-262): None = unweighted (a NULL pointer)."""
        wt = getattr(self, "weights", None)
        if wt is None:
            return None
        self.weights = np.ascontiguousarray(wt, dtype=np.float64)
        if self.weights.shape != (self.N,):
            raise ValueError("weights must have one entry per observation")
        return dptr(self.weights)

    def param_names(self):
        """Names of the stored Stan rows, continuous.hpp:3115-3204 (constrained_param_names)."""
        names = ["lp__", "accept_stat__", "stepsize__", "treedepth__", "n_leapfrog__", "divergent__", "energy__"]
        names += [f"z_beta.{i + 1}" for i in range(self.len_z_beta)]
        names += self.coef_extra_names
        names += [f"z_b.{i + 1}" for i in range(self.q)]
        names += [f"z_T.{i + 1}" for i in range(self.len_z_T)]
        names += [f"rho.{i + 1}" for i in range(self.len_rho)]
        names += [f"zeta.{i + 1}" for i in range(len(self.concentration))]
        names += [f"tau.{i + 1}" for i in range(self.t)]
        if not self.is_binary:
            names += ["aux_unscaled.1", "aux.1"]
        names += [f"beta.{i + 1}" for i in range(self.K)]
        names += [f"b.{i + 1}" for i in range(self.q)]
        names += [f"theta_L.{i + 1}" for i in range(self.len_theta_L)]
        return names

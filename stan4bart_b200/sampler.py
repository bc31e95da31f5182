"""Python mirror of the reference's host interface for the Gibbs hot path, over the C ABI.

  GpuBart    the dbarts function table stan4bart binds (src/init.cpp:54-81)
  GlmmModel  the Stan model / gradient hook (continuous.hpp; model/gradient.hpp:21-35)
  Sampler    the `.Call` routines: stan4bart_create / _run / _disengageAdaptation /
             _getParametricMean / _getBARTDataRange / _predictBART (src/init.cpp:1215-1229);
             `run` returns the same named pieces as the reference's result list
             (stan [pars x S]; bart: train [n x S], test [n_test x S], varcount [p x S], sigma [S])
"""
import ctypes as C

import numpy as np

from . import _lib
from .structs import CommonControl, c_int32_p, c_int64_p, c_uint32_p, dptr, f64

TRACE_LEN = 32


class GpuBart:
    def __init__(self, cfg, y, x, x_test=None, handle=None, shard=None):
        self.L = _lib.load()
        self.cfg = cfg
        self.n, self.p, self.nt = int(cfg.n), int(cfg.p), int(cfg.n_test)
        self._owner = handle is None
        if handle is not None:
            self.h = handle
            return
        _lib.require_device()
        y = f64(y)
        x = np.asfortranarray(x, dtype=np.float64)
        xt = np.asfortranarray(x_test, dtype=np.float64) if x_test is not None else None
        if x.shape != (self.n, self.p):
            raise ValueError("x must be n x p")
        h = C.c_void_p()
        if shard is None:
            _lib.check(self.L.gpubart_create(C.byref(cfg), dptr(y), dptr(x), dptr(xt), C.byref(h)))
        else:       # this rank's rows of an observation-sharded chain (shard.py)
            _lib.check(self.L.gpubart_create_sharded(C.byref(cfg), dptr(y), dptr(x), dptr(xt), shard.h, C.byref(h)))
        self._shard = shard
        self.h = h

    def __del__(self):
        if getattr(self, "h", None) and self._owner:
            self.L.gpubart_free(self.h)
            self.h = None

    def set_offset(self, offset, update_scale):
        o = f64(offset) if offset is not None else None
        _lib.check(self.L.gpubart_set_offset(self.h, dptr(o), int(update_scale)))

    def set_sigma(self, sigma):
        _lib.check(self.L.gpubart_set_sigma(self.h, float(sigma)))

    def set_response(self, y):
        """setResponse of the dbarts table (init.cpp:68): a new response vector, fits and offset unchanged."""
        _lib.check(self.L.gpubart_set_response(self.h, dptr(f64(y))))

    def set_keep_trees_active(self, on):
        """setControl(keepTrees = on): whether the following runs append their draw to the tree store (init.cpp:737-744)."""
        _lib.check(self.L.gpubart_set_keep_trees_active(self.h, int(bool(on))))

    def stored_scales(self, first=0, count=None):
        count = self.num_stored() - first if count is None else count
        out = np.zeros((count, 2))
        _lib.check(self.L.gpubart_get_stored_scales(self.h, int(first), int(count), dptr(out)))
        return out

    def k(self):
        """Current k of the leaf prior normal(k): the configured value, or the latest draw under the chi hyperprior."""
        v = C.c_double(0.0)
        _lib.check(self.L.gpubart_get_k(self.h, C.byref(v)))
        return v.value

    def sample_trees_from_prior(self):
        _lib.check(self.L.gpubart_sample_trees_from_prior(self.h))

    def run(self):
        train = np.zeros(self.n)
        test = np.zeros(self.nt) if self.nt else None
        vc = np.zeros(self.p, dtype=np.uint32)
        sig = C.c_double(0.0)
        _lib.check(self.L.gpubart_run_sampler_with_results(self.h, dptr(train), dptr(test), vc.ctypes.data_as(c_uint32_p), C.byref(sig)))
        return dict(train=train, test=test, varcount=vc, sigma=sig.value)

    def results(self):
        """The results of the last sweep (after run_batched: the chain's own)."""
        train = np.zeros(self.n)
        test = np.zeros(self.nt) if self.nt else None
        vc = np.zeros(self.p, dtype=np.uint32)
        sig = C.c_double(0.0)
        _lib.check(self.L.gpubart_collect_results(self.h, dptr(train), dptr(test), vc.ctypes.data_as(c_uint32_p), C.byref(sig)))
        return dict(train=train, test=test, varcount=vc, sigma=sig.value)

    @staticmethod
    def run_batched(fits, results=True):
        """One sweep step of several chains with their sweep kernels batched into ONE launch (grid.y = chain): the fits were created with
        the same shape class and max_ctas = SMs // len(fits).  Returns every chain's results (or None)."""
        L = _lib.load()
        arr = (C.c_void_p * len(fits))(*[f.h for f in fits])
        _lib.check(L.gpubart_run_batched(arr, len(fits)))
        return [f.results() for f in fits] if results else None

    def latents(self):
        out = np.zeros(self.n)
        _lib.check(self.L.gpubart_store_latents(self.h, dptr(out)))
        return out

    def data_range(self):
        out = np.zeros(3)
        _lib.check(self.L.gpubart_get_data_range(self.h, dptr(out)))
        return out

    def predict(self, x_test, offset=None):
        xt = np.asfortranarray(x_test, dtype=np.float64)
        out = np.zeros(xt.shape[0])
        o = f64(offset) if offset is not None else None
        _lib.check(self.L.gpubart_predict(self.h, dptr(xt), xt.shape[0], dptr(o), dptr(out)))
        return out

    def trees(self):
        k = C.c_int64(0)
        _lib.check(self.L.gpubart_num_nodes(self.h, C.byref(k)))
        k = k.value
        tree_no = np.zeros(k, dtype=np.int32)
        n_obs = np.zeros(k, dtype=np.int64)
        var = np.zeros(k, dtype=np.int32)
        value = np.zeros(k)
        _lib.check(self.L.gpubart_get_trees(self.h, tree_no.ctypes.data_as(c_int32_p), n_obs.ctypes.data_as(c_int64_p),
                                            var.ctypes.data_as(c_int32_p), dptr(value)))
        return dict(tree=tree_no, n=n_obs, var=var, value=value)

    # ---- keepTrees: stored draws ----
    def set_keep_trees(self, capacity):
        _lib.check(self.L.gpubart_set_keep_trees(self.h, int(capacity)))

    def num_stored(self):
        k = C.c_int64(0)
        _lib.check(self.L.gpubart_num_stored(self.h, C.byref(k)))
        return int(k.value)

    def predict_stored(self, x_test, first=0, count=None, offset=None):
        """Predictions of stored draws first .. first + count - 1 on new rows: array [rows x count]."""
        xt = np.asfortranarray(x_test, dtype=np.float64)
        count = self.num_stored() - first if count is None else count
        out = np.zeros((count, xt.shape[0]))
        o = f64(offset) if offset is not None else None
        _lib.check(self.L.gpubart_predict_stored(self.h, dptr(xt), xt.shape[0], dptr(o), int(first), int(count), dptr(out)))
        return out.T

    def stored_trees(self, sample):
        k = C.c_int64(0)
        _lib.check(self.L.gpubart_num_stored_nodes(self.h, int(sample), C.byref(k)))
        k = k.value
        tree_no = np.zeros(k, dtype=np.int32)
        n_obs = np.zeros(k, dtype=np.int64)
        var = np.zeros(k, dtype=np.int32)
        value = np.zeros(k)
        _lib.check(self.L.gpubart_get_stored_trees(self.h, int(sample), tree_no.ctypes.data_as(c_int32_p), n_obs.ctypes.data_as(c_int64_p),
                                                   var.ctypes.data_as(c_int32_p), dptr(value)))
        return dict(tree=tree_no, n=n_obs, var=var, value=value)

    def summary(self):
        """printInitialSummary: the text dbarts prints for a fit (prior, tree prior parameters, split probabilities)."""
        need = C.c_size_t(0)
        _lib.check(self.L.gpubart_summary(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _lib.check(self.L.gpubart_summary(self.h, buf, need.value, C.byref(need)))
        return buf.value.decode()

    def export_stored(self):
        """stan4bart_exportBARTState: the stored draws (and the cut points) as bytes."""
        k = C.c_int64(0)
        _lib.check(self.L.gpubart_stored_export_size(self.h, C.byref(k)))
        buf = (C.c_ubyte * k.value)()
        _lib.check(self.L.gpubart_stored_export(self.h, buf, k.value))
        return bytes(buf)

    # ---- parity instrumentation ----
    def node_assignment(self, tree):
        out = np.zeros(self.n, dtype=np.int64)
        _lib.check(self.L.gpubart_node_assignment(self.h, tree, out.ctypes.data_as(c_int64_p)))
        return out

    def leaf_stats(self, tree, max_leaves=64):
        heap = np.zeros(max_leaves, dtype=np.int64)
        cnt = np.zeros(max_leaves, dtype=np.int64)
        s = np.zeros(max_leaves)
        ss = np.zeros(max_leaves)
        nl = C.c_int(0)
        _lib.check(self.L.gpubart_leaf_stats(self.h, tree, max_leaves, heap.ctypes.data_as(c_int64_p), cnt.ctypes.data_as(c_int64_p),
                                             dptr(s), dptr(ss), C.byref(nl)))
        k = nl.value
        return heap[:k], cnt[:k], s[:k], ss[:k]

    def residual(self):
        out = np.zeros(self.n)
        _lib.check(self.L.gpubart_get_residual(self.h, dptr(out)))
        return out

    def set_trace(self, cap):
        self._trace_cap = cap
        _lib.check(self.L.gpubart_set_trace(self.h, cap))

    def trace(self):
        out = np.zeros((self._trace_cap, TRACE_LEN))
        k = C.c_size_t(0)
        _lib.check(self.L.gpubart_get_trace(self.h, dptr(out), self._trace_cap, C.byref(k)))
        return out[:min(k.value, self._trace_cap)]

    def set_tape(self, tape):
        t = f64(tape)
        _lib.check(self.L.gpubart_set_tape(self.h, dptr(t), len(t)))

    def set_record(self, cap):
        self._rec_cap = cap
        _lib.check(self.L.gpubart_set_record(self.h, cap))

    def record(self):
        out = np.zeros(self._rec_cap)
        k = C.c_size_t(0)
        _lib.check(self.L.gpubart_get_record(self.h, dptr(out), self._rec_cap, C.byref(k)))
        if k.value > self._rec_cap:
            raise _lib.S4BError("record buffer too small")
        return out[:k.value]

    def rng_counter(self):
        k = C.c_uint64(0)
        _lib.check(self.L.gpubart_rng_counter(self.h, C.byref(k)))
        return int(k.value)

    def set_use_graph(self, flag):
        _lib.check(self.L.gpubart_set_use_graph(self.h, int(flag)))

    def set_sweep_mode(self, mode):
        _lib.check(self.L.gpubart_set_sweep_mode(self.h, int(mode)))

    def sweep_mode(self):
        m = C.c_int(0)
        _lib.check(self.L.gpubart_get_sweep_mode(self.h, C.byref(m)))
        return m.value

    def set_pipeline(self, on):
        """Software-pipelined sweep kernel on / off (on by default where it applies)."""
        _lib.check(self.L.gpubart_set_pipeline(self.h, int(bool(on))))

    def pipeline(self):
        en, a, b = C.c_int(0), C.c_int64(0), C.c_int64(0)
        _lib.check(self.L.gpubart_get_pipeline(self.h, C.byref(en), C.byref(a), C.byref(b)))
        mis = (C.c_uint32 * 4)()
        _lib.check(self.L.gpubart_pipeline_misfits(self.h, mis))
        return dict(enabled=bool(en.value), sweeps_offered=int(a.value), steps_pipelined=int(b.value),
                    misfits=dict(all=int(mis[0]), tree_too_large=int(mis[1]), too_many_slots=int(mis[2]), too_many_cells=int(mis[3])))

    def time_leaf_stats(self, tree=0, reps=20):
        ms = C.c_double(0.0)
        _lib.check(self.L.gpubart_time_leaf_stats(self.h, tree, reps, C.byref(ms)))
        return ms.value

    def tree_step_ms(self, reset=True):
        ms = C.c_double(0.0)
        _lib.check(self.L.gpubart_tree_step_ms(self.h, int(reset), C.byref(ms)))
        return ms.value

    def set_profile(self, on=True):
        _lib.check(self.L.gpubart_set_profile(self.h, int(on)))

    def profile(self, reset=True):
        out = (C.c_uint64 * 24)()
        _lib.check(self.L.gpubart_get_profile(self.h, out, int(reset)))
        v = [int(x) for x in out]
        steps = max(1, v[7])
        names = ["p0", "p1", "p2", "p3", "p4", "p5", "p6"]
        fine = ["dec_summaries_accept", "dec_structure", "dec_leaf_draws", "dec_update_desc", "ctl_tree_fetch", "ctl_fill_decision_draws",
                "ctl_fill_proposal_draws", "ctl_propose"]
        return {"steps": v[7], "cycles_per_step": {k: v[i] / steps for i, k in enumerate(names)},
                "controller_cycles_per_step": {k: v[8 + i] / steps for i, k in enumerate(fine)},
                "worker_cycles_per_step": {k: v[16 + i] / steps for i, k in enumerate(["zero_bins", "accumulate", "barrier_and_row_reduce", "second_barrier"])},
                "decision_fine_cycles_per_step": {k: v[20 + i] / steps for i, k in enumerate(["slot_summaries", "stage_old_values", "ratio_accept", "-"])}}

    def num_tree_steps(self):
        k = C.c_int64(0)
        _lib.check(self.L.gpubart_num_tree_steps(self.h, C.byref(k)))
        return int(k.value)


class StoredBart:
    """stan4bart_createStoredBARTSampler: prediction from exported draws, no training data or live sampler needed."""

    def __init__(self, blob):
        self.L = _lib.load()
        _lib.require_device()
        self._blob = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        h = C.c_void_p()
        _lib.check(self.L.gpubart_stored_import(self._blob, len(blob), C.byref(h)))
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            self.L.gpubart_stored_free(self.h)
            self.h = None

    def count(self):
        k = C.c_int64(0)
        _lib.check(self.L.gpubart_stored_count(self.h, C.byref(k)))
        return int(k.value)

    def predict(self, x_test, first=0, count=None, offset=None):
        xt = np.asfortranarray(x_test, dtype=np.float64)
        count = self.count() - first if count is None else count
        out = np.zeros((count, xt.shape[0]))
        o = f64(offset) if offset is not None else None
        _lib.check(self.L.gpubart_stored_predict(self.h, dptr(xt), xt.shape[0], dptr(o), int(first), int(count), dptr(out)))
        return out.T


class GlmmModel:
    def __init__(self, stan_data, handle=None, shard=None):
        self.L = _lib.load()
        self.sd = stan_data
        self._owner = handle is None
        if handle is None:
            _lib.require_device()
            self._struct = stan_data.struct()
            h = C.c_void_p()
            if shard is None:
                _lib.check(self.L.glmm_create(C.byref(self._struct), C.byref(h)))
            else:
                _lib.check(self.L.glmm_create_sharded(C.byref(self._struct), shard.h, C.byref(h)))
            self._shard = shard
            self.h = h
        else:
            self.h = handle
        d, nc = C.c_int(0), C.c_int(0)
        _lib.check(self.L.glmm_num_params(self.h, C.byref(d), C.byref(nc)))
        self.d, self.nc = d.value, nc.value

    def __del__(self):
        if getattr(self, "h", None) and self._owner:
            self.L.glmm_free(self.h)
            self.h = None

    def set_offset(self, o):
        _lib.check(self.L.glmm_set_offset(self.h, dptr(f64(o))))

    def set_response(self, y):
        _lib.check(self.L.glmm_set_response(self.h, dptr(f64(y))))

    def log_prob_grad(self, q):
        q = f64(q)
        g = np.zeros(self.d)
        lp = C.c_double(0.0)
        st = C.c_int(0)
        _lib.check(self.L.glmm_log_prob_grad(self.h, dptr(q), C.byref(lp), dptr(g), C.byref(st)))
        return lp.value, g, st.value

    def write_array(self, q):
        out = np.zeros(self.nc)
        _lib.check(self.L.glmm_write_array(self.h, dptr(f64(q)), dptr(out)))
        return out

    def parametric_mean(self, constrained, fixed=True, random=True):
        out = np.zeros(self.sd.N)
        _lib.check(self.L.glmm_parametric_mean(self.h, dptr(f64(constrained)), dptr(out), int(fixed), int(random)))
        return out

    def data_terms(self, beta, b):
        S = C.c_double(0.0)
        gbeta = np.zeros(max(1, self.sd.K))
        gb = np.zeros(max(1, self.sd.q))
        _lib.check(self.L.glmm_data_terms(self.h, dptr(f64(beta)), dptr(f64(b)), C.byref(S), dptr(gbeta), dptr(gb)))
        return S.value, gbeta[:self.sd.K], gb[:self.sd.q]

    def stan_row_names(self):
        """Names of the stored Stan rows as the library reports them (the dimnames of the reference's `stan` result)."""
        need = C.c_size_t(0)
        _lib.check(self.L.glmm_stan_row_names(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value)
        _lib.check(self.L.glmm_stan_row_names(self.h, buf, need.value, C.byref(need)))
        return buf.value.decode().split("\n")

    def num_grad_evals(self):
        k = C.c_int64(0)
        _lib.check(self.L.glmm_num_grad_evals(self.h, C.byref(k)))
        return int(k.value)

    def num_device_passes(self):
        k = C.c_int64(0)
        _lib.check(self.L.glmm_num_device_passes(self.h, C.byref(k)))
        return int(k.value)

    def time_data_pass(self, reps=20, flush_l2=False):
        """(milliseconds per data pass on the device, whether it is the bulk-copy kernel)"""
        ms, bulk = C.c_double(0.0), C.c_int(0)
        _lib.check(self.L.glmm_time_data_pass(self.h, int(reps), int(bool(flush_l2)), C.byref(ms), C.byref(bulk)))
        return ms.value, bool(bulk.value)

    def set_mode(self, mode):
        _lib.check(self.L.glmm_set_mode(self.h, int(mode)))

    def mode(self):
        m = C.c_int(0)
        _lib.check(self.L.glmm_get_mode(self.h, C.byref(m)))
        return m.value


class StanSampler:
    """The Stan half alone: NUTS (diag_e, windowed adaptation) over a GlmmModel, like the reference's StanSampler
    (src/stan_sampler.cpp:382-489).  run(warmup) makes `ctl.skip` transitions and returns the last row
    (lp__, accept_stat__, stepsize__, treedepth__, n_leapfrog__, divergent__, energy__, constrained parameters)."""

    def __init__(self, glmm, ctl, chain_id=1, num_warmup=1000):
        self.L = _lib.load()
        self.glmm, self.ctl = glmm, ctl               # the model must outlive the sampler
        h = C.c_void_p()
        _lib.check(self.L.glmm_nuts_create(glmm.h, C.byref(ctl), int(chain_id), int(num_warmup), C.byref(h)))
        self.h = h
        k = C.c_int(0)
        _lib.check(self.L.glmm_nuts_num_pars(self.h, C.byref(k)))
        self.num_pars = k.value

    def __del__(self):
        if getattr(self, "h", None):
            self.L.glmm_nuts_free(self.h)
            self.h = None

    def run(self, warmup=True):
        out = np.zeros(self.num_pars)
        _lib.check(self.L.glmm_nuts_run(self.h, int(bool(warmup)), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def disengage_adaptation(self):
        _lib.check(self.L.glmm_nuts_disengage_adaptation(self.h))

    def stepsize(self):
        v = C.c_double(0.0)
        _lib.check(self.L.glmm_nuts_stepsize(self.h, C.byref(v)))
        return v.value


class BatchGroup:
    """Several chains of one GPU whose BART sweeps are batched into one launch per Gibbs iteration (config D: grid.y = chain).
    Attach `count` samplers (created with max_ctas = SMs // count), then call run() of each from its own host thread."""

    def __init__(self, count):
        self.L = _lib.load()
        h = C.c_void_p()
        _lib.check(self.L.s4b_batch_group_create(int(count), C.byref(h)))
        self.h, self.count = h, int(count)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.s4b_batch_group_free(self.h)
            self.h = None

    def launches(self):
        k = C.c_int64(0)
        _lib.check(self.L.s4b_batch_group_launches(self.h, C.byref(k)))
        return int(k.value)


class Sampler:
    """stan4bart_create(...) -> sampler object with run / disengage_adaptation / ... methods."""

    def __init__(self, bart_cfg, y, x_bart, x_test, stan_data, stan_ctl, warmup, iter_, keep_fits=True, sigma_init=1.0,
                 bart_offset_init=None, shard=None, offset=None, offset_type=0):
        self.L = _lib.load()
        self._shard = shard
        _lib.require_device()
        self.bcfg = bart_cfg
        self.sd = stan_data
        self._gs = stan_data.struct()
        y = f64(y)
        x = np.asfortranarray(x_bart, dtype=np.float64)
        xt = np.asfortranarray(x_test, dtype=np.float64) if x_test is not None else None
        off = f64(bart_offset_init) if bart_offset_init is not None else None
        self._user_offset = f64(offset) if offset is not None else None
        self.cc = CommonControl(warmup=warmup, iter=iter_, is_binary=int(stan_data.is_binary), keep_fits=int(keep_fits),
                                sigma_init=float(sigma_init), offset_type=int(offset_type), reserved=0, user_offset=dptr(self._user_offset))
        self.keep_fits = bool(keep_fits)
        h = C.c_void_p()
        if shard is None:
            _lib.check(self.L.s4b_sampler_create(C.byref(bart_cfg), dptr(y), dptr(x), dptr(xt), C.byref(self._gs), C.byref(stan_ctl),
                                                 C.byref(self.cc), dptr(off), C.byref(h)))
        else:
            _lib.check(self.L.s4b_sampler_create_sharded(C.byref(bart_cfg), dptr(y), dptr(x), dptr(xt), C.byref(self._gs),
                                                         C.byref(stan_ctl), C.byref(self.cc), dptr(off), shard.h, C.byref(h)))
        self.h = h
        k = C.c_int(0)
        _lib.check(self.L.s4b_sampler_num_stan_pars(self.h, C.byref(k)))
        self.num_pars = k.value
        self.n, self.nt, self.p = int(bart_cfg.n), int(bart_cfg.n_test), int(bart_cfg.p)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.s4b_sampler_free(self.h)
            self.h = None

    def bart(self):
        return GpuBart(self.bcfg, None, None, handle=C.c_void_p(self.L.s4b_sampler_bart(self.h)))

    def glmm(self):
        return GlmmModel(self.sd, handle=C.c_void_p(self.L.s4b_sampler_glmm(self.h)))

    def run(self, num_iter, is_warmup, results=True):
        """results=False runs without copying any fits back (the keep_fits = FALSE path)."""
        S = num_iter if self.keep_fits else 1
        stan = np.zeros((S, self.num_pars))
        if results:
            train = np.zeros((S, self.n))
            test = np.zeros((S, max(self.nt, 1)))
            vc = np.zeros((S, self.p), dtype=np.uint32)
        else:
            train = test = vc = None
        sigma = np.zeros(S)
        _lib.check(self.L.s4b_sampler_run(self.h, num_iter, int(is_warmup), dptr(stan), dptr(train), dptr(test),
                                          vc.ctypes.data_as(c_uint32_p) if vc is not None else c_uint32_p(), dptr(sigma)))
        out = dict(stan=stan.T)
        if results:
            out["bart"] = dict(train=train.T, test=test.T[:self.nt], varcount=vc.T, sigma=sigma)
        else:
            out["bart"] = dict(sigma=sigma)
        # the `k` row of the BART results (src/bart_util.hpp:25: present when k is modelled; here always, constant otherwise)
        k = np.zeros(num_iter)
        cnt = C.c_int(0)
        _lib.check(self.L.s4b_sampler_last_k(self.h, dptr(k), num_iter, C.byref(cnt)))
        out["bart"]["k"] = k[:cnt.value]
        return out

    def set_callback(self, fn):
        """fn(iteration, stan_row, yhat_train, yhat_test) is called on the host after every iteration (numpy views valid
        during the call only; yhat_test is None without a test sample); a truthy return value stops the run.  None removes it."""
        if fn is None:
            self._cb = None
            _lib.check(self.L.s4b_sampler_set_callback(self.h, _lib.ITERATION_CALLBACK(), None))
            return
        n, nt, npar = self.n, self.nt, self.num_pars

        def tramp(_user, it, stan_p, train_p, test_p):
            stan = np.ctypeslib.as_array(stan_p, shape=(npar,))
            train = np.ctypeslib.as_array(train_p, shape=(n,))
            test = np.ctypeslib.as_array(test_p, shape=(nt,)) if nt and test_p else None
            return 1 if fn(it, stan, train, test) else 0

        self._cb = _lib.ITERATION_CALLBACK(tramp)          # keep the trampoline alive
        _lib.check(self.L.s4b_sampler_set_callback(self.h, self._cb, None))

    def set_batch_group(self, group):
        """Join (or, with None, leave) a BatchGroup: this chain's BART sweeps are then launched together with the group's other chains."""
        _lib.check(self.L.s4b_sampler_set_batch_group(self.h, group.h if group is not None else None))
        self._batch_group = group

    def disengage_adaptation(self):
        _lib.check(self.L.s4b_sampler_disengage_adaptation(self.h))

    def data_range(self):
        out = np.zeros(2)
        _lib.check(self.L.s4b_sampler_get_bart_data_range(self.h, dptr(out)))
        return out

    def parametric_mean(self):
        out = np.zeros(self.n)
        _lib.check(self.L.s4b_sampler_get_parametric_mean(self.h, dptr(out)))
        return out

    def predict_bart(self, x_test, offset=None):
        xt = np.asfortranarray(x_test, dtype=np.float64)
        out = np.zeros(xt.shape[0])
        o = f64(offset) if offset is not None else None
        _lib.check(self.L.s4b_sampler_predict_bart(self.h, dptr(xt), xt.shape[0], dptr(o), dptr(out)))
        return out

    def means(self):
        mt = np.zeros(self.n)
        mte = np.zeros(max(self.nt, 1))
        mp = np.zeros(self.n)
        k = C.c_int64(0)
        _lib.check(self.L.s4b_sampler_get_means(self.h, dptr(mt), dptr(mte) if self.nt else dptr(None), dptr(mp), C.byref(k)))
        return dict(bart_train=mt, bart_test=mte[:self.nt], parametric=mp, num_draws=int(k.value))

    def set_host_plumbing(self, on):
        a, b = C.c_int64(0), C.c_int64(0)
        _lib.check(self.L.s4b_sampler_set_host_plumbing(self.h, int(on), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def run_into(self, num_iter, is_warmup, stan=None, train=None, test=None, varcount=None, sigma=None):
        """stan4bart_run writing into caller-provided host buffers given as raw addresses (or None)."""
        def p(addr, typ):
            return C.cast(C.c_void_p(addr), typ) if addr else typ()
        _lib.check(self.L.s4b_sampler_run(self.h, num_iter, int(is_warmup), p(stan, _lib.c_double_p), p(train, _lib.c_double_p),
                                          p(test, _lib.c_double_p), p(varcount, c_uint32_p), p(sigma, _lib.c_double_p)))

    def last_run_stats(self):
        a, b = C.c_double(0.0), C.c_double(0.0)
        g, s = C.c_int64(0), C.c_int64(0)
        _lib.check(self.L.s4b_sampler_last_run_stats(self.h, C.byref(a), C.byref(b), C.byref(g), C.byref(s)))
        return dict(ms_stan=a.value, ms_bart=b.value, grad_evals=int(g.value), tree_steps=int(s.value))

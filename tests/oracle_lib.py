"""ctypes binding of the CPU oracle (oracle/_build/libs4b_oracle.so).  TEST INFRASTRUCTURE:
imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs -- never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from stan4bart_b200.structs import (BartConfig, CommonControl, GlmmData, StanControl, c_double_p, c_int32_p,
                                    c_int64_p, c_uint32_p, dptr, f64)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libs4b_oracle.so")
FAST_LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libs4b_oracle_fast.so")     # -O3 -march=native: the build bench.py TIMES
TRACE_LEN = 32


def _cpu_signature():
    """The fast build uses -march=native, so it is only valid on the CPU model it was compiled on (the in-tree .so travels
    from the build container to the GPU box): remember which CPU that was."""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    import hashlib
                    return hashlib.md5(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    if force or any(not os.path.exists(q) or any(os.path.getmtime(s) > os.path.getmtime(q) for s in srcs) for q in (LIB_PATH, FAST_LIB_PATH)):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    stamp = os.path.join(ORACLE_DIR, "_build", "fast.cpu")
    sig = _cpu_signature()
    try:
        have = open(stamp).read().strip() if os.path.exists(stamp) else ""
        if have != sig:
            if have:        # built on another CPU model: rebuild for this one
                subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "-B", "_build/libs4b_oracle_fast.so"])
            with open(stamp, "w") as f:
                f.write(sig)
    except (OSError, subprocess.CalledProcessError):
        pass
    return LIB_PATH


def fast_build_is_native():
    """True when the -O3 -march=native build was compiled for the CPU this process runs on."""
    stamp = os.path.join(ORACLE_DIR, "_build", "fast.cpu")
    try:
        return open(stamp).read().strip() == _cpu_signature()
    except OSError:
        return False


_lib = None
_libs = {}
_fast = False


def use_fast(on):
    """Switch between the strict build (parity checks: the default) and the -O3 -march=native build (timing).  Objects
    must be used with the build that created them: switch only when none are alive."""
    global _fast, _lib
    _fast = bool(on)
    _lib = _libs.get(_fast)


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(FAST_LIB_PATH if (_fast and fast_build_is_native()) else LIB_PATH)
        vp = C.c_void_p
        sigs = {
            "or_bart_create": (vp, [C.POINTER(BartConfig), c_double_p, c_double_p, c_double_p]),
            "or_bart_free": (None, [vp]),
            "or_bart_set_tape": (None, [vp, c_double_p, C.c_size_t]),
            "or_bart_set_record": (None, [vp, c_double_p, C.c_size_t]),
            "or_bart_record_len": (C.c_size_t, [vp]),
            "or_bart_set_trace": (None, [vp, c_double_p, C.c_size_t]),
            "or_bart_trace_len": (C.c_size_t, [vp]),
            "or_bart_set_offset": (None, [vp, c_double_p, C.c_int]),
            "or_bart_set_sigma": (None, [vp, C.c_double]),
            "or_bart_sample_trees_from_prior": (None, [vp]),
            "or_bart_run": (None, [vp, c_double_p, c_double_p, c_uint32_p, c_double_p]),
            "or_bart_store_latents": (None, [vp, c_double_p]),
            "or_bart_get_range": (None, [vp, c_double_p]),
            "or_bart_node_assignment": (None, [vp, C.c_int, c_int64_p]),
            "or_bart_leaf_stats": (None, [vp, C.c_int, C.c_int, c_int64_p, c_int64_p, c_double_p, c_double_p, C.POINTER(C.c_int)]),
            "or_bart_num_nodes": (C.c_int64, [vp]),
            "or_bart_get_trees": (None, [vp, c_int32_p, c_int64_p, c_int32_p, c_double_p]),
            "or_bart_predict": (None, [vp, c_double_p, C.c_int64, c_double_p, c_double_p]),
            "or_bart_get_residual": (None, [vp, c_double_p]),
            "or_bart_rng_counter": (C.c_uint64, [vp]),
            "or_bart_get_k": (C.c_double, [vp]),
            "or_sampler_last_k": (C.c_int, [vp, c_double_p, C.c_int]),
            "or_glmm_create": (vp, [C.POINTER(GlmmData)]),
            "or_glmm_free": (None, [vp]),
            "or_glmm_num_params": (C.c_int, [vp]),
            "or_glmm_num_constrained": (C.c_int, [vp]),
            "or_glmm_set_offset": (None, [vp, c_double_p]),
            "or_glmm_set_response": (None, [vp, c_double_p]),
            "or_glmm_log_prob_grad": (C.c_int, [vp, c_double_p, c_double_p, c_double_p]),
            "or_glmm_write_array": (None, [vp, c_double_p, c_double_p]),
            "or_glmm_parametric_mean": (None, [vp, c_double_p, c_double_p, C.c_int, C.c_int]),
            "or_glmm_get_aux": (C.c_double, [vp, c_double_p]),
            "or_glmm_data_terms": (None, [vp, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
            "or_nuts_create": (vp, [vp, C.POINTER(StanControl), C.c_int, C.c_int]),
            "or_nuts_free": (None, [vp]),
            "or_nuts_num_pars": (C.c_int, [vp]),
            "or_nuts_run": (None, [vp, C.c_int, c_double_p]),
            "or_nuts_disengage_adaptation": (None, [vp]),
            "or_nuts_stepsize": (C.c_double, [vp]),
            "or_nuts_get_metric": (None, [vp, c_double_p]),
            "or_nuts_get_q": (None, [vp, c_double_p]),
            "or_nuts_num_grad_evals": (C.c_int64, [vp]),
            "or_sampler_create": (vp, [C.POINTER(BartConfig), c_double_p, c_double_p, c_double_p, C.POINTER(GlmmData),
                                       C.POINTER(StanControl), C.POINTER(CommonControl), c_double_p]),
            "or_sampler_free": (None, [vp]),
            "or_sampler_num_stan_pars": (C.c_int, [vp]),
            "or_sampler_run": (None, [vp, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_uint32_p, c_double_p]),
            "or_sampler_disengage_adaptation": (None, [vp]),
            "or_sampler_bart": (vp, [vp]),
            "or_sampler_nuts": (vp, [vp]),
            "or_sampler_get_range": (None, [vp, c_double_p]),
            "or_rng_qnorm": (C.c_double, [C.c_double]),
            "or_rng_uniforms": (None, [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int64, c_double_p]),
            "or_rng_truncnorm": (C.c_double, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_int]),
        }
        for name, (res, args) in sigs.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
        _libs[_fast] = L
    return _lib


class OracleBart:
    def __init__(self, cfg, y, x, x_test=None, handle=None, owner=True):
        self.cfg = cfg
        self.n, self.p, self.nt = int(cfg.n), int(cfg.p), int(cfg.n_test)
        self._owner = owner
        if handle is not None:
            self.h = handle
            return
        self._y = f64(y)
        self._x = np.asfortranarray(x, dtype=np.float64)
        self._xt = np.asfortranarray(x_test, dtype=np.float64) if x_test is not None else None
        self.h = lib().or_bart_create(C.byref(cfg), dptr(self._y), dptr(self._x), dptr(self._xt))
        assert self.h

    def __del__(self):
        if getattr(self, "h", None) and self._owner:
            lib().or_bart_free(self.h)
            self.h = None

    def set_tape(self, tape):
        self._tape = f64(tape)
        lib().or_bart_set_tape(self.h, dptr(self._tape), len(self._tape))

    def set_record(self, cap):
        self._rec = np.zeros(cap)
        lib().or_bart_set_record(self.h, dptr(self._rec), cap)

    def record(self):
        k = lib().or_bart_record_len(self.h)
        assert k <= len(self._rec), "record buffer too small"
        return self._rec[:k].copy()

    def set_trace(self, cap):
        self._trace = np.zeros((cap, TRACE_LEN))
        lib().or_bart_set_trace(self.h, dptr(self._trace), cap)

    def trace(self):
        return self._trace[:lib().or_bart_trace_len(self.h)].copy()

    def set_offset(self, offset, update_scale):
        o = f64(offset) if offset is not None else None
        lib().or_bart_set_offset(self.h, dptr(o), int(update_scale))

    def set_sigma(self, sigma):
        lib().or_bart_set_sigma(self.h, float(sigma))

    def sample_trees_from_prior(self):
        lib().or_bart_sample_trees_from_prior(self.h)

    def run(self):
        train = np.zeros(self.n)
        test = np.zeros(self.nt) if self.nt else None
        vc = np.zeros(self.p, dtype=np.uint32)
        sig = C.c_double(0.0)
        lib().or_bart_run(self.h, dptr(train), dptr(test), vc.ctypes.data_as(c_uint32_p), C.byref(sig))
        return dict(train=train, test=test, varcount=vc, sigma=sig.value)

    def latents(self):
        out = np.zeros(self.n)
        lib().or_bart_store_latents(self.h, dptr(out))
        return out

    def data_range(self):
        out = np.zeros(3)
        lib().or_bart_get_range(self.h, dptr(out))
        return out

    def node_assignment(self, tree):
        out = np.zeros(self.n, dtype=np.int64)
        lib().or_bart_node_assignment(self.h, tree, out.ctypes.data_as(c_int64_p))
        return out

    def leaf_stats(self, tree, max_leaves=64):
        heap = np.zeros(max_leaves, dtype=np.int64)
        cnt = np.zeros(max_leaves, dtype=np.int64)
        s = np.zeros(max_leaves)
        ss = np.zeros(max_leaves)
        nl = C.c_int(0)
        lib().or_bart_leaf_stats(self.h, tree, max_leaves, heap.ctypes.data_as(c_int64_p), cnt.ctypes.data_as(c_int64_p),
                                 dptr(s), dptr(ss), C.byref(nl))
        k = nl.value
        return heap[:k], cnt[:k], s[:k], ss[:k]

    def trees(self):
        k = lib().or_bart_num_nodes(self.h)
        tree_no = np.zeros(k, dtype=np.int32)
        n_obs = np.zeros(k, dtype=np.int64)
        var = np.zeros(k, dtype=np.int32)
        value = np.zeros(k)
        lib().or_bart_get_trees(self.h, tree_no.ctypes.data_as(c_int32_p), n_obs.ctypes.data_as(c_int64_p),
                                var.ctypes.data_as(c_int32_p), dptr(value))
        return dict(tree=tree_no, n=n_obs, var=var, value=value)

    def predict(self, x_test, offset=None):
        xt = np.asfortranarray(x_test, dtype=np.float64)
        out = np.zeros(xt.shape[0])
        o = f64(offset) if offset is not None else None
        lib().or_bart_predict(self.h, dptr(xt), xt.shape[0], dptr(o), dptr(out))
        return out

    def residual(self):
        out = np.zeros(self.n)
        lib().or_bart_get_residual(self.h, dptr(out))
        return out

    def rng_counter(self):
        return int(lib().or_bart_rng_counter(self.h))

    def k(self):
        return float(lib().or_bart_get_k(self.h))


class OracleGlmm:
    def __init__(self, stan_data):
        self.sd = stan_data
        self._struct = stan_data.struct()
        self.h = lib().or_glmm_create(C.byref(self._struct))
        assert self.h, "oracle GLMM rejected the data (unsupported branch)"
        self.d = lib().or_glmm_num_params(self.h)
        self.nc = lib().or_glmm_num_constrained(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_glmm_free(self.h)
            self.h = None

    def set_offset(self, o):
        lib().or_glmm_set_offset(self.h, dptr(f64(o)))

    def set_response(self, y):
        lib().or_glmm_set_response(self.h, dptr(f64(y)))

    def log_prob_grad(self, q):
        q = f64(q)
        g = np.zeros(self.d)
        lp = C.c_double(0.0)
        status = lib().or_glmm_log_prob_grad(self.h, dptr(q), C.byref(lp), dptr(g))
        return lp.value, g, status

    def write_array(self, q):
        out = np.zeros(self.nc)
        lib().or_glmm_write_array(self.h, dptr(f64(q)), dptr(out))
        return out

    def parametric_mean(self, constrained, fixed=True, random=True):
        out = np.zeros(self.sd.N)
        lib().or_glmm_parametric_mean(self.h, dptr(f64(constrained)), dptr(out), int(fixed), int(random))
        return out

    def data_terms(self, beta, b):
        S = C.c_double(0.0)
        gbeta = np.zeros(max(1, self.sd.K))
        gb = np.zeros(max(1, self.sd.q))
        lib().or_glmm_data_terms(self.h, dptr(f64(beta)), dptr(f64(b)), C.byref(S), dptr(gbeta), dptr(gb))
        return S.value, gbeta[:self.sd.K], gb[:self.sd.q]


class OracleNuts:
    def __init__(self, glmm, ctl, chain_id=1, num_warmup=1000):
        self.glmm = glmm
        self.ctl = ctl
        self.h = lib().or_nuts_create(glmm.h, C.byref(ctl), chain_id, num_warmup)
        self.num_pars = lib().or_nuts_num_pars(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_nuts_free(self.h)
            self.h = None

    def run(self, warmup=True):
        out = np.zeros(self.num_pars)
        lib().or_nuts_run(self.h, int(warmup), dptr(out))
        return out

    def disengage_adaptation(self):
        lib().or_nuts_disengage_adaptation(self.h)

    def stepsize(self):
        return lib().or_nuts_stepsize(self.h)

    def metric(self):
        out = np.zeros(self.glmm.d)
        lib().or_nuts_get_metric(self.h, dptr(out))
        return out

    def q(self):
        out = np.zeros(self.glmm.d)
        lib().or_nuts_get_q(self.h, dptr(out))
        return out

    def num_grad_evals(self):
        return int(lib().or_nuts_num_grad_evals(self.h))


class OracleSampler:
    """Mirror of the reference's `.Call` surface (src/init.cpp:1215-1229) on the CPU oracle."""

    def __init__(self, bart_cfg, y, x_bart, x_test, stan_data, stan_ctl, warmup, iter_, keep_fits=True, sigma_init=1.0,
                 bart_offset_init=None, offset=None, offset_type=0):
        self.bcfg = bart_cfg
        self.sd = stan_data
        self._gs = stan_data.struct()
        self._y = f64(y)
        self._x = np.asfortranarray(x_bart, dtype=np.float64)
        self._xt = np.asfortranarray(x_test, dtype=np.float64) if x_test is not None else None
        self._off = f64(bart_offset_init) if bart_offset_init is not None else None
        self._user_offset = f64(offset) if offset is not None else None
        self.cc = CommonControl(warmup=warmup, iter=iter_, is_binary=int(stan_data.is_binary), keep_fits=int(keep_fits),
                                sigma_init=float(sigma_init), offset_type=int(offset_type), reserved=0, user_offset=dptr(self._user_offset))
        self.keep_fits = keep_fits
        self.h = lib().or_sampler_create(C.byref(bart_cfg), dptr(self._y), dptr(self._x), dptr(self._xt), C.byref(self._gs),
                                         C.byref(stan_ctl), C.byref(self.cc), dptr(self._off))
        assert self.h
        self.num_pars = lib().or_sampler_num_stan_pars(self.h)
        self.n, self.nt, self.p = int(bart_cfg.n), int(bart_cfg.n_test), int(bart_cfg.p)

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_sampler_free(self.h)
            self.h = None

    def bart(self):
        return OracleBart(self.bcfg, None, None, handle=lib().or_sampler_bart(self.h), owner=False)

    def run(self, num_iter, is_warmup):
        S = num_iter if self.keep_fits else 1
        stan = np.zeros((S, self.num_pars))
        train = np.zeros((S, self.n))
        test = np.zeros((S, max(self.nt, 1)))
        vc = np.zeros((S, self.p), dtype=np.uint32)
        sigma = np.zeros(S)
        lib().or_sampler_run(self.h, num_iter, int(is_warmup), dptr(stan), dptr(train), dptr(test),
                             vc.ctypes.data_as(c_uint32_p), dptr(sigma))
        k = np.zeros(num_iter)
        m = lib().or_sampler_last_k(self.h, dptr(k), num_iter)
        return dict(stan=stan.T, bart=dict(train=train.T, test=test.T[:self.nt], varcount=vc.T, sigma=sigma, k=k[:m]))

    def disengage_adaptation(self):
        lib().or_sampler_disengage_adaptation(self.h)

    def data_range(self):
        out = np.zeros(2)
        lib().or_sampler_get_range(self.h, dptr(out))
        return out

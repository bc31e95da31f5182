"""ctypes loader for libstan4bart_b200.so (the C ABI of include/stan4bart_b200.h).

The product path fails loudly when the CUDA library is missing or no device is present:
there is no CPU fallback and nothing here ever touches oracle/."""
import ctypes as C
import os

from .structs import (BartConfig, CommonControl, GlmmData, StanControl, c_double_p, c_int32_p, c_int64_p, c_uint32_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libstan4bart_b200.so")

vp = C.c_void_p
vpp = C.POINTER(C.c_void_p)
c_size_p = C.POINTER(C.c_size_t)
c_int_p = C.POINTER(C.c_int)
c_uint64_p = C.POINTER(C.c_uint64)
c_ubyte_p = C.POINTER(C.c_ubyte)
ITERATION_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, c_double_p, c_double_p, c_double_p)

# name -> (restype, argtypes); every symbol declared in include/stan4bart_b200.h
SIGNATURES = {
    "s4b_last_error": (C.c_char_p, []),
    "s4b_device_count": (C.c_int, []),
    "s4b_set_stream": (C.c_int, [vp]),
    "s4b_set_device": (C.c_int, [C.c_int]),
    "gpubart_tree_step_ms": (C.c_int, [vp, C.c_int, c_double_p]),
    "gpubart_get_profile": (C.c_int, [vp, c_uint64_p, C.c_int]),
    "gpubart_set_profile": (C.c_int, [vp, C.c_int]),
    "gpubart_stored_export_size": (C.c_int, [vp, c_int64_p]),
    "gpubart_stored_export": (C.c_int, [vp, vp, C.c_int64]),
    "gpubart_stored_import": (C.c_int, [vp, C.c_int64, vpp]),
    "gpubart_stored_free": (C.c_int, [vp]),
    "gpubart_stored_count": (C.c_int, [vp, c_int64_p]),
    "gpubart_stored_predict": (C.c_int, [vp, c_double_p, C.c_int64, c_double_p, C.c_int64, C.c_int64, c_double_p]),
    "gpubart_summary": (C.c_int, [vp, C.c_char_p, C.c_size_t, c_size_p]),
    "gpubart_set_keep_trees": (C.c_int, [vp, C.c_int64]),
    "gpubart_num_stored": (C.c_int, [vp, c_int64_p]),
    "gpubart_predict_stored": (C.c_int, [vp, c_double_p, C.c_int64, c_double_p, C.c_int64, C.c_int64, c_double_p]),
    "gpubart_num_stored_nodes": (C.c_int, [vp, C.c_int64, c_int64_p]),
    "gpubart_get_stored_trees": (C.c_int, [vp, C.c_int64, c_int32_p, c_int64_p, c_int32_p, c_double_p]),
    "s4b_sampler_set_host_plumbing": (C.c_int, [vp, C.c_int, c_int64_p, c_int64_p]),
    "gpubart_create": (C.c_int, [C.POINTER(BartConfig), c_double_p, c_double_p, c_double_p, vpp]),
    "gpubart_free": (C.c_int, [vp]),
    "gpubart_set_offset": (C.c_int, [vp, c_double_p, C.c_int]),
    "gpubart_set_sigma": (C.c_int, [vp, C.c_double]),
    "gpubart_get_k": (C.c_int, [vp, c_double_p]),
    "gpubart_sample_trees_from_prior": (C.c_int, [vp]),
    "gpubart_run_sampler_with_results": (C.c_int, [vp, c_double_p, c_double_p, c_uint32_p, c_double_p]),
    "gpubart_run_batched": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "gpubart_collect_results": (C.c_int, [vp, c_double_p, c_double_p, c_uint32_p, c_double_p]),
    "gpubart_store_latents": (C.c_int, [vp, c_double_p]),
    "gpubart_get_data_range": (C.c_int, [vp, c_double_p]),
    "gpubart_predict": (C.c_int, [vp, c_double_p, C.c_int64, c_double_p, c_double_p]),
    "gpubart_num_nodes": (C.c_int, [vp, c_int64_p]),
    "gpubart_get_trees": (C.c_int, [vp, c_int32_p, c_int64_p, c_int32_p, c_double_p]),
    "gpubart_node_assignment": (C.c_int, [vp, C.c_int, c_int64_p]),
    "gpubart_leaf_stats": (C.c_int, [vp, C.c_int, C.c_int, c_int64_p, c_int64_p, c_double_p, c_double_p, c_int_p]),
    "gpubart_get_residual": (C.c_int, [vp, c_double_p]),
    "gpubart_set_trace": (C.c_int, [vp, C.c_size_t]),
    "gpubart_get_trace": (C.c_int, [vp, c_double_p, C.c_size_t, c_size_p]),
    "gpubart_set_tape": (C.c_int, [vp, c_double_p, C.c_size_t]),
    "gpubart_set_record": (C.c_int, [vp, C.c_size_t]),
    "gpubart_get_record": (C.c_int, [vp, c_double_p, C.c_size_t, c_size_p]),
    "gpubart_rng_counter": (C.c_int, [vp, c_uint64_p]),
    "gpubart_set_use_graph": (C.c_int, [vp, C.c_int]),
    "gpubart_set_sweep_mode": (C.c_int, [vp, C.c_int]),
    "gpubart_get_sweep_mode": (C.c_int, [vp, c_int_p]),
    "gpubart_time_leaf_stats": (C.c_int, [vp, C.c_int, C.c_int, c_double_p]),
    "gpubart_num_tree_steps": (C.c_int, [vp, c_int64_p]),
    "glmm_create": (C.c_int, [C.POINTER(GlmmData), vpp]),
    "glmm_free": (C.c_int, [vp]),
    "glmm_num_params": (C.c_int, [vp, c_int_p, c_int_p]),
    "glmm_set_offset": (C.c_int, [vp, c_double_p]),
    "glmm_set_response": (C.c_int, [vp, c_double_p]),
    "glmm_log_prob_grad": (C.c_int, [vp, c_double_p, c_double_p, c_double_p, c_int_p]),
    "glmm_write_array": (C.c_int, [vp, c_double_p, c_double_p]),
    "glmm_parametric_mean": (C.c_int, [vp, c_double_p, c_double_p, C.c_int, C.c_int]),
    "glmm_data_terms": (C.c_int, [vp, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]),
    "glmm_num_grad_evals": (C.c_int, [vp, c_int64_p]),
    "glmm_set_mode": (C.c_int, [vp, C.c_int]),
    "glmm_get_mode": (C.c_int, [vp, c_int_p]),
    "glmm_num_device_passes": (C.c_int, [vp, c_int64_p]),
    "s4b_batch_group_create": (C.c_int, [C.c_int, vpp]),
    "s4b_batch_group_free": (C.c_int, [vp]),
    "s4b_batch_group_launches": (C.c_int, [vp, c_int64_p]),
    "s4b_sampler_set_batch_group": (C.c_int, [vp, vp]),
    "glmm_nuts_create": (C.c_int, [vp, C.POINTER(StanControl), C.c_int, C.c_int, vpp]),
    "glmm_nuts_free": (C.c_int, [vp]),
    "glmm_nuts_num_pars": (C.c_int, [vp, C.POINTER(C.c_int)]),
    "glmm_nuts_run": (C.c_int, [vp, C.c_int, c_double_p]),
    "glmm_nuts_disengage_adaptation": (C.c_int, [vp]),
    "glmm_nuts_stepsize": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "glmm_time_data_pass": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "s4b_sampler_create": (C.c_int, [C.POINTER(BartConfig), c_double_p, c_double_p, c_double_p, C.POINTER(GlmmData),
                                     C.POINTER(StanControl), C.POINTER(CommonControl), c_double_p, vpp]),
    "s4b_sampler_free": (C.c_int, [vp]),
    "s4b_sampler_num_stan_pars": (C.c_int, [vp, c_int_p]),
    "s4b_sampler_run": (C.c_int, [vp, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_uint32_p, c_double_p]),
    "s4b_sampler_disengage_adaptation": (C.c_int, [vp]),
    "s4b_sampler_get_bart_data_range": (C.c_int, [vp, c_double_p]),
    "s4b_sampler_get_parametric_mean": (C.c_int, [vp, c_double_p]),
    "s4b_sampler_predict_bart": (C.c_int, [vp, c_double_p, C.c_int64, c_double_p, c_double_p]),
    "s4b_sampler_bart": (vp, [vp]),
    "s4b_sampler_glmm": (vp, [vp]),
    "s4b_sampler_get_means": (C.c_int, [vp, c_double_p, c_double_p, c_double_p, c_int64_p]),
    "s4b_sampler_last_run_stats": (C.c_int, [vp, c_double_p, c_double_p, c_int64_p, c_int64_p]),
    "s4b_sampler_set_callback": (C.c_int, [vp, ITERATION_CALLBACK, vp]),
    "s4b_sampler_last_k": (C.c_int, [vp, c_double_p, C.c_int, C.POINTER(C.c_int)]),
    "s4b_shard_create": (C.c_int, [C.c_int, C.c_int, vpp]),
    "s4b_shard_free": (C.c_int, [vp]),
    "s4b_shard_ipc_handle": (C.c_int, [vp, c_ubyte_p]),
    "s4b_shard_attach": (C.c_int, [vp, c_ubyte_p]),
    "s4b_shard_set_obs_range": (C.c_int, [vp, C.c_int64, C.c_int64]),
    "s4b_shard_allreduce": (C.c_int, [vp, c_double_p, C.c_int64, C.c_int]),
    "glmm_stan_row_names": (C.c_int, [vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "gpubart_set_pipeline": (C.c_int, [vp, C.c_int]),
    "gpubart_get_pipeline": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "gpubart_pipeline_misfits": (C.c_int, [vp, c_uint32_p]),
    "gpubart_set_keep_trees_active": (C.c_int, [vp, C.c_int]),
    "gpubart_set_response": (C.c_int, [vp, c_double_p]),
    "gpubart_get_stored_scales": (C.c_int, [vp, C.c_int64, C.c_int64, c_double_p]),
    "gpubart_stored_get_scales": (C.c_int, [vp, C.c_int64, C.c_int64, c_double_p]),
    "s4b_shard_nccl_unique_id": (C.c_int, [c_ubyte_p]),
    "s4b_shard_nccl_init": (C.c_int, [vp, c_ubyte_p]),
    "s4b_shard_use_nccl": (C.c_int, [vp, C.c_int]),
    "gpubart_create_sharded": (C.c_int, [C.POINTER(BartConfig), c_double_p, c_double_p, c_double_p, vp, vpp]),
    "glmm_create_sharded": (C.c_int, [C.POINTER(GlmmData), vp, vpp]),
    "s4b_sampler_create_sharded": (C.c_int, [C.POINTER(BartConfig), c_double_p, c_double_p, c_double_p, C.POINTER(GlmmData),
                                             C.POINTER(StanControl), C.POINTER(CommonControl), c_double_p, vp, vpp]),
}

_lib = None


class S4BError(RuntimeError):
    pass


def load():
    """Load the shared library and bind every declared symbol (raises if one is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise S4BError(f"{LIB_PATH} not found: build it with `python -m stan4bart_b200.build` "
                           "(stan4bart_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise S4BError(load().s4b_last_error().decode("utf-8", "replace"))


def require_device():
    if load().s4b_device_count() < 1:
        raise S4BError("no CUDA device visible: stan4bart_b200 has no CPU fallback")

"""CPU: the C-ABI shared library loads, exports every symbol the header declares, and refuses to
compute without a GPU (no CPU fallback, no oracle on the product path)."""
import os
import re

import numpy as np
import pytest

from stan4bart_b200 import _lib
from stan4bart_b200.structs import bart_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "stan4bart_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b((?:gpubart|glmm|s4b)_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 45
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/stan4bart_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)


def test_struct_layouts_match_the_header():
    import ctypes as C
    from stan4bart_b200.structs import BartConfig, CommonControl, GlmmData, StanControl
    assert C.sizeof(BartConfig) == 3 * 8 + 6 * 4 + 8 * 8 + 8 + 8 + 8 + 2 * 8 + 8 + 2 * 4    # + the split_probs and weights pointers, k_df, k_scale, n_cuts_var, change_symmetric + pad
    assert C.sizeof(StanControl) == 2 * 4 + 5 * 8 + 4 * 4 + 2 * 8
    assert C.sizeof(CommonControl) == 4 * 4 + 8 + 2 * 4 + 8            # + offset_type, reserved, user_offset
    assert C.sizeof(GlmmData) == 8 + 10 * 4 + 8 + 4 * 8 + 3 * 8 + 9 * 8 + 8 + 2 * 8 + 4 * 8   # + weights, prior_df, num_normals, hs hyper-parameters


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "stan4bart_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle_lib" not in text and "s4b_oracle" not in text and "libs4b_oracle" not in text, f


def test_no_device_means_loud_failure():
    L = _lib.load()
    if L.s4b_device_count() > 0:
        pytest.skip("a GPU is present")
    from stan4bart_b200.sampler import GpuBart
    x = np.asfortranarray(np.random.default_rng(0).random((20, 2)))
    with pytest.raises(_lib.S4BError, match="no CUDA device"):
        GpuBart(bart_config(20, 2, num_trees=2), np.zeros(20), x)

"""Boundary b2 as code: stan4bart's dbarts function table (BARTFunctionTable, /root/reference/src/init.cpp:54-81) bound to the
GPU sampler through glue/gpubart_shim.cpp and the shim headers include/dbarts_shim/dbarts/*.hpp.

tests/shim_driver.cpp fills the 20-entry table by name and drives it in the reference's call order (createSampler, the
warm-up and sampling runs with their setControl(keepTrees) switch, predict on the stored sampler, getTrees, printTrees,
setResponse).  Everything it saw must equal the direct C ABI (`gpubart_*`) run with the same seeds: both paths end in the same
device code, so training / test fits, latents, variable counts and trees are compared bit for bit."""
import os
import struct
import subprocess
import tempfile

import numpy as np
import pytest

from common import bart_problem
from stan4bart_b200 import build as B
from stan4bart_b200.sampler import GpuBart
from stan4bart_b200.structs import bart_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TABLE = ["initializeFit", "invalidateFit", "initializeControl", "initializeData", "invalidateData", "initializeModel", "invalidateModel",
         "createStateExpression", "initializeState", "setControl", "runSamplerWithResults", "predict", "setResponse", "setOffset", "setSigma",
         "sampleTreesFromPrior", "printInitialSummary", "storeLatents", "printTrees", "getTrees"]


def build_driver(tmp):
    exe = os.path.join(tmp, "shim_driver")
    lib_dir = os.path.dirname(B.SHIM_LIB)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include", "dbarts_shim"), "-I", os.path.join(ROOT, "glue"),
           os.path.join(ROOT, "tests", "shim_driver.cpp"), "-o", exe, "-L", lib_dir, "-lgpubart_shim", "-lstan4bart_b200", "-Wl,-rpath," + lib_dir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_shim_library_exports_the_twenty_table_entries_with_reference_signatures():
    """CPU: the library loads, every name stan4bart looks up (init.cpp:1115-1146) resolves, and a translation unit that declares the
    reference's BARTFunctionTable and assigns the shim's functions to it compiles (identical signatures)."""
    import ctypes as C
    B.build_shim()
    L = C.CDLL(B.SHIM_LIB)
    L.gpubart_shim_lookup.restype = C.c_void_p
    L.gpubart_shim_lookup.argtypes = [C.c_char_p]
    L.gpubart_shim_entry_name.restype = C.c_char_p
    assert L.gpubart_shim_num_entries() == 20
    assert sorted(L.gpubart_shim_entry_name(i).decode() for i in range(20)) == sorted(TABLE)
    for name in TABLE:
        assert L.gpubart_shim_lookup(name.encode()), name
        assert getattr(L, "gpubart_shim_" + name)
    assert not L.gpubart_shim_lookup(b"noSuchEntry")
    with tempfile.TemporaryDirectory() as tmp:       # the driver assigns every entry to a table member of the reference's type
        build_driver(tmp)


def _write_problem(path, n, p, nt, trees, binary, warm, iters, seed, kmod, y, x, xt, off0, off1, y2):
    with open(path, "wb") as f:
        f.write(struct.pack("<10q", n, p, nt, trees, int(binary), warm, iters, seed, int(kmod), 0))
        for a in (y, x.ravel(order="F"), xt.ravel(order="F") if nt else np.zeros(0), off0, off1, y2):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())


def _read_out(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            name = f.read(16)
            if len(name) < 16:
                break
            (k,) = struct.unpack("<Q", f.read(8))
            out[name.rstrip(b"\0").decode()] = np.frombuffer(f.read(8 * k), dtype=np.float64).copy()
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("binary,kmod", [(False, False), (True, False), (False, True)])
def test_table_driven_run_equals_the_direct_c_abi(binary, kmod):
    n, p, nt, trees, warm, iters, seed = 700, 5, 300, 7, 6, 5, 424242
    x, y, xt = bart_problem(n=n, p=p, n_test=nt, binary=binary, seed=12)
    rng = np.random.default_rng(3)
    off0, off1 = 0.3 * np.cos(np.arange(n) * 0.01), 0.2 * rng.standard_normal(n)
    y2 = (1.0 - y) if binary else y + 0.5 * rng.standard_normal(n)
    with tempfile.TemporaryDirectory() as tmp:
        exe = build_driver(tmp)
        prob, outp = os.path.join(tmp, "prob.bin"), os.path.join(tmp, "out.bin")
        _write_problem(prob, n, p, nt, trees, binary, warm, iters, seed, kmod, y, x, xt, off0, off1, y2)
        r = subprocess.run([exe, prob, outp], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        got = _read_out(outp)
        printed = r.stdout

    cfg = bart_config(n, p, n_test=nt, num_trees=trees, is_binary=binary, seed=seed, n_cuts=np.full(p, 100), k_df=1.25 if kmod else 0.0,
                      k_scale=0.0 if kmod else float("inf"))
    g = GpuBart(cfg, y, x, xt)
    g.set_offset(off0, True)
    if not binary:
        g.set_sigma(1.3)
    g.sample_trees_from_prior()
    first = g.run()
    assert np.array_equal(got["first_train"], first["train"]) and np.array_equal(got["first_test"], first["test"])
    if binary:
        assert np.array_equal(got["first_latents"], g.latents())
    for phase, num_iter in (("warm", warm), ("samp", iters)):
        is_warm = phase == "warm"
        if not is_warm:
            g.set_keep_trees(iters)              # setControl(keepTrees = TRUE) before the sampling run (init.cpp:737-743)
        tr, te, sg, vc, ks = [], [], [], [], []
        for it in range(num_iter):
            if not binary:
                g.set_sigma(1.0 + 0.05 * it)
            g.set_offset(off1 if it % 2 == 0 else off0, is_warm and it % (1 << (8 * it // num_iter)) == 0)
            r = g.run()
            tr.append(r["train"]); te.append(r["test"]); sg.append(r["sigma"]); vc.append(r["varcount"].astype(np.float64)); ks.append(g.k())
        assert np.array_equal(got[phase + "_train"], np.concatenate(tr)), phase
        assert np.array_equal(got[phase + "_test"], np.concatenate(te))
        assert np.array_equal(got[phase + "_sigma"], np.array(sg))
        assert np.array_equal(got[phase + "_varcount"], np.concatenate(vc))
        if kmod:
            assert np.array_equal(got[phase + "_k"], np.array(ks)) and len(set(ks)) == num_iter
        if binary:
            assert np.array_equal(got[phase + "_latents"], g.latents())
    assert np.array_equal(got["data_range"], g.data_range())
    assert got["num_stored"][0] == iters == g.num_stored()            # warm-up draws took no slots of the store
    # predict on the stored draws: the live fit's scale, and the identity scale of a re-imported stored sampler
    pred = g.predict_stored(x)                                       # [n x iters], original units
    scales = g.stored_scales()
    assert np.allclose(got["pred_live"].reshape(iters, n).T, pred, rtol=1e-12, atol=1e-12)
    internal = pred if binary else (pred - scales[:, 0]) / scales[:, 1] - 0.5
    assert np.allclose(got["pred_stored"].reshape(iters, n).T, internal, rtol=1e-12, atol=1e-12)
    if not binary:     # R un-scales with the chain's range (R/generics.R:671-674): min + (0.5 + f) range gives the original units back
        rng_ = g.data_range()
        assert np.allclose(rng_[0] + (0.5 + got["pred_stored"].reshape(iters, n).T) * rng_[2], pred, rtol=1e-10, atol=1e-10)
    # getTrees: stored draw 0 (all trees), live trees 0 and 2
    st = g.stored_trees(0)
    flat = got["trees_stored0"].reshape(-1, 4)
    assert np.array_equal(flat[:, 0], st["tree"]) and np.array_equal(flat[:, 1], st["n"]) and np.array_equal(flat[:, 2], st["var"])
    assert np.array_equal(flat[:, 3], st["value"])
    lt = g.trees()
    sel = np.isin(lt["tree"], [0, 2])
    flat = got["trees_live02"].reshape(-1, 4)
    assert np.array_equal(flat[:, 0], lt["tree"][sel]) and np.array_equal(flat[:, 2], lt["var"][sel]) and np.array_equal(flat[:, 3], lt["value"][sel])
    # printTrees / printInitialSummary wrote text
    assert "tree 2, sample 1" in printed and "mu = " in printed and "power and base for tree prior" in printed
    # setResponse + one more draw (store switched off again)
    g.set_keep_trees_active(False)
    g.set_response(y2)
    assert np.array_equal(got["after_setresp"], g.run()["train"])
    assert g.num_stored() == iters

"""The R-facing glue type-checks against R's C API: glue/init_gpu.cpp (the 12 `.Call` routines of
/root/reference/src/init.cpp:1215-1229 over the C ABI) and the R branch of glue/gpubart_shim.cpp (S4 parsers of dbartsControl /
dbartsData / dbartsModel, R_RegisterCCallable binding).  R is not in the image, so the check is `g++ -fsyntax-only` against the
declarations of tests/r_stub/ (signatures as in R >= 4.0); it proves the glue is well-formed C++ that uses the R API and our
C ABI consistently, not that it runs."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = ["-I", os.path.join(ROOT, "tests", "r_stub"), "-I", os.path.join(ROOT, "include", "dbarts_shim"), "-I", os.path.join(ROOT, "glue")]


@pytest.mark.parametrize("src", ["init_gpu.cpp", "gpubart_shim.cpp"])
def test_r_glue_type_checks_against_the_r_api(src):
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-DGPUBART_SHIM_WITH_R"] + INC + [os.path.join(ROOT, "glue", src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_call_table_has_the_reference_routines_and_arities():
    want = {"stan4bart_create": 6, "stan4bart_run": 4, "stan4bart_printInitialSummary": 1, "stan4bart_disengageAdaptation": 1,
            "stan4bart_finalize": 0, "stan4bart_exportBARTState": 1, "stan4bart_createStoredBARTSampler": 4, "stan4bart_predictBART": 3,
            "stan4bart_getParametricMean": 1, "stan4bart_getBARTDataRange": 1, "stan4bart_printTrees": 4, "stan4bart_getTrees": 5}
    text = open(os.path.join(ROOT, "glue", "init_gpu.cpp")).read()
    got = {m.group(1): int(m.group(2)) for m in re.finditer(r'DEF_FUNC\("(\w+)",\s*\w+,\s*(\d+)\)', text)}
    assert got == want

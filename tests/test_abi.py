"""CPU tests of the drop-in boundary: libgpb200.so loads and exports every symbol that
include/gpb200.h declares, the ctypes table mirrors the header, and -- with no GPU -- the library
fails loudly instead of falling back to a CPU path."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    return sorted(set(re.findall(r"GPB200_API[^;(]*?\b(gpb200_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from gp_b200 import capi
    lib = capi.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(capi.SIGNATURES) == syms
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r" T (gpb200_\w+)", out)))
    assert exported == syms
    assert lib.gpb200_version() == 100


def test_sass_is_sm100a_dmma():
    """The built library contains sm_100a code whose O(N^3) kernels use the FP64 tensor path
    (SASS DMMA) and cp.async (LDGSTS) -- and no other architecture."""
    from gp_b200 import capi
    out = subprocess.check_output(["cuobjdump", "-lelf", capi.LIB_PATH], text=True)
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    sass = subprocess.check_output(["cuobjdump", "-sass", capi.LIB_PATH], text=True)
    # split per function and look at the NT instance of the tile GEMM
    chunks = sass.split("Function : ")
    gemm = [c for c in chunks if c.startswith("_ZN3gpb16gemm_tile_kernel")]
    # three configurations x (NT, TN, NN, TT, TN + SE trace, TN + derivative trace, NT + SE trace), plus the sixteen-warp
    # fine tiles of the latency path (NT, TN, TT; 16x16 warp tiles: 16 DMMAs per unrolled k-chunk)
    assert len(gemm) == 24
    for c in gemm:
        assert c.count("DMMA.8x8x4") >= 16 and "LDGSTS" in c
    # no library GEMM/solver is linked: the O(N^3) work is ours
    ldd = subprocess.check_output(["ldd", capi.LIB_PATH], text=True)
    assert "cublas" not in ldd and "cusolver" not in ldd


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gp_b200 import capi
    with pytest.raises(capi.GpB200Error):
        capi.Handle(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU oracle", ""), os.path.join(dirpath, f)
    # the measurement/profiling tools, the examples and the R / Stan bindings do not touch it either
    for sub in ("tools", "examples", "r", "stan"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".sh", ".c", ".hpp", ".R", ".cu")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)


def test_r_shim_type_checks_against_mock_r_api():
    """R is not installed here, so r/shim.c cannot be built; it is at least type-checked against a mock
    of the R C API (r/mock/) and the real include/gpb200.h, and every .Call name used by r/R/gpb200.R
    is registered in the shim."""
    subprocess.check_call(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-cast-function-type", "-Werror",
                           "-I" + os.path.join(ROOT, "r", "mock"), "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "r", "shim.c")])
    shim = open(os.path.join(ROOT, "r", "shim.c")).read()
    rfile = open(os.path.join(ROOT, "r", "R", "gpb200.R")).read()
    registered = set(re.findall(r'\{"(gp_\w+)", \(DL_FUNC\)', shim))
    called = set(re.findall(r'\.Call\("(gp_\w+)"', rfile))
    assert called and called <= registered, called - registered
    # every gpb200_* the shim calls is declared in the header
    used = set(re.findall(r"\b(gpb200_\w+)\s*\(", shim))
    assert used <= set(header_symbols()) | {"gpb200_handle_t"}


def test_stan_header_type_checks_against_mock_stan():
    """Stan Math / Eigen are absent, so stan/gp_lml_stan.hpp cannot be built into a model here; it is
    type-checked as C++11 (the reference era's rstan toolchain) against a minimal mock of the names
    it uses (stan/mock/), for all-var, mixed and all-double argument lists."""
    subprocess.check_call(["g++", "-std=c++11", "-fsyntax-only", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "stan", "mock"),
                           os.path.join(ROOT, "stan", "mock", "check.cpp")])

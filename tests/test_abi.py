"""Drop-in boundary checks that need no GPU: the shared library loads, exports every symbol the headers
declare, and the state / operator structs have the reference's layout."""
import ctypes as C
import os
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from lobpcg_b200 import api

ROOT = Path(__file__).resolve().parent.parent
INC = ROOT / "include"
REFROOT = Path("/root/reference")

PROBE = r"""
#include <stdio.h>
#include <stddef.h>
%s
#define F(TY, f) printf(#TY "." #f " %%zu\n", offsetof(TY, f));
#define STATE(TY) printf(#TY " %%zu\n", sizeof(TY)); F(TY,S) F(TY,Cx) F(TY,Cp) F(TY,AX) F(TY,AS) F(TY,BS) \
  F(TY,eigVals) F(TY,resNorm) F(TY,signature) F(TY,wrk1) F(TY,wrk2) F(TY,wrk3) F(TY,wrk4) F(TY,rr_D) \
  F(TY,rr_eigvals) F(TY,rr_tau) F(TY,rr_VR) F(TY,rr_sig) F(TY,rr_indices) F(TY,rr_ggev) \
  F(TY,implicit_product_update) F(TY,verbosity) F(TY,iter) F(TY,nev) F(TY,converged) F(TY,size) F(TY,sizeSub) \
  F(TY,maxIter) F(TY,tol) F(TY,A) F(TY,B) F(TY,T)
#define OP(TY) printf(#TY " %%zu\n", sizeof(TY)); F(TY,rows) F(TY,cols) F(TY,matvec) F(TY,cleanup) F(TY,ctx)
int main(void) {
  STATE(s_lobpcg_t) STATE(d_lobpcg_t) STATE(c_lobpcg_t) STATE(z_lobpcg_t)
  OP(LinearOperator_s_t) OP(LinearOperator_d_t) OP(LinearOperator_c_t) OP(LinearOperator_z_t)
  printf("linop_ctx_t %%zu\n", sizeof(linop_ctx_t));
  return 0;
}
"""


def _layout(includes, flags, tmp_path, name):
    src = tmp_path / f"{name}.c"
    src.write_text(PROBE % includes)
    exe = tmp_path / name
    subprocess.run(["gcc", "-std=c11", "-w", *flags, str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    return dict(line.rsplit(" ", 1) for line in out.strip().splitlines())


def test_library_loads_and_exports_declared_symbols():
    L = api.lib()
    assert b"sm_100a" in L.lb2_version()
    hdr = (INC / "lobpcg_b200.h").read_text()
    names = set(re.findall(r"\b(lb2_[a-z0-9_]+)\s*\(", hdr))
    names = {n for n in names if "##" not in n}
    kern = set(re.findall(r"lb2_##P##_([a-z0-9_]+)\(", hdr))
    assert {"gram", "tall_nn", "residual", "col_sumsq", "fill_uniform", "spmm_stencil", "spmm_csr", "spmm_diag"} <= kern
    for p in "sdcz":
        names |= {f"lb2_{p}_{k}" for k in kern}
        names |= {f"{p}_lobpcg", f"{p}_ilobpcg", f"lb2_{p}_state_alloc", f"lb2_{p}_state_free"}
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, f"declared in include/ but not exported: {missing}"
    assert len(names) >= 80


REFERENCE_HELPERS = ["get_residual", "get_residual_norm", "rayleigh_ritz", "rayleigh_ritz_modified", "svqb", "svqb_mat",
                     "ortho_drop", "ortho_indefinite", "ortho_indefinite_mat", "indefinite_rayleigh_ritz",
                     "indefinite_rayleigh_ritz_modified", "apply_block_op", "gram_self", "gram_cross", "gram_self_mat",
                     "gram_cross_mat", "fill_random", "estimate_norm"]


def test_all_80_reference_symbols_are_exported_and_declared():
    """2 solver entry points + 18 helpers, for each of s/d/c/z (reference lobpcg.h:63-92, 98-555)."""
    L = api.lib()
    hdr = (INC / "lobpcg.h").read_text()
    want = [f"{p}_{h}" for p in "sdcz" for h in ["lobpcg", "ilobpcg"] + REFERENCE_HELPERS]
    assert len(want) == 80
    assert not [n for n in want if not hasattr(L, n)]
    for h in REFERENCE_HELPERS:
        assert f"P##_{h}(" in hdr, f"{h} is not declared in include/lobpcg.h"


@pytest.mark.skipif(not REFROOT.exists(), reason="reference tree only exists in the build container")
def test_reference_test_programs_link_against_the_product_library():
    """The reference's own tests, compiled unchanged against the reference's headers, resolve every symbol from
    liblobpcg_b200.so (they are RUN on the GPU box by tests/test_gpu_reftests.py)."""
    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "reftests"], check=True)
    out = ROOT / "oracle" / "_ref" / "reftests"
    for t in ["test_gram", "test_residual", "test_svqb", "test_ortho_drop", "test_rayleigh_ritz", "test_indefinite_rr",
              "test_lobpcg", "test_ilobpcg"]:
        exe = out / t
        assert exe.exists()
        needed = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
        assert "liblobpcg_b200.so" in needed and "libref_lobpcg" not in needed


def test_state_struct_matches_ctypes_mirror(tmp_path):
    ours = _layout('#include "lobpcg.h"', [f"-I{INC}"], tmp_path, "ours")
    for p in "sdcz":
        S = api._state_struct(p)
        assert int(ours[f"{p}_lobpcg_t"]) == C.sizeof(S)
        for f, _ in S._fields_:
            assert int(ours[f"{p}_lobpcg_t.{f}"]) == getattr(S, f).offset, (p, f)
    assert int(ours["LinearOperator_d_t"]) == C.sizeof(api.LinOpStruct)
    assert int(ours["linop_ctx_t"]) == C.sizeof(api.LinOpCtx)


@pytest.mark.skipif(not REFROOT.exists(), reason="reference tree only exists in the build container")
def test_state_struct_matches_reference_header(tmp_path):
    ours = _layout('#include "lobpcg.h"', [f"-I{INC}"], tmp_path, "ours")
    ref = _layout('#include "lobpcg/linop.h"\n#include "lobpcg.h"',
                  [f"-I{REFROOT}", f"-I{REFROOT}/include", f"-I{REFROOT}/include/lobpcg"], tmp_path, "ref")
    assert ours == ref


def test_c11_generic_front_end_compiles_and_links(tmp_path):
    """A reference-style C11 caller (lobpcg(alg) through _Generic, linop_create, <p>_lobpcg_alloc) builds
    against include/lobpcg.h and links to the shared library; it is only run on a GPU box (test_gpu_solver)."""
    src = ROOT / "tests" / "c_caller" / "caller.c"
    exe = tmp_path / "caller"
    libdir = api.LIB_PATH.parent
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", f"-I{INC}", str(src), "-o", str(exe), f"-L{libdir}",
                    "-llobpcg_b200", f"-Wl,-rpath,{libdir}", "-lm"], check=True)
    assert exe.exists()


def test_no_cpu_fallback_without_library(monkeypatch, tmp_path):
    monkeypatch.setattr(api, "_lib", None)
    monkeypatch.setattr(api, "LIB_PATH", tmp_path / "missing.so")
    with pytest.raises(api.LobpcgB200Error):
        api.lib()


def test_product_never_imports_oracle():
    for py in (ROOT / "lobpcg_b200").rglob("*.py"):
        txt = py.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", txt, flags=re.S).replace("# oracle", ""), py
    for cu in (ROOT / "lobpcg_b200" / "csrc").iterdir():
        assert "oracle/" not in cu.read_text(), cu


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.complex64, np.complex128])
def test_eigenpair_write_out_round_trips(dt, tmp_path):
    """lb2_write_mtx (host-only): Matrix Market dense array files read back by scipy bit-exactly."""
    import scipy.io
    rng = np.random.default_rng(5)
    a = rng.standard_normal((37, 5)) + (1j * rng.standard_normal((37, 5)) if np.dtype(dt).kind == "c" else 0)
    a = a.astype(dt)
    api.write_mtx(tmp_path / "x.mtx", a)
    back = scipy.io.mmread(str(tmp_path / "x.mtx"))
    assert back.shape == a.shape and np.array_equal(back.astype(dt), a)
    lam = np.sort(rng.standard_normal(7)).astype(np.dtype(dt).char.lower() if np.dtype(dt).kind == "c" else dt)
    api.write_mtx(tmp_path / "lam.mtx", lam)
    assert np.array_equal(scipy.io.mmread(str(tmp_path / "lam.mtx")).ravel().astype(lam.dtype), lam)
    with pytest.raises(api.LobpcgB200Error):
        api.write_mtx(tmp_path / "no_such_dir" / "x.mtx", a)

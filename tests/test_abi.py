"""CPU: the C-ABI library loads and exports every symbol include/bbo_b200.h declares; argument
validation paths that need no GPU behave; the product path refuses to run without CUDA."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "bbo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(bbo_[a-z0-9_]+)\s*\(", src)
    return sorted({n for n in names if n != "bbo_hyp_len"})


def test_every_declared_symbol_is_exported_and_bound():
    from bbo_b200 import _lib
    lib = _lib.load()
    declared = header_functions()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/bbo_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == declared


def test_sizes_and_padding():
    from bbo_b200 import _lib
    lib = _lib.load()
    assert lib.bbo_version() >= 100
    assert [lib.bbo_padded_n(n) for n in (1, 128, 129, 4096)] == [128, 128, 256, 4096]
    assert lib.bbo_gp_fit_workspace_bytes(0, 3, 1) == 0
    b1 = lib.bbo_gp_fit_workspace_bytes(256, 10, 1)
    b4 = lib.bbo_gp_fit_workspace_bytes(256, 10, 4)
    assert b1 >= 4 * 256 * 256 * 8 and 3.9 * b1 < b4 < 4.1 * b1
    s64 = lib.bbo_gp_score_workspace_bytes(4096, 20, 8192, 0)
    assert s64 >= 8192 * 4096 * 8
    assert lib.bbo_gp_predict_workspace_bytes(4096, 20, 8192, 1) > 1.9 * s64 - (1 << 22)


def test_struct_layouts_match_header():
    from bbo_b200 import _lib
    assert C.sizeof(_lib.GpHyper) == 24
    assert C.sizeof(_lib.Acq) == 32
    assert C.sizeof(_lib.GpState) == 24 + 8 + 5 * 8


def test_argument_validation_without_gpu():
    from bbo_b200 import _lib
    lib = _lib.load()
    h = _lib.GpHyper(kind=7, d=3, pairwise_eps=0.0, noise=1e-6)
    rc = lib.bbo_gp_fit_value_grad(C.byref(h), 10, 1, None, None, None, None, None, None, None, 0, None)
    assert rc == 1  # BBO_ERR_BAD_SHAPE: unknown kernel kind
    h.kind = 0
    rc = lib.bbo_gp_fit_value_grad(C.byref(h), 10, 1, None, None, None, None, None, None, None, 0, None)
    assert rc == 2  # BBO_ERR_BAD_ARG: null pointers
    a = _lib.Acq(kind=9)
    assert lib.bbo_acq_eval(C.byref(a), None, None, None, 0, 0, None) == 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    import bbo_b200 as B
    from bbo_b200 import _lib
    X, Y = torch.rand(8, 2, dtype=torch.float64), torch.rand(8, 1, dtype=torch.float64)
    with pytest.raises(_lib.BboError):
        B.CudaGP(X, Y)
    with pytest.raises(_lib.BboError):
        B.EI(0.0)(torch.zeros(3), torch.ones(3))
    with pytest.raises(_lib.BboError):
        B.RBFKernel()(X, X)


def test_product_code_never_imports_oracle():
    pkg = os.path.join(ROOT, "bbo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle", ""), f"{f} mentions oracle"


def test_header_is_plain_c_and_a_c_program_links(tmp_path):
    """include/bbo_b200.h is the contract a non-Python host binds (cgo / JNI / plain C): it must compile as strict C99
    and a C program must link against the shared library and call the GPU-free entry points."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from bbo_b200 import _lib
    src = tmp_path / "consumer.c"
    src.write_text(r'''
#include <stdio.h>
#include "bbo_b200.h"
int main(void) {
  bbo_gp_hyper h;
  bbo_acq a;
  h.kind = BBO_KERNEL_RBF; h.d = 3; h.pairwise_eps = 0.0; h.noise = 1e-6;
  a.kind = BBO_ACQ_EI;
  (void)a;
  /* null pointers: validation must refuse before touching the device */
  int rc = bbo_gp_fit_value_grad(&h, 10, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0);
  printf("%d %lld %lld %d %d\n", (int)bbo_version(), (long long)bbo_padded_n(129),
         (long long)bbo_tf32_rows(300), rc, (int)bbo_refine_shortlist(8));
  return 0;
}
''')
    exe = tmp_path / "consumer"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
           str(src), "-o", str(exe), "-L", libdir, "-lbbo_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    ver, pad, rows, rc, m = out.stdout.split()
    assert int(ver) >= 100 and int(pad) == 256 and int(rows) == 512 and int(rc) == 2 and int(m) == 64

"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/impes_b200.h declares, the ctypes mirror matches the C structs, and compute entry points fail
loudly (never fall back) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "impes_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(impes_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(pkg, native_lib):
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(native_lib, n), f"libimpes_b200.so does not export {n}"
    assert set(pkg.binding.EXPORTS) == set(names)


def test_struct_layout_matches_header(pkg, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include "%s"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%%zu %%zu %%zu %%zu %%zu\\n",'
                   'sizeof(impes_desc),sizeof(impes_solve_stats),sizeof(impes_dims),offsetof(impes_desc,injection),'
                   'offsetof(impes_desc,darcy_maxiter));return 0;}' % HEADER)
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", str(src), "-o", str(exe)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    abi = pkg.abi
    assert out == [C.sizeof(abi.Desc), C.sizeof(abi.SolveStats), C.sizeof(abi.Dims), abi.Desc.injection.offset,
                   abi.Desc.darcy_maxiter.offset]


def test_defaults_are_the_reference_defaults(pkg, native_lib):
    """Param::Param / DarcySolverParam / Well defaults [REF src/param.cpp:8-13,26-30,35-68]."""
    d = pkg.abi.Desc()
    native_lib.impes_desc_defaults(C.byref(d))
    assert (d.dim, d.nx, d.ny, d.nz) == (2, 100, 100, 100)
    assert (d.sx, d.sy, d.sz) == (100.0, 100.0, 100.0)
    assert (d.order_v, d.order_p, d.order_s) == (1, 1, 0)
    assert d.two_phase == 0 and d.ode_solver_type == 1
    assert (d.darcy_maxiter, d.darcy_rtol, d.darcy_atol, d.darcy_print_level) == (1000, 1e-6, 1e-10, 0)
    assert (d.inflow, d.outflow, d.saturation_source, d.Sor, d.Swc) == (1.0, -1.0, 1.0, 0.0, 0.0)
    assert (d.mu_w, d.mu_o) == (1.0, 1.0)
    assert d.injection.radius == 1.0 and d.production.height == 1.0 and list(d.injection.center) == [0, 0, 0]
    py = pkg.abi.default_desc()
    assert bytes(py) == bytes(d)


def test_bad_descriptor_is_rejected(pkg, native_lib):
    d = pkg.abi.default_desc()
    d.struct_size = 7
    h = C.c_void_p()
    K = np.ones(4)
    p = K.ctypes.data_as(C.POINTER(C.c_double))
    assert native_lib.impes_create(C.byref(d), p, p, None, None, C.byref(h)) == pkg.abi.IMPES_ERR_BAD_ARG
    assert b"struct_size" in native_lib.impes_last_error(None)


def test_no_cpu_fallback(pkg, native_lib):
    """Without a CUDA device impes_create must fail with IMPES_ERR_CUDA -- there is no CPU path to fall to."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present; the fail-loud path is exercised on CPU-only hosts")
    except ImportError:
        pass
    from incompressible_flow_b200 import workloads
    d, Kx, Ky, Kz, phi = workloads.build(dim=2, nx=8, ny=8, K="homogeneous")
    with pytest.raises(pkg.binding.ImpesError) as e:
        pkg.Solver(d, Kx, Ky, Kz, phi)
    assert e.value.code == pkg.abi.IMPES_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """The product tree must not import, link or call anything under oracle/."""
    bad = []
    pdir = os.path.join(ROOT, "incompressible-flow_b200")
    for dp, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"oracle/|libimpes_oracle|load_oracle|\borc_[a-z_]+\s*\(|impes_oracle", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, f"product files reference the oracle: {bad}"

"""ctypes wrapper of the CPU oracle (oracle/libimpes_oracle.so).

TEST INFRASTRUCTURE ONLY -- see the header of impes_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by
the product package.  PARITY UNPINNED (the reference has no tests or golden vectors and
MFEM/HYPRE cannot be built here); pinned by the analytic invariants in tests/.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from __graft_entry__ import load_package  # noqa: E402

abi = load_package().abi  # PODs only; loads no native code

_LIB = None
_dp = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_HERE, "libimpes_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("impes_oracle.c", "orc_mg.inc", "orc_q1.inc")] + [
        os.path.join(os.path.dirname(_HERE), "include", "impes_b200.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libimpes_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(abi.Desc), _dp, _dp, _dp, _dp]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_threads_det.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_dot_mode.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_mass_cg.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_arith.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_dims.argtypes = [C.c_void_p, C.POINTER(abi.Dims)]
        L.orc_pressure_solve.argtypes = [C.c_void_p, _dp, _dp, C.POINTER(abi.SolveStats)]
        L.orc_update_coefficients.argtypes = [C.c_void_p, _dp]
        L.orc_darcy_apply.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_prec_apply.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_get_rhs.argtypes = [C.c_void_p, _dp]
        L.orc_get_diag_M.argtypes = [C.c_void_p, _dp]
        L.orc_get_csr.restype = C.c_long
        L.orc_get_csr.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_int), _dp]
        L.orc_saturation_advance.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_double, C.POINTER(C.c_int)]
        L.orc_guards.argtypes = [C.c_void_p, _dp, C.POINTER(C.c_long), _dp]
        L.orc_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Oracle:
    def __init__(self, desc, Kx, Ky, Kz=None, phi=None):
        self.L = lib()
        self.desc = desc
        self._keep = [_f64(Kx), _f64(Ky), _f64(Kz), _f64(phi)]
        self.h = self.L.orc_create(C.byref(desc), *[_p(a) for a in self._keep])
        if not self.h:
            raise RuntimeError("orc_create failed (bad descriptor / unsupported orders)")
        d = abi.Dims()
        self.L.orc_get_dims(self.h, C.byref(d))
        self.n_cells, self.nv, self.np_, self.ns = d.n_cells, d.n_v, d.n_p, d.n_s
        self.nvc = list(d.n_v_comp)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_threads(self, t):
        self.L.orc_set_threads(self.h, int(t))

    def set_threads_det(self, t):
        """OpenMP threads with deterministic arithmetic: every canonical (arith == 1) path keeps its bits"""
        self.L.orc_set_threads_det(self.h, int(t))

    def set_dot_mode(self, m):
        self.L.orc_set_dot_mode(self.h, int(m))

    def set_arith(self, m):
        """0 = MFEM-like (independent formulation), 1 = canonical (bit-comparable with the GPU path)"""
        self.L.orc_set_arith(self.h, int(m))

    def set_mass_cg(self, on):
        self.L.orc_set_mass_cg(self.h, int(on))

    def _chk_S(self, S):
        if S.shape[0] != self.ns:
            raise ValueError(f"saturation vector has {S.shape[0]} entries, the DG space has {self.ns}")

    def pressure_solve(self, S):
        S = _f64(S)
        self._chk_S(S)
        x = np.zeros(self.nv + self.np_)
        st = abi.SolveStats()
        rc = self.L.orc_pressure_solve(self.h, _p(S), _p(x), C.byref(st))
        return x, st, rc

    def update_coefficients(self, S):
        S = _f64(S)
        self._chk_S(S)
        rc = self.L.orc_update_coefficients(self.h, _p(S))
        if rc:
            raise RuntimeError(self.L.orc_last_error(self.h).decode())

    def darcy_apply(self, x):
        x = _f64(x)
        y = np.zeros_like(x)
        assert self.L.orc_darcy_apply(self.h, _p(x), _p(y)) == 0
        return y

    def prec_apply(self, r):
        r = _f64(r)
        y = np.zeros_like(r)
        assert self.L.orc_prec_apply(self.h, _p(r), _p(y)) == 0
        return y

    def rhs(self):
        r = np.zeros(self.nv + self.np_)
        self.L.orc_get_rhs(self.h, _p(r))
        return r

    def diag_M(self):
        m = np.zeros(self.nv)
        assert self.L.orc_get_diag_M(self.h, _p(m)) == 0
        return m

    def csr(self, which):
        """which: 'M' | 'B' | 'S' -> scipy.sparse.csr_matrix"""
        import scipy.sparse as sp
        w = {"M": 0, "B": 1, "S": 2}[which]
        nnz = self.L.orc_get_csr(self.h, w, None, None, None)
        nr = {"M": self.nv, "B": self.np_, "S": self.np_}[which]
        nc = {"M": self.nv, "B": self.nv, "S": self.np_}[which]
        I = np.zeros(nr + 1, dtype=np.int64)
        J = np.zeros(nnz, dtype=np.int32)
        A = np.zeros(nnz)
        self.L.orc_get_csr(self.h, w, I.ctypes.data_as(C.POINTER(C.c_long)), J.ctypes.data_as(C.POINTER(C.c_int)), _p(A))
        return sp.csr_matrix((A, J, I), shape=(nr, nc))

    def saturation_advance(self, u, S, global_dt, dt_local):
        u = _f64(u)
        S = np.array(S, dtype=np.float64, copy=True)
        n = C.c_int(0)
        rc = self.L.orc_saturation_advance(self.h, _p(u), _p(S), float(global_dt), float(dt_local), C.byref(n))
        if rc:
            raise RuntimeError(self.L.orc_last_error(self.h).decode())
        return S, n.value

    def guards(self, S):
        S = _f64(S)
        bad = C.c_long(0)
        mx = C.c_double(0)
        self.L.orc_guards(self.h, _p(S), C.byref(bad), C.byref(mx))
        return bad.value, mx.value

"""ctypes binding of libimpes_b200.so (the C ABI in include/impes_b200.h).

Used by tests/, bench.py and __graft_entry__.py.  Every array crossing this boundary is a host
numpy array; device memory, streams and NCCL live inside the native library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_LIB = None
_dp = C.POINTER(C.c_double)

EXPORTS = [
    "impes_abi_version", "impes_desc_defaults", "impes_create", "impes_destroy", "impes_last_error",
    "impes_get_dims", "impes_comm_unique_id", "impes_comm_init", "impes_pressure_solve",
    "impes_saturation_advance", "impes_guards", "impes_set_field", "impes_get_field",
    "impes_update_coefficients", "impes_darcy_apply", "impes_prec_apply", "impes_get_rhs",
    "impes_get_local_dims", "impes_pressure_solve_local", "impes_saturation_advance_local",
    "impes_set_field_local", "impes_get_field_local",
    "impes_time_kernel", "impes_event_record", "impes_event_elapsed_ms", "impes_launch_count", "impes_describe", "impes_profile_enable", "impes_profile_read",
    "impes_profile_read_kernels",
]


def library_path():
    # IMPES_B200_LIB: development override (A/B builds of the same CUDA library with other tuning macros)
    return os.environ.get("IMPES_B200_LIB") or os.path.join(_HERE, "libimpes_b200.so")


def build_native(force=False, verbose=False):
    """Compile csrc/ (sm_100a) and host/ in-tree via the Makefile."""
    cmd = ["make", "-C", _HERE] + (["-B"] if force else [])
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode:
        print(out.stdout)
    if out.returncode:
        raise RuntimeError("native build failed")
    return library_path()


def load_library():
    """dlopen the native library and declare every export of include/impes_b200.h.
    Raises (never falls back) if the library or a symbol is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python __graft_entry__.py build` "
                           "(there is no CPU fallback for the IMPES hot path)")
    L = C.CDLL(path)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise RuntimeError(f"libimpes_b200.so does not export {name}")
    H = C.c_void_p
    L.impes_abi_version.restype = C.c_int
    L.impes_desc_defaults.argtypes = [C.POINTER(abi.Desc)]
    L.impes_create.argtypes = [C.POINTER(abi.Desc), _dp, _dp, _dp, _dp, C.POINTER(H)]
    L.impes_destroy.argtypes = [H]
    L.impes_last_error.restype = C.c_char_p
    L.impes_last_error.argtypes = [H]
    L.impes_get_dims.argtypes = [H, C.POINTER(abi.Dims)]
    L.impes_comm_unique_id.argtypes = [C.c_void_p]
    L.impes_comm_init.argtypes = [H, C.c_void_p]
    L.impes_pressure_solve.argtypes = [H, _dp, _dp, C.POINTER(abi.SolveStats)]
    L.impes_saturation_advance.argtypes = [H, _dp, _dp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int)]
    L.impes_guards.argtypes = [H, C.POINTER(C.c_int64), _dp]
    L.impes_get_local_dims.argtypes = [H, C.POINTER(abi.LocalDims)]
    L.impes_pressure_solve_local.argtypes = [H, _dp, _dp, C.POINTER(abi.SolveStats)]
    L.impes_saturation_advance_local.argtypes = [H, _dp, _dp, C.c_int, C.c_double, C.c_double, C.POINTER(C.c_int)]
    L.impes_set_field_local.argtypes = [H, C.c_int, _dp]
    L.impes_get_field_local.argtypes = [H, C.c_int, _dp]
    L.impes_set_field.argtypes = [H, C.c_int, _dp]
    L.impes_get_field.argtypes = [H, C.c_int, _dp]
    L.impes_update_coefficients.argtypes = [H]
    L.impes_darcy_apply.argtypes = [H, _dp, _dp]
    L.impes_prec_apply.argtypes = [H, _dp, _dp]
    L.impes_get_rhs.argtypes = [H, _dp]
    L.impes_time_kernel.argtypes = [H, C.c_int, C.c_int, C.c_int, _dp, _dp]
    L.impes_event_record.argtypes = [H, C.c_int]
    L.impes_event_elapsed_ms.argtypes = [H, C.c_int, C.c_int, _dp]
    L.impes_launch_count.restype = C.c_int64
    L.impes_launch_count.argtypes = [H]
    L.impes_describe.argtypes = [H, C.c_char_p, C.c_int]
    L.impes_profile_enable.argtypes = [H, C.c_int]
    L.impes_profile_read.argtypes = [H, _dp, C.POINTER(C.c_int64)]
    L.impes_profile_read_kernels.argtypes = [H, C.c_int, C.POINTER(C.c_char_p), _dp, C.POINTER(C.c_int64)]
    if L.impes_abi_version() != 1:
        raise RuntimeError("impes ABI version mismatch")
    _LIB = L
    return L


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class ImpesError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"impes error {code}: {msg}")
        self.code = code


class Solver:
    """Thin OO wrapper of one impes_handle."""

    def __init__(self, desc, Kx, Ky, Kz=None, phi=None):
        self.L = load_library()
        self.desc = desc
        self.h = C.c_void_p()
        arrs = [_f64(Kx), _f64(Ky), _f64(Kz), _f64(phi)]
        rc = self.L.impes_create(C.byref(desc), *[_p(a) for a in arrs], C.byref(self.h))
        if rc:
            raise ImpesError(rc, self.L.impes_last_error(None).decode())
        d = abi.Dims()
        self.L.impes_get_dims(self.h, C.byref(d))
        self.dims = d
        self.n_cells, self.nv, self.np_, self.ns = d.n_cells, d.n_v, d.n_p, d.n_s

    def close(self):
        if getattr(self, "h", None):
            self.L.impes_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc, allow=()):
        if rc and rc not in allow:
            raise ImpesError(rc, self.L.impes_last_error(self.h).decode())
        return rc

    # --- multi-GPU bootstrap
    def unique_id(self):
        buf = (C.c_char * 128)()
        self._chk(self.L.impes_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, uid: bytes):
        buf = (C.c_char * 128).from_buffer_copy(uid)
        self._chk(self.L.impes_comm_init(self.h, buf))

    # --- solver calls
    def pressure_solve(self, S=None, out=None, want_x=True, allow_not_converged=False):
        S = _f64(S)
        x = out if out is not None else (np.zeros(self.nv + self.np_) if want_x else None)
        st = abi.SolveStats()
        allow = (abi.IMPES_ERR_NOT_CONVERGED,) if allow_not_converged else ()
        self._chk(self.L.impes_pressure_solve(self.h, _p(S), _p(x), C.byref(st)), allow)
        return x, st

    def saturation_advance(self, V=None, S=None, global_ti=1, global_dt=100.0, dt_local=1.0, inplace=False):
        V = _f64(V)
        if S is not None:
            S = S if inplace else np.array(S, dtype=np.float64, copy=True)
        n = C.c_int(0)
        self._chk(self.L.impes_saturation_advance(self.h, _p(V), _p(S), int(global_ti), float(global_dt),
                                                  float(dt_local), C.byref(n)))
        return S, n.value

    # --- slab-local variants (each rank passes only what it owns; world_size == 1: local == global)
    def local_dims(self):
        d = abi.LocalDims()
        self._chk(self.L.impes_get_local_dims(self.h, C.byref(d)))
        return d

    def pressure_solve_local(self, S_local=None, out=None, allow_not_converged=False):
        """out: local [V | P] array (n_v + n_p of local_dims) or None to keep the result on the device"""
        st = abi.SolveStats()
        allow = (abi.IMPES_ERR_NOT_CONVERGED,) if allow_not_converged else ()
        self._chk(self.L.impes_pressure_solve_local(self.h, _p(_f64(S_local)), _p(out), C.byref(st)), allow)
        return out, st

    def saturation_advance_local(self, V_local=None, S_local=None, global_ti=1, global_dt=100.0, dt_local=1.0):
        """S_local is updated in place"""
        n = C.c_int(0)
        self._chk(self.L.impes_saturation_advance_local(self.h, _p(_f64(V_local)), _p(S_local), int(global_ti),
                                                        float(global_dt), float(dt_local), C.byref(n)))
        return S_local, n.value

    def set_field_local(self, which, a):
        self._chk(self.L.impes_set_field_local(self.h, which, _p(_f64(a))))

    def get_field_local(self, which):
        d = self.local_dims()
        n = {abi.FIELD_SATURATION: d.n_s, abi.FIELD_VELOCITY: d.n_v, abi.FIELD_PRESSURE: d.n_p,
             abi.FIELD_DARCY_X: d.n_v + d.n_p}[which]
        a = np.zeros(n)
        self._chk(self.L.impes_get_field_local(self.h, which, _p(a)))
        return a

    def guards(self):
        bad = C.c_int64(0)
        mx = C.c_double(0)
        self._chk(self.L.impes_guards(self.h, C.byref(bad), C.byref(mx)))
        return bad.value, mx.value

    def set_field(self, which, a):
        a = _f64(a)
        self._chk(self.L.impes_set_field(self.h, which, _p(a)))

    def get_field(self, which):
        n = {abi.FIELD_SATURATION: self.ns, abi.FIELD_VELOCITY: self.nv, abi.FIELD_PRESSURE: self.np_,
             abi.FIELD_DARCY_X: self.nv + self.np_}[which]
        a = np.zeros(n)
        self._chk(self.L.impes_get_field(self.h, which, _p(a)))
        return a

    # --- operator-level (parity tests)
    def update_coefficients(self, S=None):
        if S is not None:
            self.set_field(abi.FIELD_SATURATION, S)
        self._chk(self.L.impes_update_coefficients(self.h))

    def darcy_apply(self, x):
        x = _f64(x)
        y = np.zeros_like(x)
        self._chk(self.L.impes_darcy_apply(self.h, _p(x), _p(y)))
        return y

    def prec_apply(self, r):
        r = _f64(r)
        y = np.zeros_like(r)
        self._chk(self.L.impes_prec_apply(self.h, _p(r), _p(y)))
        return y

    def rhs(self):
        r = np.zeros(self.nv + self.np_)
        self._chk(self.L.impes_get_rhs(self.h, _p(r)))
        return r

    # --- measurement
    def time_kernel(self, kernel, reps=20, flush_l2=True):
        a = C.c_double(0)
        m = C.c_double(0)
        self._chk(self.L.impes_time_kernel(self.h, kernel, reps, int(flush_l2), C.byref(a), C.byref(m)))
        return a.value, m.value

    def event_record(self, slot):
        self._chk(self.L.impes_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_double(0)
        self._chk(self.L.impes_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(self.L.impes_launch_count(self.h))

    def describe(self):
        """one line naming the kernel variants this handle selected (impes_describe)"""
        buf = C.create_string_buffer(1024)
        self.L.impes_describe(self.h, buf, 1024)
        return buf.value.decode()

    def profile_enable(self, on=True):
        self._chk(self.L.impes_profile_enable(self.h, int(on)))

    def profile_read(self):
        ms = (C.c_double * 8)()
        n = (C.c_int64 * 8)()
        self._chk(self.L.impes_profile_read(self.h, ms, n))
        return list(ms), list(n)

    def profile_read_kernels(self):
        """-> {kernel name: (ms, launches)} since profile_enable (only kernels that launched)"""
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        n = (C.c_int64 * cap)()
        k = self.L.impes_profile_read_kernels(self.h, cap, names, ms, n)
        return {names[i].decode(): (ms[i], int(n[i])) for i in range(min(k, cap)) if n[i]}

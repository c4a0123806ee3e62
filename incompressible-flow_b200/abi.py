"""ctypes mirror of include/impes_b200.h (PODs and constants only; loads nothing)."""
import ctypes as C

IMPES_OK = 0
IMPES_ERR_BAD_ARG = 1
IMPES_ERR_NOT_CONVERGED = 2
IMPES_ERR_NONFINITE = 3
IMPES_ERR_CUDA = 4
IMPES_ERR_NCCL = 5
IMPES_ERR_UNSUPPORTED = 6

SCHUR_SGS_LEX = 0
SCHUR_MG = 1
SUBSTEPS_SERIAL = 0
SUBSTEPS_PARALLEL = 1

FIELD_SATURATION = 0
FIELD_VELOCITY = 1
FIELD_PRESSURE = 2
FIELD_DARCY_X = 3

KERNEL_DARCY_APPLY = 1
KERNEL_SAT_SUBSTEP = 2
KERNEL_MINRES_UPDATE = 3
KERNEL_PREC_APPLY = 4
KERNEL_COEFF = 5


class Well(C.Structure):
    _fields_ = [("center", C.c_double * 3), ("radius", C.c_double), ("height", C.c_double)]


class Desc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("dim", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("order_v", C.c_int32), ("order_p", C.c_int32), ("order_s", C.c_int32),
        ("sx", C.c_double), ("sy", C.c_double), ("sz", C.c_double),
        ("two_phase", C.c_int32), ("ode_solver_type", C.c_int32),
        ("Sor", C.c_double), ("Swc", C.c_double),
        ("mu_w", C.c_double), ("mu_o", C.c_double),
        ("injection", Well), ("production", Well),
        ("inflow", C.c_double), ("outflow", C.c_double),
        ("saturation_source", C.c_double),
        ("darcy_rtol", C.c_double), ("darcy_atol", C.c_double),
        ("darcy_maxiter", C.c_int32), ("darcy_print_level", C.c_int32),
        ("schur_prec", C.c_int32), ("substep_rule", C.c_int32),
        ("rank", C.c_int32), ("world_size", C.c_int32),
        ("device", C.c_int32), ("mg_smooth", C.c_int32),
    ]


class SolveStats(C.Structure):
    _fields_ = [
        ("iterations", C.c_int32), ("converged", C.c_int32),
        ("final_norm", C.c_double), ("initial_norm", C.c_double),
        ("seconds", C.c_double), ("device_ms", C.c_double),
    ]


class Dims(C.Structure):
    _fields_ = [
        ("n_cells", C.c_int64),
        ("n_v", C.c_int64), ("n_p", C.c_int64), ("n_s", C.c_int64),
        ("n_v_comp", C.c_int64 * 3),
        ("slab_lo", C.c_int32), ("slab_hi", C.c_int32), ("pad_", C.c_int32),
    ]


class LocalDims(C.Structure):
    _fields_ = [
        ("n_v", C.c_int64), ("n_p", C.c_int64), ("n_s", C.c_int64),
        ("n_v_comp", C.c_int64 * 3), ("v_comp_start", C.c_int64 * 3),
        ("p_start", C.c_int64), ("s_start", C.c_int64),
    ]


def default_desc() -> Desc:
    """Reference defaults [REF src/param.cpp:8-13,35-68]; MU_W = MU_O = 1 [REF config.hpp.in:37,43]."""
    d = Desc()
    d.struct_size = C.sizeof(Desc)
    d.dim = 2
    d.nx = d.ny = d.nz = 100
    d.sx = d.sy = d.sz = 100.0
    d.order_v, d.order_p, d.order_s = 1, 1, 0
    d.two_phase = 0
    d.ode_solver_type = 1
    d.Sor = d.Swc = 0.0
    d.mu_w = d.mu_o = 1.0
    for w in (d.injection, d.production):
        w.center[0] = w.center[1] = w.center[2] = 0.0
        w.radius = 1.0
        w.height = 1.0
    d.inflow, d.outflow, d.saturation_source = 1.0, -1.0, 1.0
    d.darcy_rtol, d.darcy_atol = 1e-6, 1e-10
    d.darcy_maxiter, d.darcy_print_level = 1000, 0
    d.schur_prec = SCHUR_SGS_LEX
    d.substep_rule = SUBSTEPS_SERIAL
    d.rank, d.world_size, d.device, d.mg_smooth = 0, 1, 0, 0
    return d

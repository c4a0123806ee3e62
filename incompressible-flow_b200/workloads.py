"""Synthetic workloads: the BASELINE.json configs C1-C5 as descriptors + K fields.

K fields follow the reference's input contract: float32 raw, x fastest then y then z
[REF scripts/random_field.py:10-18], widened to fp64 exactly like read_binary
[REF src/utilities.cpp:43-51]; z-extrusion repeats one xy-layer [REF scripts/extrude2D.sh:8-11].
"""
import numpy as np

from . import abi


def lognormal_field(n_cells, sigma=1.0, seed=12345):
    z = np.random.default_rng(seed).standard_normal(n_cells)
    return np.exp(sigma * z).astype(np.float32).astype(np.float64)


def make_desc(dim, nx, ny, nz=1, sx=None, sy=None, sz=None, order=0, two_phase=True,
              maxiter=20000, rtol=1e-6, atol=1e-10, schur_prec=abi.SCHUR_SGS_LEX, **kw):
    """Quarter-five-spot style descriptor with single-cell-column wells."""
    d = abi.default_desc()
    d.dim, d.nx, d.ny, d.nz = dim, nx, ny, (nz if dim == 3 else 1)
    d.sx = float(sx if sx is not None else (100.0 if dim == 2 else nx))
    d.sy = float(sy if sy is not None else (100.0 if dim == 2 else ny))
    d.sz = float(sz if sz is not None else (nz if dim == 3 else 100.0))
    d.order_v = d.order_p = d.order_s = order
    d.two_phase = int(two_phase)
    d.darcy_maxiter, d.darcy_rtol, d.darcy_atol = maxiter, rtol, atol
    d.schur_prec = schur_prec
    hx, hy = d.sx / nx, d.sy / ny
    if dim == 2:
        r = min(hx, hy)
        d.injection.center[0], d.injection.center[1] = 0.0, 0.0
        d.production.center[0], d.production.center[1] = d.sx, d.sy
        d.injection.radius = d.production.radius = r
    else:
        hz = d.sz / nz
        r = min(hx, hz)
        d.injection.center[0] = d.injection.center[1] = d.injection.center[2] = 0.0
        d.production.center[0], d.production.center[1], d.production.center[2] = d.sx, 0.0, d.sz
        d.injection.radius = d.production.radius = r
        d.injection.height = d.production.height = d.sy
    for k, v in kw.items():
        setattr(d, k, v)
    return d


CONFIGS = {
    "C1": dict(dim=2, nx=64, ny=64, K="homogeneous", note="2D quarter-five-spot 64x64 RT0/P0 homogeneous"),
    "C2": dict(dim=2, nx=512, ny=512, K="lognormal", sigma=1.0, note="2D 512x512 log-normal RT0/P0"),
    "C3": dict(dim=3, nx=128, ny=128, nz=32, K="lognormal-extruded", sigma=1.0, note="3D 128x128x32 extruded log-normal RT0/P0"),
    "C4": dict(dim=3, nx=60, ny=220, nz=85, K="lognormal", sigma=2.0, kz_ratio=0.1, note="SPE10-shaped 60x220x85 RT0/P0"),
    "C5": dict(dim=3, nx=256, ny=256, nz=256, K="lognormal", sigma=1.0, order=1, note="3D 256^3 RT1/Q1/DG-Q1"),
}


def build(dim, nx, ny, nz=1, K="lognormal", sigma=1.0, kz_ratio=None, order=0, seed=12345, note="", **kw):
    """-> (desc, Kx, Ky, Kz|None, phi)"""
    nz = nz if dim == 3 else 1
    n = nx * ny * nz
    if K == "homogeneous":
        Kx = np.ones(n)
    elif K == "lognormal":
        Kx = lognormal_field(n, sigma, seed)
    elif K == "lognormal-extruded":
        Kx = np.tile(lognormal_field(nx * ny, sigma, seed), nz)
    else:
        raise ValueError(K)
    Ky = Kx.copy()
    Kz = None
    if dim == 3:
        Kz = Kx.copy() if kz_ratio is None else (np.float32(kz_ratio) * Kx.astype(np.float32)).astype(np.float64)
    phi = np.ones(n)
    d = make_desc(dim, nx, ny, nz, order=order, **kw)
    return d, Kx, Ky, Kz, phi


def bench_setup(name_or_cfg):
    """-> (build() keyword dict, dt_global, dt_local) of a benchmark workload; 100 sub-steps per step everywhere.

    2-D: `-dtg 100 -dtl 1`.  3-D with S in DG_0: `-dtg 50 -dtl 0.5` (dtl = 1 is exactly CFL = 1 at the
    well cells and blows up on heterogeneous 3-D fields).  S in DG-Q1 (C5): `-sat-source 0.5 -dtg 25 -dtl 0.25`:
    forward Euler needs a smaller CFL number for linear DG, and with `-sat-source 1` the well cells sit at S = 1,
    where f(S) stops being monotone and the unlimited scheme of the reference runs away (S_max 4..490 within three
    steps, which would trip the reference's own guard [REF src/run_serial.cpp:138]); with 0.5 the well equilibrium
    is f(S) = 0.5, S stays below 0.7 and the MINRES iteration count is stable (41-45)."""
    c = dict(CONFIGS[name_or_cfg]) if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    c.pop("note", None)
    if c["dim"] == 2:
        dtg, dtl = 100.0, 1.0
    elif c.get("order", 0) == 1 and c.get("order_s", 1) == 1:
        dtg, dtl = 25.0, 0.25
        c.setdefault("saturation_source", 0.5)
    else:
        dtg, dtl = 50.0, 0.5
    return c, dtg, dtl


def config(name, **kw):
    c = dict(CONFIGS[name])
    c.update(kw)
    return build(**c)

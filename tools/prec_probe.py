"""Time the block preconditioner (Jacobi + Schur sweep) alone."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from incompressible_flow_b200 import workloads, abi
cfgs = {"C1": dict(dim=2, nx=64, ny=64, K="homogeneous"), "L256": dict(dim=2, nx=256, ny=256, K="lognormal"),
        "C2": dict(dim=2, nx=512, ny=512, K="lognormal"),
        "C3": dict(dim=3, nx=128, ny=128, nz=32, K="lognormal-extruded"),
        "C4": dict(dim=3, nx=60, ny=220, nz=85, K="lognormal", sigma=2.0, kz_ratio=0.1)}
for name in sys.argv[1:] or ["C2"]:
    d, Kx, Ky, Kz, phi = workloads.build(**cfgs[name])
    n = d.nx * d.ny * (d.nz if d.dim == 3 else 1)
    with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
        s.set_field(abi.FIELD_SATURATION, np.full(n, 0.3))
        a, m = s.time_kernel(abi.KERNEL_PREC_APPLY, 20, False)
        nst = d.nx + d.ny + (d.nz if d.dim == 3 else 1) - 2 if d.dim == 3 else d.nx + d.ny - 1
        print(f"{name}: prec apply avg {a*1e3:.1f} us min {m*1e3:.1f} us  ({2*nst} dependent stages -> {m*1e6/(2*nst):.0f} ns/stage)")

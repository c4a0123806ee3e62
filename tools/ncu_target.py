"""Small fixed workload for ncu captures: one IMPES step (pressure solve + 100 sub-steps) of a config."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from incompressible_flow_b200 import workloads, abi
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
cfgs = {"C2": dict(dim=2, nx=512, ny=512, K="lognormal"),
        "C3": dict(dim=3, nx=128, ny=128, nz=32, K="lognormal-extruded"),
        "S256": dict(dim=3, nx=256, ny=256, nz=256, K="lognormal"),
        "S384": dict(dim=3, nx=384, ny=384, nz=384, K="lognormal")}
d, Kx, Ky, Kz, phi = workloads.build(**cfgs[name], maxiter=maxit)
n = d.nx * d.ny * (d.nz if d.dim == 3 else 1)
with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
    s.set_field(abi.FIELD_SATURATION, np.zeros(n))
    _, st = s.pressure_solve(None, want_x=False, allow_not_converged=True)
    s.saturation_advance(None, None, 1, 100.0, 1.0)
    print(name, "iters", st.iterations, "launches", s.launch_count(), "solve_ms", st.device_ms)

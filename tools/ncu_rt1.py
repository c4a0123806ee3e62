"""Short fixed RT1/Q1 workload for ncu captures: a few MINRES iterations (multigrid mode) of an n^3 mesh."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from incompressible_flow_b200 import workloads, abi
n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 128
maxit = int(sys.argv[2]) if len(sys.argv) > 2 else 4
os.environ["IMPES_NO_GRAPH"] = "1"
d, Kx, Ky, Kz, phi = workloads.build(dim=3, nx=n3, ny=n3, nz=n3, K="lognormal", order=1, order_s=0,
                                     schur_prec=abi.SCHUR_MG, maxiter=maxit)
n = n3 ** 3
with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
    s.set_field(abi.FIELD_SATURATION, np.full(n, 0.3))
    _, st = s.pressure_solve(None, want_x=False, allow_not_converged=True)
    s.saturation_advance(None, None, 1, 5.0, 0.5)
    print("rt1", n3, "iters", st.iterations, "launches", s.launch_count(), "solve_ms", st.device_ms)

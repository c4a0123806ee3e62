"""ncu target: a few K6 / K1 launches of the RT0 path on a streaming-size mesh."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from incompressible_flow_b200 import workloads, abi
n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d, Kx, Ky, Kz, phi = workloads.build(dim=3, nx=n3, ny=n3, nz=n3, K="lognormal")
with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
    s.set_field(abi.FIELD_SATURATION, np.full(n3 ** 3, 0.3))
    s.set_field(abi.FIELD_DARCY_X, np.random.default_rng(0).standard_normal(s.nv + s.np_))
    print("K6", s.time_kernel(abi.KERNEL_SAT_SUBSTEP, 3, True))
    print("K1", s.time_kernel(abi.KERNEL_DARCY_APPLY, 3, True))

"""Time K1 / K6 / K2 alone on a streaming-size mesh (L2 flushed before every launch)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package
pkg = load_package()
from incompressible_flow_b200 import workloads, abi
dims = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "384x384x384").split("x")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = dict(dim=len(dims), nx=dims[0], ny=dims[1], K="lognormal")
if len(dims) == 3:
    cfg["nz"] = dims[2]
t0 = time.time()
d, Kx, Ky, Kz, phi = workloads.build(**cfg)
n = int(np.prod(dims))
with pkg.Solver(d, Kx, Ky, Kz, phi) as s:
    s.set_field(abi.FIELD_SATURATION, np.full(n, 0.3))
    x = np.random.default_rng(0).standard_normal(s.nv + s.np_)
    s.set_field(abi.FIELD_DARCY_X, x)
    print(f"setup {time.time()-t0:.1f}s  cells={n:.3e} dofs={s.nv+s.np_:.3e}", flush=True)
    nd = s.nv + s.np_
    for kid, name, b in ((abi.KERNEL_DARCY_APPLY, "K1 darcy_apply+dot", (88 if d.dim == 3 else 64) * n),
                         (abi.KERNEL_SAT_SUBSTEP, "K6 sat_substep", (48 if d.dim == 3 else 40) * n),
                         (abi.KERNEL_MINRES_UPDATE, "K2 minres b+c", 13 * 8 * nd)):
        for fl in (1,):
            a, m = s.time_kernel(kid, reps, fl)
            print(f"{name:22s} flush={fl}: avg {a*1e3:9.1f} us  min {m*1e3:9.1f} us -> {b/(a*1e-3)/1e9:7.0f} GB/s algorithmic "
                  f"({b/(a*1e-3)/1e9/6548.2:.3f} of measured HBM peak)", flush=True)
